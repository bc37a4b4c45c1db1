// Decoder-side building blocks (the consumer right behind the encoder path): first correct CUDA path,
// fp32 on the CUDA cores (parity on a B200: tests/test_gpu_decoder.py, first run round 2).
//
// Reference behaviour replaced: TransformerMultiTask.decode(), /root/reference/stac-st/modules/TransformerMultiTask.py
// :234-271, and the decoder half of forward() :185-209 - NormalizedEmbedding + positional encoding, then SpeechBrain's
// TransformerDecoder (pre-LN layers of causal self-attention, cross-attention over the encoder output, feed-forward),
// which returns the head-averaged cross-attention weights of the last layer next to the prediction.  The GEMMs and
// LayerNorms of that stack are the encoder's entry points (stac_gemm_f32, stac_layernorm); what the encoder kernels do
// not cover is here:
//   * stac_embed_scale_pe : emb[token] * sqrt(d_model) + pe[position]
//   * stac_attention_f32  : attention with separate query and key/value tensors, causal and padding masks, optional
//                           head-averaged weights (torch's need_weights=True output)
// Both are what the reference's forward_step costs per call (the whole prefix, every step, mutitask_decoder.py:119-128);
// decoder.DecoderCache drives the same kernels one token at a time over cached keys / values (the key / value tensor
// is addressed with a batch stride and a row stride so that a time-major cache works); tensor-core versions are the
// next step of this row.
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int kHd = 64;
constexpr int kThreads = 128;

__global__ void __launch_bounds__(256)
embed_scale_pe_kernel(const long long* __restrict__ tokens, const float* __restrict__ emb,
                      const float* __restrict__ pe, long long rows, int seq_len, int d_model, int vocab, float scale,
                      float* __restrict__ out) {
  const int d4 = d_model >> 2;
  const long long n = rows * d4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / d4;
    const int c = (int)(i % d4);
    long long tok = tokens[row];
    tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);          // (the reference raises on such ids; no fault here)
    const float4 e = __ldg(reinterpret_cast<const float4*>(emb + tok * d_model) + c);
    const float4 p = __ldg(reinterpret_cast<const float4*>(pe + (row % seq_len) * d_model) + c);
    reinterpret_cast<float4*>(out + row * d_model)[c] =
        make_float4(fmaf(e.x, scale, p.x), fmaf(e.y, scale, p.y), fmaf(e.z, scale, p.z), fmaf(e.w, scale, p.w));
  }
}

// CTA = one query position of one row; loops over the heads (head_dim 64) so that the head average of the weights is a
// plain sum in shared memory.  Dynamic shared memory: sc[lk] | wacc[lk] | q[64] | part[128] | red[40].
__global__ void __launch_bounds__(kThreads)
attention_f32_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ k,
                     const float* __restrict__ v, long long kv_bs, long long kv_rs, int lq, int lk, int n_head, int mem_rows_div,
                     int causal, const int* __restrict__ kv_len, const long long* __restrict__ key_tokens,
                     long long pad_idx, float* __restrict__ ctx, long long ldctx, float* __restrict__ weights) {
  extern __shared__ float sm[];
  float* sc = sm;
  float* wacc = sm + lk;
  float* qs = wacc + lk;
  float* part = qs + kHd;
  float* red = part + kThreads;
  const long long qrow = blockIdx.x;                 // r * lq + i
  const long long r = qrow / lq;
  const int i = (int)(qrow % lq);
  const long long rb = r / mem_rows_div;             // batch index of the key / value tensor
  int n_keys = lk;
  if (kv_len != nullptr) n_keys = min(max(kv_len[r], 0), lk);
  if (causal) n_keys = min(n_keys, i + 1);
  const int tid = threadIdx.x;
  for (int j = tid; j < lk; j += kThreads) wacc[j] = 0.f;
  const float inv_heads = 1.0f / (float)n_head;
  for (int h = 0; h < n_head; ++h) {
    __syncthreads();                                 // qs / sc / part of the previous head are no longer read
    if (tid < kHd) qs[tid] = q[qrow * ldq + h * kHd + tid];
    __syncthreads();
    float mx = -INFINITY;
    for (int j = tid; j < lk; j += kThreads) {
      float s = -INFINITY;
      const bool masked = j >= n_keys || (key_tokens != nullptr && key_tokens[r * lk + j] == pad_idx);
      if (!masked) {
        const float4* kp = reinterpret_cast<const float4*>(k + rb * kv_bs + j * kv_rs + h * kHd);
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < kHd / 4; ++c) {
          const float4 kk = __ldg(kp + c);
          a = fmaf(qs[4 * c], kk.x, a);
          a = fmaf(qs[4 * c + 1], kk.y, a);
          a = fmaf(qs[4 * c + 2], kk.z, a);
          a = fmaf(qs[4 * c + 3], kk.w, a);
        }
        s = a;
      }
      sc[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = block_max(mx, red);
    float sum = 0.f;
    for (int j = tid; j < lk; j += kThreads) {
      const float s = sc[j];
      const float p = s == -INFINITY ? 0.f : expf(s - mx);
      sc[j] = p;
      sum += p;
    }
    sum = block_sum(sum, red);                       // (its barriers also publish sc[])
    const float inv = 1.0f / sum;                    // every key masked: 0 * inf = NaN, as torch's softmax of -inf
    const int dim = tid & (kHd - 1), half = tid >> 6;
    float acc = 0.f;
    for (int j = half; j < n_keys; j += kThreads / kHd)
      acc = fmaf(sc[j], __ldg(v + rb * kv_bs + j * kv_rs + h * kHd + dim), acc);
    part[tid] = acc;
    __syncthreads();
    if (tid < kHd) ctx[qrow * ldctx + h * kHd + tid] = (part[tid] + part[tid + kHd]) * inv;
    if (weights != nullptr)
      for (int j = tid; j < lk; j += kThreads) wacc[j] += sc[j] * inv * inv_heads;
  }
  if (weights != nullptr)
    for (int j = tid; j < lk; j += kThreads) weights[qrow * lk + j] = wacc[j];
}

// Cross-attention of one decoding step, all hypothesis rows of an utterance in ONE CTA (round 2): in a KV-cached beam
// search every row asks one query and the `group` rows of an utterance attend the SAME encoder keys / values; with a CTA
// per row (attention_f32_kernel) a step of 640 rows read 640 x 1.5 MB of keys and values per layer through L2.  Here a
// key row is loaded once and used for every row of the group.  Same arithmetic as attention_f32_kernel with lq = 1,
// no causal / token masks.  Dynamic shared memory: qs[16][64] | p[lk][gp] | wacc[lk][gp] (if weights) | red[8][16][64].
constexpr int kBeamThreads = 256, kMaxGroup = 16;

__global__ void __launch_bounds__(kBeamThreads)
attention_beam_f32_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ k,
                          const float* __restrict__ v, long long kv_bs, long long kv_rs, int group, int gp, int lk,
                          int n_head, const int* __restrict__ kv_len, float* __restrict__ ctx, long long ldctx,
                          float* __restrict__ weights) {
  extern __shared__ float sm[];
  float* qs = sm;                                   // [kMaxGroup][64]
  float* pr = qs + kMaxGroup * kHd;                 // [lk][gp]: scores, then probabilities
  float* wacc = pr + (size_t)lk * gp;               // [lk][gp] (only if weights)
  float* red = wacc + (weights != nullptr ? (size_t)lk * gp : 0);      // [8][kMaxGroup][64]
  __shared__ int nk[kMaxGroup];
  __shared__ float inv_s[kMaxGroup];
  const int u = blockIdx.x;                         // memory block (utterance)
  const long long row0 = (long long)u * group;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* kb = k + (long long)u * kv_bs;
  const float* vb = v + (long long)u * kv_bs;
  if (tid < kMaxGroup) nk[tid] = tid < group ? (kv_len != nullptr ? min(max(kv_len[row0 + tid], 0), lk) : lk) : 0;
  if (weights != nullptr)
    for (int i = tid; i < lk * gp; i += kBeamThreads) wacc[i] = 0.f;
  const float inv_heads = 1.0f / (float)n_head;
  // without the averaged weights the heads are independent: one CTA per (utterance, head) (gridDim.y = heads)
  const int h_begin = gridDim.y > 1 ? (int)blockIdx.y : 0, h_end = gridDim.y > 1 ? (int)blockIdx.y + 1 : n_head;
  for (int h = h_begin; h < h_end; ++h) {
    __syncthreads();                                // qs / pr / red of the previous head are no longer read
    for (int i = tid; i < group * kHd; i += kBeamThreads)
      qs[i] = q[(row0 + i / kHd) * ldq + h * kHd + (i & (kHd - 1))];
    __syncthreads();
    // scores: one key per thread, its 64 values in registers, every row of the group against it
    for (int j = tid; j < lk; j += kBeamThreads) {
      float4 kk[kHd / 4];
      const float4* kp = reinterpret_cast<const float4*>(kb + (long long)j * kv_rs + h * kHd);
#pragma unroll
      for (int c = 0; c < kHd / 4; ++c) kk[c] = __ldg(kp + c);
      for (int g = 0; g < group; ++g) {
        const float4* q4 = reinterpret_cast<const float4*>(qs + g * kHd);
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < kHd / 4; ++c) {
          const float4 qq = q4[c];
          a = fmaf(qq.x, kk[c].x, a);
          a = fmaf(qq.y, kk[c].y, a);
          a = fmaf(qq.z, kk[c].z, a);
          a = fmaf(qq.w, kk[c].w, a);
        }
        pr[(size_t)j * gp + g] = j < nk[g] ? a : -INFINITY;
      }
    }
    __syncthreads();
    // softmax of every row: warp w takes rows w, w + 8
    for (int g = warp; g < group; g += kBeamThreads / 32) {
      float mx = -INFINITY;
      for (int j = lane; j < lk; j += 32) mx = fmaxf(mx, pr[(size_t)j * gp + g]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int j = lane; j < lk; j += 32) {
        const float sc = pr[(size_t)j * gp + g];
        const float p = sc == -INFINITY ? 0.f : expf(sc - mx);
        pr[(size_t)j * gp + g] = p;
        sum += p;
      }
      sum = warp_sum(sum);
      if (lane == 0) inv_s[g] = 1.0f / sum;         // every key masked: 0 * inf = NaN, as torch's softmax of -inf
    }
    __syncthreads();
    // P . V: thread = (four value columns, one of sixteen key subsets), four accumulators per row of the group; the two
    // subsets of a warp meet by shuffle, the eight warps through shared memory.  (The first version gave a thread one
    // column and a quarter of the keys: 188 dependent global loads per head and thread - 300 us per launch.)
    const int dq = tid & 15, ks = tid >> 4;
    float4 acc[kMaxGroup];
#pragma unroll
    for (int g = 0; g < kMaxGroup; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int j = ks; j < lk; j += kBeamThreads / 16) {
      const float4 vj = __ldg(reinterpret_cast<const float4*>(vb + (long long)j * kv_rs + h * kHd) + dq);
      const float4* p4 = reinterpret_cast<const float4*>(pr + (size_t)j * gp);
#pragma unroll
      for (int g4 = 0; g4 < kMaxGroup / 4; ++g4) {
        if (4 * g4 < gp) {
          const float4 pp = p4[g4];
          const float pv[4] = {pp.x, pp.y, pp.z, pp.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            acc[4 * g4 + e].x = fmaf(pv[e], vj.x, acc[4 * g4 + e].x);
            acc[4 * g4 + e].y = fmaf(pv[e], vj.y, acc[4 * g4 + e].y);
            acc[4 * g4 + e].z = fmaf(pv[e], vj.z, acc[4 * g4 + e].z);
            acc[4 * g4 + e].w = fmaf(pv[e], vj.w, acc[4 * g4 + e].w);
          }
        }
      }
    }
#pragma unroll
    for (int g = 0; g < kMaxGroup; ++g) {
      if (g < gp) {
        acc[g].x += __shfl_xor_sync(0xffffffffu, acc[g].x, 16);
        acc[g].y += __shfl_xor_sync(0xffffffffu, acc[g].y, 16);
        acc[g].z += __shfl_xor_sync(0xffffffffu, acc[g].z, 16);
        acc[g].w += __shfl_xor_sync(0xffffffffu, acc[g].w, 16);
        if (lane < 16) reinterpret_cast<float4*>(red + (warp * kMaxGroup + g) * kHd)[dq] = acc[g];
      }
    }
    __syncthreads();
    for (int i = tid; i < group * kHd; i += kBeamThreads) {
      const int g = i / kHd, d2 = i & (kHd - 1);
      float tot = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < kBeamThreads / 32; ++w2) tot += red[(w2 * kMaxGroup + g) * kHd + d2];
      ctx[(row0 + g) * ldctx + h * kHd + d2] = tot * inv_s[g];
    }
    if (weights != nullptr)
      for (int i = tid; i < lk * gp; i += kBeamThreads) {
        const int g = i % gp;
        if (g < group) wacc[i] += pr[i] * inv_s[g] * inv_heads;
      }
  }
  if (weights != nullptr) {
    __syncthreads();
    for (int g = 0; g < group; ++g)
      for (int j = tid; j < lk; j += kBeamThreads) weights[(row0 + g) * lk + j] = wacc[(size_t)j * gp + g];
  }
}

}  // namespace

extern "C" int stac_embed_scale_pe(const int64_t* tokens, const float* emb, const float* pe, int64_t rows,
                                   int64_t seq_len, int64_t d_model, int64_t vocab, float scale, float* out,
                                   void* stream) {
  STAC_REQUIRE(tokens && emb && pe && out && rows > 0 && seq_len > 0 && vocab > 0 && d_model > 0);
  if (d_model % 4 != 0 || vocab >= (1ll << 31) || seq_len >= (1ll << 31)) return STAC_ERR_UNSUPPORTED_SHAPE;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(rows * (d_model / 4), 256), 148 * 16);
  embed_scale_pe_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const long long*>(tokens), emb, pe,
                                                            (long long)rows, (int)seq_len, (int)d_model, (int)vocab,
                                                            scale, out);
  STAC_LAUNCH_CHECK();
}

extern "C" int stac_attention_f32(const float* q, int64_t ldq, const float* k, const float* v, int64_t kv_batch_stride,
                                  int64_t kv_row_stride, int64_t rows, int64_t lq, int64_t lk, int64_t n_head,
                                  int64_t mem_rows_div,
                                  int causal, const int32_t* kv_len, const int64_t* key_tokens, int64_t pad_idx,
                                  float* ctx, int64_t ldctx, float* weights, void* stream) {
  STAC_REQUIRE(q && k && v && ctx && rows > 0 && lq > 0 && lk > 0 && n_head > 0 && mem_rows_div > 0);
  STAC_REQUIRE(ldq >= n_head * kHd && kv_row_stride >= n_head * kHd && kv_batch_stride > 0 && ldctx >= n_head * kHd);
  // float4 key loads: 16-byte aligned rows
  if (kv_row_stride % 4 != 0 || kv_batch_stride % 4 != 0 || (reinterpret_cast<uintptr_t>(k) & 15) != 0)
    return STAC_ERR_UNSUPPORTED_SHAPE;
  if (rows * lq >= (1ll << 31) || lk >= (1 << 24) || n_head > 65535) return STAC_ERR_UNSUPPORTED_SHAPE;
  const size_t smem = (size_t)(2 * lk + kHd + kThreads + 40) * sizeof(float);
  if (smem > 200 * 1024) return STAC_ERR_UNSUPPORTED_SHAPE;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  attention_f32_kernel<<<(unsigned)(rows * lq), kThreads, smem, as_stream(stream)>>>(
      q, (long long)ldq, k, v, (long long)kv_batch_stride, (long long)kv_row_stride, (int)lq, (int)lk, (int)n_head, (int)mem_rows_div, causal, kv_len,
      reinterpret_cast<const long long*>(key_tokens), (long long)pad_idx, ctx, (long long)ldctx, weights);
  STAC_LAUNCH_CHECK();
}

extern "C" int stac_attention_beam_f32(const float* q, int64_t ldq, const float* k, const float* v,
                                       int64_t kv_batch_stride, int64_t kv_row_stride, int64_t rows, int64_t group,
                                       int64_t lk, int64_t n_head, const int32_t* kv_len, float* ctx, int64_t ldctx,
                                       float* weights, void* stream) {
  STAC_REQUIRE(q && k && v && ctx && rows > 0 && group > 0 && lk > 0 && n_head > 0 && rows % group == 0);
  STAC_REQUIRE(ldq >= n_head * kHd && kv_row_stride >= n_head * kHd && kv_batch_stride > 0 && ldctx >= n_head * kHd);
  if (group > kMaxGroup) return STAC_ERR_UNSUPPORTED_SHAPE;          // wider beams: stac_attention_f32
  if (kv_row_stride % 4 != 0 || kv_batch_stride % 4 != 0 || (reinterpret_cast<uintptr_t>(k) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(q) & 3) != 0)
    return STAC_ERR_UNSUPPORTED_SHAPE;
  if (rows / group >= (1ll << 31) || lk >= (1 << 24) || n_head > 65535) return STAC_ERR_UNSUPPORTED_SHAPE;
  const int gp = (int)((group + 3) / 4 * 4);
  const size_t smem = ((size_t)kMaxGroup * kHd + (size_t)lk * gp * (weights ? 2 : 1) + (size_t)8 * kMaxGroup * kHd) * sizeof(float);
  if (smem > 220 * 1024) return STAC_ERR_UNSUPPORTED_SHAPE;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(attention_beam_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  attention_beam_f32_kernel<<<dim3((unsigned)(rows / group), weights ? 1u : (unsigned)n_head), kBeamThreads, smem,
                              as_stream(stream)>>>(
      q, (long long)ldq, k, v, (long long)kv_batch_stride, (long long)kv_row_stride, (int)group, gp, (int)lk, (int)n_head,
      kv_len, ctx, (long long)ldctx, weights);
  STAC_LAUNCH_CHECK();
}
