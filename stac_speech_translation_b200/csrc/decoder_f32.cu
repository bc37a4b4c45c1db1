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
