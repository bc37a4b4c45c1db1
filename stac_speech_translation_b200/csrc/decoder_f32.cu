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
                      const int* __restrict__ pos_dev, float* __restrict__ out) {
  const int d4 = d_model >> 2;
  if (pos_dev != nullptr) pe += (long long)(*pos_dev) * d_model;      // position counter kept on the device (graph replay)
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

// One decoding step over the cache (lq = 1, no weights, no token mask): one WARP per (row, head) - lane = key for the
// scores (the probabilities of a 256-key chunk stay in eight registers per lane), lane = two value columns for P . V,
// shuffles instead of barriers; caches longer than 256 keys go chunk by chunk with a running maximum / sum (one chunk:
// the plain two-pass softmax).  (The CTA-per-row kernel above spends a step's 640 x 4 tiny problems on block reductions
// and on 16 dependent loads per thread: 37 us per launch at 32 keys.)
// row_map (or NULL): int32 [lk][rows]; key / value j of hypothesis row r lives in cache row row_map[j * rows + r] - the
// lazy form of a beam re-ordering (DecoderCache.reorder permutes this small table instead of moving the cache).
constexpr int kStepWarps = 8, kStepMaxKeys = 256;

__global__ void __launch_bounds__(kStepWarps * 32)
attention_step_warp_kernel(const float* __restrict__ q, long long ldq, const float* k, const float* v, long long kv_bs,
                           long long kv_rs, int lk, int n_head, int mem_rows_div, const int* __restrict__ kv_len,
                           const int* __restrict__ row_map, const float* __restrict__ k_new,
                           const float* __restrict__ v_new, long long ld_new, const int* __restrict__ t_dev,
                           float* __restrict__ ctx, long long ldctx, long long rows) {
  __shared__ __align__(16) float qs[kStepWarps][kHd];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * kStepWarps + warp;
  if (item >= rows * n_head) return;                 // (whole warps leave: no CTA barrier below)
  const long long row = item / n_head;
  const int h = (int)(item % n_head);
  const int rb = (int)(row / mem_rows_div);
  reinterpret_cast<float2*>(qs[warp])[lane] = *reinterpret_cast<const float2*>(q + row * ldq + h * kHd + 2 * lane);
  __syncwarp();
  float m_run = -INFINITY, l_run = 0.f;
  float2 acc = make_float2(0.f, 0.f);
  // append mode (t_dev given): the position counter lives on the device, the keys / values of position t = *t_dev come
  // from the projection's output rows (k_new / v_new) and are stored into slab t of the cache here; the step attends
  // the cached keys 0 .. t-1 and the new one.  One launch sequence then serves every step of a search (DecoderCache
  // replays it as a CUDA graph).  The new key opens the running softmax: score by the whole warp, weight 1.
  if (t_dev != nullptr) {
    const int t_new = min(*t_dev, lk - 1);           // (lk = capacity of the cache in this mode)
    lk = t_new;                                      // cached keys to attend
    const long long o_new = row * ld_new + h * kHd + 2 * lane;
    const long long o_cache = row * kv_bs + (long long)t_new * kv_rs + h * kHd + 2 * lane;
    const float2 kn = *reinterpret_cast<const float2*>(k_new + o_new);
    const float2 vn = *reinterpret_cast<const float2*>(v_new + o_new);
    *reinterpret_cast<float2*>(const_cast<float*>(k) + o_cache) = kn;
    *reinterpret_cast<float2*>(const_cast<float*>(v) + o_cache) = vn;
    float s_new = fmaf(qs[warp][2 * lane], kn.x, qs[warp][2 * lane + 1] * kn.y);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s_new += __shfl_xor_sync(0xffffffffu, s_new, o);
    m_run = s_new;
    l_run = 1.f;
    acc = vn;
  }
  int n_keys = lk;
  if (kv_len != nullptr) n_keys = min(max(kv_len[row], 0), lk);
  const float* kb = k + h * kHd;
  const float* vb = v + h * kHd + 2 * lane;
  for (int base = 0; base < n_keys; base += kStepMaxKeys) {
    float p[kStepMaxKeys / 32];
    int src[kStepMaxKeys / 32];                      // cache row of my key in each 32-key group
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < kStepMaxKeys / 32; ++i) {
      const int j = base + 32 * i + lane;
      float a = -INFINITY;
      src[i] = rb;
      if (j < n_keys) {
        if (row_map != nullptr) src[i] = __ldg(row_map + (long long)j * rows + row);
        const float4* kp = reinterpret_cast<const float4*>(kb + (long long)src[i] * kv_bs + (long long)j * kv_rs);
        a = 0.f;
#pragma unroll
        for (int c = 0; c < kHd / 4; ++c) {
          const float4 kk = __ldg(kp + c);
          const float4 qq = reinterpret_cast<const float4*>(qs[warp])[c];
          a = fmaf(qq.x, kk.x, a);
          a = fmaf(qq.y, kk.y, a);
          a = fmaf(qq.z, kk.z, a);
          a = fmaf(qq.w, kk.w, a);
        }
      }
      p[i] = a;
      mx = fmaxf(mx, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float m_new = fmaxf(m_run, mx);
    const float scale = m_run == -INFINITY ? 0.f : expf(m_run - m_new);      // (first chunk: nothing to rescale)
    acc.x *= scale;
    acc.y *= scale;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kStepMaxKeys / 32; ++i) {
      p[i] = p[i] == -INFINITY ? 0.f : expf(p[i] - m_new);
      sum += p[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    l_run = l_run * scale + sum;
    m_run = m_new;
    // P . V without a branch per key: keys past the end carry p = 0 and read the last valid row (finite values), so
    // the sixteen loads of an unrolled group are issued back to back
#pragma unroll
    for (int i = 0; i < kStepMaxKeys / 32; ++i) {
      if (base + 32 * i < n_keys) {                  // (warp-uniform)
#pragma unroll 16
        for (int jj = 0; jj < 32; ++jj) {
          const float pj = __shfl_sync(0xffffffffu, p[i], jj);
          const int jl = min(base + 32 * i + jj, n_keys - 1);
          const int rj = __shfl_sync(0xffffffffu, src[i], jl & 31);
          const float2 vv = __ldg(reinterpret_cast<const float2*>(vb + (long long)rj * kv_bs + (long long)jl * kv_rs));
          acc.x = fmaf(pj, vv.x, acc.x);
          acc.y = fmaf(pj, vv.y, acc.y);
        }
      }
    }
  }
  const float inv = 1.0f / l_run;                    // every key masked: 0 * inf = NaN, as the kernel above
  *reinterpret_cast<float2*>(ctx + row * ldctx + h * kHd + 2 * lane) = make_float2(acc.x * inv, acc.y * inv);
}

// Cross-attention of one decoding step, all hypothesis rows of an utterance in ONE CTA (round 2): in a KV-cached beam
// search every row asks one query and the `group` rows of an utterance attend the SAME encoder keys / values; with a CTA
// per row (attention_f32_kernel) a step of 640 rows read 640 x 1.5 MB of keys and values per layer through L2.  Here a
// key row is loaded once and used for every row of the group.  Same arithmetic as attention_f32_kernel with lq = 1,
// no causal / token masks.  Dynamic shared memory: qs[16][64] | p[lk][gp] | wacc[lk][gp] (if weights) | red[8][16][64].
constexpr int kBeamThreads = 256, kMaxGroup = 16;

template <int kG4>                                  // rows of the group, rounded up to a multiple of four, / 4
__global__ void __launch_bounds__(kBeamThreads)
attention_beam_f32_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ k,
                          const float* __restrict__ v, long long kv_bs, long long kv_rs, int group, int lk,
                          int n_head, const int* __restrict__ kv_len, float* __restrict__ ctx, long long ldctx,
                          float* __restrict__ weights, float* __restrict__ head_p, long long rows) {
  constexpr int gp = 4 * kG4;
  extern __shared__ float sm[];
  float* qs = sm;                                   // [kMaxGroup][64]
  float* pr = qs + kMaxGroup * kHd;                 // [lk][gp]: scores, then probabilities
  float* wacc = pr + (size_t)lk * gp;               // [lk][gp] (only if weights)
  float* red = wacc + (weights != nullptr ? (size_t)lk * gp : 0);      // [8][kMaxGroup][64]; its head: [8][16] exchange
  __shared__ int nk[kMaxGroup];
  __shared__ float mx_s[kMaxGroup], inv_s[kMaxGroup];
  const int u = blockIdx.x;                         // memory block (utterance)
  const long long row0 = (long long)u * group;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* kb = k + (long long)u * kv_bs;
  const float* vb = v + (long long)u * kv_bs;
  if (tid < kMaxGroup) nk[tid] = tid < group ? (kv_len != nullptr ? min(max(kv_len[row0 + tid], 0), lk) : lk) : 0;
  if (weights != nullptr)
    for (int i = tid; i < lk * gp; i += kBeamThreads) wacc[i] = 0.f;
  const float inv_heads = 1.0f / (float)n_head;
  // without the in-CTA head average the heads are independent: one CTA per (utterance, head) (gridDim.y = heads)
  const int h_begin = gridDim.y > 1 ? (int)blockIdx.y : 0, h_end = gridDim.y > 1 ? (int)blockIdx.y + 1 : n_head;
  for (int h = h_begin; h < h_end; ++h) {
    __syncthreads();                                // qs / pr / red of the previous head are no longer read
    for (int i = tid; i < gp * kHd; i += kBeamThreads) {
      const int g = i / kHd;
      qs[i] = g < group ? q[(row0 + g) * ldq + h * kHd + (i & (kHd - 1))] : 0.f;
    }
    __syncthreads();
    // scores: one key per thread, its 64 values in registers, four rows of the group at a time against it (four
    // independent FMA chains; one chain per row ran at the FMA latency: 43 % of the kernel's stall samples)
    float mxl[gp];
#pragma unroll
    for (int g = 0; g < gp; ++g) mxl[g] = -INFINITY;
    for (int j = tid; j < lk; j += kBeamThreads) {
      float4 kk[kHd / 4];
      const float4* kp = reinterpret_cast<const float4*>(kb + (long long)j * kv_rs + h * kHd);
#pragma unroll
      for (int c = 0; c < kHd / 4; ++c) kk[c] = __ldg(kp + c);
#pragma unroll
      for (int g4 = 0; g4 < kG4; ++g4) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < kHd / 4; ++c) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 qq = reinterpret_cast<const float4*>(qs + (4 * g4 + e) * kHd)[c];
            a[e] = fmaf(qq.x, kk[c].x, a[e]);
            a[e] = fmaf(qq.y, kk[c].y, a[e]);
            a[e] = fmaf(qq.z, kk[c].z, a[e]);
            a[e] = fmaf(qq.w, kk[c].w, a[e]);
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          a[e] = j < nk[4 * g4 + e] ? a[e] : -INFINITY;
          mxl[4 * g4 + e] = fmaxf(mxl[4 * g4 + e], a[e]);
        }
        reinterpret_cast<float4*>(pr + (size_t)j * gp)[g4] = make_float4(a[0], a[1], a[2], a[3]);
      }
    }
    // softmax, every thread on its own keys: row maxima and row sums meet through shuffles and red[8][16]
#pragma unroll
    for (int g = 0; g < gp; ++g) {
      const float m = warp_max(mxl[g]);
      if (lane == 0) red[warp * kMaxGroup + g] = m;
    }
    __syncthreads();
    if (tid < gp) {
      float m = red[tid];
      for (int w2 = 1; w2 < kBeamThreads / 32; ++w2) m = fmaxf(m, red[w2 * kMaxGroup + tid]);
      mx_s[tid] = m;
    }
    __syncthreads();
    float sl[gp];
#pragma unroll
    for (int g = 0; g < gp; ++g) {
      mxl[g] = mx_s[g];
      sl[g] = 0.f;
    }
    for (int j = tid; j < lk; j += kBeamThreads) {
#pragma unroll
      for (int g4 = 0; g4 < kG4; ++g4) {
        float4 sc = reinterpret_cast<const float4*>(pr + (size_t)j * gp)[g4];
        sc.x = sc.x == -INFINITY ? 0.f : expf(sc.x - mxl[4 * g4]);
        sc.y = sc.y == -INFINITY ? 0.f : expf(sc.y - mxl[4 * g4 + 1]);
        sc.z = sc.z == -INFINITY ? 0.f : expf(sc.z - mxl[4 * g4 + 2]);
        sc.w = sc.w == -INFINITY ? 0.f : expf(sc.w - mxl[4 * g4 + 3]);
        sl[4 * g4] += sc.x;
        sl[4 * g4 + 1] += sc.y;
        sl[4 * g4 + 2] += sc.z;
        sl[4 * g4 + 3] += sc.w;
        reinterpret_cast<float4*>(pr + (size_t)j * gp)[g4] = sc;
      }
    }
#pragma unroll
    for (int g = 0; g < gp; ++g) {
      const float t = warp_sum(sl[g]);
      if (lane == 0) red[warp * kMaxGroup + g] = t;
    }
    __syncthreads();
    if (tid < gp) {
      float t = red[tid];
      for (int w2 = 1; w2 < kBeamThreads / 32; ++w2) t += red[w2 * kMaxGroup + tid];
      inv_s[tid] = 1.0f / t;                        // every key masked: 0 * inf = NaN, as torch's softmax of -inf
    }
    __syncthreads();
    // P . V: thread = (four value columns, one of sixteen key subsets), four accumulators per row of the group, eight
    // value loads in flight; the two subsets of a warp meet by shuffle, the eight warps through shared memory.  (The
    // first version gave a thread one column and a quarter of the keys: 188 dependent global loads per head and thread -
    // 300 us per launch.)
    const int dq = tid & 15, ks = tid >> 4;
    constexpr int kFly = 8, kSub = kBeamThreads / 16;
    float4 acc[gp];
#pragma unroll
    for (int g = 0; g < gp; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j0 = ks; j0 < lk; j0 += kSub * kFly) {
      float4 vj[kFly];
#pragma unroll
      for (int f = 0; f < kFly; ++f) {
        const int j = j0 + kSub * f;
        vj[f] = j < lk ? __ldg(reinterpret_cast<const float4*>(vb + (long long)j * kv_rs + h * kHd) + dq)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int f = 0; f < kFly; ++f) {
        const int j = j0 + kSub * f;
        if (j < lk) {
#pragma unroll
          for (int g4 = 0; g4 < kG4; ++g4) {
            const float4 pp = reinterpret_cast<const float4*>(pr + (size_t)j * gp)[g4];
            const float pv[4] = {pp.x, pp.y, pp.z, pp.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              acc[4 * g4 + e].x = fmaf(pv[e], vj[f].x, acc[4 * g4 + e].x);
              acc[4 * g4 + e].y = fmaf(pv[e], vj[f].y, acc[4 * g4 + e].y);
              acc[4 * g4 + e].z = fmaf(pv[e], vj[f].z, acc[4 * g4 + e].z);
              acc[4 * g4 + e].w = fmaf(pv[e], vj[f].w, acc[4 * g4 + e].w);
            }
          }
        }
      }
    }
#pragma unroll
    for (int g = 0; g < gp; ++g) {
      acc[g].x += __shfl_xor_sync(0xffffffffu, acc[g].x, 16);
      acc[g].y += __shfl_xor_sync(0xffffffffu, acc[g].y, 16);
      acc[g].z += __shfl_xor_sync(0xffffffffu, acc[g].z, 16);
      acc[g].w += __shfl_xor_sync(0xffffffffu, acc[g].w, 16);
      if (lane < 16) reinterpret_cast<float4*>(red + (warp * kMaxGroup + g) * kHd)[dq] = acc[g];
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
    if (head_p != nullptr)                           // heads on separate CTAs: head_average_kernel adds them up
      for (int g = 0; g < group; ++g)
        for (int j = tid; j < lk; j += kBeamThreads)
          head_p[((long long)h * rows + row0 + g) * lk + j] = pr[(size_t)j * gp + g] * inv_s[g];
  }
  if (weights != nullptr) {
    __syncthreads();
    for (int g = 0; g < group; ++g)
      for (int j = tid; j < lk; j += kBeamThreads) weights[(row0 + g) * lk + j] = wacc[(size_t)j * gp + g];
  }
}

// weights[i] = sum over the heads, in head order, of head_p[h][i] / heads (the order and the rounding of the
// one-CTA-for-all-heads path above)
__global__ void __launch_bounds__(256)
head_average_kernel(const float* __restrict__ head_p, long long n, int n_head, float* __restrict__ weights) {
  const float inv_heads = 1.0f / (float)n_head;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int h = 0; h < n_head; ++h) a += head_p[(long long)h * n + i] * inv_heads;
    weights[i] = a;
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
                                                            scale, nullptr, out);
  STAC_LAUNCH_CHECK();
}

extern "C" int stac_embed_step(const int64_t* tokens, const float* emb, const float* pe, int64_t rows, int64_t d_model,
                               int64_t vocab, float scale, const int32_t* pos_dev, float* out, void* stream) {
  STAC_REQUIRE(tokens && emb && pe && out && pos_dev && rows > 0 && vocab > 0 && d_model > 0);
  if (d_model % 4 != 0 || vocab >= (1ll << 31)) return STAC_ERR_UNSUPPORTED_SHAPE;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(rows * (d_model / 4), 256), 148 * 16);
  embed_scale_pe_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const long long*>(tokens), emb, pe,
                                                            (long long)rows, 1, (int)d_model, (int)vocab, scale,
                                                            pos_dev, out);
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
  if (lq == 1 && weights == nullptr && key_tokens == nullptr && lk <= kStepMaxKeys && kv_row_stride % 2 == 0 &&
      ldq % 2 == 0 && ldctx % 2 == 0 && (reinterpret_cast<uintptr_t>(q) & 7) == 0 &&
      (reinterpret_cast<uintptr_t>(v) & 7) == 0 && (reinterpret_cast<uintptr_t>(ctx) & 7) == 0) {
    // (lq = 1: "causal" only caps the keys at 1, which a one-token prefix already is)
    const long long items = (long long)rows * n_head;
    const int lk_eff = causal ? 1 : (int)lk;
    attention_step_warp_kernel<<<(unsigned)((items + kStepWarps - 1) / kStepWarps), kStepWarps * 32, 0, as_stream(stream)>>>(
        q, (long long)ldq, k, v, (long long)kv_batch_stride, (long long)kv_row_stride, lk_eff, (int)n_head,
        (int)mem_rows_div, kv_len, nullptr, nullptr, nullptr, 0, nullptr, ctx, (long long)ldctx, (long long)rows);
    STAC_LAUNCH_CHECK();
  }
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

extern "C" int stac_attention_step_f32(const float* q, int64_t ldq, float* k, float* v, int64_t kv_row_stride,
                                       int64_t kv_time_stride, int64_t rows, int64_t lk, int64_t n_head,
                                       const int32_t* row_map, const float* k_new, const float* v_new, int64_t ld_new,
                                       const int32_t* t_dev, float* ctx, int64_t ldctx, void* stream) {
  STAC_REQUIRE(q && k && v && ctx && rows > 0 && lk > 0 && n_head > 0);
  STAC_REQUIRE(ldq >= n_head * kHd && kv_row_stride >= n_head * kHd && kv_time_stride > 0 && ldctx >= n_head * kHd);
  STAC_REQUIRE(t_dev == nullptr || (k_new && v_new && ld_new >= n_head * kHd));
  if (kv_row_stride % 4 != 0 || kv_time_stride % 4 != 0 || ldq % 2 != 0 || ldctx % 2 != 0 ||
      (reinterpret_cast<uintptr_t>(k) & 15) != 0 || (reinterpret_cast<uintptr_t>(v) & 7) != 0 ||
      (reinterpret_cast<uintptr_t>(q) & 7) != 0 || (reinterpret_cast<uintptr_t>(ctx) & 7) != 0)
    return STAC_ERR_UNSUPPORTED_SHAPE;
  if (t_dev != nullptr && (ld_new % 4 != 0 || (reinterpret_cast<uintptr_t>(k_new) & 15) != 0 ||
                           (reinterpret_cast<uintptr_t>(v_new) & 7) != 0))
    return STAC_ERR_UNSUPPORTED_SHAPE;
  if (rows >= (1ll << 31) || lk >= (1 << 24) || n_head > 65535) return STAC_ERR_UNSUPPORTED_SHAPE;
  const long long items = (long long)rows * n_head;
  attention_step_warp_kernel<<<(unsigned)((items + kStepWarps - 1) / kStepWarps), kStepWarps * 32, 0, as_stream(stream)>>>(
      q, (long long)ldq, k, v, (long long)kv_row_stride, (long long)kv_time_stride, (int)lk, (int)n_head, 1, nullptr,
      row_map, k_new, v_new, (long long)ld_new, t_dev, ctx, (long long)ldctx, (long long)rows);
  STAC_LAUNCH_CHECK();
}

extern "C" int stac_attention_beam_f32(const float* q, int64_t ldq, const float* k, const float* v,
                                       int64_t kv_batch_stride, int64_t kv_row_stride, int64_t rows, int64_t group,
                                       int64_t lk, int64_t n_head, const int32_t* kv_len, float* ctx, int64_t ldctx,
                                       float* weights, float* head_scratch, void* stream) {
  STAC_REQUIRE(q && k && v && ctx && rows > 0 && group > 0 && lk > 0 && n_head > 0 && rows % group == 0);
  STAC_REQUIRE(ldq >= n_head * kHd && kv_row_stride >= n_head * kHd && kv_batch_stride > 0 && ldctx >= n_head * kHd);
  if (group > kMaxGroup) return STAC_ERR_UNSUPPORTED_SHAPE;          // wider beams: stac_attention_f32
  if (kv_row_stride % 4 != 0 || kv_batch_stride % 4 != 0 || (reinterpret_cast<uintptr_t>(k) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(q) & 3) != 0)
    return STAC_ERR_UNSUPPORTED_SHAPE;
  if (rows / group >= (1ll << 31) || lk >= (1 << 24) || n_head > 65535) return STAC_ERR_UNSUPPORTED_SHAPE;
  const int gp = (int)((group + 3) / 4 * 4);
  // the averaged weights: with a scratch of heads x rows x lk floats the heads still run on separate CTAs and a second
  // kernel adds them up; without one, one CTA per utterance walks the heads and sums in shared memory
  const bool serial_heads = weights != nullptr && head_scratch == nullptr;
  const size_t smem =
      ((size_t)kMaxGroup * kHd + (size_t)lk * gp * (serial_heads ? 2 : 1) + (size_t)8 * kMaxGroup * kHd) * sizeof(float);
  if (smem > 220 * 1024) return STAC_ERR_UNSUPPORTED_SHAPE;
  const dim3 grid((unsigned)(rows / group), serial_heads ? 1u : (unsigned)n_head);
  float* w_serial = serial_heads ? weights : nullptr;
  float* w_heads = weights != nullptr && !serial_heads ? head_scratch : nullptr;
#define STAC_BEAM_LAUNCH(G4)                                                                                          \
  do {                                                                                                                \
    if (smem > 48 * 1024) {                                                                                           \
      cudaError_t e = cudaFuncSetAttribute(attention_beam_f32_kernel<G4>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           (int)smem);                                                                \
      if (e != cudaSuccess) return (int)e;                                                                            \
    }                                                                                                                 \
    attention_beam_f32_kernel<G4><<<grid, kBeamThreads, smem, as_stream(stream)>>>(                                   \
        q, (long long)ldq, k, v, (long long)kv_batch_stride, (long long)kv_row_stride, (int)group, (int)lk,           \
        (int)n_head, kv_len, ctx, (long long)ldctx, w_serial, w_heads, (long long)rows);                              \
  } while (0)
  switch (gp / 4) {
    case 1: STAC_BEAM_LAUNCH(1); break;
    case 2: STAC_BEAM_LAUNCH(2); break;
    case 3: STAC_BEAM_LAUNCH(3); break;
    default: STAC_BEAM_LAUNCH(4); break;
  }
#undef STAC_BEAM_LAUNCH
  if (weights != nullptr && !serial_heads) {
    const long long n = (long long)rows * lk;
    head_average_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, as_stream(stream)>>>(
        head_scratch, n, (int)n_head, weights);
  }
  STAC_LAUNCH_CHECK();
}
