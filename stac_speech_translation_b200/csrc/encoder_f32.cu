// a5-a9 building blocks that run on the CUDA cores in fp32: row LayerNorm, the fp32-mode GEMM
// and attention, CTC log-softmax/argmax, fp32->bf16 casts.
// Reference behaviour: SpeechBrain TransformerEncoder (pre-LN, eps 1e-6) as built by
//   /root/reference/stac-st/modules/TransformerMultiTask.py:111-128 and called at :304-308;
//   CTC head /root/reference/stac-st/inference.py:104-107, greedy argmax :54-56.
#include <algorithm>
#include "common.cuh"
#include "gemm_simt.cuh"

namespace {

// ---- LayerNorm: one warp per kRows rows, dim <= 1024, dim % 128 == 0 ---------------------------
// (kRows rows per warp: all their loads are issued before the first reduction, so a warp keeps kRows x dim x 4 bytes in
//  flight instead of one row's - with one row per warp the d_model = 256 launch ran at 4.9 TB/s, latency-bound)
template <int kVec, int kRows>  // float4 per lane and row = dim / 128
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int64_t rows, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, float* __restrict__ out_f32,
                 __nv_bfloat16* __restrict__ out_bf16) {
  constexpr int dim = kVec * 128;
  const int lane = threadIdx.x & 31;
  const int64_t row0 = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * kRows;
  if (row0 >= rows) return;
  float4 v[kRows][kVec];
  float s[kRows];
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    const int64_t row = min(row0 + r, rows - 1);               // rows past the end repeat the last one (not stored)
    const float4* xr = reinterpret_cast<const float4*>(x + row * dim);
#pragma unroll
    for (int i = 0; i < kVec; ++i) v[r][i] = __ldg(xr + lane + 32 * i);
  }
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    s[r] = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) s[r] += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
  }
  float mean[kRows], q[kRows];
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    mean[r] = s[r] * (1.0f / dim);
    q[r] = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const float a = v[r][i].x - mean[r], b = v[r][i].y - mean[r], c = v[r][i].z - mean[r], d = v[r][i].w - mean[r];
      q[r] += (a * a + b * b) + (c * c + d * d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
  }
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int col = (lane + 32 * i) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + col));
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const int64_t row = row0 + r;
      if (row >= rows) break;
      const float rstd = rsqrtf(q[r] * (1.0f / dim) + eps);
      const float y0 = (v[r][i].x - mean[r]) * rstd * g.x + bb.x, y1 = (v[r][i].y - mean[r]) * rstd * g.y + bb.y;
      const float y2 = (v[r][i].z - mean[r]) * rstd * g.z + bb.z, y3 = (v[r][i].w - mean[r]) * rstd * g.w + bb.w;
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * dim + col) = make_float4(y0, y1, y2, y3);
      if (out_bf16)
        *reinterpret_cast<uint2*>(out_bf16 + row * dim + col) = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
    }
  }
}

// ---- fp32 attention: CTA = (64-query tile, head, utterance); thread = one query row ----------
constexpr int kHd = 64, kQTile = 64, kKTile = 32;

__global__ void __launch_bounds__(kQTile)
mha_f32_kernel(const float* __restrict__ qkv, const int* __restrict__ kv_len, int seq_len,
               int d_model, float* __restrict__ ctx) {
  __shared__ float ks[kKTile][kHd];
  __shared__ float vs[kKTile][kHd];
  const int b = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * kQTile + threadIdx.x;
  const int64_t ld = 3 * (int64_t)d_model;
  const float* base = qkv + (int64_t)b * seq_len * ld + h * kHd;
  const int n_keys = min(max(kv_len[b], 1), seq_len);
  float q[kHd], o[kHd];
  const bool active = qi < seq_len;
#pragma unroll
  for (int i = 0; i < kHd; i += 4) {
    float4 t = active ? __ldg(reinterpret_cast<const float4*>(base + (int64_t)qi * ld + i))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
    q[i] = t.x; q[i + 1] = t.y; q[i + 2] = t.z; q[i + 3] = t.w;
    o[i] = o[i + 1] = o[i + 2] = o[i + 3] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < n_keys; k0 += kKTile) {
    __syncthreads();
    for (int i = threadIdx.x; i < kKTile * (kHd / 4); i += kQTile) {
      const int r = i / (kHd / 4), c4 = (i % (kHd / 4)) * 4;
      float4 kv4 = make_float4(0.f, 0.f, 0.f, 0.f), vv4 = kv4;
      if (k0 + r < n_keys) {
        const float* p = base + (int64_t)(k0 + r) * ld;
        kv4 = __ldg(reinterpret_cast<const float4*>(p + d_model + c4));
        vv4 = __ldg(reinterpret_cast<const float4*>(p + 2 * d_model + c4));
      }
      *reinterpret_cast<float4*>(&ks[r][c4]) = kv4;
      *reinterpret_cast<float4*>(&vs[r][c4]) = vv4;
    }
    __syncthreads();
    const int nk = min(kKTile, n_keys - k0);
    float sc[kKTile];
    float tile_max = -INFINITY;
#pragma unroll
    for (int j = 0; j < kKTile; ++j) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < kHd; ++i) a = fmaf(q[i], ks[j][i], a);
      sc[j] = j < nk ? a : -INFINITY;
      tile_max = fmaxf(tile_max, sc[j]);
    }
    const float m_new = fmaxf(m, tile_max);
    const float alpha = expf(m - m_new);
    l *= alpha;
#pragma unroll
    for (int i = 0; i < kHd; ++i) o[i] *= alpha;
#pragma unroll
    for (int j = 0; j < kKTile; ++j) {
      const float p = expf(sc[j] - m_new);
      l += p;
#pragma unroll
      for (int i = 0; i < kHd; ++i) o[i] = fmaf(p, vs[j][i], o[i]);
    }
    m = m_new;
  }
  if (active) {
    const float inv = 1.0f / l;
    float* dst = ctx + ((int64_t)b * seq_len + qi) * d_model + h * kHd;
#pragma unroll
    for (int i = 0; i < kHd; i += 4)
      *reinterpret_cast<float4*>(dst + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
  }
}

// ---- log-softmax over the vocabulary (+ greedy argmax): one CTA per row ----------------------
__global__ void __launch_bounds__(256)
log_softmax_kernel(const float* logits, int vocab, float* out,      // in-place calls alias logits and out
                   int* __restrict__ argmax) {
  __shared__ float red[40];
  __shared__ int red_i[8];
  __shared__ float red_v[8];
  const int64_t r = blockIdx.x;
  const float* xr = logits + r * vocab;
  float mx = -INFINITY;
  int mi = 0;
  for (int i = threadIdx.x; i < vocab; i += 256) {
    const float v = xr[i];
    if (v > mx) { mx = v; mi = i; }
  }
  // warp arg-max (first index wins on ties, like torch.argmax on CPU/CUDA for distinct maxima)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (ov > mx || (ov == mx && oi < mi)) { mx = ov; mi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { red_v[threadIdx.x >> 5] = mx; red_i[threadIdx.x >> 5] = mi; }
  __syncthreads();
  float bm = red_v[0];
  int bi = red_i[0];
#pragma unroll
  for (int w = 1; w < 8; ++w)
    if (red_v[w] > bm || (red_v[w] == bm && red_i[w] < bi)) { bm = red_v[w]; bi = red_i[w]; }
  float s = 0.f;
  for (int i = threadIdx.x; i < vocab; i += 256) s += expf(xr[i] - bm);
  const float lse = bm + logf(block_sum(s, red));
  for (int i = threadIdx.x; i < vocab; i += 256) out[r * vocab + i] = xr[i] - lse;
  if (argmax && threadIdx.x == 0) argmax[r] = bi;
}

// valid key count per utterance from the reference's own fp32 expressions (see include/stac_b200.h)
__global__ void kv_lengths_kernel(const float* __restrict__ wav_len, int batch, int t2, int round_rule,
                                  int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch) return;
  int n = t2;
  if (wav_len != nullptr) {
    const float x = __fmul_rn(wav_len[i], (float)t2);
    n = round_rule ? (int)rintf(x) : (int)floorf(x) + 1;      // torch.round is round-half-to-even
  }
  out[i] = min(max(n, 1), t2);
}

__global__ void cast_bf16_kernel(const float* __restrict__ x, int64_t n, __nv_bfloat16* __restrict__ out) {
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n;
       i += (int64_t)gridDim.x * blockDim.x * 4) {
    if (i + 3 < n) {
      const float4 v = *reinterpret_cast<const float4*>(x + i);
      *reinterpret_cast<uint2*>(out + i) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    } else {
      for (int64_t j = i; j < n; ++j) out[j] = __float2bfloat16_rn(x[j]);
    }
  }
}

__global__ void cast_f32_kernel(const __nv_bfloat16* __restrict__ x, int64_t n, float* __restrict__ out) {
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n;
       i += (int64_t)gridDim.x * blockDim.x * 4) {
    if (i + 3 < n) {
      const uint2 v = *reinterpret_cast<const uint2*>(x + i);
      *reinterpret_cast<float4*>(out + i) = make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u),
                                                        __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
    } else {
      for (int64_t j = i; j < n; ++j)
        out[j] = __uint_as_float((uint32_t)reinterpret_cast<const uint16_t*>(x)[j] << 16);
    }
  }
}

}  // namespace

extern "C" int stac_layernorm(const float* x, int64_t rows, int64_t dim, const float* gamma,
                              const float* beta, float eps, float* out_f32, uint16_t* out_bf16,
                              void* stream) {
  STAC_REQUIRE(x && gamma && beta && rows > 0 && (out_f32 || out_bf16));
  if (dim % 128 != 0 || dim > 1024 || dim <= 0) return STAC_ERR_UNSUPPORTED_SHAPE;
  __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  cudaStream_t st = as_stream(stream);
  // rows per warp: 32 floats of row data per lane at most
  switch (dim / 128) {
#define LN_CASE(V, R) case V: layernorm_kernel<V,R><<<(unsigned)ceil_div64(rows, 8 * R), 256, 0, st>>>(x, rows, gamma, beta, eps, out_f32, ob); break;
    LN_CASE(1, 4) LN_CASE(2, 4) LN_CASE(3, 2) LN_CASE(4, 2) LN_CASE(5, 1) LN_CASE(6, 1) LN_CASE(7, 1) LN_CASE(8, 1)
#undef LN_CASE
    default: return STAC_ERR_UNSUPPORTED_SHAPE;
  }
  STAC_LAUNCH_CHECK();
}

extern "C" int stac_gemm_f32(const float* a, const float* w, const float* bias, const float* resid,
                             int64_t resid_period, int act, float* c, int64_t m, int64_t n,
                             int64_t k, void* stream) {
  STAC_REQUIRE(a && w && c && m > 0 && n > 0 && k > 0 && resid_period >= 0);
  STAC_REQUIRE(act == STAC_ACT_NONE || act == STAC_ACT_GELU_ERF);
  if (k % 4 != 0 || n > (1 << 22)) return STAC_ERR_UNSUPPORTED_SHAPE;
  simt::RowMajorA ld{a, m, k};
  simt::LinearEpilogue ep{bias, resid, resid_period, act, c, n};
  dim3 grid((unsigned)ceil_div64(m, simt::BM), (unsigned)ceil_div64(n, simt::BN));
  simt::gemm_kernel<<<grid, simt::THREADS, 0, as_stream(stream)>>>(ld, w, ep, m, (int)n, (int)k);
  STAC_LAUNCH_CHECK();
}

extern "C" int stac_mha_f32(const float* qkv, const int32_t* kv_len, int64_t batch, int64_t seq_len,
                            int64_t d_model, int64_t n_head, float* ctx, void* stream) {
  STAC_REQUIRE(qkv && kv_len && ctx && batch > 0 && batch < 65536 && seq_len > 0);
  if (d_model != n_head * kHd || n_head > 65535) return STAC_ERR_UNSUPPORTED_SHAPE;
  dim3 grid((unsigned)ceil_div64(seq_len, kQTile), (unsigned)n_head, (unsigned)batch);
  mha_f32_kernel<<<grid, kQTile, 0, as_stream(stream)>>>(qkv, kv_len, (int)seq_len, (int)d_model, ctx);
  STAC_LAUNCH_CHECK();
}

extern "C" int stac_log_softmax(const float* logits, int64_t rows, int64_t vocab, float* out,
                                int32_t* argmax, void* stream) {
  STAC_REQUIRE(logits && out && rows > 0 && rows < (1ll << 31) && vocab > 0 && vocab < (1 << 30));
  log_softmax_kernel<<<(unsigned)rows, 256, 0, as_stream(stream)>>>(logits, (int)vocab, out, argmax);
  STAC_LAUNCH_CHECK();
}

extern "C" int stac_kv_lengths(const float* wav_len, int64_t batch, int64_t t2, int round_rule, int32_t* out,
                               void* stream) {
  STAC_REQUIRE(out && batch > 0 && batch < (1 << 30) && t2 > 0 && t2 < (1 << 24));
  kv_lengths_kernel<<<(unsigned)ceil_div64(batch, 128), 128, 0, as_stream(stream)>>>(wav_len, (int)batch, (int)t2,
                                                                                    round_rule, out);
  STAC_LAUNCH_CHECK();
}

extern "C" int stac_cast_f32(const uint16_t* x, int64_t n, float* out, void* stream) {
  STAC_REQUIRE(x && out && n > 0);
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(n, 256 * 4), 148 * 32);
  cast_f32_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), n, out);
  STAC_LAUNCH_CHECK();
}

extern "C" int stac_cast_bf16(const float* x, int64_t n, uint16_t* out, void* stream) {
  STAC_REQUIRE(x && out && n > 0);
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(n, 256 * 4), 148 * 32);
  cast_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, n, reinterpret_cast<__nv_bfloat16*>(out));
  STAC_LAUNCH_CHECK();
}
