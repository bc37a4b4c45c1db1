// a4: ConvolutionFrontEnd = 2 x { reflect-pad 1, Conv2d 3x3 stride 2, LayerNorm(F,C), LeakyReLU }.
// Reference behaviour: SpeechBrain ConvolutionFrontEnd as configured at
//   /root/reference/stac-st/hparams/transformer_multitask.yaml:173-180 (call: inference.py:99).
// Layout convention of SpeechBrain's Conv2d: input [B,T,F(,C)], H = freq, W = time.
#include <algorithm>
#include "common.cuh"
#include "gemm_simt.cuh"

namespace {

constexpr int kMel = 80, kF1 = 40, kF2 = 20, kC = 256;
constexpr float kLnEps = 1e-5f, kSlope = 0.01f;

__device__ __forceinline__ void tc_fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) return -i;
  if (i >= n) return 2 * (n - 1) - i;
  return i;
}

// ---- block 0, fp32 output: one CTA per (b, t1); thread = output channel -------------------
__global__ void __launch_bounds__(kC)
conv0_ln_lrelu_f32_kernel(const float* __restrict__ feats, const float* __restrict__ w0,
                          const float* __restrict__ b0, const float* __restrict__ ln_g,
                          const float* __restrict__ ln_b, int frames, int t1_len, float* __restrict__ out_all) {
  __shared__ float in_s[3][kMel + 2];
  __shared__ float red[40];
  const int c = threadIdx.x;
  const int t1 = blockIdx.x, b = blockIdx.y;
  for (int i = c; i < 3 * (kMel + 1); i += kC) {
    const int kt = i / (kMel + 1), fi = i - kt * (kMel + 1);      // fi = f + 1, f in [-1, 79]
    const int t = reflect_idx(2 * t1 + kt - 1, frames);
    const int f = fi == 0 ? 1 : fi - 1;
    in_s[kt][fi] = __ldg(feats + ((int64_t)b * frames + t) * kMel + f);
  }
  float w[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) w[i] = __ldg(w0 + c * 9 + i);  // [kf][kt]
  const float bias = __ldg(b0 + c);
  __syncthreads();

  float acc[kF1];
  float lsum = 0.f;
#pragma unroll
  for (int f1 = 0; f1 < kF1; ++f1) {
    float a = bias;
#pragma unroll
    for (int kf = 0; kf < 3; ++kf)
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) a = fmaf(w[kf * 3 + kt], in_s[kt][2 * f1 + kf], a);
    acc[f1] = a;
    lsum += a;
  }
  const float mean = block_sum(lsum, red) * (1.0f / (kF1 * kC));
  float lsq = 0.f;
#pragma unroll
  for (int f1 = 0; f1 < kF1; ++f1) { const float d = acc[f1] - mean; lsq = fmaf(d, d, lsq); }
  const float var = block_sum(lsq, red) * (1.0f / (kF1 * kC));
  const float rstd = rsqrtf(var + kLnEps);
  float* out = out_all + ((int64_t)b * t1_len + t1) * (kF1 * kC);
#pragma unroll
  for (int f1 = 0; f1 < kF1; ++f1) {
    float y = (acc[f1] - mean) * rstd * __ldg(ln_g + f1 * kC + c) + __ldg(ln_b + f1 * kC + c);
    out[f1 * kC + c] = y > 0.f ? y : kSlope * y;
  }
}

// ---- block 0, bf16 output in the reflect-padded parity-split layout ----------------------
// CTA = (32 consecutive t1, b), 2 CTAs per SM.  Thread = (channel pair, freq parity): 2 channels x 20 freq
// bins.  The LayerNorm affine (2 x 10240 fp32 = 80 KB) is staged ONCE per CTA in shared memory: re-reading
// it from L2 for every time step was the whole cost of the first version (7.7 GB of L2 traffic per batch).
// Results are written to shared memory as conflict-free bf16x2 words into two staging planes that are
// exactly the two contiguous global runs of this time step (even / odd padded freq index); each run leaves
// the SM as one bulk async copy (cp.async.bulk shared -> global, full-line writes).  The three input feature
// rows of the next time step are prefetched into registers while the current one is computed.
constexpr int kRowsPerCta = 32;
constexpr int kPlane0Rows = 21, kPlane1Rows = 20;          // fp = 0,2,..,40  /  fp = 1,3,..,39
constexpr int kStageElems = (kPlane0Rows + kPlane1Rows) * kC;
constexpr int kInRow = kMel + 2;

struct Conv0Smem {
  float gamma[kF1 * kC];
  float beta[kF1 * kC];
  __nv_bfloat16 stage[kStageElems];
  float in_s[2][3][kInRow];
  float red[40];
};

__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(reinterpret_cast<uint64_t>(gdst)), "r"(ssrc), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(kC, 2)
conv0_ln_lrelu_bf16_kernel(const float* __restrict__ feats, const float* __restrict__ w0,
                           const float* __restrict__ b0, const float* __restrict__ ln_g,
                           const float* __restrict__ ln_b, int frames, int t1_len,
                           __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Conv0Smem& sm = *reinterpret_cast<Conv0Smem*>(smem_raw);
  const int tid = threadIdx.x;
  const int cp = tid & 127, fpar = tid >> 7;        // channels 2cp, 2cp+1; f1 = 2*j + fpar
  const int b = blockIdx.y;
  const int tp2 = (t1_len + 3) >> 1;
  const int t1_begin = blockIdx.x * kRowsPerCta;
  const int t1_end = min(t1_begin + kRowsPerCta, t1_len);

  for (int i = tid * 4; i < kF1 * kC; i += kC * 4) {
    *reinterpret_cast<float4*>(sm.gamma + i) = __ldg(reinterpret_cast<const float4*>(ln_g + i));
    *reinterpret_cast<float4*>(sm.beta + i) = __ldg(reinterpret_cast<const float4*>(ln_b + i));
  }
  float w[2][9], bias[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
#pragma unroll
    for (int i = 0; i < 9; ++i) w[q][i] = __ldg(w0 + (2 * cp + q) * 9 + i);
    bias[q] = __ldg(b0 + 2 * cp + q);
  }
  // input element this thread fetches for a time step: (kt, fi), fi = f + 1 in [0, 80]
  const bool loader = tid < 3 * (kMel + 1);
  const int l_kt = tid / (kMel + 1), l_fi = tid - l_kt * (kMel + 1);
  const int l_f = l_fi == 0 ? 1 : l_fi - 1;
  auto fetch = [&](int t1) -> float {
    const int t = reflect_idx(2 * t1 + l_kt - 1, frames);
    return __ldg(feats + ((int64_t)b * frames + t) * kMel + l_f);
  };
  if (loader) sm.in_s[0][l_kt][l_fi] = fetch(t1_begin);
  __syncthreads();

  for (int t1 = t1_begin; t1 < t1_end; ++t1) {
    const int cur = (t1 - t1_begin) & 1;
    float nxt = 0.f;
    if (loader && t1 + 1 < t1_end) nxt = fetch(t1 + 1);
    float acc[2][20];
    float lsum = 0.f;
#pragma unroll
    for (int j = 0; j < 20; ++j) {
      const int f1 = 2 * j + fpar;
      float x[9];
#pragma unroll
      for (int kf = 0; kf < 3; ++kf)
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) x[kf * 3 + kt] = sm.in_s[cur][kt][2 * f1 + kf];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float a = bias[q];
#pragma unroll
        for (int i = 0; i < 9; ++i) a = fmaf(w[q][i], x[i], a);
        acc[q][j] = a;
        lsum += a;
      }
    }
    const float mean = block_sum(lsum, sm.red) * (1.0f / (kF1 * kC));
    float lsq = 0.f;
#pragma unroll
    for (int j = 0; j < 20; ++j)
#pragma unroll
      for (int q = 0; q < 2; ++q) { const float d = acc[q][j] - mean; lsq = fmaf(d, d, lsq); }
    // the previous time step's bulk copies must have finished reading the staging planes
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    const float var = block_sum(lsq, sm.red) * (1.0f / (kF1 * kC));   // (its barriers also publish the wait)
    const float rstd = rsqrtf(var + kLnEps);

    // plane 0 (even fp): row i <-> fp = 2i <-> f1 = 2i-1 (row 0 = reflection of f1 = 1); odd f1 -> fpar = 1
    // plane 1 (odd fp):  row i <-> fp = 2i+1 <-> f1 = 2i;                              even f1 -> fpar = 0
    uint32_t* base = reinterpret_cast<uint32_t*>(sm.stage);
#pragma unroll
    for (int j = 0; j < 20; ++j) {
      const int f1 = 2 * j + fpar;
      const float2 g = *reinterpret_cast<const float2*>(sm.gamma + f1 * kC + 2 * cp);
      const float2 be = *reinterpret_cast<const float2*>(sm.beta + f1 * kC + 2 * cp);
      float y0 = (acc[0][j] - mean) * rstd * g.x + be.x;
      float y1 = (acc[1][j] - mean) * rstd * g.y + be.y;
      y0 = y0 > 0.f ? y0 : kSlope * y0;
      y1 = y1 > 0.f ? y1 : kSlope * y1;
      const uint32_t pk = pack_bf16x2(y0, y1);
      if (fpar == 1) {
        base[(j + 1) * (kC / 2) + cp] = pk;                 // plane 0, row j+1
        if (j == 0) base[cp] = pk;                           // fp = 0 mirrors f1 = 1
      } else {
        base[(kPlane0Rows + j) * (kC / 2) + cp] = pk;       // plane 1, row j
      }
    }
    if (loader && t1 + 1 < t1_end) sm.in_s[cur ^ 1][l_kt][l_fi] = nxt;
    tc_fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      int tps[3];
      int n_tp = 0;
      tps[n_tp++] = t1 + 1;
      if (t1 == 1) tps[n_tp++] = 0;
      if (t1 == t1_len - 2) tps[n_tp++] = t1_len + 1;
      const uint32_t s0 = static_cast<uint32_t>(__cvta_generic_to_shared(sm.stage));
      for (int i = 0; i < n_tp; ++i) {
        const int tp = tps[i];
        const int64_t plane_t = (int64_t)b * 4 + (tp & 1) * 2;
        bulk_store(out + ((plane_t + 0) * tp2 + (tp >> 1)) * (21 * kC), s0, kPlane0Rows * kC * 2);
        bulk_store(out + ((plane_t + 1) * tp2 + (tp >> 1)) * (21 * kC), s0 + kPlane0Rows * kC * 2,
                   kPlane1Rows * kC * 2);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- LayerNorm over a whole row (F*C elements) + LeakyReLU ------------------------------
template <bool kOutBf16>
__global__ void __launch_bounds__(256)
group_ln_lrelu_kernel(const float* __restrict__ x, int dim, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, float slope, void* __restrict__ out_v) {
  extern __shared__ float rowbuf[];
  __shared__ float red[40];
  const int64_t r = blockIdx.x;
  const float* xr = x + r * dim;
  float lsum = 0.f;
  for (int i = threadIdx.x * 4; i < dim; i += 256 * 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(xr + i));
    *reinterpret_cast<float4*>(rowbuf + i) = v;
    lsum += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = block_sum(lsum, red) / (float)dim;
  float lsq = 0.f;
  for (int i = threadIdx.x * 4; i < dim; i += 256 * 4) {
    const float4 v = *reinterpret_cast<const float4*>(rowbuf + i);
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    lsq += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(block_sum(lsq, red) / (float)dim + eps);
  for (int i = threadIdx.x * 4; i < dim; i += 256 * 4) {
    const float4 v = *reinterpret_cast<const float4*>(rowbuf + i);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + i));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + i));
    float y[4] = {(v.x - mean) * rstd * g.x + bb.x, (v.y - mean) * rstd * g.y + bb.y,
                  (v.z - mean) * rstd * g.z + bb.z, (v.w - mean) * rstd * g.w + bb.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = y[j] > 0.f ? y[j] : slope * y[j];
    if constexpr (kOutBf16) {
      uint2 pk = make_uint2(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]));
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out_v) + r * dim + i) = pk;
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out_v) + r * dim + i) =
          make_float4(y[0], y[1], y[2], y[3]);
    }
  }
}

// ---- block 1 convolution as fp32 implicit GEMM: row = (b, t2, f2), k = (kf*3+kt)*256 + cin ----
struct Conv1Loader {
  const float* x;   // [B, T1, 40, 256]
  int64_t m;        // B*T2*20
  int t1_len, t2_len;
  __device__ __forceinline__ float4 load4(int64_t row, int kk) const {
    if (row >= m) return make_float4(0.f, 0.f, 0.f, 0.f);
    const int f2 = (int)(row % kF2);
    const int64_t bt = row / kF2;
    const int t2 = (int)(bt % t2_len);
    const int64_t b = bt / t2_len;
    const int tap = kk >> 8, cin = kk & 255;
    const int kf = tap / 3, kt = tap - 3 * kf;
    const int t = reflect_idx(2 * t2 + kt - 1, t1_len);
    const int f = reflect_idx(2 * f2 + kf - 1, kF1);
    return __ldg(reinterpret_cast<const float4*>(x + ((b * t1_len + t) * kF1 + f) * (int64_t)kC + cin));
  }
};

}  // namespace

extern "C" int64_t stac_conv0_padded_elems(int64_t batch, int64_t t1) {
  return batch * 4 * ((t1 + 3) / 2) * 21 * kC;
}

extern "C" int stac_conv0_ln_lrelu(const float* feats, const float* w0, const float* b0,
                                   const float* ln_g, const float* ln_b, int64_t batch,
                                   int64_t frames, void* out, int out_mode, void* stream) {
  STAC_REQUIRE(feats && w0 && b0 && ln_g && ln_b && out);
  STAC_REQUIRE(batch > 0 && batch < 65536 && frames >= 3 && frames < (1 << 30));
  const int t1 = (int)((frames - 1) / 2 + 1);
  if (out_mode == STAC_DT_F32) {
    dim3 grid((unsigned)t1, (unsigned)batch);
    conv0_ln_lrelu_f32_kernel<<<grid, kC, 0, as_stream(stream)>>>(feats, w0, b0, ln_g, ln_b, (int)frames, t1,
                                                                  reinterpret_cast<float*>(out));
  } else if (out_mode == STAC_DT_BF16) {
    static bool attr = false;
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(conv0_ln_lrelu_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(Conv0Smem));
      if (e != cudaSuccess) return (int)e;
      attr = true;
    }
    dim3 grid((unsigned)ceil_div64(t1, kRowsPerCta), (unsigned)batch);
    conv0_ln_lrelu_bf16_kernel<<<grid, kC, sizeof(Conv0Smem), as_stream(stream)>>>(
        feats, w0, b0, ln_g, ln_b, (int)frames, t1, reinterpret_cast<__nv_bfloat16*>(out));
  } else {
    return STAC_ERR_INVALID_ARGUMENT;
  }
  STAC_LAUNCH_CHECK();
}

extern "C" int stac_group_ln_lrelu(const float* x, int64_t rows, int64_t dim, const float* gamma,
                                   const float* beta, float eps, float slope, void* out,
                                   int out_dtype, void* stream) {
  STAC_REQUIRE(x && gamma && beta && out && rows > 0 && rows < (1ll << 31));
  STAC_REQUIRE(dim > 0 && dim % 4 == 0 && dim <= 12288);
  const size_t smem = (size_t)dim * sizeof(float);
  if (out_dtype == STAC_DT_BF16) {
    static bool set_b = false;
    if (!set_b) { cudaFuncSetAttribute(group_ln_lrelu_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152); set_b = true; }
    group_ln_lrelu_kernel<true><<<(unsigned)rows, 256, smem, as_stream(stream)>>>(x, (int)dim, gamma, beta, eps, slope, out);
  } else if (out_dtype == STAC_DT_F32) {
    static bool set_f = false;
    if (!set_f) { cudaFuncSetAttribute(group_ln_lrelu_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152); set_f = true; }
    group_ln_lrelu_kernel<false><<<(unsigned)rows, 256, smem, as_stream(stream)>>>(x, (int)dim, gamma, beta, eps, slope, out);
  } else {
    return STAC_ERR_INVALID_ARGUMENT;
  }
  STAC_LAUNCH_CHECK();
}

extern "C" int stac_conv1_f32(const float* x, const float* w1, const float* b1, int64_t batch,
                              int64_t t1, float* out, void* stream) {
  STAC_REQUIRE(x && w1 && b1 && out && batch > 0 && t1 >= 2 && t1 < (1 << 30));
  const int t2 = (int)((t1 - 1) / 2 + 1);
  const int64_t m = batch * t2 * kF2;
  Conv1Loader ld{x, m, (int)t1, t2};
  simt::LinearEpilogue ep{b1, nullptr, 0, STAC_ACT_NONE, out, kC};
  dim3 grid((unsigned)ceil_div64(m, simt::BM), (unsigned)(kC / simt::BN));
  simt::gemm_kernel<<<grid, simt::THREADS, 0, as_stream(stream)>>>(ld, w1, ep, m, kC, 9 * kC);
  STAC_LAUNCH_CHECK();
}
