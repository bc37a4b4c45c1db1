// a4: ConvolutionFrontEnd = 2 x { reflect-pad 1, Conv2d 3x3 stride 2, LayerNorm(F,C), LeakyReLU }.
// Reference behaviour: SpeechBrain ConvolutionFrontEnd as configured at
//   /root/reference/stac-st/hparams/transformer_multitask.yaml:173-180 (call: inference.py:99).
// Layout convention of SpeechBrain's Conv2d: input [B,T,F(,C)], H = freq, W = time.
#include <algorithm>
#include "common.cuh"
#include "gemm_simt.cuh"

int stac_conv0_tc_launch(const float* feats, const float* w0, const float* b0, const float* ln_g, const float* ln_b,
                         int64_t batch, int64_t frames, int t1, uint16_t* out, cudaStream_t st,
                         const uint32_t* utt_max, int per_utt, float top_db, const float* mean, const float* std);

namespace {

constexpr int kMel = 80, kF1 = 40, kF2 = 20, kC = 256;
constexpr float kLnEps = 1e-5f, kSlope = 0.01f;

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
    // tensor-core version (conv0_tc.cu)
    return stac_conv0_tc_launch(feats, w0, b0, ln_g, ln_b, batch, frames, t1, reinterpret_cast<uint16_t*>(out),
                                as_stream(stream), nullptr, 1, 0.f, nullptr, nullptr);
  } else {
    return STAC_ERR_INVALID_ARGUMENT;
  }
  STAC_LAUNCH_CHECK();
}

extern "C" int stac_conv0_topdb_norm_bf16(const float* logmel_db, const uint32_t* utt_max_ordered, int per_utterance,
                                          float top_db, const float* mean, const float* std, const float* w0,
                                          const float* b0, const float* ln_g, const float* ln_b, int64_t batch,
                                          int64_t frames, uint16_t* out, void* stream) {
  STAC_REQUIRE(logmel_db && utt_max_ordered && w0 && b0 && ln_g && ln_b && out);
  STAC_REQUIRE((mean == nullptr) == (std == nullptr));
  STAC_REQUIRE(batch > 0 && batch < 65536 && frames >= 3 && frames < (1 << 30));
  const int t1 = (int)((frames - 1) / 2 + 1);
  return stac_conv0_tc_launch(logmel_db, w0, b0, ln_g, ln_b, batch, frames, t1, out, as_stream(stream),
                              utt_max_ordered, per_utterance, top_db, mean, std);
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
