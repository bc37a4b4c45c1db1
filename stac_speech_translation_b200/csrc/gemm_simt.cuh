// fp32 CUDA-core GEMM template used by the fp32 (verification-precision) mode:
//   C[M,N] = epilogue( sum_k A(row,k) * W[n,k] ),  W row-major [N,K] (nn.Linear layout).
// A is supplied by a loader functor so the same kernel serves plain row-major activations and
// the implicit-GEMM view of the second convolution block.  64x64x16 tiles, 4x4 per thread.
#pragma once
#include "common.cuh"

namespace simt {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256;

struct RowMajorA {
  const float* a;
  int64_t m, k;
  __device__ __forceinline__ float4 load4(int64_t row, int kk) const {
    if (row >= m) return make_float4(0.f, 0.f, 0.f, 0.f);
    return __ldg(reinterpret_cast<const float4*>(a + row * k + kk));
  }
};

struct LinearEpilogue {
  const float* bias;
  const float* resid;
  int64_t resid_period;
  int act;
  float* c;
  int64_t n;
  __device__ __forceinline__ void store(int64_t row, int col, float v) const {
    if (bias) v += __ldg(bias + col);
    if (act == STAC_ACT_GELU_ERF) v = gelu_erf(v);
    if (resid) {
      const int64_t rr = resid_period > 0 ? row % resid_period : row;
      v += resid[rr * n + col];
    }
    c[row * n + col] = v;
  }
};

template <class ALoader, class Epilogue>
__global__ void __launch_bounds__(THREADS)
gemm_kernel(ALoader A, const float* __restrict__ w, Epilogue ep, int64_t m, int n, int k) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int col0 = blockIdx.y * BN;
  const int lr = tid >> 2, lk = (tid & 3) * 4;  // loader coordinates: 64 rows x 4 float4
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < k; k0 += BK) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), wv = av;
    if (k0 + lk < k) {
      av = A.load4(row0 + lr, k0 + lk);
      if (col0 + lr < n) wv = __ldg(reinterpret_cast<const float4*>(w + (int64_t)(col0 + lr) * k + k0 + lk));
    }
    __syncthreads();
    As[lk + 0][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
    Ws[lk + 0][lr] = wv.x; Ws[lk + 1][lr] = wv.y; Ws[lk + 2][lr] = wv.z; Ws[lk + 3][lr] = wv.w;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 w4 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float a_[4] = {a4.x, a4.y, a4.z, a4.w};
      const float w_[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a_[i], w_[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = row0 + ty * 4 + i;
    if (row >= m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + tx * 4 + j;
      if (col < n) ep.store(row, col, acc[i][j]);
    }
  }
}

}  // namespace simt
