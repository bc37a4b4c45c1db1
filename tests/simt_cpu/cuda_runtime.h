// TEST INFRASTRUCTURE ONLY: a CPU emulation of the CUDA execution model, good enough to run the library's plain SIMT
// kernels (no tensor cores, TMA or mbarriers) from their real source in the CPU test-suite.  tests/simt_cpu/build.py
// compiles selected csrc/*.cu files against this header instead of the CUDA runtime (this directory shadows
// <cuda_runtime.h> and <cuda_bf16.h>) after rewriting `kernel<<<grid, block, smem, stream>>>(args)` into simt::launch.
//
// Model: the blocks of a launch run one after another; the threads of a block are OS threads; __syncthreads() and the
// warp-synchronous primitives are barriers over the block / over 32 consecutive threads; `__shared__` becomes `static`
// (one copy, reused by the next block); dynamic shared memory is one buffer per launch.  Nothing here is ever linked
// into the product.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n)

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

struct __nv_bfloat16 { uint16_t v; };
struct __nv_bfloat162 { uint16_t lo, hi; };
inline uint16_t simt_bf16_rn(float f) {
  uint32_t u; std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline __nv_bfloat162 __floats2bfloat162_rn(float lo, float hi) { return {simt_bf16_rn(lo), simt_bf16_rn(hi)}; }
inline __nv_bfloat16 __float2bfloat16(float f) { return {simt_bf16_rn(f)}; }
inline __nv_bfloat16 __float2bfloat16_rn(float f) { return {simt_bf16_rn(f)}; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline float __expf(float x) { return std::exp(x); }
inline float __logf(float x) { return std::log(x); }
template <class T> inline T __ldg(const T* p) { return *p; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline unsigned atomicMax(unsigned* p, unsigned v) {       // (threads of a block are OS threads)
  unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
using std::max;
using std::min;

namespace simt {
struct Block {
  std::barrier<> all;
  std::vector<std::unique_ptr<std::barrier<>>> warp;
  std::vector<uint32_t> scratch;       // one slot per thread (shuffle / ballot exchange)
  explicit Block(int n) : all(n), scratch(n) {
    for (int w = 0; w * 32 < n; ++w) warp.emplace_back(new std::barrier<>(std::min(32, n - w * 32)));
  }
};
inline thread_local Block* cur = nullptr;
inline thread_local int linear_tid = 0;
inline dim3 g_block_dim, g_grid_dim;
inline unsigned char* dyn_smem = nullptr;

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, F body);
}  // namespace simt

inline thread_local uint3 threadIdx, blockIdx;
#define blockDim (simt::g_block_dim)
#define gridDim (simt::g_grid_dim)

inline void __syncthreads() { simt::cur->all.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { simt::cur->warp[simt::linear_tid >> 5]->arrive_and_wait(); }
inline uint32_t simt_exchange(uint32_t v, int src_lane) {
  simt::Block* b = simt::cur;
  const int base = simt::linear_tid & ~31;
  b->scratch[simt::linear_tid] = v;
  b->warp[simt::linear_tid >> 5]->arrive_and_wait();
  const uint32_t r = b->scratch[base + src_lane];
  b->warp[simt::linear_tid >> 5]->arrive_and_wait();
  return r;
}
inline float __shfl_xor_sync(unsigned, float v, int o) {
  return __uint_as_float(simt_exchange(__float_as_uint(v), (simt::linear_tid & 31) ^ o));
}
inline float __shfl_sync(unsigned, float v, int src) { return __uint_as_float(simt_exchange(__float_as_uint(v), src & 31)); }
inline int __shfl_sync(unsigned, int v, int src) { return (int)simt_exchange((uint32_t)v, src & 31); }
inline int __shfl_xor_sync(unsigned, int v, int o) { return (int)simt_exchange((uint32_t)v, (simt::linear_tid & 31) ^ o); }
inline unsigned __ballot_sync(unsigned, bool pred) {
  simt::Block* b = simt::cur;
  const int base = simt::linear_tid & ~31;
  b->scratch[simt::linear_tid] = pred ? 1u : 0u;
  b->warp[simt::linear_tid >> 5]->arrive_and_wait();
  unsigned m = 0;
  const int n = std::min<int>(32, (int)b->scratch.size() - base);
  for (int l = 0; l < n; ++l) m |= b->scratch[base + l] << l;
  b->warp[simt::linear_tid >> 5]->arrive_and_wait();
  return m;
}
inline bool __any_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) != 0; }

template <class F>
inline void simt::launch(dim3 grid, dim3 block, size_t smem, F body) {
  g_block_dim = block;
  g_grid_dim = grid;
  const int n = (int)(block.x * block.y * block.z);
  std::vector<unsigned char> dyn(smem + 16);
  dyn_smem = dyn.data();
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        Block blk(n);
        std::vector<std::thread> threads;
        for (int t = 0; t < n; ++t)
          threads.emplace_back([&, t] {
            cur = &blk;
            linear_tid = t;
            threadIdx = {(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
            blockIdx = {bx, by, bz};
            body();
          });
        for (auto& th : threads) th.join();
      }
  dyn_smem = nullptr;
}
