// Write-only HBM bandwidth on B200 by store form (what bounds conv0, the QKV projection and the CTC head's second pass):
// plain st.global.v4, st.global.cs (streaming), st.global with an L2 evict_first / no_allocate policy, and bulk async
// copies shared -> global (the form the kernels' epilogues use) with and without an L2 cache hint.  2 GiB per launch.
// Build + run (on the GPU box):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_write tools/ubench_write.cu && /tmp/ubench_write
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int kMode>
__global__ void __launch_bounds__(256) st_kernel(uint4* out, size_t n16) {
  const uint4 v = make_uint4(1, 2, 3, threadIdx.x);
  uint64_t pol = 0;
  if (kMode == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  if (kMode == 3) asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(pol));
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    if (kMode == 0) out[i] = v;
    else if (kMode == 1) asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(out + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(out + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
  }
}

// one warp per CTA issues bulk copies of kChunk bytes out of a shared-memory tile; 4 in flight
template <int kHint>
__global__ void __launch_bounds__(128) bulk_kernel(unsigned char* out, size_t bytes, int chunk) {
  extern __shared__ __align__(128) unsigned char tile[];
  for (int i = threadIdx.x; i < chunk / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(tile)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x != 0) return;
  uint64_t pol = 0;
  if (kHint == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const size_t n_chunks = bytes / chunk;
  int inflight = 0;
  for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    if (kHint == 0)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * chunk), "r"(smem_u32(tile)), "r"(chunk) : "memory");
    else
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(out + c * chunk), "r"(smem_u32(tile)), "r"(chunk), "l"(pol) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (++inflight >= 8) { asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); inflight = 4; }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class F>
static float timed(F f) {
  for (int i = 0; i < 2; ++i) f();
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  const size_t bytes = 2ull << 30;
  unsigned char* buf;
  if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  const size_t n16 = bytes / 16;
  float ms;
  ms = timed([&] { cudaMemsetAsync(buf, 0, bytes); });
  printf("cudaMemset                         %7.3f ms  %6.0f GB/s\n", ms, bytes / ms / 1e6);
  ms = timed([&] { st_kernel<0><<<148 * 16, 256>>>((uint4*)buf, n16); });
  printf("st.global.v4                       %7.3f ms  %6.0f GB/s\n", ms, bytes / ms / 1e6);
  ms = timed([&] { st_kernel<1><<<148 * 16, 256>>>((uint4*)buf, n16); });
  printf("st.global.cs.v4                    %7.3f ms  %6.0f GB/s\n", ms, bytes / ms / 1e6);
  ms = timed([&] { st_kernel<2><<<148 * 16, 256>>>((uint4*)buf, n16); });
  printf("st.global.v4 L2 evict_first        %7.3f ms  %6.0f GB/s\n", ms, bytes / ms / 1e6);
  ms = timed([&] { st_kernel<3><<<148 * 16, 256>>>((uint4*)buf, n16); });
  printf("st.global.v4 L2 evict_unchanged    %7.3f ms  %6.0f GB/s\n", ms, bytes / ms / 1e6);
  for (int chunk : {4096, 16384, 65536}) {
    cudaFuncSetAttribute(bulk_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk);
    cudaFuncSetAttribute(bulk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk);
    for (int ctas : {148, 296, 592}) {
      ms = timed([&] { bulk_kernel<0><<<ctas, 128, chunk>>>(buf, bytes, chunk); });
      printf("bulk s->g %6d B x %3d CTAs        %7.3f ms  %6.0f GB/s\n", chunk, ctas, ms, bytes / ms / 1e6);
      ms = timed([&] { bulk_kernel<1><<<ctas, 128, chunk>>>(buf, bytes, chunk); });
      printf("bulk s->g %6d B x %3d CTAs hint   %7.3f ms  %6.0f GB/s\n", chunk, ctas, ms, bytes / ms / 1e6);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
