// Latency of the synchronisation instructions the warp-specialised kernels sit on (B200, sm_100a), one warp alone on an
// SM and four warps on one scheduler: a satisfied mbarrier wait, the tcgen05 fences, tcgen05.st + wait::st,
// tcgen05.ld + wait::ld, mbarrier.arrive, fence.proxy.async.  clock64 around 64 repetitions of each.
// Build + run (on the GPU box):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_sync tools/ubench_sync.cu && /tmp/ubench_sync
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define TIME(slot, body)                                                    \
  do {                                                                      \
    __syncwarp();                                                           \
    const long long t0 = clock64();                                         \
    _Pragma("unroll 1") for (int i = 0; i < 64; ++i) { body }               \
    const long long t1 = clock64();                                         \
    if (threadIdx.x == 0) out[slot] = (t1 - t0) / 64;                       \
  } while (0)

__global__ void sync_kernel(long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[1])), "r"(1 << 19) : "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[0])) : "memory");     // phase 0 complete
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t b0 = smem_u32(&bar[0]), b1 = smem_u32(&bar[1]);
  uint32_t v[8] = {1, 2, 3, 4, 5, 6, 7, 8};
  uint32_t acc = 0;
  if (warp >= (int)(blockDim.x >> 5)) return;
  long long* o = out;
  (void)o;
  TIME(0, { acc += i; });                                                                     // loop overhead
  TIME(1, {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(b0), "r"(0) : "memory");
    acc += ok;
  });                                                                                         // satisfied try_wait
  TIME(2, { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); });
  TIME(3, { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); });
  TIME(4, {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(tmem), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  });                                                                                         // st.x8 + wait::st
  TIME(5, {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(tmem) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc += v[0];
  });                                                                                         // ld.x8 + wait::ld
  TIME(6, { if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b1) : "memory"); __syncwarp(); });
  TIME(7, { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); });
  TIME(8, {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(tmem), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b1) : "memory");
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(b0), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    acc += ok;
  });                                                                                         // a producer's whole hand-off
  TIME(9, { asm volatile("bar.sync 1, 32;" ::: "memory"); });
  if (acc == 0x7fffffffu) sink[0] = acc + v[1];
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

int main() {
  long long* out;
  uint32_t* sink;
  cudaMalloc(&out, 16 * sizeof(long long));
  cudaMalloc(&sink, 4);
  const char* names[10] = {"loop overhead", "try_wait on a completed phase", "tcgen05.fence::after_thread_sync",
                           "tcgen05.fence::before_thread_sync", "tcgen05.st.x8 + wait::st", "tcgen05.ld.x8 + wait::ld",
                           "mbarrier.arrive (lane 0) + syncwarp", "fence.proxy.async.shared::cta",
                           "st + wait::st + fence + arrive + try_wait + fence", "bar.sync (32 threads)"};
  for (int threads : {32, 128, 512}) {
    cudaMemset(out, 0, 16 * sizeof(long long));
    sync_kernel<<<1, threads>>>(out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[16];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%d threads (warp 0's view), clocks per repetition:\n", threads);
    for (int i = 0; i < 10; ++i) printf("  %-52s %5lld\n", names[i], h[i]);
  }
  return 0;
}
