// Probe of the tensor-memory conventions csrc/attention_tc2.cu relies on (run on a B200 when its parity test is red):
//   1. SS-form MMA with a 128-row B tile: S[128 x 128] = Q[128 x 64] . K[128 x 64]^T, both K-major, 128-byte swizzle;
//   2. TS-form MMA: O[128 x 64] = P[128 x 128] . V[128 x 64] with P in tensor memory, under three layouts of P:
//        H1 (the kernel's): 32-bit column c of lane r holds keys (2c, 2c + 1), even key in the low half; K = 16 step kk
//                           reads columns 8 kk .. 8 kk + 7
//        H2: the same with the halves swapped
//        H3: one bf16 per 32-bit column (low half), 16 columns per K = 16 step
//      V is the MN-major B operand of the kernel ([keys][64] rows of 128 bytes, swizzled, + 2048 bytes per 16 keys).
// All values are small integers, so every product is exact: a layout either matches bit for bit or it does not.
// Build + run (on the GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 --expt-relaxed-constexpr -o /tmp/probe tools/probe_ts_mma.cu && /tmp/probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../stac_speech_translation_b200/csrc/tc_common.cuh"

int stac_grid_limit() { return 148; }      // (declared by common.cuh; unused here)

using namespace tc;

__device__ __forceinline__ void tmem_st32_(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__host__ __device__ inline int qv(int r, int c) { return (r * 3 + c * 5) % 7 - 3; }
__host__ __device__ inline int kv(int r, int c) { return (r * 5 + c * 3) % 5 - 2; }
__host__ __device__ inline int pv(int r, int k) { return (r * 7 + k * 3) % 13 - 6; }
__host__ __device__ inline int vv(int k, int d) { return (k * 5 + d) % 11 - 5; }

__device__ __forceinline__ uint32_t bf16_bits(int x) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16((float)x));
}

// out: [0 .. 128*128) S ; then three blocks of 128*64 for H1, H2, H3
__global__ void __launch_bounds__(128, 1) probe_kernel(float* out) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* sp = smem_raw + (sbase - smem_u32(smem_raw));
  // smem: Q 16 KB | K 16 KB | V 16 KB | barrier | tmem slot
  const uint32_t q_s = sbase, k_s = sbase + 16384, v_s = sbase + 32768, bar = sbase + 49152, slot = bar + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // K-major tiles: row r = 128 bytes (64 bf16), 16-byte chunk j stored at chunk j ^ (r & 7)
  for (int i = tid; i < 128 * 64; i += 128) {
    const int r = i >> 6, c = i & 63;
    const uint32_t off = r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2;
    *reinterpret_cast<uint16_t*>(sp + off) = (uint16_t)bf16_bits(qv(r, c));
    *reinterpret_cast<uint16_t*>(sp + 16384 + off) = (uint16_t)bf16_bits(kv(r, c));
    *reinterpret_cast<uint16_t*>(sp + 32768 + off) = (uint16_t)bf16_bits(vv(r, c));     // V[key r][d c], same tile shape
  }
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  const int r = warp * 32 + lane;
  uint32_t phase = 0;

  // ---- 1. S = Q K^T (columns 0..127) ----
  if (warp == 0 && elect_one()) {
    const uint64_t qd = make_smem_desc_sw128(q_s), kd = make_smem_desc_sw128(k_s);
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, qd + 2 * k, kd + 2 * k, make_idesc_bf16(128, 128), k != 0);
    umma_commit(bar);
  }
  mbar_wait(bar, phase); phase ^= 1;
  tc_fence_after();
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    tmem_ld32(tmem + c * 32 + lane_off, v);
    tmem_ld_wait();
    for (int e = 0; e < 32; ++e) out[r * 128 + c * 32 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();

  // ---- 2. O = P V with P in tensor memory (P at columns 128.., O at columns 384..) ----
  for (int hyp = 0; hyp < 3; ++hyp) {
    const int cols = hyp == 2 ? 128 : 64;
    for (int c0 = 0; c0 < cols; c0 += 32) {
      uint32_t v[32];
      for (int e = 0; e < 32; ++e) {
        const int c = c0 + e;
        if (hyp == 0) v[e] = bf16_bits(pv(r, 2 * c)) | (bf16_bits(pv(r, 2 * c + 1)) << 16);
        else if (hyp == 1) v[e] = bf16_bits(pv(r, 2 * c + 1)) | (bf16_bits(pv(r, 2 * c)) << 16);
        else v[e] = bf16_bits(pv(r, c));
      }
      tmem_st32_(tmem + 128 + c0 + lane_off, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0 && elect_one()) {
      const uint64_t vd = make_smem_desc_sw128(v_s);
      const uint32_t idesc = make_idesc_bf16(128, 64) | (1u << 16);          // B (V) MN-major
      for (int k = 0; k < 8; ++k)
        umma_ts(tmem + 384, tmem + 128 + (hyp == 2 ? 16 : 8) * k, vd + 128 * k, idesc, k != 0);
      umma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem + 384 + c * 32 + lane_off, v);
      tmem_ld_wait();
      for (int e = 0; e < 32; ++e) out[128 * 128 + hyp * 128 * 64 + r * 64 + c * 32 + e] = __uint_as_float(v[e]);
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  const size_t n = 128 * 128 + 3 * 128 * 64;
  float* d_out = nullptr;
  cudaMalloc(&d_out, n * sizeof(float));
  cudaMemset(d_out, 0xff, n * sizeof(float));
  const int smem = 49152 + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> h(n);
  cudaMemcpy(h.data(), d_out, n * sizeof(float), cudaMemcpyDeviceToHost);
  long bad_s = 0;
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < 128; ++c) {
      int want = 0;
      for (int k = 0; k < 64; ++k) want += qv(r, k) * kv(c, k);
      bad_s += h[r * 128 + c] != (float)want;
    }
  printf("SS  S = Q K^T, 128-row B tile (K-major, swizzle 128): %s (%ld of 16384 wrong)\n", bad_s ? "MISMATCH" : "exact", bad_s);
  const char* names[3] = {"H1 packed pairs, even key low (the kernel's layout)", "H2 packed pairs, even key high",
                          "H3 one bf16 per column"};
  for (int hyp = 0; hyp < 3; ++hyp) {
    long bad = 0;
    for (int r = 0; r < 128; ++r)
      for (int d = 0; d < 64; ++d) {
        int want = 0;
        for (int k = 0; k < 128; ++k) want += pv(r, k) * vv(k, d);
        bad += h[128 * 128 + hyp * 128 * 64 + r * 64 + d] != (float)want;
      }
    printf("TS  O = P V, %-52s: %s (%ld of 8192 wrong)\n", names[hyp], bad ? "mismatch" : "EXACT", bad);
  }
  return 0;
}
