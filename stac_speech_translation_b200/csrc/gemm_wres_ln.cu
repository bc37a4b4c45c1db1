// Output projection of the attention block fused with the residual add AND the LayerNorm that follows it (bf16 mode,
// d_model = 256):   x += ctx . W_o^T + b_o   (fp32 residual stream, in place)   and   h = LayerNorm(x) (bf16),
// i.e. the "dropout1(out) + src" and "norm2(src)" lines of SpeechBrain's pre-LN TransformerEncoderLayer, reached from
// /root/reference/stac-st/modules/TransformerMultiTask.py:304-308.  north_star: "fused LayerNorm + ..." - the LayerNorm
// is folded into the kernel that PRODUCES the residual stream, whose CTA tile holds whole rows (N = 256 = d_model).
//
// Why: stac_gemm_bf16 (weight-resident, TMA reduce-add into x) followed by stac_layernorm moves ctx 25 MB + x 49 MB read
// + 49 MB write (the reduce) + x 49 MB read + h 25 MB write = 197 MB per layer at the benchmark shape in 27 + 15 us.
// Here the epilogue loads the old x rows itself, forms x_new in registers, keeps it in the accumulator's own TMEM
// columns (tcgen05.st) while the row statistics are completed, and emits both x_new (fp32) and LayerNorm(x_new) (bf16):
// 148 MB, one launch, and the next GEMM's A operand is produced directly.
//
// Structure = gemm_wres.cu (one resident 256 x 256 block of W_o per CTA, activations through a 3-stage ring, two TMEM
// accumulators); the epilogue differs.  Warps 2-9 (TMEM lane quarter = warp & 3, column half = (warp - 2) >> 2):
//   pass 1, per 32-column chunk: acc (TMEM) + bias + old x (plain 16-byte loads, thread = row) -> x_new: back into TMEM,
//           into the staging tile -> TMA store to x; running sum / sum of squares of the row half;
//   the two warps that share a row exchange their partial statistics through shared memory (one named barrier);
//   pass 2, per chunk: x_new (TMEM) -> (x - mean) rstd gamma + beta -> bf16 -> staging tile -> TMA store to h.
#include <algorithm>
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int kK = 256, BM = 128, BK = 64, BN = 256;
constexpr int kABytes = BM * BK * 2;              // 16 KB
constexpr int kBBlock = BN * BK * 2;              // 32 KB: one k-block of the resident weight tile
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kAStages = 3;
constexpr int kOffB = 0;
constexpr int kOffA = 4 * kBBlock;
constexpr int kOffStage = kOffA + kAStages * kABytes;          // 8 x 4 KB staging tiles
constexpr int kOffVec = kOffStage + kEpiWarps * 4096;          // float bias[256] | gamma[256] | beta[256]
constexpr int kOffStat = kOffVec + 3 * BN * 4;                 // float2 [2 tile parities][2 column halves][128 rows]
constexpr int kOffBar = kOffStat + 2 * 2 * BM * 8;
constexpr int kNumBars = 1 + 2 * kAStages + 4;
constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16 + 1024;
static_assert(kSmemBytes <= 232448, "shared memory budget");

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// tcgen05.st: thread i of the warp writes TMEM lane (base_lane + i), 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(kThreads, 1)
outproj_ln_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_h,
                  const float* __restrict__ bias, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                  const float eps, const float* x_in, const int m_rows, const int num_m_tiles) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bars = sbase + kOffBar;
  auto b_full = [&]() { return bars; };
  auto a_full = [&](int s) { return bars + 8u * (1 + s); };
  auto a_empty = [&](int s) { return bars + 8u * (1 + kAStages + s); };
  auto t_full = [&](int i) { return bars + 8u * (1 + 2 * kAStages + i); };
  auto t_empty = [&](int i) { return bars + 8u * (3 + 2 * kAStages + i); };
  const uint32_t tmem_slot = bars + 8u * kNumBars;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* vec_s = reinterpret_cast<float*>(sptr + kOffVec);
  float2* stat_s = reinterpret_cast<float2*>(sptr + kOffStat);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_h);
    mbar_init(b_full(), 1);
    for (int s = 0; s < kAStages; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(t_full(i), 1); mbar_init(t_empty(i), kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 2 * BN); tmem_relinquish(); }
  for (int i = threadIdx.x; i < BN; i += kThreads) {
    vec_s[i] = bias ? __ldg(bias + i) : 0.f;
    vec_s[BN + i] = __ldg(ln_g + i);
    vec_s[2 * BN + i] = __ldg(ln_b + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(b_full(), 4 * kBBlock);
      for (int kb = 0; kb < 4; ++kb) tma_load_2d(sbase + kOffB + kb * kBBlock, &tmap_b, b_full(), kb * BK, 0);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int m_tile = blockIdx.x; m_tile < num_m_tiles; m_tile += gridDim.x) {
      for (int kb = 0; kb < kK / BK; ++kb) {
        mbar_wait(a_empty(stage), phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(a_full(stage), kABytes);
          tma_load_2d(sbase + kOffA + stage * kABytes, &tmap_a, a_full(stage), kb * BK, m_tile * BM);
        }
        __syncwarp();
        if (++stage == kAStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
    mbar_wait(b_full(), 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int m_tile = blockIdx.x; m_tile < num_m_tiles; m_tile += gridDim.x) {
      mbar_wait(t_empty(acc), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < kK / BK; ++kb) {
        mbar_wait(a_full(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a_desc = make_smem_desc_sw128(sbase + kOffA + stage * kABytes);
          const uint64_t b_desc = make_smem_desc_sw128(sbase + kOffB + kb * kBBlock);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(a_empty(stage));
          if (kb == kK / BK - 1) umma_commit(t_full(acc));
        }
        __syncwarp();
        if (++stage == kAStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue: 8 warps; thread = row, 128 columns (four 32-column chunks) each =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int cgrp = ew >> 2;              // column half of the 256-wide row
    const int r = quarter * 32 + lane;     // row inside the tile
    const uint32_t stage_buf = sbase + kOffStage + ew * 4096;
    const uint32_t my_row = stage_buf + lane * 128;
    const int sw = lane & 7;
    const float* bias_s = vec_s;
    const float* g_s = vec_s + BN;
    const float* be_s = vec_s + 2 * BN;
    int acc = 0, n_done = 0;
    uint32_t acc_phase = 0;
    for (int m_tile = blockIdx.x; m_tile < num_m_tiles; m_tile += gridDim.x, ++n_done) {
      const int row0 = m_tile * BM + quarter * 32;       // first row of this warp's 32-row slab
      const int grow = row0 + lane;
      const bool row_ok = grow < m_rows;
      const float* xrow = x_in + (int64_t)grow * BN + cgrp * 128;
      // the old rows do not depend on this tile's MMAs: the first two chunks are requested before the accumulator is
      // waited for, every later chunk two chunks ahead of its use, so that the loads' latency is off the warp's chain
      float4 xq[2][8];
#pragma unroll
      for (int pre = 0; pre < 2; ++pre)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          xq[pre][j] = row_ok ? *reinterpret_cast<const float4*>(xrow + pre * 32 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
      mbar_wait(t_full(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * BN + cgrp * 128 + ((uint32_t)(quarter * 32) << 16);
      float s1 = 0.f, s2 = 0.f;
      // ---- pass 1: x_new = acc + bias + x_old -> TMEM (row buffer), x (fp32, TMA store), row statistics ----
#pragma unroll
      for (int chunk = 0; chunk < 4; ++chunk) {
        float4 xo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) xo[j] = xq[chunk & 1][j];
        if (chunk + 2 < 4) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            xq[chunk & 1][j] = row_ok ? *reinterpret_cast<const float4*>(xrow + (chunk + 2) * 32 + 4 * j)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        uint32_t v[32];
        tmem_ld32(t_addr + chunk * 32, v);
        tmem_ld_wait();
        const float* bc = bias_s + cgrp * 128 + chunk * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bv = *reinterpret_cast<const float4*>(bc + 4 * j);
          const float a0 = __uint_as_float(v[4 * j]) + bv.x + xo[j].x;
          const float a1 = __uint_as_float(v[4 * j + 1]) + bv.y + xo[j].y;
          const float a2 = __uint_as_float(v[4 * j + 2]) + bv.z + xo[j].z;
          const float a3 = __uint_as_float(v[4 * j + 3]) + bv.w + xo[j].w;
          s1 += (a0 + a1) + (a2 + a3);
          s2 = fmaf(a0, a0, fmaf(a1, a1, fmaf(a2, a2, fmaf(a3, a3, s2))));
          v[4 * j] = __float_as_uint(a0); v[4 * j + 1] = __float_as_uint(a1);
          v[4 * j + 2] = __float_as_uint(a2); v[4 * j + 3] = __float_as_uint(a3);
        }
        tmem_st32(t_addr + chunk * 32, v);
        // 32 fp32 columns = one 128-byte staging row; the previous TMA store out of this tile must have read it
        bulk_wait_read0();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) st_shared_v4(my_row + ((j ^ sw) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (row0 < m_rows && elect_one()) {
          tma_store_2d(&tmap_x, stage_buf, cgrp * 128 + chunk * 32, row0);       // rows past m_rows are clipped
          bulk_commit();
        }
        __syncwarp();
      }
      tmem_st_wait();
      // ---- the two warps of a row (column halves) exchange their partial sums ----
      float2* st = stat_s + (n_done & 1) * (2 * BM);
      st[cgrp * BM + r] = make_float2(s1, s2);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float2 other = st[(cgrp ^ 1) * BM + r];
      const float mean = (s1 + other.x) * (1.0f / BN);
      const float var = fmaxf((s2 + other.y) * (1.0f / BN) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + eps);
      const float shift = -mean * rstd;
      // ---- pass 2: h = (x_new - mean) rstd gamma + beta (bf16), 64 columns per staging row ----
#pragma unroll 1
      for (int chunk = 0; chunk < 4; ++chunk) {
        uint32_t v[32];
        tmem_ld32(t_addr + chunk * 32, v);
        tmem_ld_wait();
        if (chunk == 3) {
          // the accumulator stage has been read for the last time: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(t_empty(acc));
        }
        const int lcol = cgrp * 128 + chunk * 32;
        const int half = chunk & 1;
        if (half == 0) { bulk_wait_read0(); __syncwarp(); }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float y[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float u = fmaf(__uint_as_float(v[8 * j + e]), rstd, shift);
            y[e] = fmaf(u, g_s[lcol + 8 * j + e], be_s[lcol + 8 * j + e]);
          }
          st_shared_v4(my_row + (((half * 4 + j) ^ sw) << 4), pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]),
                       pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
        }
        if (half == 1) {
          fence_proxy_async_smem();
          __syncwarp();
          if (row0 < m_rows && elect_one()) {
            tma_store_2d(&tmap_h, stage_buf, lcol - 32, row0);
            bulk_commit();
          }
          __syncwarp();
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    bulk_wait0();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * BN); }
}

}  // namespace

extern "C" int stac_outproj_ln_bf16(const uint16_t* a, const uint16_t* w, const float* bias, float* x, const float* ln_g,
                                    const float* ln_b, float eps, uint16_t* h, int64_t m, void* stream) {
  STAC_REQUIRE(a && w && x && ln_g && ln_b && h && m > 0 && m < (1ll << 31) - 256 && eps > 0.f);
  CUtensorMap ta, tb, tx, th;
  {
    const uint64_t dims[2] = {(uint64_t)kK, (uint64_t)m};
    const uint64_t str[1] = {(uint64_t)kK * 2};
    const uint32_t box[2] = {BK, BM};
    int r = encode_map(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)kK, (uint64_t)BN};
    const uint64_t str[1] = {(uint64_t)kK * 2};
    const uint32_t box[2] = {BK, BN};
    int r = encode_map(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)BN, (uint64_t)m};
    const uint64_t str[1] = {(uint64_t)BN * 4};
    const uint32_t box[2] = {32u, 32u};
    int r = encode_map(&tx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, x, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)BN, (uint64_t)m};
    const uint64_t str[1] = {(uint64_t)BN * 2};
    const uint32_t box[2] = {64u, 32u};
    int r = encode_map(&th, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, h, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(outproj_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  const int m_tiles = (int)ceil_div64(m, BM);
  const int grid = std::min(stac_grid_limit(), m_tiles);
  outproj_ln_kernel<<<grid, kThreads, kSmemBytes, as_stream(stream)>>>(ta, tb, tx, th, bias, ln_g, ln_b, eps, x, (int)m,
                                                                      m_tiles);
  STAC_LAUNCH_CHECK();
}
