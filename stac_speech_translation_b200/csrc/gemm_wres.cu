// Weight-resident tcgen05 GEMM for the d_model = 256 projections of the encoder layer (bf16 mode):
//   C[M, N] (+)= A[M, 256] . W[N, 256]^T + bias        N % 256 == 0   (QKV: N = 768, bf16 out; out-proj: N = 256, fp32 +=)
// Reference behaviour replaced: the in_proj / out_proj nn.Linear calls of torch.nn.MultiheadAttention inside
// SpeechBrain's TransformerEncoderLayer, reached from /root/reference/stac-st/modules/TransformerMultiTask.py:304-308.
//
// Why a second GEMM kernel: with K = 256 a 128 x 256 output tile is only 2 k cycles of MMA, and the general kernel
// (gemm_tc.cu) re-streams the 128 KB weight tile from L2 for every one of them - 216 MB of L2 -> smem traffic per QKV
// launch on top of the 99 MB that really have to move, which is what bounds it (30 us against a 15 us HBM floor).
// Here every CTA owns ONE 256-column block of W for its whole life: the 128 KB are loaded once and stay in shared
// memory, and only the activations stream (four 16 KB k-blocks per tile through a 4-stage ring).  CTA c serves column
// block c % n_blocks and every (grid / n_blocks)-th row tile.
// Roles: warp 0 TMA producer, warp 1 MMA issuer (two 256-column accumulators in TMEM, so the epilogue of tile i
// overlaps the MMAs of tile i+1), warps 2-9 epilogue (TMEM lane quarter = warp & 3, 128 columns each, four 32-column
// chunks -> bias -> swizzled staging tile -> TMA store / reduce-add, so the residual stream is never loaded).
#include <algorithm>
#include <cstdlib>
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int kK = 256, BM = 128, BK = 64;
constexpr int kABytes = BM * BK * 2;              // 16 KB
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kSmemLimit = 232448;                // 227 KB per CTA

// BN = columns of W resident per CTA.  256 (default): each row tile of the activations is read N/256 times from L2 and
// four ring stages fit beside the 128 KB block; 128: 64 KB block, eight ring stages.  Measured on the QKV projection
// (48064 x 768): ring of 2 / 3 / 4 stages at BN = 256 -> 37 / 28 / 25 us; BN = 128 with 8 stages -> 25 us as well.
// (25 us is neither the tensor floor, 13.6 us, nor the HBM floor, 15 us: this pool writes 6.3 TB/s, tools/ubench_write.cu.)
template <int BN>
struct Cfg {
  static constexpr int kBBlock = BN * BK * 2;               // one k-block of the resident weight tile
  static constexpr int kColsPerWarp = BN / (kEpiWarps / 4);
  static constexpr int kChunks = kColsPerWarp / 32;
  static constexpr int kAStages = BN == 256 ? 4 : 8;
  static constexpr int kOffB = 0;
  static constexpr int kOffA = 4 * kBBlock;
  static constexpr int kOffStage = kOffA + kAStages * kABytes;       // 8 x 4 KB output staging
  static constexpr int kOffBias = kOffStage + kEpiWarps * 4096;
  static constexpr int kOffBar = kOffBias + BN * 4;
  static constexpr int kNumBars = 1 + 2 * kAStages + 4;
  static constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16 + 1024;
  static_assert(kSmemBytes <= kSmemLimit, "shared memory budget");
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_wres_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_c, const float* __restrict__ bias, const int c_bf16,
                 const int reduce_add, const int m_rows, const int num_m_tiles, const int num_n_blocks) {
  using C = Cfg<BN>;
  constexpr int kAStages = C::kAStages, kBBlock = C::kBBlock, kOffA = C::kOffA, kOffB = C::kOffB;
  constexpr int kOffStage = C::kOffStage, kOffBias = C::kOffBias, kOffBar = C::kOffBar, kNumBars = C::kNumBars;
  constexpr int kColsPerWarp = C::kColsPerWarp, kChunks = C::kChunks;
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
  const int n_blk = blockIdx.x % num_n_blocks;              // the column block this CTA keeps resident
  const int m_first = blockIdx.x / num_n_blocks;
  const int m_step = gridDim.x / num_n_blocks;               // host: gridDim.x % num_n_blocks == 0
  float* bias_s = reinterpret_cast<float*>(sptr + kOffBias);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_c);
    mbar_init(b_full(), 1);
    for (int s = 0; s < kAStages; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(t_full(i), 1); mbar_init(t_empty(i), kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 2 * BN); tmem_relinquish(); }
  for (int i = threadIdx.x; i < BN; i += kThreads) bias_s[i] = bias ? __ldg(bias + n_blk * BN + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(b_full(), 4 * kBBlock);
      for (int kb = 0; kb < 4; ++kb) tma_load_2d(sbase + kOffB + kb * kBBlock, &tmap_b, b_full(), kb * BK, n_blk * BN);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int m_tile = m_first; m_tile < num_m_tiles; m_tile += m_step) {
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
    for (int m_tile = m_first; m_tile < num_m_tiles; m_tile += m_step) {
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
    // ===================== epilogue: 8 warps, 128 columns each =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;          // TMEM lane quarter this warp may read
    const int cgrp = ew >> 2;              // column group of the 256-wide tile
    const uint32_t stage_buf = sbase + kOffStage + ew * 4096;
    const uint32_t my_row = stage_buf + lane * 128;
    const int sw = lane & 7;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int m_tile = m_first; m_tile < num_m_tiles; m_tile += m_step) {
      const int row0 = m_tile * BM + quarter * 32;
      mbar_wait(t_full(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * BN + cgrp * kColsPerWarp + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int chunk = 0; chunk < kChunks; ++chunk) {
        uint32_t v[32];
        tmem_ld32(t_addr + chunk * 32, v);
        tmem_ld_wait();
        if (chunk == kChunks - 1) {
          // the whole accumulator has been read: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(t_empty(acc));
        }
#ifdef WRES_NO_EPI        // timing experiment: accumulators read, nothing staged or written
        if (v[0] != 0x7fc12345u) continue;
#endif
        const int lcol = cgrp * kColsPerWarp + chunk * 32;                 // column inside the 256-wide block
        const float* bc = bias_s + lcol;
        if (c_bf16) {
          // 64 bf16 columns = one 128-byte staging row; an even chunk fills 16-byte pieces 0..3, an odd one 4..7
          const int half = chunk & 1;
          if (half == 0) { bulk_wait_read0(); __syncwarp(); }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b0 = *reinterpret_cast<const float4*>(bc + 8 * j);
            const float4 b1 = *reinterpret_cast<const float4*>(bc + 8 * j + 4);
            st_shared_v4(my_row + (((half * 4 + j) ^ sw) << 4),
                         pack_bf16x2(__uint_as_float(v[8 * j]) + b0.x, __uint_as_float(v[8 * j + 1]) + b0.y),
                         pack_bf16x2(__uint_as_float(v[8 * j + 2]) + b0.z, __uint_as_float(v[8 * j + 3]) + b0.w),
                         pack_bf16x2(__uint_as_float(v[8 * j + 4]) + b1.x, __uint_as_float(v[8 * j + 5]) + b1.y),
                         pack_bf16x2(__uint_as_float(v[8 * j + 6]) + b1.z, __uint_as_float(v[8 * j + 7]) + b1.w));
          }
          if (half == 1) {
            fence_proxy_async_smem();
            __syncwarp();
#ifndef WRES_NO_STORE     // timing experiment: tiles staged but not written
            if (row0 < m_rows && elect_one()) {
              tma_store_2d(&tmap_c, stage_buf, n_blk * BN + lcol - 32, row0);
              bulk_commit();
            }
#endif
            __syncwarp();
          }
        } else {
          // 32 fp32 columns = one 128-byte staging row
          bulk_wait_read0();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bv = *reinterpret_cast<const float4*>(bc + 4 * j);
            st_shared_v4(my_row + ((j ^ sw) << 4), __float_as_uint(__uint_as_float(v[4 * j]) + bv.x),
                         __float_as_uint(__uint_as_float(v[4 * j + 1]) + bv.y),
                         __float_as_uint(__uint_as_float(v[4 * j + 2]) + bv.z),
                         __float_as_uint(__uint_as_float(v[4 * j + 3]) + bv.w));
          }
          fence_proxy_async_smem();
          __syncwarp();
#ifndef WRES_NO_STORE
          if (row0 < m_rows && elect_one()) {
            if (reduce_add) tma_reduce_add_2d(&tmap_c, stage_buf, n_blk * BN + lcol, row0);
            else tma_store_2d(&tmap_c, stage_buf, n_blk * BN + lcol, row0);
            bulk_commit();
          }
#endif
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

template <int BN>
int launch_wres(const CUtensorMap& ta, const uint16_t* w, const CUtensorMap& tcm, const float* bias, int c_bf16,
                int reduce_add, int64_t m, int64_t n, cudaStream_t st) {
  using C = Cfg<BN>;
  CUtensorMap tb;
  {
    const uint64_t dims[2] = {(uint64_t)kK, (uint64_t)n};
    const uint64_t str[1] = {(uint64_t)kK * 2};
    const uint32_t box[2] = {BK, BN};
    int r = encode_map(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  const int n_blocks = (int)(n / BN);
  const int m_tiles = (int)ceil_div64(m, BM);
  int grid = std::min(stac_grid_limit(), m_tiles * n_blocks);
  grid -= grid % n_blocks;
  if (grid < n_blocks) return STAC_ERR_UNSUPPORTED_SHAPE;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_wres_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  gemm_wres_kernel<BN><<<grid, kThreads, C::kSmemBytes, st>>>(ta, tb, tcm, bias, c_bf16, reduce_add, (int)m, m_tiles,
                                                             n_blocks);
  STAC_LAUNCH_CHECK();
}

}  // namespace

// Called by launch_linear (gemm_tc.cu) with the tensor maps it has built for A (box {64, 128}) and C (box {64 bf16 |
// 32 fp32, 32}).  Returns STAC_ERR_UNSUPPORTED_SHAPE when the shape is not this kernel's.
int stac_gemm_wres_launch(const CUtensorMap& ta, const uint16_t* w, const CUtensorMap& tcm, const float* bias,
                          int c_bf16, int reduce_add, int64_t m, int64_t n, int64_t k, cudaStream_t st) {
  if (k != kK || n % 256 != 0 || n / 256 > 8) return STAC_ERR_UNSUPPORTED_SHAPE;
  static int bn = 0;
  if (bn == 0) {
    const char* e = getenv("STAC_WRES_BN");          // timing experiments: force the resident block width
    bn = (e && atoi(e) == 256) ? 256 : (e && atoi(e) == 128) ? 128 : -1;
  }
  const int use = bn > 0 ? bn : 256;
  if (use == 256) return launch_wres<256>(ta, w, tcm, bias, c_bf16, reduce_add, m, n, st);
  return launch_wres<128>(ta, w, tcm, bias, c_bf16, reduce_add, m, n, st);
}
