// tcgen05 / TMEM / TMA GEMM for the bf16 mode:
//   C[M,N] = epilogue( A[M,K] . W[N,K]^T )   bf16 operands, fp32 accumulation in tensor memory.
// One persistent warp-specialised kernel serves every dense projection of the encoder
// (src-linear + PE, QKV, out-proj + residual, FFN1 + GELU, FFN2 + residual, CTC logits) and, with
// a different A-tile address generator, the second convolution block as an implicit GEMM
// (A tile = 6 time steps x 20 freq bins x 64 channels fetched by one 5-D TMA box per filter tap from
// the reflect-padded parity-split output of block 0).
//
// Reference behaviour replaced: nn.Linear / nn.Conv2d calls inside SpeechBrain reached from
//   /root/reference/stac-st/modules/TransformerMultiTask.py:296,304-308 and inference.py:99,106.
//
// Roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2.. =
// epilogue.  Two accumulator stages in TMEM (2 x 256 columns) let the epilogue of tile i overlap
// the MMAs of tile i+1.
//   linear mode: 16 epilogue warps (4 per TMEM lane quarter, 64 columns each).  The encoder GEMMs
//     have K = 256..1024, so a 128x256 tile is only 2-8 k cycles of MMA and the epilogue is the
//     critical path: results go TMEM -> registers -> fused bias/GELU -> 128B-swizzled staging tile in
//     shared memory -> one TMA store per 32-row x 128-byte box (full-line writes, bounds clipped by the
//     tensor map).  An in-place residual (C += ...) is a TMA reduce-add, so the residual stream is never
//     loaded into the SM.
//   conv mode: 4 epilogue warps, direct stores (K = 2304: the MMAs hide the epilogue).
#include <algorithm>
#include <cstdlib>
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int BLOCK_M = 128, BLOCK_N = 256, BLOCK_K = 64, UMMA_K = 16;
constexpr int kConvRows = 120, kConvT = 6, kConvF = 20;  // conv A tile: 6 time steps x 20 freq bins
constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kTmemCols = 2 * BLOCK_N;

template <bool kConv>
struct Cfg {
  static constexpr int kStages = 3;
  static constexpr int kEpiWarps = kConv ? 4 : 16;
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
  // linear: one 4 KB TMA-store staging tile per epilogue warp.
  // conv:   LayerNorm affine [2][20][256] fp32 (40 KB) | bias (1 KB) | row statistics (1 KB) | 2 x 16 KB output staging
  static constexpr int kStagingBytes = kConv ? (40960 + 1024 + 1024 + 2 * 16384) : kEpiWarps * 4096;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct EpiParams {
  const float* bias;
  const float* resid;      // direct-load residual (period > 0 -> row % period), NULL if none / reduce-add
  int64_t resid_period;
  int reduce_add;          // 1: C += result through TMA reduce-add (in-place residual)
  int act;
  void* c;
  int c_bf16;
  int64_t m, n;
  // V^T side output of a packed QKV projection
  __nv_bfloat16* vt;
  int64_t vt_cols, seq_len, t_pad;
  // CTC head: pass 1 (stats != NULL) reduces every (row, 64-column group) to (max, sum exp, argmax) and stores nothing
  // to C; pass 2 (row_sub != NULL) subtracts the row's log-sum-exp
  float* stats;       // [2][n_groups][m]: max | sum of exp(x - max)
  int64_t stats_plane;   // n_groups * m
  const float* row_sub;  // [m] log-sum-exp
  const float* row_max;  // [m] row maximum (pass 2 finds the greedy id: first column equal to it)
  int* argmax;           // [m], pre-set to INT_MAX by the reduce kernel
  // conv mode
  int t2_len;         // output time steps per utterance
  int tiles_per_utt;  // ceil(T2 / 6)
  const float* ln_g;  // [20*256] LayerNorm affine of the second block
  const float* ln_b;
};

// bf16-mode GELU: x * Phi(x) with Phi(x) = 0.5 (1 + erf(x / sqrt 2)) ~ 0.5 (1 + tanh(x (a + b x^2 + c x^4))); a, b, c are a
// minimax fit to the exact-erf GELU (max |error| 2.5e-5 over [-8, 8] before the 2^-11 relative error of tanh.approx,
// well under the bf16 rounding of the stored activation).  6 FMA-pipe instructions + 1 MUFU per element: the exact
// erff() (or a long polynomial) makes the FFN1 epilogue issue-bound - slower than its MMAs and its HBM traffic.
__device__ __forceinline__ float gelu_fast(float x) {
  const float u = fminf(x * x, 64.0f);      // the fit is for |x| <= 8; beyond it tanh is saturated anyway
  float p = fmaf(-3.51516785e-04f, u, 3.70056460e-02f);
  p = fmaf(p, u, 7.97507884e-01f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * p));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <bool kConv>
__global__ void __launch_bounds__(Cfg<kConv>::kThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_c, const EpiParams ep, const int num_m_tiles,
                 const int num_n_tiles, const int num_k_blocks) {
  using C = Cfg<kConv>;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t staging_base = smem_base + C::kStages * kStageBytes;
  const uint32_t bar_base = staging_base + C::kStagingBytes;
  // barriers: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]; then the TMEM base address
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = num_m_tiles * num_n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_c);
    for (int s = 0; s < C::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), C::kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================== TMA producer =====================
    // (the whole warp runs the loop; one elected lane issues: warp-uniform operands stay in uniform registers)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_tile = tile / num_n_tiles, n_tile = tile - m_tile * num_n_tiles;
        const int conv_b = kConv ? m_tile / ep.tiles_per_utt : 0;
        const int conv_t0 = kConv ? (m_tile - conv_b * ep.tiles_per_utt) * kConvT : 0;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t a_dst = smem_base + stage * kStageBytes;
          const uint32_t b_dst = a_dst + kABytes;
          if (elect_one()) {
            if (kConv) {
              mbar_arrive_expect_tx(full_bar(stage), kConvRows * BLOCK_K * 2 + kBBytes);
              const int tap = kb >> 2, c0 = (kb & 3) * BLOCK_K;
              const int kf = tap / 3, kt = tap - 3 * kf;
              tma_load_5d(a_dst, &tmap_a, full_bar(stage), c0, kf >> 1, conv_t0 + (kt >> 1),
                          (kt & 1) * 2 + (kf & 1), conv_b);
              tma_load_2d(b_dst, &tmap_b, full_bar(stage), c0, tap * 256 + n_tile * BLOCK_N);
            } else {
              mbar_arrive_expect_tx(full_bar(stage), kStageBytes);
              tma_load_2d(a_dst, &tmap_a, full_bar(stage), kb * BLOCK_K, m_tile * BLOCK_M);
              tma_load_2d(b_dst, &tmap_b, full_bar(stage), kb * BLOCK_K, n_tile * BLOCK_N);
            }
          }
          __syncwarp();
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
    {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_base + stage * kStageBytes;
            const uint64_t a_desc = make_smem_desc_sw128(a_addr);
            const uint64_t b_desc = make_smem_desc_sw128(a_addr + kABytes);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              // +32 B per UMMA_K step inside the 128-B swizzle atom (descriptor address unit = 16 B)
              umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            }
            umma_commit(empty_bar(stage));   // frees the smem slot when these MMAs retire
            if (kb == num_k_blocks - 1) umma_commit(tfull_bar(acc));
          }
          __syncwarp();
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if constexpr (kConv) {
    // ===================== conv epilogue: 4 warps, fused bias + LayerNorm(20 x 256) + LeakyReLU ==========
    // Thread = accumulator row (t_local, f2); a time step's LayerNorm group is 20 rows x 256 columns, all inside
    // this CTA's tile.  Pass 1 reads the tile from TMEM for the row sums, pass 2 re-reads it, normalises and
    // stages bf16 64-channel slabs (128B-swizzled) that leave through TMA stores into [B, T2, 20*256].
    const int quarter = warp & 3;
    const int r_local = quarter * 32 + lane;
    const int tid_e = (warp - 2) * 32 + lane;
    const int tl = r_local / kConvF, f2 = r_local - tl * kConvF;
    unsigned char* sgen = smem_raw + (staging_base - smem_u32(smem_raw));
    float* gam_s = reinterpret_cast<float*>(sgen);
    float* bet_s = gam_s + kConvF * 256;
    float* bias_s = bet_s + kConvF * 256;
    float2* stats_s = reinterpret_cast<float2*>(bias_s + 256);
    const uint32_t out_stage = staging_base + 40960 + 2048;
    // affine rows are rotated by 4 floats per freq bin so that lanes (= consecutive bins) hit distinct banks
    for (int i = tid_e; i < kConvF * 256; i += 128) {
      const int f = i >> 8, c = i & 255;
      gam_s[f * 256 + ((c + 4 * f) & 255)] = __ldg(ep.ln_g + i);
      bet_s[f * 256 + ((c + 4 * f) & 255)] = __ldg(ep.ln_b + i);
    }
    for (int i = tid_e; i < 256; i += 128) bias_s[i] = __ldg(ep.bias + i);
    epi_bar_sync();
    const float* g_row = gam_s + f2 * 256;
    const float* b_row = bet_s + f2 * 256;
    const int rot = 4 * f2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_tile = tile / num_n_tiles;
      const int conv_b = m_tile / ep.tiles_per_utt;
      const int t0 = (m_tile - conv_b * ep.tiles_per_utt) * kConvT;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * BLOCK_N + ((uint32_t)(quarter * 32) << 16);
      // ---- pass 1: row sum and sum of squares of (acc + bias) ----
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
        uint32_t v[32];
        tmem_ld32(t_addr + ch * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 bv = *reinterpret_cast<const float4*>(bias_s + ch * 32 + i);
          const float y0 = __uint_as_float(v[i]) + bv.x, y1 = __uint_as_float(v[i + 1]) + bv.y;
          const float y2 = __uint_as_float(v[i + 2]) + bv.z, y3 = __uint_as_float(v[i + 3]) + bv.w;
          s1 += (y0 + y1) + (y2 + y3);
          s2 = fmaf(y0, y0, s2); s2 = fmaf(y1, y1, s2); s2 = fmaf(y2, y2, s2); s2 = fmaf(y3, y3, s2);
        }
      }
      stats_s[r_local] = make_float2(s1, s2);
      epi_bar_sync();
      float mean = 0.f, rstd = 0.f;
      if (tl < kConvT) {
        float a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int f = 0; f < kConvF; ++f) {
          const float2 p = stats_s[tl * kConvF + f];
          a1 += p.x; a2 += p.y;
        }
        mean = a1 * (1.0f / (kConvF * 256));
        const float var = fmaxf(a2 * (1.0f / (kConvF * 256)) - mean * mean, 0.f);
        rstd = rsqrtf(var + 1e-5f);
      }
      // ---- pass 2: normalise, affine, LeakyReLU, bf16, 64-channel slabs through TMA ----
#pragma unroll 1
      for (int slab = 0; slab < 4; ++slab) {
        const uint32_t buf = out_stage + (slab & 1) * 16384;
        if (tid_e == 0) bulk_wait_read1();     // the store issued two slabs ago has finished reading this buffer
        epi_bar_sync();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int ch = slab * 2 + half;
          uint32_t v[32];
          tmem_ld32(t_addr + ch * 32, v);
          tmem_ld_wait();
          if (ch == BLOCK_N / 32 - 1) {
            // last read of this accumulator stage: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const int c = ch * 32 + i;
            const float4 bv = *reinterpret_cast<const float4*>(bias_s + c);
            const float4 g = *reinterpret_cast<const float4*>(g_row + ((c + rot) & 255));
            const float4 be = *reinterpret_cast<const float4*>(b_row + ((c + rot) & 255));
            float y[4] = {__uint_as_float(v[i]) + bv.x, __uint_as_float(v[i + 1]) + bv.y,
                          __uint_as_float(v[i + 2]) + bv.z, __uint_as_float(v[i + 3]) + bv.w};
            const float ga[4] = {g.x, g.y, g.z, g.w}, bb[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float a = rstd * ga[q];
              float o = fmaf(y[q], a, fmaf(-mean, a, bb[q]));
              y[q] = fmaxf(o, 0.01f * o);        // LeakyReLU(0.01)
            }
            pk[i >> 1] = pack_bf16x2(y[0], y[1]);
            pk[(i >> 1) + 1] = pack_bf16x2(y[2], y[3]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            st_shared_v4(buf + r_local * 128 + (((half * 4 + j) ^ (r_local & 7)) << 4), pk[4 * j], pk[4 * j + 1],
                         pk[4 * j + 2], pk[4 * j + 3]);
        }
        fence_proxy_async_smem();
        epi_bar_sync();
        if (tid_e == 0) { tma_store_3d(&tmap_c, buf, slab * 64, t0 * kConvF, conv_b); bulk_commit(); }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (tid_e == 0) bulk_wait0();
    __syncwarp();
  } else {
    // ===================== linear epilogue: 16 warps =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;          // TMEM lane quarter this warp may read
    const int cgrp = ew >> 2;              // 64-column group of the 256-wide tile
    const uint32_t stage_buf = staging_base + ew * 4096;
    const uint32_t my_row = stage_buf + lane * 128;
    const int sw = lane & 7;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_tile = tile / num_n_tiles, n_tile = tile - m_tile * num_n_tiles;
      const int row0 = m_tile * BLOCK_M + quarter * 32;
      const int64_t row = (int64_t)row0 + lane;
      const bool row_ok = row < ep.m;
      const int colg = n_tile * BLOCK_N + cgrp * 64;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * BLOCK_N + cgrp * 64 + ((uint32_t)(quarter * 32) << 16);
      float st_m = -INFINITY, st_s = 0.f;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        float x[32];
        {
          uint32_t v[32];
          tmem_ld32(t_addr + half * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(v[i]);
        }
        if (half == 1) {
          // both halves are in registers / staged: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        const int col0 = colg + half * 32;
        if (col0 >= ep.n) continue;                    // warp-uniform
        if (ep.bias) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            // n % 8 == 0, so a 4-column group is either fully inside or fully outside the matrix
            if (col0 + i < ep.n) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + i));
              x[i] += bv.x; x[i + 1] += bv.y; x[i + 2] += bv.z; x[i + 3] += bv.w;
            }
          }
        }
        if (ep.act == STAC_ACT_GELU_ERF) {
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = gelu_fast(x[i]);
        }
        if (ep.stats != nullptr) {
          // CTC pass 1: online (max, sum exp) of this thread's row over its 64-column group
          constexpr float kL2e = 1.4426950408889634f;
          if (col0 + 32 > ep.n) {                        // warp-uniform: only the ragged last group masks columns
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + i >= ep.n) x[i] = -INFINITY;
          }
          float hm0 = fmaxf(x[0], x[1]), hm1 = fmaxf(x[2], x[3]);
#pragma unroll
          for (int i = 4; i < 32; i += 4) { hm0 = fmaxf(hm0, fmaxf(x[i], x[i + 1])); hm1 = fmaxf(hm1, fmaxf(x[i + 2], x[i + 3])); }
          const float hm = fmaxf(hm0, hm1);
          if (hm > st_m) { st_s *= ex2_approx((st_m - hm) * kL2e); st_m = hm; }
          const float ms = st_m * kL2e;
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            s0 += ex2_approx(fmaf(x[i], kL2e, -ms));
            s1 += ex2_approx(fmaf(x[i + 1], kL2e, -ms));
          }
          st_s += s0 + s1;
          if ((half == 1 || col0 + 32 >= ep.n) && row_ok) {
            const int64_t g = (int64_t)(colg >> 6) * ep.m + row;
            ep.stats[g] = st_m;
            ep.stats[ep.stats_plane + g] = st_s;
          }
          continue;
        }
        if (ep.row_sub != nullptr) {
          // CTC pass 2: log-probabilities; the greedy id is the first column holding the row maximum
          const float l = row_ok ? __ldg(ep.row_sub + row) : 0.f;
          const float rmax = row_ok ? __ldg(ep.row_max + row) : INFINITY;
          if (col0 + 32 > ep.n) {                        // ragged last group (its stores are clipped by the map)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + i >= ep.n) x[i] = -INFINITY;
          }
          float h0 = fmaxf(x[0], x[1]), h1 = fmaxf(x[2], x[3]);
#pragma unroll
          for (int i = 4; i < 32; i += 4) { h0 = fmaxf(h0, fmaxf(x[i], x[i + 1])); h1 = fmaxf(h1, fmaxf(x[i + 2], x[i + 3])); }
          if (fmaxf(h0, h1) == rmax && ep.argmax != nullptr) {      // rare: this chunk holds the row maximum
            int first = 0x7fffffff;
#pragma unroll
            for (int i = 31; i >= 0; --i)
              if (x[i] == rmax && col0 + i < ep.n) first = col0 + i;
            if (first != 0x7fffffff) atomicMin(ep.argmax + row, first);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] -= l;
        }
        if (ep.resid && row_ok) {
          const int64_t rrow = ep.resid_period > 0 ? row % ep.resid_period : row;
          const float* rp = ep.resid + rrow * ep.n + col0;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            if (col0 + i < ep.n) {
              const float4 rv = __ldg(reinterpret_cast<const float4*>(rp + i));
              x[i] += rv.x; x[i + 1] += rv.y; x[i + 2] += rv.z; x[i + 3] += rv.w;
            }
          }
        }
        if (ep.vt != nullptr && col0 >= ep.n - ep.vt_cols) {
          // V columns of a packed QKV projection -> V^T [B*H][64][t_pad] (keys contiguous)
          if (row_ok) {
            const int64_t b = row / ep.seq_len, t = row - b * ep.seq_len;
            const int vcol = (int)(col0 - (ep.n - ep.vt_cols));
            const int64_t heads = ep.vt_cols >> 6;
            __nv_bfloat16* dst = ep.vt + ((b * heads + (vcol >> 6)) * 64 + (vcol & 63)) * ep.t_pad + t;
#pragma unroll
            for (int i = 0; i < 32; ++i) dst[(int64_t)i * ep.t_pad] = __float2bfloat16_rn(x[i]);
          }
          continue;
        }
        if (ep.c_bf16) {
          // 64 bf16 columns = one 128-byte staging row; this half fills 16-byte chunks half*4 .. +3
          if (half == 0) { bulk_wait_read0(); __syncwarp(); }     // (every lane: the elected issuer is one of them)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            st_shared_v4(my_row + (((half * 4 + j) ^ sw) << 4), pack_bf16x2(x[8 * j], x[8 * j + 1]),
                         pack_bf16x2(x[8 * j + 2], x[8 * j + 3]), pack_bf16x2(x[8 * j + 4], x[8 * j + 5]),
                         pack_bf16x2(x[8 * j + 6], x[8 * j + 7]));
          }
          const bool last_half = half == 1 || colg + 32 >= ep.n;
          if (last_half) {
            fence_proxy_async_smem();
            __syncwarp();
            if (row0 < ep.m && elect_one()) { tma_store_2d(&tmap_c, stage_buf, colg, row0); bulk_commit(); }
          }
        } else {
          // 32 fp32 columns = one 128-byte staging row
          bulk_wait_read0();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            st_shared_v4(my_row + ((j ^ sw) << 4), __float_as_uint(x[4 * j]), __float_as_uint(x[4 * j + 1]),
                         __float_as_uint(x[4 * j + 2]), __float_as_uint(x[4 * j + 3]));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (row0 < ep.m && elect_one()) {
            if (ep.reduce_add) tma_reduce_add_2d(&tmap_c, stage_buf, col0, row0);
            else tma_store_2d(&tmap_c, stage_buf, col0, row0);
            bulk_commit();
          }
        }
      }
    }
    bulk_wait0();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

int num_sms() { return stac_grid_limit(); }

template <bool kConv>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tcm, const EpiParams& ep, int m_tiles,
           int n_tiles, int k_blocks, cudaStream_t st) {
  using C = Cfg<kConv>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<kConv>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  const int grid = std::min(m_tiles * n_tiles, num_sms());
  gemm_bf16_kernel<kConv><<<grid, C::kThreads, C::kSmemBytes, st>>>(ta, tb, tcm, ep, m_tiles, n_tiles, k_blocks);
  STAC_LAUNCH_CHECK();
}

}  // namespace

// gemm_wres.cu: the weight-resident kernel for K = 256, N % 256 == 0 (same tensor maps)
int stac_gemm_wres_launch(const CUtensorMap& ta, const uint16_t* w, const CUtensorMap& tcm, const float* bias,
                          int c_bf16, int reduce_add, int64_t m, int64_t n, int64_t k, cudaStream_t st);

static bool wres_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("STAC_WRES");      // STAC_WRES=0: always the general kernel (A/B measurements)
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

// tensor maps + launch of the linear-mode kernel; `ep` carries the epilogue options (c/m/n are filled in here)
static int launch_linear(const uint16_t* a, const uint16_t* w, void* c, int c_dtype, int64_t m, int64_t n, int64_t k,
                         EpiParams ep, void* stream) {
  if (k % BLOCK_K != 0 || n % 8 != 0 || m >= (1ll << 31) - 256) return STAC_ERR_UNSUPPORTED_SHAPE;
  CUtensorMap ta, tb, tcm;
  {
    const uint64_t dims[2] = {(uint64_t)k, (uint64_t)m};
    const uint64_t str[1] = {(uint64_t)k * 2};
    const uint32_t box[2] = {BLOCK_K, BLOCK_M};
    int r = encode_map(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)k, (uint64_t)n};
    const uint64_t str[1] = {(uint64_t)k * 2};
    const uint32_t box[2] = {BLOCK_K, BLOCK_N};
    int r = encode_map(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  {
    const bool bf = c_dtype == STAC_DT_BF16;
    const uint64_t dims[2] = {(uint64_t)n, (uint64_t)m};
    const uint64_t str[1] = {(uint64_t)n * (bf ? 2 : 4)};
    const uint32_t box[2] = {bf ? 64u : 32u, 32u};
    int r = encode_map(&tcm, bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, c, 2, dims,
                       str, box);
    if (r != STAC_OK) return r;
  }
  ep.c = c; ep.c_bf16 = c_dtype == STAC_DT_BF16; ep.m = m; ep.n = n;
  ep.t2_len = 0; ep.tiles_per_utt = 1;
  // plain projections of a d_model = 256 layer with enough row tiles to amortise a resident weight block
  if (k == 256 && n % BLOCK_N == 0 && n / BLOCK_N <= 8 && ep.act == STAC_ACT_NONE && !ep.vt && !ep.resid && !ep.stats &&
      !ep.row_sub && (!ep.reduce_add || !ep.c_bf16) && ceil_div64(m, BLOCK_M) * (n / BLOCK_N) >= 2 * num_sms() &&
      wres_enabled())
    return stac_gemm_wres_launch(ta, w, tcm, ep.bias, ep.c_bf16, ep.reduce_add, m, n, k, as_stream(stream));
  return launch<false>(ta, tb, tcm, ep, (int)ceil_div64(m, BLOCK_M), (int)ceil_div64(n, BLOCK_N),
                       (int)(k / BLOCK_K), as_stream(stream));
}

extern "C" int stac_gemm_bf16(const uint16_t* a, const uint16_t* w, const float* bias, const float* resid,
                              int64_t resid_period, int act, void* c, int c_dtype, int64_t m, int64_t n,
                              int64_t k, uint16_t* vt_out, int64_t vt_cols, int64_t seq_len, int64_t t_pad,
                              void* stream) {
  STAC_REQUIRE(a && w && c && m > 0 && n > 0 && k > 0 && resid_period >= 0);
  STAC_REQUIRE(act == STAC_ACT_NONE || act == STAC_ACT_GELU_ERF);
  STAC_REQUIRE(c_dtype == STAC_DT_F32 || c_dtype == STAC_DT_BF16);
  if (vt_out) {
    STAC_REQUIRE(vt_cols > 0 && vt_cols % 64 == 0 && vt_cols <= n && (n - vt_cols) % 64 == 0);
    STAC_REQUIRE(seq_len > 0 && t_pad >= seq_len && t_pad % 8 == 0 && m % seq_len == 0);
  }
  EpiParams ep{};
  ep.bias = bias; ep.act = act;
  // an in-place, row-aligned fp32 residual becomes a TMA reduce-add; anything else is loaded directly
  ep.reduce_add = (resid != nullptr && resid == c && resid_period == 0 && c_dtype == STAC_DT_F32) ? 1 : 0;
  ep.resid = ep.reduce_add ? nullptr : resid;
  ep.resid_period = resid_period;
  ep.vt = reinterpret_cast<__nv_bfloat16*>(vt_out); ep.vt_cols = vt_cols; ep.seq_len = seq_len; ep.t_pad = t_pad;
  return launch_linear(a, w, c, c_dtype, m, n, k, ep, stream);
}

namespace {
// CTC head, between the two GEMM passes: combine the per-group statistics of every row.  One CTA = 32 rows, the groups
// dealt over its 8 warps (a thread per row was 188 CTAs of long dependent load chains: 34 us for 45 MB), the 8 partial
// (max, sum) pairs of a row merged through shared memory.
__global__ void __launch_bounds__(256)
ctc_reduce_kernel(const float* __restrict__ stats, int64_t plane, int n_groups, int64_t m, float* __restrict__ lse,
                  float* __restrict__ row_max, int* __restrict__ argmax) {
  __shared__ float part_m[8][33], part_s[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * 32 + lane;
  float mx = -INFINITY, s = 0.f;
  if (row < m) {
    // two passes of independent, coalesced loads (a single online pass is one long dependent chain per row)
    const float* pm = stats + row;
#pragma unroll 4
    for (int g = w; g < n_groups; g += 8) mx = fmaxf(mx, __ldg(pm + (int64_t)g * m));
#pragma unroll 4
    for (int g = w; g < n_groups; g += 8)
      s = fmaf(__ldg(pm + plane + (int64_t)g * m), __expf(__ldg(pm + (int64_t)g * m) - mx), s);
  }
  part_m[w][lane] = mx;
  part_s[w][lane] = s;
  __syncthreads();
  if (w == 0 && row < m) {
    float big = part_m[0][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) big = fmaxf(big, part_m[k][lane]);
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float pk = part_m[k][lane];
      if (pk != -INFINITY) tot = fmaf(part_s[k][lane], __expf(pk - big), tot);     // a warp without groups holds -inf
    }
    lse[row] = big + logf(tot);
    row_max[row] = big;
    if (argmax) argmax[row] = 0x7fffffff;
  }
}
}  // namespace

extern "C" int64_t stac_ctc_head_workspace_floats(int64_t m, int64_t vocab) {
  return 2 * ceil_div64(vocab, 64) * m + 2 * m;
}

extern "C" int stac_ctc_head_bf16(const uint16_t* enc, const uint16_t* w, const float* bias, int64_t m, int64_t vocab,
                                  int64_t d_model, float* workspace, void* log_probs, int out_dtype, int32_t* argmax,
                                  void* stream) {
  STAC_REQUIRE(enc && w && workspace && log_probs && m > 0 && vocab > 0 && d_model > 0);
  STAC_REQUIRE(out_dtype == STAC_DT_F32 || out_dtype == STAC_DT_BF16);
  const int64_t n_groups = ceil_div64(vocab, 64);
  const int64_t plane = n_groups * m;
  float* lse = workspace + 2 * plane;
  float* row_max = lse + m;
  EpiParams ep{};
  ep.bias = bias;
  ep.stats = workspace; ep.stats_plane = plane;
  int r = launch_linear(enc, w, log_probs, out_dtype, m, vocab, d_model, ep, stream);   // pass 1: statistics only
  if (r != STAC_OK) return r;
  ctc_reduce_kernel<<<(unsigned)ceil_div64(m, 32), 256, 0, as_stream(stream)>>>(workspace, plane, (int)n_groups, m,
                                                                                lse, row_max, argmax);
  if (cudaPeekAtLastError() != cudaSuccess) return (int)cudaGetLastError();
  EpiParams ep2{};
  ep2.bias = bias;
  ep2.row_sub = lse; ep2.row_max = row_max; ep2.argmax = argmax;
  return launch_linear(enc, w, log_probs, out_dtype, m, vocab, d_model, ep2, stream);  // pass 2: logits - lse
}

extern "C" int stac_conv1_bf16(const uint16_t* xpad, const uint16_t* w1_packed, const float* b1,
                               const float* ln_g, const float* ln_b, int64_t batch, int64_t t1, uint16_t* out,
                               void* stream) {
  STAC_REQUIRE(xpad && w1_packed && b1 && ln_g && ln_b && out && batch > 0 && t1 >= 2 && t1 < (1 << 30));
  const int64_t t2 = (t1 - 1) / 2 + 1, tp2 = (t1 + 3) / 2;
  const int tiles_per_utt = (int)ceil_div64(t2, kConvT);
  if (batch * tiles_per_utt >= (1ll << 31)) return STAC_ERR_UNSUPPORTED_SHAPE;
  CUtensorMap ta, tb, tcm;
  {
    // [B][4 planes][Tp2][21][256] bf16, innermost first
    const uint64_t dims[5] = {256, 21, (uint64_t)tp2, 4, (uint64_t)batch};
    const uint64_t str[4] = {256 * 2, 21 * 256 * 2, (uint64_t)tp2 * 21 * 256 * 2, (uint64_t)4 * tp2 * 21 * 256 * 2};
    const uint32_t box[5] = {BLOCK_K, kConvF, kConvT, 1, 1};
    int r = encode_map(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, xpad, 5, dims, str, box);
    if (r != STAC_OK) return r;
  }
  {
    const uint64_t dims[2] = {256, 9 * 256};
    const uint64_t str[1] = {256 * 2};
    const uint32_t box[2] = {BLOCK_K, 256};
    int r = encode_map(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w1_packed, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  {
    // output [B][T2*20 rows][256 ch] bf16; rows past an utterance's end are clipped by the map
    const uint64_t dims[3] = {256, (uint64_t)(t2 * kConvF), (uint64_t)batch};
    const uint64_t str[2] = {256 * 2, (uint64_t)(t2 * kConvF) * 256 * 2};
    const uint32_t box[3] = {64, kConvRows, 1};
    int r = encode_map(&tcm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, out, 3, dims, str, box);
    if (r != STAC_OK) return r;
  }
  EpiParams ep{};
  ep.bias = b1; ep.c = out; ep.c_bf16 = 1; ep.m = batch * t2 * kConvF; ep.n = 256;
  ep.t2_len = (int)t2; ep.tiles_per_utt = tiles_per_utt; ep.ln_g = ln_g; ep.ln_b = ln_b;
  return launch<true>(ta, tb, tcm, ep, (int)(batch * tiles_per_utt), 1, 36, as_stream(stream));
}
