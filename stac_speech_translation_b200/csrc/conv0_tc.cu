// a4, block 0 of the ConvolutionFrontEnd in bf16 mode, on tcgen05 tensor cores:
//   reflect-pad 1, Conv2d(1 -> 256, 3x3, stride 2) + bias, LayerNorm over (40, 256) eps 1e-5, LeakyReLU(0.01),
//   bf16 output in the reflect-padded parity-split layout the tensor-core conv1 reads (include/stac_b200.h).
// Reference behaviour: SpeechBrain ConvolutionFrontEnd block 0 as configured at
//   /root/reference/stac-st/hparams/transformer_multitask.yaml:173-180 (call: stac-st/inference.py:99).
//
// The stage's floor is its 2 GB output write (0.31 ms at the 6.3 TB/s bulk stores reach on this pool; the kernel takes
// 0.40 ms), so the job of the kernel is to keep the CUDA cores out of the way:
//   * the 9-tap convolution is one tiny GEMM per tile, D[channel][position] = W[channel][k] . X[position][k], with
//     near-fp32 accuracy from a bf16 hi/lo split packed along K (x_hi.w_hi + x_lo.w_hi + x_hi.w_lo, the bias and
//     the LayerNorm shift fill all 32 K slots: two tcgen05.mma K=16 steps per 128-channel half);
//   * the LayerNorm statistics never touch the 10240 outputs of a time step: sum and sum of squares over
//     (40 freq x 256 channels) are a linear and a quadratic form of the 40 x 9 input patches (G = W^T W is
//     built once per CTA), evaluated in fp32 by the producer warps, which then scale the patch rows by rstd and
//     append -mean*rstd as extra K slots: the accumulator already holds the NORMALISED value;
//   * with channels on the TMEM lanes each epilogue thread owns ONE channel: its 40 LayerNorm gammas and betas
//     live in registers for the whole kernel and the epilogue is a single pass
//     TMEM -> fma(g, u, beta) -> bf16x2 LeakyReLU -> staging smem;
//   * a time step's two output planes leave the SM as bulk async copies (full 512-byte rows) issued by a
//     dedicated warp; staging buffers, B tiles and accumulators are all multi-buffered and every hand-off is an
//     mbarrier, so no role ever waits at a CTA-wide barrier in the steady state.
// Roles (16 warps): 0-7 epilogue (warp & 3 = TMEM lane quarter, warp >> 2 = channel half), 8-10 and 11-13 two producer
// groups (patch rows + statistics) that take alternate tiles, 14 MMA issuer + TMEM owner, 15 output bulk-copy issuer.
// (Two producer groups since round 2: the stage was believed to sit on a 3.9 TB/s write ceiling; tools/ubench_write.cu
// measures 6.3 TB/s for bulk shared -> global copies and 6.9 TB/s for plain stores on this pool, i.e. the kernel was
// at 55 % of what the memory takes and its ONE producer group - a chain of three barriers, 110 FMAs and a 40-term sum
// per tile - was the critical path.)
#include <algorithm>
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int kMel = 80, kF1 = 40, kC = 256;
constexpr float kLnEps = 1e-5f, kSlope = 0.01f;
constexpr int kTs = 2;                       // time steps per tile
constexpr int kN = kTs * kF1;                // 80 positions = UMMA N
constexpr int kStages = 3;                   // B tiles / accumulators in flight
constexpr int kOutStages = 4;                // staged time steps in flight
constexpr int kThreads = 16 * 32;
constexpr int kProducerThreads = 96;
constexpr int kProducerGroups = 2;
constexpr int kWarpMma = 8 + 3 * kProducerGroups, kWarpOut = kWarpMma + 1;

constexpr int kWBytes = 2 * 128 * 128;                 // two 128-channel halves, 128-byte (64 x bf16) rows
constexpr int kBBytes = kN * 128;                      // 10240
constexpr int kPlane0Rows = 21, kPlane1Rows = 20;
constexpr int kOutBytes = (kPlane0Rows + kPlane1Rows) * kC * 2;   // 20992 per time step
constexpr int kFRow = kMel + 2;

constexpr int kOffW = 0;
constexpr int kOffB = kOffW + kWBytes;
constexpr int kOffOut = kOffB + kStages * kBBytes;
constexpr int kOffWraw = kOffOut + kOutStages * kOutBytes;       // fp32 w0 [256][9] + b0 [256] (setup only)
constexpr int kOffFeat = kOffWraw + (kC * 9 + kC) * 4;            // float [groups][5][82]
constexpr int kOffRowStat = kOffFeat + kProducerGroups * 5 * kFRow * 4;   // float2 [groups][80]
constexpr int kOffTab = kOffRowStat + kProducerGroups * kN * 8;                     // float [128]: wsum[9] G[81] bw[9] bsum bb
constexpr int kOffNorm = kOffTab + 128 * 4;                       // float [80] mean | float [80] 1 / std | float floor_all
constexpr int kOffBar = kOffNorm + (2 * kMel + 4) * 4;
constexpr int kNumBars = 4 * kStages + 2 * kOutStages;
constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16 + 1024;

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) return -i;
  if (i >= n) return 2 * (n - 1) - i;
  return i;
}
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(reinterpret_cast<uint64_t>(gdst)), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t addr, unsigned short v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ unsigned short bf16_bits(float x) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(x));
}
__device__ __forceinline__ float bf16_val(unsigned short b) { return __uint_as_float((uint32_t)b << 16); }

// staging row (512 bytes each) of conv-0 frequency bin f1: plane 0 holds even padded bins fp = f1 + 1
// (row fp / 2; row 0 mirrors f1 = 1), plane 1 the odd ones.
__host__ __device__ constexpr int stage_row(int f1) { return (f1 & 1) ? (f1 + 1) / 2 : kPlane0Rows + f1 / 2; }

__global__ void __launch_bounds__(kThreads, 1)
conv0_tc_kernel(const float* __restrict__ feats, const float* __restrict__ w0, const float* __restrict__ b0,
                const float* __restrict__ ln_g, const float* __restrict__ ln_b, int frames, int t1_len,
                int tiles_per_utt, int n_tiles, __nv_bfloat16* __restrict__ out,
                // fused a2 tail + a3 (nullptr utt_max: feats are already clamped and normalised)
                const unsigned int* __restrict__ utt_max, int per_utt, float top_db, const float* __restrict__ nmean,
                const float* __restrict__ nstd, int batch) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bars = sbase + kOffBar;
  auto b_full = [&](int s) { return bars + 8u * s; };
  auto b_empty = [&](int s) { return bars + 8u * (kStages + s); };
  auto acc_full = [&](int s) { return bars + 8u * (2 * kStages + s); };
  auto acc_empty = [&](int s) { return bars + 8u * (3 * kStages + s); };
  auto out_full = [&](int s) { return bars + 8u * (4 * kStages + s); };
  auto out_empty = [&](int s) { return bars + 8u * (4 * kStages + kOutStages + s); };
  const uint32_t tmem_slot = bars + 8u * kNumBars;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* wraw = reinterpret_cast<float*>(sptr + kOffWraw);
  float* tab = reinterpret_cast<float*>(sptr + kOffTab);

  // ---------------- one-time setup: barriers, TMEM, weight tile, statistics tables ----------------
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(b_full(s), 3); mbar_init(b_empty(s), 1); mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 8);
    }
    for (int s = 0; s < kOutStages; ++s) { mbar_init(out_full(s), 8); mbar_init(out_empty(s), 1); }
    fence_barrier_init();
  }
  if (warp == kWarpMma) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  for (int i = tid; i < kC * 9; i += kThreads) wraw[i] = __ldg(w0 + i);
  for (int i = tid; i < kC; i += kThreads) wraw[kC * 9 + i] = __ldg(b0 + i);
  float* normtab = reinterpret_cast<float*>(sptr + kOffNorm);
  const bool fused_norm = utt_max != nullptr;
  if (fused_norm) {
    for (int i = tid; i < kMel; i += kThreads) {
      normtab[i] = nmean != nullptr ? __ldg(nmean + i) : 0.f;
      normtab[kMel + i] = nstd != nullptr ? 1.0f / __ldg(nstd + i) : 1.f;      // reciprocal: the loader multiplies
    }
    if (warp == kWarpOut && !per_utt) {         // batch-global top-dB (older SpeechBrain): one maximum for everybody
      float mx = -INFINITY;
      for (int i = lane; i < batch; i += 32) mx = fmaxf(mx, ordered_to_float(__ldg(utt_max + i)));
      mx = warp_max(mx);
      if (lane == 0) normtab[2 * kMel] = mx - top_db;
    }
  }
  __syncthreads();
  if (tid < kC) {
    // weight row of channel c: [w_hi(9) | w_hi(9) | w_lo(9) | b_hi | b_lo | b_hi | 1 | 1 | 0 ...] bf16, 128-byte
    // swizzled row; the patch rows carry [xs_hi | xs_lo | xs_hi | r_hi | r_hi | r_lo | s_hi | s_lo] with
    // xs = rstd * x, r = rstd, s = -mean * rstd, so the accumulator is (conv + bias - mean) * rstd
    // row r of half h holds channel 2 r + h: an epilogue thread (= TMEM lane r) then owns the ADJACENT channels 2 r, 2 r + 1
    // (one in each half's accumulator) and stores them as one 32-bit word
    const int c = tid, r = c >> 1;
    unsigned short kv[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) kv[i] = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float w = wraw[c * 9 + k];
      const unsigned short hi = bf16_bits(w);
      kv[k] = hi; kv[9 + k] = hi; kv[18 + k] = bf16_bits(w - bf16_val(hi));
    }
    const float bb = wraw[kC * 9 + c];
    kv[27] = bf16_bits(bb);
    kv[28] = bf16_bits(bb - bf16_val(kv[27]));
    kv[29] = kv[27];
    kv[30] = 0x3f80;
    kv[31] = 0x3f80;
    const uint32_t row = sbase + kOffW + (c & 1) * 16384 + r * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sts_v4(row + ((j ^ (r & 7)) << 4), kv[8 * j] | ((uint32_t)kv[8 * j + 1] << 16),
             kv[8 * j + 2] | ((uint32_t)kv[8 * j + 3] << 16), kv[8 * j + 4] | ((uint32_t)kv[8 * j + 5] << 16),
             kv[8 * j + 6] | ((uint32_t)kv[8 * j + 7] << 16));
  } else if (tid < kC + 101) {
    // tab: [0,9) wsum_k = sum_c w_ck ; [9,90) G_kk' = sum_c w_ck w_ck' ; [90,99) bw_k = sum_c b_c w_ck ;
    //      [99] bsum = sum_c b_c ; [100] bb = sum_c b_c^2
    const int i = tid - kC;
    float acc = 0.f;
    for (int c = 0; c < kC; ++c) {
      const float bc = wraw[kC * 9 + c];
      float v;
      if (i < 9) v = wraw[c * 9 + i];
      else if (i < 90) v = wraw[c * 9 + (i - 9) / 9] * wraw[c * 9 + (i - 9) % 9];     // (k, k2 > k: stored doubled below)
      else if (i < 99) v = bc * wraw[c * 9 + (i - 90)];
      else if (i == 99) v = bc;
      else v = bc * bc;
      acc += v;
    }
    // G is symmetric: the producers evaluate x^T G x over the upper triangle only, with the off-diagonal terms doubled
    tab[i] = (i >= 9 && i < 90 && (i - 9) % 9 > (i - 9) / 9) ? 2.0f * acc : acc;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int tp2 = (t1_len + 3) >> 1;

  if (warp < 8) {
    // ============================ epilogue: thread = two adjacent output channels, half of the frequency bins ============
    // (round 2: with one channel per thread the stage was bound by its own issue slots and the shared-memory pipe - 41
    //  two-byte stores and their byte extractions per time step and thread; a thread now takes channels 2 l and 2 l + 1
    //  from the two halves' accumulators on ITS TMEM lane, packs them into one word and stores 128 contiguous bytes per
    //  warp; the 2 x 2 x 20 LayerNorm gammas / betas of its bins still live in registers)
    const int fh = warp >> 2, quarter = warp & 3;          // fh: frequency bins [20 fh, 20 fh + 20)
    const int l = quarter * 32 + lane;                     // TMEM lane = channel pair
    constexpr int kFh = kF1 / 2;
    float g0[kFh], g1[kFh], be0[kFh], be1[kFh];
#pragma unroll
    for (int f = 0; f < kFh; ++f) {
      const float2 gg = __ldg(reinterpret_cast<const float2*>(ln_g + (fh * kFh + f) * kC + 2 * l));
      const float2 bb = __ldg(reinterpret_cast<const float2*>(ln_b + (fh * kFh + f) * kC + 2 * l));
      g0[f] = gg.x; g1[f] = gg.y; be0[f] = bb.x; be1[f] = bb.y;
    }
    const uint32_t my_col = sbase + kOffOut + l * 4;
    const __nv_bfloat162 slope2 = __floats2bfloat162_rn(kSlope, kSlope);
    int n = 0, as = 0;
    uint32_t acc_phase = 0;
    int t1_0 = (blockIdx.x % tiles_per_utt) * kTs;
    const int t1_step = (gridDim.x % tiles_per_utt) * kTs, t1_wrap = tiles_per_utt * kTs;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
      mbar_wait(acc_full(as), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * (2 * kN) + fh * kFh;
#pragma unroll
      for (int tl = 0; tl < kTs; ++tl) {
        const int m = n * kTs + tl, os = m & (kOutStages - 1);
        const bool valid = t1_0 + tl < t1_len;
        mbar_wait(out_empty(os), ((m >> 2) & 1) ^ 1);
        if (valid) {
          uint32_t ua[16], ub[4], va[16], vb[4];             // channel 2 l (half 0) | channel 2 l + 1 (half 1), 20 bins each
          tmem_ld16(t_addr + tl * kF1, ua);
          tmem_ld4(t_addr + tl * kF1 + 16, ub);
          tmem_ld16(t_addr + kN + tl * kF1, va);
          tmem_ld4(t_addr + kN + tl * kF1 + 16, vb);
          tmem_ld_wait();
          if (tl == kTs - 1) {
            // last read of this accumulator stage: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(as));
          }
#ifndef C0_NO_EPI          // timing experiment: accumulators read, nothing computed or staged
          const uint32_t col = my_col + os * kOutBytes;
#pragma unroll
          for (int f = 0; f < kFh; ++f) {
            const float u0 = __uint_as_float(f < 16 ? ua[f] : ub[f - 16]);
            const float u1 = __uint_as_float(f < 16 ? va[f] : vb[f - 16]);
            __nv_bfloat162 y = __floats2bfloat162_rn(fmaf(g0[f], u0, be0[f]), fmaf(g1[f], u1, be1[f]));
            y = __hmax2(y, __hmul2(y, slope2));            // LeakyReLU on the packed pair
            const uint32_t bits = *reinterpret_cast<uint32_t*>(&y);
            const int f1 = fh * kFh + f;
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(col + stage_row(f1) * (kC * 2)), "r"(bits) : "memory");
            if (f1 == 1) asm volatile("st.shared.b32 [%0], %1;" ::"r"(col), "r"(bits) : "memory");   // padded bin 0 mirrors f1 = 1
          }
#else
          if (ua[0] == 0x12345678u && vb[0] == 1u && ub[0] == 2u && va[0] == 3u) sts_u16(my_col, 1);
#endif
        } else if (tl == kTs - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(as));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(out_full(os));
      }
      if (++as == kStages) { as = 0; acc_phase ^= 1; }
      t1_0 += t1_step;
      if (t1_0 >= t1_wrap) t1_0 -= t1_wrap;
    }
  } else if (warp < kWarpMma) {
    // ============================ producers: LayerNorm statistics + scaled patch rows ============================
    // group grp builds the tiles with local ordinal n = grp, grp + 2, ... (B stage n % kStages)
    const int grp = (warp - 8) / 3;
    const int p = tid - (8 + 3 * grp) * 32;          // 0..95; rows 0..79 are real positions
    float* fbuf = reinterpret_cast<float*>(sptr + kOffFeat) + grp * 5 * kFRow;
    float2* rowstat = reinterpret_cast<float2*>(sptr + kOffRowStat) + grp * kN;
    const int bar_id = 2 + grp;
    // position of this thread: warp 0 of the group = time step 0, bins 0-31; warp 1 = time step 1, bins 0-31; warp 2 =
    // bins 32-39 of both steps (lanes 0-7 / 8-15): a time step's 40 row statistics are reduced with shuffles inside
    // a warp plus one exchange through shared memory (every thread used to add all 40 in a serial chain)
    const int pw = p >> 5;
    const int tl = pw < 2 ? pw : (lane >> 3) & 1, f1 = pw < 2 ? lane : 32 + (lane & 7);
    const bool has_pos = pw < 2 || lane < 16;
    const int prow = tl * kF1 + f1;                  // row of the B tile
    // input rows t = 2*t1_0 - 1 .. 2*t1_0 + 3 (reflected at the batch edges), bins -1..79 (bin -1 mirrors bin 1);
    // the loads of tile n+1 are issued while tile n is built.  Staging slot i of this thread: row ld_r[i], bin ld_c[i]
    float ld[5];
    int ld_r[5], ld_c[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int idx = p + i * kProducerThreads;         // 0 .. 5*81-1
      ld_r[i] = idx < 5 * (kMel + 1) ? idx / (kMel + 1) : -1;
      const int fi = idx - max(ld_r[i], 0) * (kMel + 1);
      ld_c[i] = fi == 0 ? 1 : fi - 1;
    }
    float ld_floor = 0.f;                            // (fused mode) top-dB floor of the utterance the loads belong to
    auto fetch = [&](int tile) {
      const int b = tile / tiles_per_utt;
      const int t1_0 = (tile - b * tiles_per_utt) * kTs;
      const float* frow = feats + (int64_t)b * frames * kMel;
      // fused mode: the features are the raw dB values of the Fbank kernel; the top-dB clamp against the utterance
      // maximum and (x - mean) / std are applied when the values are CONSUMED (one tile later: these loads are the
      // prefetch of the next tile and must not be waited for here), with the expressions of topdb_norm_kernel
      // (fbank.cu), so the normalised [B, T, 80] tensor never exists in HBM
      if (fused_norm) ld_floor = per_utt ? ordered_to_float(__ldg(utt_max + b)) - top_db : normtab[2 * kMel];
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        ld[i] = 0.f;
        if (ld_r[i] >= 0) {
          int t = reflect_idx(2 * t1_0 - 1 + ld_r[i], frames);
          t = min(max(t, 0), frames - 1);               // rows of an invalid second time step are never used
          ld[i] = __ldg(frow + (int64_t)t * kMel + ld_c[i]);
        }
      }
    };
    const int tile0 = (int)blockIdx.x + grp * (int)gridDim.x, tile_step = kProducerGroups * (int)gridDim.x;
    if (tile0 < n_tiles) fetch(tile0);
    int n = grp;
    for (int tile = tile0; tile < n_tiles; tile += tile_step, n += kProducerGroups) {
      const int s = n % kStages;
      const uint32_t phase = (uint32_t)(n / kStages) & 1;
      named_bar_sync(bar_id, kProducerThreads);         // previous tile's patches have been read from fbuf
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        if (ld_r[i] >= 0) {
          float v = ld[i];
          if (fused_norm) {
            // (a true division here costs the producer warps 0.1 ms per batch; the product with the reciprocal differs
            // from topdb_norm_kernel's quotient by at most one fp32 ulp)
            v = (fmaxf(v, ld_floor) - normtab[ld_c[i]]) * normtab[kMel + ld_c[i]];
          }
          const int idx = p + i * kProducerThreads;
          fbuf[idx + ld_r[i]] = v;                      // = fbuf[r * kFRow + fi]: kFRow = 81 + 1
        }
      }
      named_bar_sync(bar_id, kProducerThreads);
      if (tile + tile_step < n_tiles) fetch(tile + tile_step);
      float x[9];
      float s1 = 0.f, s2 = 0.f;
      if (has_pos) {
#pragma unroll
        for (int kf = 0; kf < 3; ++kf)
#pragma unroll
          for (int kt = 0; kt < 3; ++kt) x[kf * 3 + kt] = fbuf[(2 * tl + kt) * kFRow + 2 * f1 + kf];
        // row sum and sum of squares over the 256 channels as forms of the 9-tap patch (G: upper triangle, doubled)
        s1 = tab[99];
        s2 = tab[100];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          float q = 2.0f * tab[90 + k];
#pragma unroll
          for (int k2 = k; k2 < 9; ++k2) q = fmaf(tab[9 + k * 9 + k2], x[k2], q);
          s1 = fmaf(tab[k], x[k], s1);
          s2 = fmaf(q, x[k], s2);
        }
      }
      // sums over the 40 bins of a time step: inside the warp (warps 0 / 1: 32 bins; warp 2: two segments of 8 lanes) ...
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t1 = __shfl_xor_sync(0xffffffffu, s1, o), t2 = __shfl_xor_sync(0xffffffffu, s2, o);
        if (pw < 2 || o < 8) { s1 += t1; s2 += t2; }
      }
      // ... and the two partial sums of a step through shared memory: slots [tl][0] (warp tl) and [tl][1] (warp 2)
      if (pw < 2) { if (lane == 0) rowstat[2 * pw] = make_float2(s1, s2); }
      else if (lane == 0 || lane == 8) rowstat[2 * tl + 1] = make_float2(s1, s2);
      named_bar_sync(bar_id, kProducerThreads);
      mbar_wait(b_empty(s), phase ^ 1);
      if (has_pos) {
        const float2 pa = rowstat[2 * tl], pb = rowstat[2 * tl + 1];
        const float a1 = pa.x + pb.x, a2 = pa.y + pb.y;
        const float mean = a1 * (1.0f / (kF1 * kC));
        const float var = fmaxf(a2 * (1.0f / (kF1 * kC)) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + kLnEps);
        const float shift = -mean * rstd;
        unsigned short hi[9], lo[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const float xs = x[k] * rstd;
          hi[k] = bf16_bits(xs);
          lo[k] = bf16_bits(xs - bf16_val(hi[k]));
        }
        const unsigned short r_hi = bf16_bits(rstd), r_lo = bf16_bits(rstd - bf16_val(r_hi));
        const unsigned short s_hi = bf16_bits(shift), s_lo = bf16_bits(shift - bf16_val(s_hi));
        auto pk = [](unsigned short a, unsigned short bb) { return (uint32_t)a | ((uint32_t)bb << 16); };
        const uint32_t row = sbase + kOffB + s * kBBytes + prow * 128;
        const int sw = prow & 7;
        sts_v4(row + ((0 ^ sw) << 4), pk(hi[0], hi[1]), pk(hi[2], hi[3]), pk(hi[4], hi[5]), pk(hi[6], hi[7]));
        sts_v4(row + ((1 ^ sw) << 4), pk(hi[8], lo[0]), pk(lo[1], lo[2]), pk(lo[3], lo[4]), pk(lo[5], lo[6]));
        sts_v4(row + ((2 ^ sw) << 4), pk(lo[7], lo[8]), pk(hi[0], hi[1]), pk(hi[2], hi[3]), pk(hi[4], hi[5]));
        sts_v4(row + ((3 ^ sw) << 4), pk(hi[6], hi[7]), pk(hi[8], r_hi), pk(r_hi, r_lo), pk(s_hi, s_lo));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_full(s));
    }
  } else if (warp == kWarpMma) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, kN);
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        mbar_wait(acc_empty(s), ph ^ 1);
        mbar_wait(b_full(s), ph);
        tc_fence_after();
        const uint64_t bd = make_smem_desc_sw128(sbase + kOffB + s * kBBytes);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint64_t ad = make_smem_desc_sw128(sbase + kOffW + half * 16384);
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_bf16(tmem_base + s * (2 * kN) + half * kN, ad + 2 * k, bd + 2 * k, idesc, k);
        }
        umma_commit(b_empty(s));
        umma_commit(acc_full(s));
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ============================ output: one bulk copy per plane and padded time index ============================
    if (lane == 0) {
      int n = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
        const int b = tile / tiles_per_utt;
        const int t1_0 = (tile - b * tiles_per_utt) * kTs;
        for (int tl = 0; tl < kTs; ++tl) {
          const int m = n * kTs + tl, os = m & (kOutStages - 1);
          const int t1 = t1_0 + tl;
          mbar_wait(out_full(os), (m >> 2) & 1);
          if (t1 < t1_len) {
            int tps[3];
            int n_tp = 0;
            tps[n_tp++] = t1 + 1;
            if (t1 == 1) tps[n_tp++] = 0;                          // padded time 0 mirrors t1 = 1
            if (t1 == t1_len - 2) tps[n_tp++] = t1_len + 1;        // padded time T1+1 mirrors t1 = T1-2
            const uint32_t s0 = sbase + kOffOut + os * kOutBytes;
            for (int i = 0; i < n_tp; ++i) {
              const int tp = tps[i];
              const int64_t plane_t = (int64_t)b * 4 + (tp & 1) * 2;
#ifndef C0_NO_STORE        // timing experiment: the staged planes are not written
              bulk_store(out + ((plane_t + 0) * tp2 + (tp >> 1)) * (21 * kC), s0, kPlane0Rows * kC * 2);
              bulk_store(out + ((plane_t + 1) * tp2 + (tp >> 1)) * (21 * kC), s0 + kPlane0Rows * kC * 2,
                         kPlane1Rows * kC * 2);
#endif
            }
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          // at most 2 groups still reading: the buffer staged two steps ago is free again (the epilogue will
          // want it two steps from now)
          asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
          if (m >= 2) mbar_arrive(out_empty((m - 2) & (kOutStages - 1)));
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int num_sms_conv0() { return stac_grid_limit(); }

}  // namespace

// called by stac_conv0_ln_lrelu (conv_frontend.cu) for out_mode == STAC_DT_BF16
int stac_conv0_tc_launch(const float* feats, const float* w0, const float* b0, const float* ln_g, const float* ln_b,
                         int64_t batch, int64_t frames, int t1, uint16_t* out, cudaStream_t st,
                         const uint32_t* utt_max, int per_utt, float top_db, const float* mean, const float* std) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(conv0_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  const int tiles_per_utt = (t1 + kTs - 1) / kTs;
  const int64_t n_tiles = batch * tiles_per_utt;
  if (n_tiles >= (1ll << 30)) return STAC_ERR_UNSUPPORTED_SHAPE;
  const int grid = (int)std::min<int64_t>(n_tiles, num_sms_conv0());
  conv0_tc_kernel<<<grid, kThreads, kSmemBytes, st>>>(feats, w0, b0, ln_g, ln_b, (int)frames, t1, tiles_per_utt,
                                                      (int)n_tiles, reinterpret_cast<__nv_bfloat16*>(out), utt_max, per_utt,
                                                      top_db, mean, std, (int)batch);
  STAC_LAUNCH_CHECK();
}
