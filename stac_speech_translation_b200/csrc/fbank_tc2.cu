// a2 in bf16 mode, second design: STFT as a tensor-core GEMM whose A operand never touches shared memory.
// Reference behaviour: SpeechBrain Fbank as configured at
//   /root/reference/stac-st/hparams/transformer_multitask.yaml:299-302, called at stac-st/inference.py:95.
//
// Same arithmetic as fbank_tc.cu (hamming window, 400-point real DFT folded on its symmetry, fp16 operands, fp32
// accumulation, power, mel, dB, per-utterance maximum); what changes is where the bytes move.  ncu on the first kernel:
// 22 k clk per 128-frame tile against a 7.4 k clk HBM budget, the shared-memory pipe alone needing ~12 k (the windowed /
// folded A operand was built INTO shared memory with 2-way conflicts, read back by the MMA, and 213 KB of twiddles were
// streamed from L2 for every tile).  Here:
//   * thread = frame.  A producer thread reads its own frame from the PCM tile with 128-bit loads (the tile arrives by
//     tensor-map TMA as rows of 32 samples with the 128-byte swizzle; a frame starts every 5 rows and 5 is odd, so the
//     eight threads of a quarter warp hit eight different 16-byte bank groups: conflict-free), folds and windows in
//     registers (one add and one subtract per sample pair: the window is folded into the twiddles) and writes the fp16
//     A operand straight into TENSOR MEMORY (tcgen05.st); the MMAs take A from TMEM (TS form).  No A stores to / MMA
//     reads from smem.
//   * K is cut into 7 stages of 32 samples n = 32 i .. 32 i + 31; stage i is ONE twiddle tile [bins x 64]: columns
//     0-31 w[n] cos(2 pi k n / 400), columns 32-63 -w[n] sin(...).  e[n] = x[n] + x[400 - n] feeds the cos accumulator,
//     o[n] = x[n] - x[400 - n] the sin accumulator; both come out of the same two loads.
//   * pair mode (default): two CTAs of a cluster work as ONE tcgen05 cta_group::2 instance - 2 x 128 frames (M = 256),
//     each CTA holds its own frames' A operand and accumulators and HALF of every twiddle tile (104 bins), so a tile's
//     twiddle traffic from L2 and its shared-memory footprint halve (46 KB per 128 frames instead of 213), which is
//     what makes room for a double-buffered PCM tile.  Only the leader CTA issues MMAs; completion is multicast to both
//     CTAs' barriers; the follower's producers / epilogue arrive on the leader's barriers through the cluster window.
//   * single mode (STAC_FBANK_PAIR=0 on the host side): the same kernel without the cluster.
// Warps: 0-7 epilogue (thread = frame: power from TMEM, mel accumulators in registers, mel weights as constant operands,
// 10 log10 via lg2, row maximum, rows staged and written coalesced; warps 0-3 own mel filters 0-59 = DFT bins 0-99,
// warps 4-7 filters 60-79 = bins 96-200: no filter straddles the cut, so the two halves never talk and the time the
// single-buffered accumulators are held halves), 8-15 A producers (TMEM lane quadrant warp % 4; warps 8-11 build
// the even stages, 12-15 the odd ones), 16 twiddle TMA, 17 PCM loader, 18 TMEM allocation + MMA issue,
// 19 idle (register reallocation works on groups of four warps).
// First version of this file (one bulk copy per 160-sample segment into a padded layout, cluster-scope release on the
// remote arrives): correct on its first run and exactly as slow as the first kernel - 130 bulk copies took 8-10 k clk per
// tile to ISSUE, and `mbarrier.arrive.release.cluster` compiles to MEMBAR.ALL.GPU (profiles/r4/).
#include <algorithm>
#include <utility>
#include <cuda_fp16.h>
#include "tc_common.cuh"
#include "fbank_mel_structure.h"

namespace {

using namespace tc;

constexpr int kNfft = 400, kHop = 160, kMel = 80;
constexpr int kRows = 128;                 // frames per CTA tile
constexpr int kBins = 208;                 // 201 bins padded to a multiple of 16 (UMMA N)
constexpr int kStages = 7;                 // K stages of 32 samples per tile (the last one holds n = 192..207 only)
constexpr int kASlots = 3;                 // A ring in tensor memory: 3 x (16 columns e + 16 columns o)
constexpr int kBSlots = 3;                 // twiddle tiles in flight
// PCM tile in shared memory: rows of 32 samples starting at global sample 160 (t0 - 2) (a multiple of 32), i.e. kLead
// samples before the first frame's first sample; 4 TMA boxes of 168 rows (a multiple of 8: every box starts on a
// 1024-byte swizzle period)
constexpr int kLead = 2 * kHop - kNfft / 2;                   // 120
constexpr int kPcmBoxRows = 168, kPcmBoxes = 4;
constexpr int kPcmRows = kPcmBoxRows * kPcmBoxes;             // 672 >= (127 * 160 + 399 + 120) / 32 + 1 = 652
constexpr int kPcmBytes = kPcmRows * 128;                     // 86016 = 84 KB
constexpr int kPcmBufBytes = kPcmBytes;
static_assert(kPcmBufBytes % 1024 == 0 && kPcmBoxRows % 8 == 0 && kPcmBoxRows <= 256, "PCM tile geometry");
constexpr int kOutRows = 32, kOutStride = 84;                 // output staging: one epilogue warp's rows at a time; the
                                                              // 336-byte pitch makes thread = row 128-bit stores conflict-free
constexpr int kEpiWarps = 8, kProdWarps = 8;
constexpr int kWarpB = 16, kWarpPcm = 17, kWarpMma = 18;
constexpr int kThreads = 20 * 32;
constexpr int kMelCut = 60;                                   // epilogue warps 0-3: filters [0, 60), warps 4-7: [60, 80)
constexpr int kCutChunkA = 13, kCutChunkB = 12;               // ... = 8-bin chunks [0, 13) (bins 0-103) and [12, 25) (96-199)
// launch: 640 threads x 96 registers = 61 440, which is also the pool setmaxnreg redistributes (NOT the SM's 65 536: an
// increase beyond what the CTA was launched with blocks for ever): 256 x 144 (epilogue) + 256 x 80 (producers) + 128 x 32
constexpr int kRegsEpi = 136, kRegsProd = 88, kRegsCtl = 32;
static_assert(256 * kRegsEpi + 256 * kRegsProd + 128 * kRegsCtl <= kThreads * 96, "register pool");
constexpr int kNumBars = 2 * kASlots + 2 * kBSlots + 2 + 4;
constexpr int kTabFloats = 2 * kBins;                         // per-bin mel weights [208][2]

// tensor-memory columns: cos accumulator 0..207, sin accumulator 256..463, the six 16-column halves of the A ring between
__device__ __forceinline__ uint32_t a_half_col(int h) { return h < 3 ? 208u + 16u * h : 464u + 16u * (h - 3); }
constexpr uint32_t kColCos = 0, kColSin = 256;

template <bool kPair>
struct Cfg {
  static constexpr int kCtas = kPair ? 2 : 1;
  static constexpr int kBRows = kBins / kCtas;
  static constexpr int kBBytes = kBRows * 128;            // [rows x 64 fp16], 128-byte swizzle (multiple of 1024)
  static constexpr int kPcmBufs = kPair ? 2 : 1;
  static constexpr int kOffB = 0;
  static constexpr int kOffPcm = kBSlots * kBBytes;
  static constexpr int kOffOut = kOffPcm + kPcmBufs * kPcmBufBytes;
  static constexpr int kOffW = ((kOffOut + kOutRows * kOutStride * 4 + 15) / 16) * 16;    // per-bin mel weights [208][2]
  static constexpr int kOffBar = kOffW + 2 * kBins * 4;
  static constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16 + 1024;
};
static_assert(Cfg<true>::kSmemBytes <= 232448 && Cfg<false>::kSmemBytes <= 232448, "shared memory budget");

// kind::f16 instruction descriptor: D fp32, A/B fp16, both K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- cluster / cta_group::2 forms ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// arrive on a barrier anywhere in the cluster (address from map_to_cta)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  // (default semantics, as CUTLASS' ClusterBarrier::arrive: `.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR per
  //  arrive, ~2 k clk; what the arrive publishes here lives in tensor memory and is ordered by tcgen05.wait::st and the
  //  tcgen05 fences on both sides)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
template <bool kPair>
__device__ __forceinline__ void tmem_alloc_g(uint32_t dst_smem) {
  if constexpr (kPair) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst_smem) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst_smem) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
}
template <bool kPair>
__device__ __forceinline__ void tmem_dealloc_g(uint32_t taddr) {
  if constexpr (kPair)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T : A = 128 lanes (frames) x 8 columns of packed fp16 pairs per K = 16 step
template <bool kPair>
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  if constexpr (kPair)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on `bar` (same shared-memory offset in both CTAs of the pair) when the MMAs issued so far have completed
template <bool kPair>
__device__ __forceinline__ void umma_commit_g(uint32_t bar) {
  if constexpr (kPair)
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// twiddle tile load; in pair mode the completion goes to the LEADER's barrier (cluster address)
template <bool kPair>
__device__ __forceinline__ void tma_load_b(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  if constexpr (kPair)
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1) : "memory");
  else
    tma_load_2d(dst, m, bar, c0, c1);
}

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  const __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 16 bytes holding tile sample Q + 32 * (rowb - 5 r) of the frame that starts at row 5 r (Q compile-time, a multiple of
// 4): row q of the tile keeps its 16-byte piece p at piece p ^ (q & 7) (TMA SWIZZLE_128B).  The stage index only moves
// whole rows (32 samples per stage), so it travels in rowb and the stage loop needs no unrolling.
template <int Q>
__device__ __forceinline__ const unsigned char* pcm_piece(const unsigned char* buf, int rowb) {
  constexpr int R = Q >> 5, P = (Q & 31) >> 2;
  const int row = rowb + R;
  return buf + row * 128 + ((P ^ (row & 7)) << 4);
}

// ---- A producer: one stage half = 16 columns n0 .. n0 + 15 (n0 = 32 i + 16 HH) of e and o for this thread's frame ----
struct StageRegs {
  float4 a[4], b[4];
  float carry;
};
template <int HH, int J = 0>
__device__ __forceinline__ void stage_load(const unsigned char* buf, int r5, int i, StageRegs& r) {
  if constexpr (J < 4) {
    r.a[J] = *reinterpret_cast<const float4*>(pcm_piece<kLead + 16 * HH + 4 * J>(buf, r5 + i));         // x[n0 + 4 j .. + 3]
    r.b[J] = *reinterpret_cast<const float4*>(pcm_piece<kLead + 396 - 16 * HH - 4 * J>(buf, r5 - i));   // x[396 - n0 - 4 j ..]
    stage_load<HH, J + 1>(buf, r5, i, r);
  } else {
    r.carry = *reinterpret_cast<const float*>(pcm_piece<kLead + 400 - 16 * HH>(buf, r5 - i));           // x[400 - n0]
  }
}
template <int HH>
__device__ __forceinline__ void stage_fold(const StageRegs& r, int i, uint32_t (&ev)[8], uint32_t (&ov)[8]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float av[4] = {r.a[j].x, r.a[j].y, r.a[j].z, r.a[j].w};
    // mirror samples x[400 - n] for n = n0 + 4 j + {0, 1, 2, 3}
    const float bv[4] = {j == 0 ? r.carry : r.b[j - 1].x, r.b[j].w, r.b[j].z, r.b[j].y};
    float e[4], o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // the window lives in the twiddles (w[n] cos, -w[n] sin); n = 0 and n = 200 have no mirror partner (their sine
      // twiddles are zero, so o only has to be finite); columns n > 200 meet zero twiddle rows
      float m = bv[k];
      if (HH == 0 && j == 0 && k == 0) m = i == 0 ? 0.f : m;                 // n = 0
      if (HH == 0 && j == 2 && k == 0) m = i == kStages - 1 ? 0.f : m;       // n = 200
      e[k] = av[k] + m;
      o[k] = av[k] - bv[k];
    }
    ev[2 * j] = pack_f16x2(e[0], e[1]);
    ev[2 * j + 1] = pack_f16x2(e[2], e[3]);
    ov[2 * j] = pack_f16x2(o[0], o[1]);
    ov[2 * j + 1] = pack_f16x2(o[2], o[3]);
  }
}

// ---- epilogue: one DFT bin (compile-time K) into those of the mel accumulators [M0, M1) it feeds ----
// (weights: the first version read them as __constant__ operands; ptxas turns those into LDCU.128 -> uniform register ->
//  FFMA, with too few uniform registers to run ahead: 440 clk per 16 bins.  Here they come from shared memory, one
//  chunk ahead, into ordinary registers.)
template <int K, int M0, int M1>
__device__ __forceinline__ void mel_bin(uint32_t re, uint32_t im, float w0, float w1, float (&acc)[M1 - M0]) {
  constexpr int first = kMelFirst[K], cnt = kMelCnt[K];
  constexpr bool f0 = cnt > 0 && first >= M0 && first < M1, f1 = cnt > 1 && first + 1 >= M0 && first + 1 < M1;
  if constexpr (f0 || f1) {
    const float a = __uint_as_float(re), b = __uint_as_float(im);
    const float p = fmaf(a, a, b * b);
    if constexpr (f0) acc[first - M0] = fmaf(p, w0, acc[first - M0]);
    if constexpr (f1) acc[first + 1 - M0] = fmaf(p, w1, acc[first + 1 - M0]);
  }
}
struct MelRegs {
  uint32_t re[8], im[8];
  float4 w[4];
};
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr) : "memory");
}
template <int K0>
__device__ __forceinline__ void mel_fetch(uint32_t t_cos, uint32_t t_sin, const float4* __restrict__ wsm, MelRegs& m) {
  tmem_ld8(t_cos + K0, m.re);
  tmem_ld8(t_sin + K0, m.im);
#pragma unroll
  for (int j = 0; j < 4; ++j) m.w[j] = wsm[K0 / 2 + j];
}
template <int K0, int M0, int M1, int... I>
__device__ __forceinline__ void mel_chunk(const MelRegs& m, float (&acc)[M1 - M0], std::integer_sequence<int, I...>) {
  const float wv[16] = {m.w[0].x, m.w[0].y, m.w[0].z, m.w[0].w, m.w[1].x, m.w[1].y, m.w[1].z, m.w[1].w,
                        m.w[2].x, m.w[2].y, m.w[2].z, m.w[2].w, m.w[3].x, m.w[3].y, m.w[3].z, m.w[3].w};
  (mel_bin<K0 + I, M0, M1>(m.re[I], m.im[I], wv[2 * I], wv[2 * I + 1], acc), ...);
}
// TMEM read-out of the 8-bin chunks [C, CEnd) with the next chunk's loads in flight behind the current chunk's arithmetic
template <int C, int CEnd, int M0, int M1>
__device__ __forceinline__ void mel_all(uint32_t t_cos, uint32_t t_sin, const float4* __restrict__ wsm, MelRegs (&m)[2],
                                        float (&acc)[M1 - M0]) {
  if constexpr (C < CEnd) {
    tmem_ld_wait();                                           // chunk C has landed
    if constexpr (C + 1 < CEnd) mel_fetch<(C + 1) * 8>(t_cos, t_sin, wsm, m[(C + 1) & 1]);
    mel_chunk<C * 8, M0, M1>(m[C & 1], acc, std::make_integer_sequence<int, 8>{});
    mel_all<C + 1, CEnd, M0, M1>(t_cos, t_sin, wsm, m, acc);
  }
}

#ifdef FBANK2_TRACE    // timing experiment only (tools/trace_fbank.py): CTA 0 logs clock64 per (role, tile ordinal, event)
__device__ unsigned int* g_fb2_trace = nullptr;
#define BTRACE(role, ev, n) do { if (blockIdx.x == 0 && g_fb2_trace != nullptr && (n) < 8) g_fb2_trace[((role) * 8 + (n)) * 16 + (ev)] = (unsigned int)clock64(); } while (0)
#else
#define BTRACE(role, ev, n) do {} while (0)
#endif

template <bool kPair>
__global__ void __launch_bounds__(kThreads, 1)
fbank_tc2_kernel(const __grid_constant__ CUtensorMap tmap_tw, const __grid_constant__ CUtensorMap tmap_pcm,
                 int n_frames, int tiles_per_utt, int num_tiles, const float* __restrict__ tables,
                 float* __restrict__ out, unsigned int* __restrict__ utt_max) {
  using C = Cfg<kPair>;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bars = sbase + C::kOffBar;
  auto a_full = [&](int s) { return bars + 8u * s; };                                  // leader's copy is the live one
  auto a_empty = [&](int s) { return bars + 8u * (kASlots + s); };
  auto b_full = [&](int s) { return bars + 8u * (2 * kASlots + s); };                  // leader's copy
  auto b_empty = [&](int s) { return bars + 8u * (2 * kASlots + kBSlots + s); };
  const uint32_t misc = bars + 8u * (2 * kASlots + 2 * kBSlots);
  const uint32_t tfull_bar = misc, tempty_bar = misc + 8u;                             // tempty: leader's copy
  auto pcm_full = [&](int b) { return misc + 16u + 8u * b; };
  auto pcm_free = [&](int b) { return misc + 32u + 8u * b; };
  const uint32_t tmem_slot = bars + 8u * kNumBars;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int unit = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;                  // pair (or CTA) index
  const int n_units = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int unit_tiles = (num_tiles + C::kCtas - 1) / C::kCtas;

  if (tid == 0) {
    prefetch_tmap(&tmap_tw);
    prefetch_tmap(&tmap_pcm);
    for (int s = 0; s < kASlots; ++s) { mbar_init(a_full(s), (kProdWarps / 2) * C::kCtas); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < kBSlots; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, kEpiWarps * C::kCtas);
    for (int b = 0; b < 2; ++b) { mbar_init(pcm_full(b), 1); mbar_init(pcm_free(b), kProdWarps); }
    fence_barrier_init();
  }
  if (warp == kWarpMma) tmem_alloc_g<kPair>(tmem_slot);
  for (int i = tid; i < 2 * kBins; i += kThreads) reinterpret_cast<float*>(sptr + C::kOffW)[i] = __ldg(tables + i);
  tc_fence_before();
  if constexpr (kPair) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // barriers that live in the leader CTA, as seen from this CTA
  auto leader_bar = [&](uint32_t local) { return kPair ? map_to_cta(local, 0) : local; };

  if (warp >= kEpiWarps && warp < kEpiWarps + kProdWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsProd));
    // ============================ A producers: fold + window in registers, fp16 A operand into TMEM ============================
    const int q = warp & 3, half = (warp - kEpiWarps) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t a_full_remote0 = leader_bar(a_full(0));
    // this warp builds the stages i with i % 2 == half, all 32 columns (the stage loop is a real loop: the first version
    // unrolled all seven stages, and a quarter of the kernel's stall samples were instruction fetch)
    {
      int tn = 0;
      for (int ut = unit; ut < unit_tiles; ut += n_units, ++tn) {
        const int buf = tn % C::kPcmBufs;
        const unsigned char* pbuf = sptr + C::kOffPcm + buf * kPcmBufBytes;
        mbar_wait(pcm_full(buf), (uint32_t)(tn / C::kPcmBufs) & 1);
        if (tid == kEpiWarps * 32) BTRACE(0, 0, tn);
#pragma unroll 1
        for (int i = half; i < kStages; i += 2) {
          const int g = kStages * tn + i;                   // running stage count: ring slot and its parity
          const int slot = g % kASlots;
          const uint32_t phase = (uint32_t)(g / kASlots) & 1;
          // fold first (the PCM tile is there), then wait for the slot: the other warp set is doing the same for the
          // neighbouring stage, so the store / announce latency of one stage hides behind the other's
          uint32_t ev[2][8], ov[2][8];
#ifndef FB2_NO_BUILD      // timing experiment: no loads, no arithmetic, no TMEM stores
          {
            StageRegs regs;
            stage_load<0>(pbuf, 5 * r, i, regs);
            stage_fold<0>(regs, i, ev[0], ov[0]);
            if (i < kStages - 1) {                          // the last stage holds 16 real columns
              stage_load<1>(pbuf, 5 * r, i, regs);
              stage_fold<1>(regs, i, ev[1], ov[1]);
            }
          }
#endif
          if (i == 2 && tid == kEpiWarps * 32) BTRACE(0, 8, tn);
          mbar_wait(a_empty(slot), phase ^ 1);
          tc_fence_after();
          if (i == 2 && tid == kEpiWarps * 32) BTRACE(0, 9, tn);
#ifndef FB2_NO_BUILD
          const uint32_t e_addr = tmem_base + lane_off + a_half_col(2 * slot), o_addr = tmem_base + lane_off + a_half_col(2 * slot + 1);
          tmem_st8(e_addr, ev[0]);
          tmem_st8(o_addr, ov[0]);
          if (i < kStages - 1) {
            tmem_st8(e_addr + 8u, ev[1]);
            tmem_st8(o_addr + 8u, ov[1]);
          }
          tmem_st_wait();
#endif
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (kPair) mbar_arrive_remote(a_full_remote0 + 8u * slot); else mbar_arrive(a_full(slot));
          }
          if (tid == kEpiWarps * 32) BTRACE(0, 1 + i, tn);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(pcm_free(buf));           // this warp no longer reads the PCM tile
      }
    }
  } else if (warp >= kWarpB) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsCtl));
    if (warp == kWarpPcm) {
    // ============================ PCM loader: 4 boxes of 168 rows x 32 samples per tile ============================
    // rows before the utterance (negative coordinates) and past its end are zero-filled by the TMA unit: the STFT's
    // centre padding costs nothing
    int tn = 0;
    for (int ut = unit; ut < unit_tiles; ut += n_units, ++tn) {
      const int tile = min(ut * C::kCtas + (int)rank, num_tiles - 1);   // an odd tile count: the follower repeats the last tile
      const int b = tile / tiles_per_utt;
      const int t0 = (tile - b * tiles_per_utt) * kRows;
      const int buf = tn % C::kPcmBufs;
      if (lane == 0) BTRACE(3, 0, tn);
      mbar_wait(pcm_free(buf), ((uint32_t)(tn / C::kPcmBufs) & 1) ^ 1);
      if (lane == 0) BTRACE(3, 1, tn);
      if (elect_one()) {
#ifdef FB2_NO_PCM         // timing experiment: the producers read whatever the buffer holds
        mbar_arrive(pcm_full(buf));
#else
        mbar_arrive_expect_tx(pcm_full(buf), (uint32_t)kPcmBytes);
        const uint32_t dst = sbase + C::kOffPcm + buf * kPcmBufBytes;
        for (int k = 0; k < kPcmBoxes; ++k)
          tma_load_3d(dst + k * kPcmBoxRows * 128, &tmap_pcm, pcm_full(buf), 0, 5 * (t0 - 2) + k * kPcmBoxRows, b);
#endif
      }
      __syncwarp();
      if (lane == 0) BTRACE(3, 3, tn);
    }
    } else if (warp == kWarpB) {
    // ============================ twiddle tiles: this CTA's bins of every stage ============================
    const uint32_t bfull_remote0 = leader_bar(b_full(0));
    int slot = 0;
    uint32_t phase = 0;
    for (int ut = unit; ut < unit_tiles; ut += n_units) {
      for (int i = 0; i < kStages; ++i) {
        mbar_wait(b_empty(slot), phase ^ 1);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(b_full(slot), (uint32_t)(C::kBBytes * C::kCtas));
          tma_load_b<kPair>(sbase + C::kOffB + slot * C::kBBytes, &tmap_tw, bfull_remote0 + 8u * slot, 0,
                            i * kBins + (int)rank * C::kBRows);
        }
        __syncwarp();
        if (++slot == kBSlots) { slot = 0; phase ^= 1; }
      }
    }
    } else if (warp == kWarpMma) {
    // ============================ MMA issuer (leader CTA only) ============================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_f16(kRows * C::kCtas, kBins);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int tn = 0;
      for (int ut = unit; ut < unit_tiles; ut += n_units, ++tn) {
        mbar_wait(tempty_bar, ((uint32_t)tn & 1) ^ 1);
        if (lane == 0) BTRACE(1, 0, tn);
        tc_fence_after();
        for (int i = 0; i < kStages; ++i) {
          mbar_wait(b_full(sb), pb);
          mbar_wait(a_full(sa), pa);
          if (lane == 0) BTRACE(1, 1 + i, tn);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t bd = make_smem_desc_sw128(sbase + C::kOffB + sb * C::kBBytes);
            const uint32_t ae = tmem_base + a_half_col(2 * sa), ao = tmem_base + a_half_col(2 * sa + 1);
            const int steps = i == kStages - 1 ? 1 : 2;
            for (int k = 0; k < steps; ++k) {
              umma_f16_ts<kPair>(tmem_base + kColCos, ae + 8u * k, bd + 2 * k, idesc, (i | k) != 0);
              umma_f16_ts<kPair>(tmem_base + kColSin, ao + 8u * k, bd + 4 + 2 * k, idesc, (i | k) != 0);
            }
            umma_commit_g<kPair>(a_empty(sa));
            umma_commit_g<kPair>(b_empty(sb));
            if (i == kStages - 1) umma_commit_g<kPair>(tfull_bar);
          }
          __syncwarp();
          if (++sa == kASlots) { sa = 0; pa ^= 1; }
          if (++sb == kBSlots) { sb = 0; pb ^= 1; }
        }
      }
    }
    }
  } else if (warp < kEpiWarps) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpi));
    // ============================ epilogue: thread = frame, warp group = half of the mel filters ============================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tempty_remote = leader_bar(tempty_bar);
    float* stage_out = reinterpret_cast<float*>(sptr + C::kOffOut);
    const float4* wsm = reinterpret_cast<const float4*>(sptr + C::kOffW);
    auto run = [&](auto G) {
      constexpr int g = decltype(G)::value;
      constexpr int M0 = g == 0 ? 0 : kMelCut, M1 = g == 0 ? kMelCut : kMel;
      constexpr int C0 = g == 0 ? 0 : kCutChunkB, C1 = g == 0 ? kCutChunkA : 25;
      int tn = 0;
      for (int ut = unit; ut < unit_tiles; ut += n_units, ++tn) {
        const int tile_raw = ut * C::kCtas + (int)rank;
        const bool active = tile_raw < num_tiles;
        const int tile = min(tile_raw, num_tiles - 1);
        const int b = tile / tiles_per_utt;
        const int t0 = (tile - b * tiles_per_utt) * kRows;
        const int valid = active ? min(kRows, n_frames - t0) : 0;
        if (tid == 0) BTRACE(2, 0, tn);
        mbar_wait(tfull_bar, (uint32_t)tn & 1);
        if (tid == 0) BTRACE(2, 1, tn);
        tc_fence_after();
        float acc[M1 - M0];
#pragma unroll
        for (int m = 0; m < M1 - M0; ++m) acc[m] = 0.f;
#ifndef FB2_NO_EPI        // timing experiment: accumulators released unread
        {
          MelRegs m[2];
          mel_fetch<C0 * 8>(tmem_base + lane_off + kColCos, tmem_base + lane_off + kColSin, wsm, m[C0 & 1]);
          mel_all<C0, C1, M0, M1>(tmem_base + lane_off + kColCos, tmem_base + lane_off + kColSin, wsm, m, acc);
        }
#endif
        // this warp has consumed its share of the accumulators; when all eight (sixteen) have, the next tile's MMAs start
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kPair) mbar_arrive_remote(tempty_remote); else mbar_arrive(tempty_bar);
        }
        if (tid == 0) BTRACE(2, 2, tn);
        float vmax = -INFINITY;
#pragma unroll
        for (int m = 0; m < M1 - M0; ++m) {
          acc[m] = 3.0102999566398120f * lg2_approx(fmaxf(acc[m], 1e-10f));     // 10 log10(x) = 10 log10(2) lg2(x)
          if (r < valid) vmax = fmaxf(vmax, acc[m]);
        }
        vmax = warp_max(vmax);
        if (lane == 0 && vmax > -INFINITY) atomicMax(utt_max + b, float_to_ordered(vmax));
        if (tid == 0) BTRACE(2, 3, tn);
        // rows leave through a 32-row staging tile, one lane quadrant at a time (the two warps of the quadrant write
        // their filters' columns), in 16-byte pieces; [rows][80] floats are contiguous in the output
        float4* dst4 = reinterpret_cast<float4*>(out + ((int64_t)b * n_frames + t0) * kMel);
#pragma unroll 1
        for (int part = 0; part < 4; ++part) {
          if (q == part) {
#pragma unroll
            for (int j = 0; j < (M1 - M0) / 4; ++j)
              *reinterpret_cast<float4*>(stage_out + lane * kOutStride + M0 + 4 * j) =
                  make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
#ifndef FB2_NO_STORE      // timing experiment: rows staged but not written
          const int n_out = (min(valid, part * kOutRows + kOutRows) - part * kOutRows) * (kMel / 4);   // <= 0 past the end
#pragma unroll
          for (int it = 0; it < (kOutRows * (kMel / 4) + kEpiWarps * 32 - 1) / (kEpiWarps * 32); ++it) {
            const int gi = tid + it * kEpiWarps * 32;
            const int rr = gi / (kMel / 4), j = gi - rr * (kMel / 4);
            if (gi < n_out)
              dst4[part * kOutRows * (kMel / 4) + gi] = *reinterpret_cast<const float4*>(stage_out + rr * kOutStride + 4 * j);
          }
#endif
          asm volatile("bar.sync 1, 256;" ::: "memory");          // the staging tile is rewritten next
        }
        if (tid == 0) BTRACE(2, 4, tn);
      }
    };
    if (warp < 4) run(std::integral_constant<int, 0>{}); else run(std::integral_constant<int, 1>{});
  }

  tc_fence_before();
  if constexpr (kPair) cluster_sync_all(); else __syncthreads();
  if (warp == kWarpMma) { tc_fence_after(); tmem_dealloc_g<kPair>(tmem_base); }
}

template <bool kPair>
int launch_fbank_tc2(const CUtensorMap& tw, const CUtensorMap& pm, int n_frames, int tiles_per_utt, int num_tiles,
                     const float* tables, float* out, unsigned int* utt_max, cudaStream_t stream) {
  using C = Cfg<kPair>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(fbank_tc2_kernel<kPair>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  const int units = (num_tiles + C::kCtas - 1) / C::kCtas;
  const int grid_units = std::min(units, std::max(1, stac_grid_limit() / C::kCtas));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(grid_units * C::kCtas));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C::kCtas;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, fbank_tc2_kernel<kPair>, tw, pm, n_frames, tiles_per_utt, num_tiles, tables,
                                     out, utt_max);
  if (e != cudaSuccess) return (int)e;
  STAC_LAUNCH_CHECK();
}

}  // namespace

#ifdef FBANK2_TRACE
extern "C" int stac_fbank2_trace(unsigned int* buf) { cudaMemcpyToSymbol(g_fb2_trace, &buf, sizeof(buf)); return 0; }
#endif

extern "C" int stac_fbank_tc2_tables_floats(void) { return kTabFloats; }
extern "C" int stac_fbank_tc2_twiddle_halfs(void) { return kStages * kBins * 64; }

extern "C" int stac_fbank_logmel_tc2(const float* pcm, int64_t batch, int64_t n_samples, int64_t pcm_row_stride,
                                     const float* tables, const uint16_t* twiddles, float* logmel_db,
                                     uint32_t* utt_max_ordered, int pair, void* stream) {
  STAC_REQUIRE(pcm && tables && twiddles && logmel_db && utt_max_ordered);
  STAC_REQUIRE(batch > 0 && batch < 65536 && n_samples > 0 && pcm_row_stride >= n_samples);
  // the PCM tile travels as tensor-map boxes of 32-sample rows: 16-byte aligned rows, a sample count that is a multiple
  // of 32 (other lengths: stac_fbank_logmel_tc / stac_fbank_logmel), 16-byte aligned output rows
  if (n_samples % 32 != 0 || pcm_row_stride % 4 != 0 || (reinterpret_cast<uintptr_t>(pcm) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(logmel_db) & 15) != 0)
    return STAC_ERR_UNSUPPORTED_SHAPE;
  const int64_t n_frames = 1 + n_samples / kHop;
  const int64_t tiles_per_utt = ceil_div64(n_frames, kRows);
  if (n_frames >= (1ll << 30) || batch * tiles_per_utt >= (1ll << 30)) return STAC_ERR_UNSUPPORTED_SHAPE;
  CUtensorMap tw;
  {
    // twiddles fp16 [7 stages x 208 bins][64 columns: cos n = 32 i .. + 31 | -sin same n]
    const uint64_t dims[2] = {64, (uint64_t)(kStages * kBins)};
    const uint64_t str[1] = {64 * 2};
    const uint32_t box[2] = {64, (uint32_t)(pair ? kBins / 2 : kBins)};
    int r = encode_map(&tw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, twiddles, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  CUtensorMap pm;
  {
    // PCM fp32 [B][n_samples / 32 rows][32]: rows outside an utterance read as zeros
    const uint64_t dims[3] = {32, (uint64_t)(n_samples / 32), (uint64_t)batch};
    const uint64_t str[2] = {128, (uint64_t)pcm_row_stride * 4};
    const uint32_t box[3] = {32, (uint32_t)kPcmBoxRows, 1};
    int r = encode_map(&pm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, pcm, 3, dims, str, box);
    if (r != STAC_OK) return r;
  }
  cudaStream_t st = as_stream(stream);
  cudaError_t ce = cudaMemsetAsync(utt_max_ordered, 0, (size_t)batch * sizeof(uint32_t), st);
  if (ce != cudaSuccess) return (int)ce;
  const int num_tiles = (int)(batch * tiles_per_utt);
  if (pair)
    return launch_fbank_tc2<true>(tw, pm, (int)n_frames, (int)tiles_per_utt, num_tiles, tables, logmel_db,
                                  utt_max_ordered, st);
  return launch_fbank_tc2<false>(tw, pm, (int)n_frames, (int)tiles_per_utt, num_tiles, tables, logmel_db,
                                 utt_max_ordered, st);
}
