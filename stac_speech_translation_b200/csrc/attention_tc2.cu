// Second decomposition of the tcgen05 flash attention (stac_mha_bf16_v2): P in tensor memory, one thread per query row,
// scores DOUBLE-BUFFERED per query group so that the softmax warps never wait for the tensor pipe.
//
// Reference behaviour replaced: the same as attention_tc.cu (torch.nn.MultiheadAttention slow path with a -inf
// key-padding mask, reached from /root/reference/stac-st/modules/TransformerMultiTask.py:304-308).
//
// History.  The first kernel (attention_tc.cu: two threads per row, 64-key tiles, P through shared
// memory) balances two halves that are both too slow.  The first version of this file (round 2, first GPU call: parity
// green, 83 us against 87 us) moved to the FlashAttention-4 layout - 128-key tiles, thread = row, P written back into
// TMEM over its own scores and taken from there as the A operand of P.V - but kept ONE score buffer per query group.
// Its clock trace (profiles/r3/r3a_mha2_trace.log) shows what that costs: per 128-key step a group spends ~2000 clk in
// its exponentials and then ~1850 clk waiting for P.V(g) and S(g+1) to come back from the tensor pipe (both groups end
// up in phase, so they also share the MUFU while they compute and leave it idle while they wait): 4980 clk per step
// against a MUFU floor of 2048.  This version removes the wait instead of shortening it:
//   * 96-key tiles, so that TWO score buffers per group fit in tensor memory next to O (4 x 96 + 2 x 64 = 512 columns);
//     S(g+1) and S(g+2) are issued while softmax(g) runs, and the softmax warps go from one tile straight to the next;
//   * O is single-buffered (one item boundary per 8 steps; the epilogue warpgroup drains it in a few hundred clocks);
//   * everything else as before: ONE thread per query row (no row-max exchange), P never touches shared memory
//     (tcgen05.st over the scores, TS-form tcgen05.mma), lazy rescaling (O is only touched by the CUDA cores when a row
//     maximum rises by more than 2^40 - then behind a wait for the P.V in flight), O / l -> bf16 -> smem -> TMA store by
//     its own warpgroup, optional share of the exponentials on the FMA pipe (-DMHA2_POLY=n, default 0: measured slower).
// Budget per 96-key step of one CTA (2 x 128 query rows): MUFU 2 x 128 x 96 / 16 = 1536 clk, tensor pipe
// 2 x (192 + 192) = 768 clk.
//
// The TMEM conventions this relies on were probed with exact integer data on a B200 (tools/probe_ts_mma.cu,
// profiles/r3/r3a_first_call_verification.log): thread = lane = query row for 32x32b loads / stores, P as packed bf16
// pairs (keys 2c, 2c+1 in 32-bit column c, even key in the LOW half) on top of its own scores, 8 columns per K = 16 step
// of the TS-form MMA.
//
// TMEM (512 columns): S/P buffer (group w, buffer sb) at (w * 2 + sb) * 96 (96 fp32 score columns; P = 48 columns of
// packed bf16 pairs on top of the first 48); O[group] at 384 + group * 64.
// Threads (512): warps 0-3 / 4-7 softmax group 0 / 1 (one thread per row), 8-11 epilogue warpgroup, 12/13 MMA issuers of
// group 0/1, 14 TMA producer, 15 idle; setmaxnreg 176 / 80.  (-DMHA2_SPLIT: 768 threads - warps 0-7 softmax group 0, 0-3
// keys 0-47 and 4-7 keys 48-95 of every tile, 8-15 group 1, 16-19 epilogue, 20/21 issuers, 22 producer; setmaxnreg 96 /
// 48.)  warp & 3 = TMEM lane quarter in every role that touches tensor memory.
#include <algorithm>
#include <type_traits>
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int kHd = 64, kQTile = 128, kKTile = 96;
#ifndef MHA2_KV_STAGES
#define MHA2_KV_STAGES 4
#endif
constexpr int kKvStages = MHA2_KV_STAGES;            // 24 KB each (K 12 KB + V 12 KB); S runs two tiles ahead of P.V: >= 3
constexpr int kKBytes = kKTile * kHd * 2;             // one K (or V) tile
constexpr int kStageBytes = 2 * kKBytes;
constexpr int kSCols = kKTile;                         // fp32 score columns of one S buffer
constexpr int kOCol = 4 * kSCols;                      // first O column
// -DMHA2_SPLIT (measured, not adopted): TWO softmax threads per query row, each owning 48 of the tile's 96 keys - 16
// softmax warps, four per scheduler instead of two, one exchange of the row maximum per step between the two warps of
// a row (shared memory + a 64-thread named barrier).  Parity green; 73.0 us against 69.5 us at the benchmark shape
// (profiles/r3/r3i_mha_split.log): the exponential phase becomes MUFU-bound (4 x 48 x 8 = 1536 clk, measured ~1650), but
// all four warps of a scheduler still move in lockstep (same score barrier, pair barrier), so the ~1800 clk of barrier /
// TMEM load / maximum / store-wait per step stay un-overlapped and the step is as long as before.
#ifdef MHA2_SPLIT
constexpr int kSoftWarps = 16;
constexpr int kRegsSoftmax = 96, kRegsOther = 48;      // 512 * 96 + 256 * 48 = 61440 <= 65536 (launch: 768 * 80)
#else
constexpr int kSoftWarps = 8;
constexpr int kRegsSoftmax = 176, kRegsOther = 80;     // 256 * 176 + 256 * 80 = 512 * 128
#endif
constexpr int kEpiWarp0 = kSoftWarps;                  // 4 epilogue warps, then 2 MMA issuers, the TMA producer, 1 idle
constexpr int kIssuer0 = kSoftWarps + 4, kProducer = kSoftWarps + 6;
constexpr int kThreads = (kSoftWarps + 8) * 32;
#ifndef MHA2_POLY
#define MHA2_POLY 0
#endif
#ifndef MHA2_OFFSET
#define MHA2_OFFSET 0
#endif
#ifndef MHA2_MAXCHAINS
#define MHA2_MAXCHAINS 2
#endif
// shared memory map (bytes, from a 1024-aligned base)
constexpr int kOffQ = 0;                               // [2 bufs][2 groups] x 16 KB
constexpr int kOffOut = 65536;                         // [2 groups] x 16 KB: normalised bf16 O tile for the TMA store
constexpr int kOffKV = 98304;                          // [stages] x (K 16 KB + V 16 KB)
constexpr int kOffX = kOffKV + kKvStages * kStageBytes; // float [2 groups][2 halves][128 rows]: (partial) row sums
constexpr int kOffMax = kOffX + 2 * 2 * 128 * 4;       // float [2 step parities][2 groups][2 halves][128 rows]: row maxima (MHA2_SPLIT)
constexpr int kOffLen = kOffMax + 2 * 2 * 2 * 128 * 4; // int4 [kLenCache]: (b, h, q0, n_keys) of this CTA's first work items
constexpr int kLenCache = 128;
constexpr int kOffBar = kOffLen + kLenCache * 16;
constexpr int kNumBars = 8 + 2 * kKvStages + 2 * 9;
constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16 + 1024;
static_assert(kKvStages >= 3 && kSmemBytes <= 232448, "K/V ring does not fit in shared memory");
static_assert(kOCol + 2 * kHd <= 512 && kKTile % 32 == 0 && kKBytes % 1024 == 0, "tensor-memory / tile layout");
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_mufu(float x) {
#ifdef MHA2_NOEXP     // timing experiment only: results are wrong
  return x;
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}
// 2^x on the FMA pipe for x in [-126, 126]: round-to-nearest split x = n + f, |f| <= 0.5, minimax cubic for 2^f with
// p(0) = 1 (maximum relative error 1.0e-4, checked in float32 over [-126, 40]; bf16 resolution is 3.9e-3), exponent
// added as an integer
__device__ __forceinline__ float ex2_fma(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;                     // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);
  float p = 0.0550129f;
  p = fmaf(p, f, 0.24221165f);
  p = fmaf(p, f, 0.69328244f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// the same on a pair, in packed fp32x2 arithmetic (FADD2 / FFMA2): 10 instructions for two exponentials
__device__ __forceinline__ float2 ex2_fma2(float2 x) {
  x.x = fmaxf(x.x, -126.0f);
  x.y = fmaxf(x.y, -126.0f);
  const float2 t = __fadd2_rn(x, make_float2(12582912.0f, 12582912.0f));
  const float2 tf = __fadd2_rn(t, make_float2(-12582912.0f, -12582912.0f));
  const float2 f = __ffma2_rn(tf, make_float2(-1.0f, -1.0f), x);
  float2 p = __ffma2_rn(make_float2(0.0550129f, 0.0550129f), f, make_float2(0.24221165f, 0.24221165f));
  p = __ffma2_rn(p, f, make_float2(0.69328244f, 0.69328244f));
  p = __ffma2_rn(p, f, make_float2(1.0f, 1.0f));
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)));
}

__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void named_bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
// tcgen05.st: thread i of the warp writes TMEM lane (base_lane + i), 16 / 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] . B[smem]: A = 128 lanes (rows) x 8 columns of packed bf16 pairs per K = 16 step
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

#ifdef MHA2_TRACE    // timing experiment only (tools/trace_mha2.py): CTA 0 logs clock64 per (role, step, event) to global memory
__device__ unsigned int* g_trace2 = nullptr;
#define TRACE2(role, ev, step)                                                                                   \
  do {                                                                                                           \
    if (blockIdx.x == 0 && g_trace2 != nullptr && (step) < 64)                                                    \
      g_trace2[((role) * 64 + (step)) * 8 + (ev)] = (unsigned int)clock64();                                     \
  } while (0)
#else
#define TRACE2(role, ev, step) do {} while (0)
#endif
// roles: 0 / 1 MMA issuer of group 0 / 1 (0 S-start, 1 operands ready, 2 S issued, 3 PV-start, 4 p_full passed,
// 5 PV issued), 2 / 3 softmax warp 0 of group 0 / 1 (0 start, 1 s_full passed, 2 scores in registers, 3 maximum done,
// 4 exponentials + P stores issued, 5 p_full arrived), 4 epilogue (per item: 0 start, 1 l_full, 2 o_full, 3 O read,
// 4 staged, 5 store issued)

struct Item {
  int b, h, q0, n_keys, n_kt;
  bool active1;      // the second query tile of the block exists
};

// The work items of a CTA are decoded once, in parallel, at kernel start (item_cache): every role then reads 16 bytes of
// shared memory per item instead of running four integer divisions on its critical path (the clock trace showed a
// ~1000 clk bubble in the softmax warps at every item boundary).  CTAs with more than kLenCache items decode the rest here.
__device__ __forceinline__ Item decode_item(int item, int ordinal, const int4* item_cache, int n_qblk, int n_head,
                                            int seq_len, const int* __restrict__ kv_len) {
  Item it;
  if (ordinal < kLenCache) {
    const int4 c = item_cache[ordinal];
    it.b = c.x; it.h = c.y; it.q0 = c.z; it.n_keys = c.w;
  } else {
    const int qb = item % n_qblk;
    const int bh = item / n_qblk;
    it.h = bh % n_head;
    it.b = bh / n_head;
    it.q0 = qb * 2 * kQTile;
    it.n_keys = min(max(__ldg(kv_len + it.b), 1), seq_len);
  }
  it.n_kt = (it.n_keys + kKTile - 1) / kKTile;
  it.active1 = it.q0 + kQTile < seq_len;
  return it;
}

__global__ void __launch_bounds__(kThreads, 1)
mha2_bf16_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                 const __grid_constant__ CUtensorMap tmap_ctx, const int* __restrict__ kv_len, int seq_len,
                 int d_model, int n_head, int n_qblk, int n_items) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bars = sbase + kOffBar;
  auto q_full = [&](int buf, int w) { return bars + 8u * (buf * 2 + w); };
  auto q_empty = [&](int buf, int w) { return bars + 8u * (4 + buf * 2 + w); };
  auto kv_full = [&](int s) { return bars + 8u * (8 + s); };
  auto kv_empty = [&](int s) { return bars + 8u * (8 + kKvStages + s); };
  // Per group (9 each).  s_full / p_full / pv_done exist per score buffer sb = step & 1; use number (step >> 1) of a buffer
  // has parity (step >> 1) & 1.  Per-buffer barriers cannot run ahead of their waiter: S(g + 2) is only issued behind
  // P.V(g), i.e. after the issuer has passed p_full(g), i.e. after the softmax group has consumed s_full(g).
  // o_full / l_full / o_free: one phase per work item in which the group has a query tile (parity = use & 1).
  const uint32_t gb = bars + 8u * (8 + 2 * kKvStages);
  auto s_full = [&](int w, int sb) { return gb + 8u * (w * 9 + sb); };        // MMA commit: S(step) is in TMEM
  auto p_full = [&](int w, int sb) { return gb + 8u * (w * 9 + 2 + sb); };    // 4 softmax warps: P(step) is in TMEM
  auto pv_done = [&](int w, int sb) { return gb + 8u * (w * 9 + 4 + sb); };   // MMA commit: P.V(step) has retired
  auto o_full = [&](int w) { return gb + 8u * (w * 9 + 6); };                 // MMA commit: last P.V of the item retired
  auto l_full = [&](int w) { return gb + 8u * (w * 9 + 7); };                 // 4 softmax warps: 1 / l is in smem
  auto o_free = [&](int w) { return gb + 8u * (w * 9 + 8); };                 // 4 epilogue warps: O and 1 / l were read
  const uint32_t tmem_slot = bars + 8u * kNumBars;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int4* len_cache = reinterpret_cast<int4*>(sptr + kOffLen);
  for (int n = tid; n < kLenCache; n += kThreads) {
    const long long item = (long long)blockIdx.x + (long long)n * gridDim.x;
    if (item < n_items) {
      const int qb = (int)(item % n_qblk), bh = (int)(item / n_qblk);
      const int b = bh / n_head;
      len_cache[n] = make_int4(b, bh % n_head, qb * 2 * kQTile, min(max(__ldg(kv_len + b), 1), seq_len));
    }
  }
  if (tid == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_kv);
    prefetch_tmap(&tmap_ctx);
    for (int i = 0; i < 4; ++i) { mbar_init(q_full(i >> 1, i & 1), 1); mbar_init(q_empty(i >> 1, i & 1), 1); }
    for (int s = 0; s < kKvStages; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 2); }
    for (int w = 0; w < 2; ++w) {
      for (int sb = 0; sb < 2; ++sb) { mbar_init(s_full(w, sb), 1); mbar_init(p_full(w, sb), kSoftWarps / 2); mbar_init(pv_done(w, sb), 1); }
      mbar_init(o_full(w), 1);
      mbar_init(l_full(w), kSoftWarps / 2);
      mbar_init(o_free(w), 4);
    }
    fence_barrier_init();
  }
  if (warp == kIssuer0) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // (setmaxnreg at the top of each role's branch: ptxas derives the register limit of a region from the setmaxnreg that
  // dominates it; the two 80-register warpgroups release before the softmax warpgroups can grow)
  if (warp >= kIssuer0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsOther));
    if (warp == kProducer) {
      // ============================ TMA producer ============================
      if (lane == 0) {
        int stage = 0;
        uint32_t kv_phase = 0;
        int n_done = 0;
        uint32_t q_par = 0;                         // bit (buf * 2 + w): parity of the fills of that Q buffer
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_done) {
          const Item it = decode_item(item, n_done, len_cache, n_qblk, n_head, seq_len, kv_len);
          const int buf = n_done & 1;
          const int row_base = it.b * seq_len;
          for (int w = 0; w < 2; ++w) {
            if (w == 1 && !it.active1) continue;
            const uint32_t qph = (q_par >> (buf * 2 + w)) & 1;
            q_par ^= 1u << (buf * 2 + w);
            mbar_wait(q_empty(buf, w), qph ^ 1);
            mbar_arrive_expect_tx(q_full(buf, w), kQTile * kHd * 2);
            tma_load_2d(sbase + kOffQ + (buf * 2 + w) * 16384, &tmap_q, q_full(buf, w), it.h * kHd,
                        row_base + it.q0 + w * kQTile);
          }
          for (int j = 0; j < it.n_kt; ++j) {
            mbar_wait(kv_empty(stage), kv_phase ^ 1);
            const uint32_t kdst = sbase + kOffKV + stage * kStageBytes;
            mbar_arrive_expect_tx(kv_full(stage), kStageBytes);
            // rows past the utterance (or past the batch: zero fill) carry finite values and meet P = 0
            tma_load_2d(kdst, &tmap_kv, kv_full(stage), d_model + it.h * kHd, row_base + j * kKTile);
            tma_load_2d(kdst + kKBytes, &tmap_kv, kv_full(stage), 2 * d_model + it.h * kHd, row_base + j * kKTile);
            if (++stage == kKvStages) { stage = 0; kv_phase ^= 1; }
          }
        }
      }
      __syncwarp();
    } else if (warp < kProducer) {
      // ============================ MMA issuers: warp 12 -> group 0, warp 13 -> group 1 ============================
      // Order per group: S(0) S(1) | P.V(0) S(2) | P.V(1) S(3) | ...  S(g + 2) lands in the buffer P(g) lives in, so it
      // is issued behind P.V(g): the tcgen05.mma of one thread execute in issue order, which is what keeps S(g + 2) from
      // overwriting P(g) before P.V(g) has read it.  p_full(g) also says that every softmax thread has pulled S(g) out of
      // TMEM.
      auto issuer = [&](auto group) {
        constexpr int w = decltype(group)::value;
        constexpr uint32_t idesc_s = make_idesc_bf16(128, kKTile);
        constexpr uint32_t idesc_o = make_idesc_bf16(128, kHd) | (1u << 16);     // V: MN-major B operand
        struct Cursor {
          int item, j, n_done;      // work item, key tile inside it, ordinal of the item on this CTA
          int stage;                // K/V ring stage of the flattened key-tile sequence of this CTA
          uint32_t phase;
          int n_kt;
          bool valid;
          bool virt;                // group 1 has no query tile in this item: its stages are walked, not used
        };
        auto load_item = [&](Cursor& c) {
          c.valid = c.item < n_items;
          c.virt = false;
          if (c.valid) {
            const Item it = decode_item(c.item, c.n_done, len_cache, n_qblk, n_head, seq_len, kv_len);
            c.n_kt = it.n_kt;
            c.virt = w == 1 && !it.active1;
          }
        };
        auto advance = [&](Cursor& c) {
          if (++c.stage == kKvStages) { c.stage = 0; c.phase ^= 1; }
          if (++c.j == c.n_kt) {
            c.j = 0;
            c.item += gridDim.x;
            ++c.n_done;
            load_item(c);
          }
        };
        Cursor sc;
        sc.item = blockIdx.x; sc.j = 0; sc.n_done = 0; sc.stage = 0; sc.phase = 0; sc.n_kt = 1;
        load_item(sc);
        Cursor pc = sc;
        int g_s = 0, g_p = 0;                 // per-group step indices of the next S and the next P.V
        uint32_t q_fill0 = 0, q_fill1 = 0;    // consumed fills of Q buffers 0 / 1 of this group
        uint32_t uses = 0;                    // items of this group whose P.V sequence has finished being issued
        const uint32_t q_base = sbase + kOffQ + w * 16384;
        const uint32_t o_tmem = tmem_base + kOCol + w * kHd;
        while (pc.valid) {
          if (w == 1 && pc.virt) {
            // an item without a query tile for this group: every K/V stage of it is still observed and handed back in
            // order (a consumer that jumps over uses of a parity-tracked mbarrier can alias a pending phase).
            // The S cursor never passes a virtual item, so it stands at the start of this one.
            const int n = pc.n_kt;
            for (int k = 0; k < n; ++k) {
              mbar_wait(kv_full(pc.stage), pc.phase);
              if (lane == 0) mbar_arrive(kv_empty(pc.stage));
              __syncwarp();
              advance(pc);
            }
            sc = pc;
            continue;
          }
          // ---- S(g_s) = Q K^T, up to two tiles ahead of the P.V sequence ----
          // (g_s == g_p: nothing else can make progress, wait for the tile; otherwise take it only if it has landed,
          // so that a late K/V tile never delays a P.V whose probabilities are ready)
          if (sc.valid && !(w == 1 && sc.virt) && g_s < g_p + 2 &&
              (g_s == g_p || mbar_test_wait(kv_full(sc.stage), sc.phase))) {
            const int buf = sc.n_done & 1;
            const int sb = g_s & 1;
            if (lane == 0) TRACE2(w, 0, g_s);
            mbar_wait(kv_full(sc.stage), sc.phase);
            if (sc.j == 0) mbar_wait(q_full(buf, w), (buf ? q_fill1 : q_fill0) & 1);
            if (lane == 0) TRACE2(w, 1, g_s);
            tc_fence_after();
            const bool last_of_item = sc.j == sc.n_kt - 1;
            if (elect_one()) {
              const uint32_t s_tmem = tmem_base + (w * 2 + sb) * kSCols;
              const uint64_t qd = make_smem_desc_sw128(q_base + buf * 32768);
              const uint64_t kd = make_smem_desc_sw128(sbase + kOffKV + sc.stage * kStageBytes);
#pragma unroll
              for (int k = 0; k < kHd / 16; ++k) umma_bf16(s_tmem, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
              umma_commit(s_full(w, sb));
              if (last_of_item) umma_commit(q_empty(buf, w));
            }
            __syncwarp();
            if (last_of_item) { if (buf) ++q_fill1; else ++q_fill0; }
            if (lane == 0) TRACE2(w, 2, g_s);
            ++g_s;
            advance(sc);
            continue;
          }
          // ---- O += P(g_p) V ----
          {
            const int sb = g_p & 1;
            if (lane == 0) TRACE2(w, 3, g_p);
            mbar_wait(p_full(w, sb), (uint32_t)(g_p >> 1) & 1);
            if (pc.j == 0) mbar_wait(o_free(w), (uses & 1) ^ 1);      // the epilogue has drained the previous item's O
            if (lane == 0) TRACE2(w, 4, g_p);
            tc_fence_after();
            const bool last_of_item = pc.j == pc.n_kt - 1;
            if (elect_one()) {
              const uint32_t p_tmem = tmem_base + (w * 2 + sb) * kSCols;
              const uint64_t vd = make_smem_desc_sw128(sbase + kOffKV + pc.stage * kStageBytes + kKBytes);
#pragma unroll
              for (int k = 0; k < kKTile / 16; ++k) {
                // A: 16 keys = 8 TMEM columns of bf16 pairs; B: 16 keys = 16 rows of 128 bytes of the MN-major V tile
                umma_bf16_ts(o_tmem, p_tmem + 8 * k, vd + 128 * k, idesc_o, (k | pc.j) != 0);
              }
              umma_commit(kv_empty(pc.stage));                 // the second arrival comes from the other group's issuer
              umma_commit(pv_done(w, sb));
              if (last_of_item) umma_commit(o_full(w));
            }
            __syncwarp();
            if (last_of_item) ++uses;
            if (lane == 0) TRACE2(w, 5, g_p);
            ++g_p;
            advance(pc);
          }
        }
      };
      if (warp == kIssuer0) issuer(std::integral_constant<int, 0>{});
      else issuer(std::integral_constant<int, 1>{});
      __syncwarp();
    }
  } else if (warp >= kEpiWarp0) {
    // ============================ epilogue warpgroup ============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsOther));
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int sw = r & 7;
    const float* xch = reinterpret_cast<const float*>(sptr + kOffX);
    uint32_t uses[2] = {0, 0};
    int ordinal = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ordinal) {
      const Item it = decode_item(item, ordinal, len_cache, n_qblk, n_head, seq_len, kv_len);
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        if (w == 1 && !it.active1) continue;
        const uint32_t ph = uses[w] & 1;
        ++uses[w];
        const bool tr = warp == kEpiWarp0 && lane == 0;
        if (tr) TRACE2(4, 0, ordinal * 2 + w);
        mbar_wait(l_full(w), ph);
        if (tr) TRACE2(4, 1, ordinal * 2 + w);
#ifdef MHA2_SPLIT
        const float inv = 1.0f / (xch[(w * 2) * 128 + r] + xch[(w * 2 + 1) * 128 + r]);     // the two halves' partial sums
#else
        const float inv = xch[w * 128 + r];
#endif
        mbar_wait(o_full(w), ph);
        if (tr) TRACE2(4, 2, ordinal * 2 + w);
        tc_fence_after();
        // the previous TMA store out of this group's staging tile must have read it
        if (warp == kEpiWarp0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        named_bar_sync(1, 128);
        const uint32_t srow = sbase + kOffOut + w * 16384 + r * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t o[32];
          tmem_ld32(tmem_base + kOCol + w * kHd + half * 32 + lane_off, o);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t a0 = pack_bf16x2(__uint_as_float(o[8 * c]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
            const uint32_t a1 = pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
            const uint32_t a2 = pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
            const uint32_t a3 = pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                         ::"r"(srow + (((half * 4 + c) ^ sw) << 4)), "r"(a0), "r"(a1), "r"(a2), "r"(a3) : "memory");
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free(w));         // O and the 1 / l slot may be reused
        if (tr) TRACE2(4, 3, ordinal * 2 + w);
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (tr) TRACE2(4, 4, ordinal * 2 + w);
        if (warp == kEpiWarp0 && lane == 0) {
          // rows past the end of the utterance are clipped by the 3-D map
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmap_ctx)), "r"(sbase + kOffOut + w * 16384),
                         "r"(it.h * kHd), "r"(it.q0 + w * kQTile), "r"(it.b) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (tr) TRACE2(4, 5, ordinal * 2 + w);
      }
    }
    if (warp == kEpiWarp0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
#ifdef MHA2_SPLIT
    // ============================ softmax groups, two threads per query row ============================
    // warps 0-3 / 4-7: group 0, keys 0-47 / 48-95 of every tile; warps 8-11 / 12-15: group 1.  warp & 3 = TMEM lane quarter,
    // so the two warps of a row sit on the same scheduler.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsSoftmax));
    const int w = warp >> 3;                       // group / query tile
    const int half = (warp >> 2) & 1;              // which 48 keys of a tile
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;             // row inside the tile = TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t o_tmem = tmem_base + kOCol + w * kHd + half * 32 + lane_off;    // my 32 of the row's 64 O columns
    float* xch = reinterpret_cast<float*>(sptr + kOffX);
    float* xmax = reinterpret_cast<float*>(sptr + kOffMax);
    const int pair_bar = 2 + w * 4 + quarter;      // named barrier of the two warps that share these 32 rows
    constexpr int kHalf = kKTile / 2;              // 48 keys
    uint32_t g = 0;                                // per-group step index
    uint32_t uses = 0;
    int ordinal = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ordinal) {
      const Item it = decode_item(item, ordinal, len_cache, n_qblk, n_head, seq_len, kv_len);
      if (w == 1 && !it.active1) continue;
      constexpr float kRaise = 40.0f;              // see the one-thread-per-row form below
      float m_ref = -INFINITY, l_run = 0.f;
      for (int j = 0; j < it.n_kt; ++j, ++g) {
        const bool tr = (warp & 7) == 0 && lane == 0;
        const int sb = g & 1;
        const uint32_t s_tmem = tmem_base + (w * 2 + sb) * kSCols + lane_off;
        if (tr) TRACE2(2 + w, 0, g);
        mbar_wait(s_full(w, sb), (g >> 1) & 1);
        if (tr) TRACE2(2 + w, 1, g);
        tc_fence_after();
        uint32_t v[kHalf / 16][16];
#pragma unroll
        for (int c = 0; c < kHalf / 16; ++c) tmem_ld16(s_tmem + half * kHalf + c * 16, v[c]);
        tmem_ld_wait();
        if (tr) TRACE2(2 + w, 2, g);
        const int valid = it.n_keys - j * kKTile - half * kHalf;      // my columns < valid are real keys
        if (valid < kHalf) {
#pragma unroll
          for (int c = 0; c < kHalf; ++c)
            if (c >= valid) v[c >> 4][c & 15] = 0xff800000u;         // -inf: exp2 gives exactly 0
        }
        float tm0 = -INFINITY, tm1 = -INFINITY;
#pragma unroll
        for (int c = 0; c < kHalf; c += 4) {
          tm0 = max3(tm0, __uint_as_float(v[c >> 4][c & 15]), __uint_as_float(v[c >> 4][(c & 15) + 1]));
          tm1 = max3(tm1, __uint_as_float(v[c >> 4][(c & 15) + 2]), __uint_as_float(v[c >> 4][(c & 15) + 3]));
        }
        // row maximum of the tile = max over the two halves: exchanged through shared memory (slot = step parity: the
        // partner cannot be two steps ahead, it needs this warp at the barrier of every step)
        float* slot = xmax + ((sb * 2 + w) * 2) * 128;
        slot[half * 128 + r] = fmaxf(tm0, tm1);
        named_bar_sync(pair_bar, 64);
        const float tile_max = fmaxf(fmaxf(tm0, tm1), slot[(half ^ 1) * 128 + r]);
        if (j == 0) {
          m_ref = tile_max;                            // O is overwritten by the first P.V of the item
        } else {
          const bool raise = (tile_max - m_ref) * kLog2e > kRaise;     // same value in both threads of the row
          if (__any_sync(0xffffffffu, raise)) {
            // (rare; see the one-thread-per-row form for why this occasional wait cannot alias)  Each half rescales
            // its own 32 O columns; both apply the same factor to their partial row sum.
            mbar_wait(pv_done(w, sb ^ 1), ((g - 1) >> 1) & 1);
            tc_fence_after();
            const float factor = raise ? ex2_mufu((m_ref - tile_max) * kLog2e) : 1.0f;
            if (raise) m_ref = tile_max;
            l_run *= factor;
#pragma unroll 1
            for (int c = 0; c < kHd / 2; c += 8) {
              uint32_t o[8];
              tmem_ld8(o_tmem + c, o);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * factor);
              tmem_st8(o_tmem + c, o);
            }
          }
        }
        if (tr) TRACE2(2 + w, 3, g);
        const float m_scaled = m_ref * kLog2e;
        float2 l0 = make_float2(0.f, 0.f), l1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < kHalf / 16; ++c) {
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[c][2 * e]), __uint_as_float(v[c][2 * e + 1])),
                                        make_float2(kLog2e, kLog2e), make_float2(-m_scaled, -m_scaled));
            const float p0 = ex2_mufu(x.x), p1 = ex2_mufu(x.y);
            if (e & 1) l1 = __fadd2_rn(l1, make_float2(p0, p1));
            else l0 = __fadd2_rn(l0, make_float2(p0, p1));
            pk[e] = pack_bf16x2(p0, p1);
          }
          // P column k holds keys 2k, 2k + 1: my 48 keys are P columns half * 24 + [0, 24)
          tmem_st8(s_tmem + half * (kHalf / 2) + c * 8, pk);
        }
        l_run += (l0.x + l0.y) + (l1.x + l1.y);
        if (tr && l_run != 123.f) TRACE2(2 + w, 4, g);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(w, sb));
        if (tr) TRACE2(2 + w, 5, g);
      }
      // end of the item: hand my partial row sum to the epilogue warpgroup (it adds the two halves)
      mbar_wait(o_free(w), (uses & 1) ^ 1);
      xch[(w * 2 + half) * 128 + r] = l_run;
      __syncwarp();
      if (lane == 0) mbar_arrive(l_full(w));
      ++uses;
    }
#else
    // ============================ softmax groups (thread = query row) ============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsSoftmax));
    const int w = warp >> 2;                       // group / query tile
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;             // row inside the tile = TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t o_tmem = tmem_base + kOCol + w * kHd + lane_off;
    float* xch = reinterpret_cast<float*>(sptr + kOffX);
    uint32_t g = 0;                                // per-group step index
    uint32_t uses = 0;
    int ordinal = 0;
#if MHA2_OFFSET > 0
    // Timing experiment: group 1 starts MHA2_OFFSET clocks late, so that the two groups' exponential phases (which
    // otherwise run in lockstep: same start, same step time) overlap each other's barrier / load / maximum phases.
    if (w == 1) {
      const long long t0 = clock64();
      while (clock64() - t0 < MHA2_OFFSET) {}
    }
#endif
    uint32_t v[kKTile / 32][32];
    bool have_scores = false;                      // (MHA2_EARLY_LD) this step's score load was issued a step early
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ordinal) {
      const Item it = decode_item(item, ordinal, len_cache, n_qblk, n_head, seq_len, kv_len);
      if (w == 1 && !it.active1) continue;
#ifdef MHA2_OFFSET_ITEM
      // Timing experiment: group 1 enters EVERY item MHA2_OFFSET_ITEM clocks late (the groups re-align at item boundaries)
      if (w == 1) {
        const long long t0 = clock64();
        while (clock64() - t0 < MHA2_OFFSET_ITEM) {}
      }
#endif
      // m_ref: the maximum the exponents are taken against.  It is only raised (and O / l rescaled) when the running
      // maximum exceeds it by more than 2^kRaise: P <= 2^kRaise stays far inside bf16 / fp32 range, and O in TMEM is
      // touched by the CUDA cores only on those rare steps.
      constexpr float kRaise = 40.0f;
      float m_ref = -INFINITY, l_run = 0.f;
      for (int j = 0; j < it.n_kt; ++j, ++g) {
        const bool tr = (warp & 3) == 0 && lane == 0;
        const int sb = g & 1;
        const uint32_t s_tmem = tmem_base + (w * 2 + sb) * kSCols + lane_off;
        if (tr) TRACE2(2 + w, 0, g);
        if (!have_scores) {
          mbar_wait(s_full(w, sb), (g >> 1) & 1);
          if (tr) TRACE2(2 + w, 1, g);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < kKTile / 32; ++c) tmem_ld32(s_tmem + c * 32, v[c]);
        }
        tmem_ld_wait();
        have_scores = false;
        if (tr) TRACE2(2 + w, 2, g);
        const int valid = it.n_keys - j * kKTile;      // my columns < valid are real keys
        if (valid < kKTile) {
          // last tile of the utterance only (a real branch: the full tiles must not pay the compare / select pairs)
#pragma unroll
          for (int c = 0; c < kKTile; ++c)
            if (c >= valid) v[c >> 5][c & 31] = 0xff800000u;        // -inf: exp2 gives exactly 0
        }
#if MHA2_MAXCHAINS == 4
        float tm0 = -INFINITY, tm1 = -INFINITY, tm2 = -INFINITY, tm3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < kKTile; c += 8) {
          tm0 = max3(tm0, __uint_as_float(v[c >> 5][c & 31]), __uint_as_float(v[c >> 5][(c & 31) + 1]));
          tm1 = max3(tm1, __uint_as_float(v[c >> 5][(c & 31) + 2]), __uint_as_float(v[c >> 5][(c & 31) + 3]));
          tm2 = max3(tm2, __uint_as_float(v[c >> 5][(c & 31) + 4]), __uint_as_float(v[c >> 5][(c & 31) + 5]));
          tm3 = max3(tm3, __uint_as_float(v[c >> 5][(c & 31) + 6]), __uint_as_float(v[c >> 5][(c & 31) + 7]));
        }
        const float tile_max = fmaxf(fmaxf(tm0, tm1), fmaxf(tm2, tm3));
#else
        float tm0 = -INFINITY, tm1 = -INFINITY;
#pragma unroll
        for (int c = 0; c < kKTile; c += 4) {
          tm0 = max3(tm0, __uint_as_float(v[c >> 5][c & 31]), __uint_as_float(v[c >> 5][(c & 31) + 1]));
          tm1 = max3(tm1, __uint_as_float(v[c >> 5][(c & 31) + 2]), __uint_as_float(v[c >> 5][(c & 31) + 3]));
        }
        const float tile_max = fmaxf(tm0, tm1);
#endif
        if (j == 0) {
          m_ref = tile_max;                            // O is overwritten by the first P.V of the item
        } else {
          const bool raise = (tile_max - m_ref) * kLog2e > kRaise;
          if (__any_sync(0xffffffffu, raise)) {
            // O must hold every earlier step of the item and nothing may be in flight on it: P.V(g - 1) is the last one
            // issued (P.V(g) needs the probabilities this thread has yet to write).  Its barrier is the other buffer's
            // pv_done; completion number (g - 1) >> 1 of it is either the current phase or the one just finished
            // (P.V(g + 1) cannot be issued before this step is over, P.V(g - 3) retired before S(g) did), so this
            // occasional wait cannot alias even though the steps in between never look at the barrier.
            mbar_wait(pv_done(w, sb ^ 1), ((g - 1) >> 1) & 1);
            tc_fence_after();
            const float factor = raise ? ex2_mufu((m_ref - tile_max) * kLog2e) : 1.0f;
            if (raise) m_ref = tile_max;
            l_run *= factor;
            // (rare path: 8 columns at a time so that the live scores are not spilled around it)
#pragma unroll 1
            for (int c = 0; c < kHd; c += 8) {
              uint32_t o[8];
              tmem_ld8(o_tmem + c, o);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * factor);
              tmem_st8(o_tmem + c, o);
            }
          }
        }
        if (tr) TRACE2(2 + w, 3, g);
        const float m_scaled = m_ref * kLog2e;
        float2 l0 = make_float2(0.f, 0.f), l1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < kKTile / 32; ++c) {
          // P columns 16 c .. 16 c + 15 cover score columns that are already in registers (16 c + 15 < 32 (c + 1))
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            // packed fp32x2 arithmetic (FFMA2 / FADD2): half the issue slots of the scalar forms
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[c][2 * e]), __uint_as_float(v[c][2 * e + 1])),
                                        make_float2(kLog2e, kLog2e), make_float2(-m_scaled, -m_scaled));
            // MHA2_POLY of every 8 exponentials (whole pairs) run on the FMA pipe instead of the MUFU, the unit this
            // kernel is bound by (ncu: XU pipe 64 %, tensor 31 %, FMA 19 %, ALU 28 % of their peaks).  The masked last
            // tile of an utterance keeps every exponential on the MUFU: exp2(-inf) must be exactly 0.
            const bool fma_pipe = (e & 3) < MHA2_POLY / 2 && valid >= kKTile;
            float p0, p1;
            if (fma_pipe) {
              const float2 pp = ex2_fma2(x);
              p0 = pp.x; p1 = pp.y;
            } else {
              p0 = ex2_mufu(x.x); p1 = ex2_mufu(x.y);
            }
            if (e & 1) l1 = __fadd2_rn(l1, make_float2(p0, p1));
            else l0 = __fadd2_rn(l0, make_float2(p0, p1));
            pk[e] = pack_bf16x2(p0, p1);
          }
          tmem_st16(s_tmem + c * 16, pk);
        }
        l_run += (l0.x + l0.y) + (l1.x + l1.y);
        if (tr && l_run != 123.f) TRACE2(2 + w, 4, g);
#ifdef MHA2_EARLY_LD
        // The next step's scores (same item) are normally complete by now - S runs two tiles ahead: start their load
        // before waiting for the P stores, so that its latency overlaps the store wait, the fence and the arrive.
        // Never by blocking: the issuer may need this step's p_full before it can issue S(g + 1).
        if (j + 1 < it.n_kt &&
            __all_sync(0xffffffffu, mbar_test_wait(s_full(w, sb ^ 1), ((g + 1) >> 1) & 1))) {
          tc_fence_after();
          const uint32_t s_next = tmem_base + (w * 2 + (sb ^ 1)) * kSCols + lane_off;
#pragma unroll
          for (int c = 0; c < kKTile / 32; ++c) tmem_ld32(s_next + c * 32, v[c]);
          have_scores = true;
        }
#endif
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(w, sb));
        if (tr) TRACE2(2 + w, 5, g);
      }
      // end of the item: hand 1 / l to the epilogue warpgroup.  The slot was last used one item of this group ago;
      // o_free says the epilogue is done with it.
      mbar_wait(o_free(w), (uses & 1) ^ 1);
      xch[w * 128 + r] = 1.0f / l_run;
      __syncwarp();
      if (lane == 0) mbar_arrive(l_full(w));
      ++uses;
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kIssuer0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace

#ifdef MHA2_TRACE
extern "C" int stac_mha2_trace(unsigned int* buf) {
  cudaMemcpyToSymbol(g_trace2, &buf, sizeof(buf));
  return 0;
}
#endif

extern "C" int stac_mha_bf16_v2(const uint16_t* qkv, const int32_t* kv_len, int64_t batch, int64_t seq_len,
                                int64_t d_model, int64_t n_head, uint16_t* ctx, void* stream) {
  STAC_REQUIRE(qkv && kv_len && ctx && batch > 0 && batch < 65536 && seq_len > 0);
  if (d_model != n_head * kHd || n_head > 65535 || batch * seq_len >= (1ll << 31)) return STAC_ERR_UNSUPPORTED_SHAPE;
  const int64_t n_qblk = ceil_div64(seq_len, 2 * kQTile);
  const int64_t n_items = batch * n_head * n_qblk;
  if (n_items >= (1ll << 31)) return STAC_ERR_UNSUPPORTED_SHAPE;
  CUtensorMap tq, tkv, tctx;
  const uint64_t dims[2] = {(uint64_t)(3 * d_model), (uint64_t)(batch * seq_len)};
  const uint64_t str[1] = {(uint64_t)(3 * d_model) * 2};
  {
    const uint32_t box[2] = {kHd, kQTile};
    int r = encode_map(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  {
    const uint32_t box[2] = {kHd, kKTile};
    int r = encode_map(&tkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  {
    // ctx [B][T][d_model]: a tile that runs past the end of its utterance is clipped in the T dimension
    const uint64_t cdims[3] = {(uint64_t)d_model, (uint64_t)seq_len, (uint64_t)batch};
    const uint64_t cstr[2] = {(uint64_t)d_model * 2, (uint64_t)seq_len * d_model * 2};
    const uint32_t cbox[3] = {kHd, kQTile, 1};
    int r = encode_map(&tctx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, ctx, 3, cdims, cstr, cbox);
    if (r != STAC_OK) return r;
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mha2_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  const int grid = (int)std::min<int64_t>(n_items, stac_grid_limit());
  mha2_bf16_kernel<<<grid, kThreads, kSmemBytes, as_stream(stream)>>>(tq, tkv, tctx, kv_len, (int)seq_len,
                                                                    (int)d_model, (int)n_head, (int)n_qblk,
                                                                    (int)n_items);
  STAC_LAUNCH_CHECK();
}
