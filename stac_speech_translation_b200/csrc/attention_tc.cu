// Flash-style multi-head self-attention on tcgen05 tensor cores (bf16 operands, fp32 softmax).
// Key-padding is by per-utterance valid length (no dense mask tensor, no B*H*T*T score matrix).
//
// Reference behaviour replaced: torch.nn.MultiheadAttention slow path (baddbmm + softmax + bmm with a
// -inf key-padding mask) reached through SpeechBrain's TransformerEncoderLayer from
//   /root/reference/stac-st/modules/TransformerMultiTask.py:304-308 (mask built at :289-294 / :225-226).
//
// head_dim 64 makes this kernel MUFU-bound (one ex2 per score against 4 MMA flops per score per 64-wide d), so the
// design goal is to keep the exponential pipe of every SM sub-partition busy, not the tensor core:
//   * persistent, warp-specialised, one CTA per SM, 608 threads:
//       warps 0-7   softmax group 0 (query tile 0: 128 rows; warp & 3 = 32-row quarter, (warp >> 2) & 1 = which 32 of
//                   the 64 key columns: two threads per row, so every SM sub-partition holds FOUR softmax warps whose
//                   max / exp / pack / store phases interleave and keep its exp unit fed)
//       warps 8-15  softmax group 1 (query tile 1)
//       warps 16-17 MMA issuers (one thread per softmax group)
//       warp  18    TMA producer (Q tiles double-buffered per work item, K / V tiles in a 5-stage ring)
//       warp  19    output: one TMA store per (item, group) of the normalised O tile, staged in the item's dead Q
//                   buffer (row-per-thread global stores cost ~2800 cycles per item on the softmax critical path)
//   * work item = (utterance, head, block of 256 queries); the two query tiles share every K / V tile;
//   * key tiles are 64 wide and every group owns TWO score buffers in TMEM: S(j+1) = Q K_{j+1}^T is issued as soon
//     as the softmax warps have pulled S(j-1) into registers, i.e. it is already there when they finish tile j, and
//     P is double-buffered in shared memory the same way, so neither MMA latency sits on the softmax critical path;
//   * O accumulates in TMEM over the key tiles (P.V with accumulate), the running-max correction is lazy (only
//     when the maximum grows by more than 2^8), so the CUDA cores touch O on those rare steps and once per item;
//   * V is consumed straight from the packed QKV projection as an MN-major B operand (the same [keys][64] tile
//     shape TMA delivers for K), so no transposed copy of V is ever written.  A K-major V^T tensor is still
//     accepted (v_t != NULL) for callers that have one.
#include <algorithm>
#include <type_traits>
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int kHd = 64, kQTile = 128, kKTile = 64;
constexpr int kKvStages = 5;
constexpr int kThreads = 640;
// shared memory map (bytes, from a 1024-aligned base)
constexpr int kOffQ = 0;                         // [2 bufs][2 groups] x 16 KB
constexpr int kOffP = 65536;                     // [2 groups][2 bufs] x 16 KB (128 rows x 64 keys bf16, K-major)
constexpr int kOffKV = 131072;                   // [stages] x (K 8 KB + V 8 KB)
constexpr int kOffX = kOffKV + kKvStages * 16384;      // float [2 groups][2 bufs][2 halves][128 rows]: row-max exchange
constexpr int kOffLen = kOffX + 2 * 2 * 2 * 128 * 4;   // int [kLenCache]
constexpr int kOffBar = kOffLen + 128 * 4;
// One o_staged barrier per (Q buffer, group), not one per group: tools/model_check_mha1.py shows that with one barrier
// per group a softmax group that gets two short work items ahead of the store warp completes TWO phases of it before the
// store warp looks, the parity wait aliases, and the kernel deadlocks (mbarrier time-out trap); a per-buffer barrier
// cannot run ahead, because the buffer itself is only refilled after the store warp has released it.  Verified on a
// B200 (round 2: parity suite + the many-short-items stress shape, 86.1 us vs 87.4 us).  DESIGN.md section 9.
constexpr int kNumOStaged = 4;
constexpr int kNumBars = 8 + 2 * kKvStages + 18 + kNumOStaged;
#ifdef MHA_TRACE
constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16 + 1024 + 5 * 64 * 8 * 4;
#else
constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16 + 1024;
#endif
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
#ifdef MHA_NOEXP      // timing experiment only (tools/bench_mha.py): results are wrong
  return x;
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}

// tcgen05.st: 32 lanes x 32 consecutive fp32 columns (thread i writes TMEM lane base_lane + i)
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

// 3-input max (one instruction on sm_100)
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void named_bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}

#ifdef MHA_TRACE     // timing experiment only (tools/trace_mha.py): CTA 0 logs clock32 per (role, step, event) in smem
__device__ unsigned int* g_trace = nullptr;
#define TRACE(role, ev, step)                                                                        \
  do {                                                                                               \
    if (blockIdx.x == 0 && (step) < 64) trace_s[((role) * 64 + (step)) * 8 + (ev)] = (unsigned int)clock64(); \
  } while (0)
#else
#define TRACE(role, ev, step) do {} while (0)
#endif

struct Item {
  int b, h, q0, n_keys, n_kt;
  bool active1;      // the second query tile of the block exists
};

constexpr int kLenCache = 128;     // key counts of the first work items of a CTA, staged in smem at kernel start

__device__ __forceinline__ Item decode_item(int item, int ordinal, const int* len_cache, int n_qblk, int n_head,
                                            int seq_len, const int* __restrict__ kv_len) {
  Item it;
  const int qb = item % n_qblk;
  const int bh = item / n_qblk;
  it.h = bh % n_head;
  it.b = bh / n_head;
  it.q0 = qb * 2 * kQTile;
  // (a global load here would sit on the critical path of every role at every item boundary)
  it.n_keys = min(max(ordinal < kLenCache ? len_cache[ordinal] : __ldg(kv_len + it.b), 1), seq_len);
  it.n_kt = (it.n_keys + kKTile - 1) / kKTile;
  it.active1 = it.q0 + kQTile < seq_len;
  return it;
}

template <bool kVT>
__global__ void __launch_bounds__(kThreads, 1)
mha_bf16_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                const __grid_constant__ CUtensorMap tmap_vt, const __grid_constant__ CUtensorMap tmap_ctx,
                const int* __restrict__ kv_len, int seq_len, int d_model, int n_head, int n_qblk, int n_items) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bars = sbase + kOffBar;
  auto q_full = [&](int buf, int w) { return bars + 8u * (buf * 2 + w); };
  auto q_empty = [&](int buf, int w) { return bars + 8u * (4 + buf * 2 + w); };
  auto kv_full = [&](int s) { return bars + 8u * (8 + s); };
  auto kv_empty = [&](int s) { return bars + 8u * (8 + kKvStages + s); };
  // per-group barriers (9 each), buffer i = step & 1.  p_free(w, i) completes when P.V of a step has retired: it
  // frees the P buffer AND tells the softmax group that O holds that step (parity = (step >> 1) & 1).
  const uint32_t gb = bars + 8u * (8 + 2 * kKvStages);
  auto s_full = [&](int w, int i) { return gb + 8u * (w * 9 + i); };
  auto s_free = [&](int w, int i) { return gb + 8u * (w * 9 + 2 + i); };
  auto p_full = [&](int w, int i) { return gb + 8u * (w * 9 + 4 + i); };
  auto p_free = [&](int w, int i) { return gb + 8u * (w * 9 + 6 + i); };
  auto o_free = [&](int w) { return gb + 8u * (w * 9 + 8); };
  // the group's O tile is staged in smem (in Q buffer `buf`) for the store warp
  auto o_staged = [&](int buf, int w) { return gb + 8u * (18 + buf * 2 + w); };
  const uint32_t tmem_slot = bars + 8u * kNumBars;
#ifdef MHA_TRACE
  unsigned int* trace_s = reinterpret_cast<unsigned int*>(sptr + kOffBar + kNumBars * 8 + 16);
#endif

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int* len_cache = reinterpret_cast<int*>(sptr + kOffLen);
  for (int n = tid; n < kLenCache; n += kThreads) {
    const long long item = (long long)blockIdx.x + (long long)n * gridDim.x;
    if (item < n_items) len_cache[n] = __ldg(kv_len + (int)(item / n_qblk) / n_head);
  }

  if (tid == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_kv);
    if (kVT) prefetch_tmap(&tmap_vt);
    prefetch_tmap(&tmap_ctx);
    // q_empty: the last S of the item has retired (commit) AND the item's output tile, staged in the same buffer, has
    // been read by its TMA store (store warp)
    for (int i = 0; i < 4; ++i) { mbar_init(q_full(i >> 1, i & 1), 1); mbar_init(q_empty(i >> 1, i & 1), 2); }
    for (int s = 0; s < kKvStages; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 2); }
    for (int w = 0; w < 2; ++w) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(s_full(w, i), 1); mbar_init(s_free(w, i), 8); mbar_init(p_full(w, i), 8); mbar_init(p_free(w, i), 1);
      }
      mbar_init(o_free(w), 8);
      for (int buf = 0; buf < kNumOStaged / 2; ++buf) mbar_init(o_staged(buf, w), 8);
    }
    fence_barrier_init();
  }
  if (warp == 16) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 18) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t kv_phase = 0;
      int n_done = 0;
      uint32_t q_uses[2][2] = {{0, 0}, {0, 0}};   // fills of Q buffer (buf, w): parity source
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_done) {
        const Item it = decode_item(item, n_done, len_cache, n_qblk, n_head, seq_len, kv_len);
        const int buf = n_done & 1;
        const int row_base = it.b * seq_len;
        for (int w = 0; w < 2; ++w) {
          if (w == 1 && !it.active1) continue;
          const uint32_t qph = (q_uses[buf][w]++) & 1;
          mbar_wait(q_empty(buf, w), qph ^ 1);
          mbar_arrive_expect_tx(q_full(buf, w), kQTile * kHd * 2);
          tma_load_2d(sbase + kOffQ + (buf * 2 + w) * 16384, &tmap_q, q_full(buf, w), it.h * kHd,
                      row_base + it.q0 + w * kQTile);
        }
        for (int j = 0; j < it.n_kt; ++j) {
          mbar_wait(kv_empty(stage), kv_phase ^ 1);
          TRACE(0, 0, (n_done * it.n_kt + j));
          const uint32_t kdst = sbase + kOffKV + stage * 16384;
          mbar_arrive_expect_tx(kv_full(stage), 16384);
          tma_load_2d(kdst, &tmap_kv, kv_full(stage), d_model + it.h * kHd, row_base + j * kKTile);
          if (kVT) tma_load_3d(kdst + 8192, &tmap_vt, kv_full(stage), j * kKTile, 0, it.b * n_head + it.h);
          else tma_load_2d(kdst + 8192, &tmap_kv, kv_full(stage), 2 * d_model + it.h * kHd, row_base + j * kKTile);
          if (++stage == kKvStages) { stage = 0; kv_phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 19) {
    // ============================ output store warp ============================
    uint32_t st_par = 0;      // bit (barrier index): parity of the uses of that o_staged barrier
    int n_done = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_done) {
      const Item it = decode_item(item, n_done, len_cache, n_qblk, n_head, seq_len, kv_len);
      const int buf = n_done & 1;
      for (int w = 0; w < 2; ++w) {
        if (w == 1 && !it.active1) continue;
        const int bit = buf * 2 + w;
        mbar_wait(o_staged(buf, w), (st_par >> bit) & 1);
        st_par ^= 1u << bit;
        if (elect_one()) {
          // rows past the end of the utterance are clipped by the 3-D map
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmap_ctx)), "r"(sbase + kOffQ + (buf * 2 + w) * 16384),
                         "r"(it.h * kHd), "r"(it.q0 + w * kQTile), "r"(it.b) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive(q_empty(buf, w));               // the Q buffer may be refilled
        }
        __syncwarp();
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (warp >= 16) {
    // ============================ MMA issuers: warp 16 -> group 0, warp 17 -> group 1 ============================
    // Per group the order is fixed: S(0) S(1) | P.V(0) S(2) | P.V(1) S(3) | ...  (the scores run two tiles ahead of
    // the softmax; s_free(g) always arrives before p_full(g), so blocking waits in this order never stall a ready
    // operation).  The two groups are independent instruction streams on different sub-partitions.
    // (the group index is a compile-time constant of the issuer body: group 0 carries none of the idle-item logic)
    auto issuer = [&](auto group) {
      // every lane runs the (warp-uniform) control flow; one elected lane issues the tcgen05 instructions, which lets
      // ptxas keep descriptors in uniform registers instead of wrapping each MMA in a broadcast loop
      constexpr int w = decltype(group)::value;
      constexpr uint32_t idesc_s = make_idesc_bf16(128, kKTile);
      // P.V: N = head_dim; V straight from the QKV projection is an MN-major B operand (bit 16)
      constexpr uint32_t idesc_o = make_idesc_bf16(128, kHd) | (kVT ? 0u : (1u << 16));
      struct Cursor {
        int item, j, n_done;      // work item, key tile inside it, ordinal of the item on this CTA
        int stage;                // K/V ring stage of the flattened key-tile sequence of this CTA
        uint32_t phase;
        int n_kt;
        bool active1, valid;
        bool virt;                // group 1 has no query tile in this item: its stages are walked, not used
      };
      auto load_item = [&](Cursor& c) {
        c.valid = c.item < n_items;
        c.virt = false;
        if (c.valid) { const Item it = decode_item(c.item, c.n_done, len_cache, n_qblk, n_head, seq_len, kv_len); c.n_kt = it.n_kt; c.active1 = it.active1; c.virt = w == 1 && !it.active1; }
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
      sc.item = blockIdx.x; sc.j = 0; sc.n_done = 0; sc.stage = 0; sc.phase = 0;
      load_item(sc);
      Cursor pc = sc;
      // Group 1 has no query tile in some items (last block of an utterance with <= 128 rows left).  Its issuer still
      // walks every K/V stage of those items as virtual steps - it waits for the fill and hands the stage back - instead
      // of jumping over them: a warp that skips uses of a parity-tracked mbarrier can later test it while an EARLIER use
      // is still pending, and the parity aliases to "complete".
      // The S cursor stops in front of such an item; when the P.V cursor has caught up, both walk it together.
      int g_s = 0, g_p = 0;                 // per-group step indices of the next S and the next P.V
      uint32_t q_fill0 = 0, q_fill1 = 0;    // consumed fills of Q buffers 0 / 1 of this group
      uint32_t items_started = 0;
      const uint32_t q_base = sbase + kOffQ + w * 16384;
      const uint32_t s_tmem = tmem_base + w * 2 * kKTile;
      const uint32_t o_tmem = tmem_base + 256 + w * kHd;
      while (pc.valid) {
        if (w == 1 && pc.virt) {
          // both cursors stand at the first stage of an item without a query tile for this group: observe every fill
          // and hand the stage back, in order
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
        while (sc.valid && !(w == 1 && sc.virt) && g_s < g_p + 2) {
          // ---- S(g_s) = Q K^T into score buffer g_s & 1 ----
          const int i = g_s & 1, buf = sc.n_done & 1;
          TRACE(1 + w, 0, g_s);
          mbar_wait(kv_full(sc.stage), sc.phase);
          TRACE(1 + w, 1, g_s);
          mbar_wait(s_free(w, i), ((uint32_t)(g_s >> 1) & 1) ^ 1);
          TRACE(1 + w, 2, g_s);
          if (sc.j == 0) mbar_wait(q_full(buf, w), (buf ? q_fill1 : q_fill0) & 1);
          tc_fence_after();
          const bool last_of_item = sc.j == sc.n_kt - 1;
          if (elect_one()) {
            const uint64_t qd = make_smem_desc_sw128(q_base + buf * 32768);
            const uint64_t kd = make_smem_desc_sw128(sbase + kOffKV + sc.stage * 16384);
#pragma unroll
            for (int k = 0; k < kHd / 16; ++k) umma_bf16(s_tmem + i * kKTile, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
            umma_commit(s_full(w, i));
            if (last_of_item) umma_commit(q_empty(buf, w));
          }
          __syncwarp();
          if (last_of_item) { if (buf) ++q_fill1; else ++q_fill0; }
          TRACE(1 + w, 3, g_s);
          ++g_s;
          advance(sc);
        }
        // ---- O += P(g_p) V ----
        {
          const int i = g_p & 1;
          TRACE(1 + w, 4, g_p);
          mbar_wait(p_full(w, i), (uint32_t)(g_p >> 1) & 1);
          TRACE(1 + w, 5, g_p);
          if (pc.j == 0) { mbar_wait(o_free(w), (items_started & 1) ^ 1); ++items_started; }
          tc_fence_after();
          if (elect_one()) {
            const uint64_t pd = make_smem_desc_sw128(sbase + kOffP + (w * 2 + i) * 16384);
            const uint64_t vd = make_smem_desc_sw128(sbase + kOffKV + pc.stage * 16384 + 8192);
#pragma unroll
            for (int k = 0; k < kKTile / 16; ++k) {
              // K-major V^T: 16 keys = 32 bytes inside the swizzled row (+2); MN-major V: 16 keys = 16 rows of 128 bytes (+128)
              umma_bf16(o_tmem, pd + 2 * k, vd + (kVT ? 2 * k : 128 * k), idesc_o, (k | pc.j) != 0);
            }
            umma_commit(p_free(w, i));
            umma_commit(kv_empty(pc.stage));                 // the second arrival comes from the other group's issuer
          }
          __syncwarp();
          TRACE(1 + w, 6, g_p);
          ++g_p;
          advance(pc);
        }
      }
    };
    if (warp == 16) issuer(std::integral_constant<int, 0>{});
    else issuer(std::integral_constant<int, 1>{});
    __syncwarp();
  } else {
    // ============================ softmax groups ============================
    const int w = warp >> 3;                       // group / query tile
    const int quarter = warp & 3, ch = (warp >> 2) & 1;   // 32-row quarter (= TMEM lane quarter), column half
    const int r = quarter * 32 + lane;             // row inside the tile
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t tmem_o = tmem_base + 256 + w * kHd + ch * 32 + lane_off;
    const int sw = r & 7;
    const int pair_bar = 1 + w * 4 + quarter;      // named barrier of the two warps that share these 32 rows
    float* xch = reinterpret_cast<float*>(sptr + kOffX) + w * 512;
    int g = 0;                                     // per-group step index
    int ordinal = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ordinal) {
      const Item it = decode_item(item, ordinal, len_cache, n_qblk, n_head, seq_len, kv_len);
      if (w == 1 && !it.active1) continue;
      // m_ref: the maximum the exponents are taken against.  It is only raised (and O / l rescaled) when the
      // running maximum exceeds it by more than 2^kRaise: P <= 2^kRaise stays far inside bf16 / fp32 range, and the
      // O accumulator in TMEM is touched by the CUDA cores only on those rare steps and once at the end.
      constexpr float kRaise = 40.0f;
      float m_ref = -INFINITY, l_run = 0.f;
#ifndef MHA_STAGGER_NS
#define MHA_STAGGER_NS 350
#endif
#if MHA_STAGGER_NS > 0
      // the two groups share every sub-partition's exp unit: start group 1 half a tile late so that its exp phase
      // falls into group 0's max / barrier phase instead of on top of its exp phase
      if (w == 1) __nanosleep(MHA_STAGGER_NS);
#endif
      for (int j = 0; j < it.n_kt; ++j, ++g) {
        const int i = g & 1;
        const uint32_t u = (uint32_t)(g >> 1) & 1;
        if ((warp & 7) == 0 && lane == 0) TRACE(3 + w, 0, g);
        mbar_wait(s_full(w, i), u);
        if ((warp & 7) == 0 && lane == 0) TRACE(3 + w, 1, g);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tmem_base + (w * 2 + i) * kKTile + ch * 32 + lane_off, v);
        tmem_ld_wait();
        // the scores are in registers: the MMA warp may overwrite this buffer with S(g + 2)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free(w, i));
        if ((warp & 7) == 0 && lane == 0) TRACE(3 + w, 2, g);
        const int valid = it.n_keys - j * kKTile - ch * 32;    // my columns < valid are real keys
        float tm0 = -INFINITY, tm1 = -INFINITY;
#ifdef MHA_NOMAX
        tm0 = tm1 = 0.f;
        if (true) {
        } else
#endif
        if (valid >= 32) {
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            tm0 = max3(tm0, __uint_as_float(v[c]), __uint_as_float(v[c + 1]));
            tm1 = max3(tm1, __uint_as_float(v[c + 2]), __uint_as_float(v[c + 3]));
          }
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            if (c >= valid) v[c] = 0xff800000u;        // -inf: exp2 gives exactly 0
            tm0 = fmaxf(tm0, __uint_as_float(v[c]));
          }
        }
        // row maximum over both column halves (the partner warp holds the other 32 columns of these rows)
        float* xrow = xch + i * 256 + r;
#ifdef MHA_NOMAX
        const float tile_max = 0.f;
#else
        xrow[ch * 128] = fmaxf(tm0, tm1);
        named_bar_sync(pair_bar, 64);
        const float tile_max = fmaxf(fmaxf(tm0, tm1), xrow[(ch ^ 1) * 128]);
#endif
        // P buffer i was last read by P.V(g - 2); once that has retired, O holds every step up to g - 2
        mbar_wait(p_free(w, i), u ^ 1);
        if ((warp & 7) == 0 && lane == 0) TRACE(3 + w, 3, g);
        if (j == 0) {
          m_ref = tile_max;                          // O is overwritten by the first P.V of the item
        } else {
          const bool raise = (tile_max - m_ref) * kLog2e > kRaise;
          if (__any_sync(0xffffffffu, raise)) {
            const float factor = raise ? ex2_approx((m_ref - tile_max) * kLog2e) : 1.0f;
            if (raise) m_ref = tile_max;
            l_run *= factor;
            mbar_wait(p_free(w, i ^ 1), (uint32_t)((g - 1) >> 1) & 1);   // P.V of the previous step has landed in TMEM
            tc_fence_after();
            uint32_t o[32];
            tmem_ld32(tmem_o, o);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * factor);
            tmem_st32(tmem_o, o);
            tmem_st_wait();
            tc_fence_before();
          }
        }
        const float m_scaled = m_ref * kLog2e;
        unsigned char* prow = sptr + kOffP + (w * 2 + i) * 16384 + r * 128;
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(v[q * 8 + 2 * e]), kLog2e, -m_scaled));
            const float p1 = ex2_approx(fmaf(__uint_as_float(v[q * 8 + 2 * e + 1]), kLog2e, -m_scaled));
            l0 += p0;
            l1 += p1;
            pk[e] = pack_bf16x2(p0, p1);
          }
#ifdef MHA_NOPSTORE
          if (pk[0] == 0x12345678u && pk[3] == 0x9abcdef0u)
#endif
          *reinterpret_cast<uint4*>(prow + (((ch * 4 + q) ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        l_run += l0 + l1;
        if ((warp & 7) == 0 && lane == 0) TRACE(3 + w, 4, g);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(w, i));
        if ((warp & 7) == 0 && lane == 0) TRACE(3 + w, 5, g);
      }
      // epilogue of the item: O (TMEM) / l -> ctx; the row sum is the total over both column halves
      float* xrow = xch + (g & 1) * 256 + r;
      xrow[ch * 128] = l_run;
      named_bar_sync(pair_bar, 64);
      const float inv = 1.0f / (l_run + xrow[(ch ^ 1) * 128]);
      named_bar_sync(pair_bar, 64);                  // the exchange slot is reused by the next item's first steps
      mbar_wait(p_free(w, (g - 1) & 1), (uint32_t)((g - 1) >> 1) & 1);   // the last P.V of the item has retired
      if ((warp & 7) == 0 && lane == 0) TRACE(3 + w, 6, g - 1);
      tc_fence_after();
      {
        uint32_t o[32];
        tmem_ld32(tmem_o, o);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free(w));       // the next item's first P.V may overwrite O_w
        // stage my 32 columns of row r in the item's Q buffer (dead since the last S), 128-byte swizzled rows
        const uint32_t srow = sbase + kOffQ + ((ordinal & 1) * 2 + w) * 16384 + r * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t a0 = pack_bf16x2(__uint_as_float(o[8 * c]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
          const uint32_t a1 = pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
          const uint32_t a2 = pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
          const uint32_t a3 = pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                       ::"r"(srow + (((ch * 4 + c) ^ sw) << 4)), "r"(a0), "r"(a1), "r"(a2), "r"(a3) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_staged(ordinal & 1, w));
      }
      if ((warp & 7) == 0 && lane == 0) TRACE(3 + w, 7, g - 1);
    }
  }

  tc_fence_before();
  __syncthreads();
#ifdef MHA_TRACE
  if (blockIdx.x == 0 && g_trace != nullptr)
    for (int k = tid; k < 5 * 64 * 8; k += kThreads) g_trace[k] = trace_s[k];
#endif
  if (warp == 16) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int num_sms_attn() { return stac_grid_limit(); }

template <bool kVT>
int launch_mha(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& tv, const CUtensorMap& tctx,
               const int32_t* kv_len, int seq_len, int d_model, int n_head, int n_qblk, int n_items, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mha_bf16_kernel<kVT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  const int grid = std::min(n_items, num_sms_attn());
  mha_bf16_kernel<kVT><<<grid, kThreads, kSmemBytes, st>>>(tq, tkv, tv, tctx, kv_len, seq_len, d_model, n_head, n_qblk,
                                                          n_items);
  STAC_LAUNCH_CHECK();
}

}  // namespace

#ifdef MHA_TRACE
extern "C" int stac_mha_trace(unsigned int* buf) {
  cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf));
  return 0;
}
#endif

extern "C" int stac_mha_bf16(const uint16_t* qkv, const uint16_t* v_t, const int32_t* kv_len, int64_t batch,
                             int64_t seq_len, int64_t t_pad, int64_t d_model, int64_t n_head, uint16_t* ctx,
                             void* stream) {
  STAC_REQUIRE(qkv && kv_len && ctx && batch > 0 && batch < 65536 && seq_len > 0);
  if (v_t) STAC_REQUIRE(t_pad >= seq_len && t_pad % 8 == 0);
  if (d_model != n_head * kHd || n_head > 65535 || batch * seq_len >= (1ll << 31)) return STAC_ERR_UNSUPPORTED_SHAPE;
  const int64_t n_qblk = ceil_div64(seq_len, 2 * kQTile);
  const int64_t n_items = batch * n_head * n_qblk;
  if (n_items >= (1ll << 31)) return STAC_ERR_UNSUPPORTED_SHAPE;
  CUtensorMap tq, tkv, tv;
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
  CUtensorMap tctx;
  {
    // ctx [B][T][d_model]: a tile that runs past the end of its utterance is clipped in the T dimension
    const uint64_t cdims[3] = {(uint64_t)d_model, (uint64_t)seq_len, (uint64_t)batch};
    const uint64_t cstr[2] = {(uint64_t)d_model * 2, (uint64_t)seq_len * d_model * 2};
    const uint32_t cbox[3] = {kHd, kQTile, 1};
    int r = encode_map(&tctx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, ctx, 3, cdims, cstr, cbox);
    if (r != STAC_OK) return r;
  }
  tv = tkv;
  if (v_t) {
    // V^T [B*H][64][t_pad], innermost = keys
    const uint64_t vdims[3] = {(uint64_t)t_pad, kHd, (uint64_t)(batch * n_head)};
    const uint64_t vstr[2] = {(uint64_t)t_pad * 2, (uint64_t)t_pad * kHd * 2};
    const uint32_t vbox[3] = {kKTile, kHd, 1};
    int r = encode_map(&tv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, v_t, 3, vdims, vstr, vbox);
    if (r != STAC_OK) return r;
    return launch_mha<true>(tq, tkv, tv, tctx, kv_len, (int)seq_len, (int)d_model, (int)n_head, (int)n_qblk,
                            (int)n_items, as_stream(stream));
  }
  return launch_mha<false>(tq, tkv, tv, tctx, kv_len, (int)seq_len, (int)d_model, (int)n_head, (int)n_qblk,
                           (int)n_items, as_stream(stream));
}
