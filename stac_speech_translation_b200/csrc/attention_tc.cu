// Flash-style multi-head self-attention on tcgen05 tensor cores (bf16 operands, fp32 softmax).
// Key-padding is by per-utterance valid length (no dense mask tensor, no B*H*T*T score matrix).
//
// Reference behaviour replaced: torch.nn.MultiheadAttention slow path (baddbmm + softmax + bmm with a
// -inf key-padding mask) reached through SpeechBrain's TransformerEncoderLayer from
//   /root/reference/stac-st/modules/TransformerMultiTask.py:304-308 (mask built at :289-294 / :225-226).
//
// Persistent, warp-specialised kernel; one CTA per SM, 320 threads:
//   warps 0-3  softmax group 0 (query tile 0: 128 rows, thread = row)
//   warps 4-7  softmax group 1 (query tile 1)
//   warp  8    TMA producer  (Q tiles double-buffered per work item, K / V^T tiles in a 3-stage ring)
//   warp  9    MMA issuer    (one thread)
// Work item = (utterance, head, block of 256 queries); the two query tiles share every K/V tile.
// Per key tile j and group w:
//   S_w = Q_w K_j^T   tcgen05.mma 128x128x64 -> TMEM columns [128w, 128w+128)
//   softmax           tcgen05.ld S_w -> registers, online max / sum (exp2, fp32), P_w -> smem as bf16 in
//                     the K-major 128B-swizzled UMMA layout
//   O_w += P_w V_j    tcgen05.mma 128x64x128 accumulating in TMEM columns [256+64w, +64) over the key tiles;
//                     the running-max correction is lazy (only when the maximum grows by more than 2^8), so
//                     the CUDA cores touch O only on those steps and once per item for the final 1/l scale.
// While one group does its softmax on the CUDA cores the tensor core runs the other group's MMAs.
#include <algorithm>
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int kHd = 64, kTile = 128;
constexpr int kKvStages = 3;
constexpr int kThreads = 320;
// shared memory map (bytes, from a 1024-aligned base)
constexpr int kOffQ = 0;                         // [2 bufs][2 tiles] x 16 KB
constexpr int kOffP = 65536;                     // [2 groups] x 32 KB (two 64-key K-blocks of 16 KB)
constexpr int kOffKV = 131072;                   // [3 stages] x (K 16 KB + V^T 16 KB)
constexpr int kOffBar = kOffKV + kKvStages * 32768;
constexpr int kSmemBytes = kOffBar + 256 + 1024;
constexpr float kLog2e = 1.4426950408889634f;


__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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

// exp2 of one 32-column chunk relative to the row's reference maximum; P goes to smem as bf16 (K-major,
// 128B swizzle: columns ch*32.. -> K-block ch>>1, 16-byte chunks (ch&1)*4 .. +3).  Returns the chunk's row sum.
template <bool kFull>
__device__ __forceinline__ float exp_chunk(const uint32_t (&v)[32], int ch, int valid, float m_scaled,
                                           unsigned char* prow, int sw) {
  uint32_t pk[16];
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), kLog2e, -m_scaled));
    float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), kLog2e, -m_scaled));
    if (!kFull) {
      if (ch * 32 + i >= valid) p0 = 0.f;
      if (ch * 32 + i + 1 >= valid) p1 = 0.f;
    }
    l0 += p0;
    l1 += p1;
    pk[i >> 1] = pack_bf16x2(p0, p1);
  }
  unsigned char* blk = prow + (ch >> 1) * 16384;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int chunk = ((ch & 1) * 4 + q) ^ sw;
    *reinterpret_cast<uint4*>(blk + chunk * 16) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  }
  return l0 + l1;
}

template <bool kFull>
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32], int ch, int valid, float m) {
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (kFull || ch * 32 + i < valid) m = fmaxf(m, __uint_as_float(v[i]));
  return m;
}

struct Item {
  int b, h, q0, n_keys, n_kt;
  bool active[2];
};

__device__ __forceinline__ Item decode_item(int item, int n_qblk, int n_head, int seq_len,
                                            const int* __restrict__ kv_len) {
  Item it;
  const int qb = item % n_qblk;
  const int bh = item / n_qblk;
  it.h = bh % n_head;
  it.b = bh / n_head;
  it.q0 = qb * 2 * kTile;
  it.n_keys = min(max(__ldg(kv_len + it.b), 1), seq_len);
  it.n_kt = (it.n_keys + kTile - 1) / kTile;
  it.active[0] = true;
  it.active[1] = it.q0 + kTile < seq_len;
  return it;
}

// 10 warps = 3 on two of the four SM sub-partitions, whose 16 K-register files cap the kernel at 168 regs/thread
__global__ void __launch_bounds__(kThreads, 1)
mha_bf16_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_vt,
                const int* __restrict__ kv_len, int seq_len, int d_model, int n_head, int n_qblk,
                int n_items, __nv_bfloat16* __restrict__ ctx) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bars = sbase + kOffBar;
  auto q_full = [&](int buf, int w) { return bars + 8u * (buf * 2 + w); };
  auto q_empty = [&](int buf, int w) { return bars + 8u * (4 + buf * 2 + w); };
  auto kv_full = [&](int s) { return bars + 8u * (8 + s); };
  auto kv_empty = [&](int s) { return bars + 8u * (11 + s); };
  auto s_full = [&](int w) { return bars + 8u * (14 + w); };
  auto p_full = [&](int w) { return bars + 8u * (16 + w); };
  auto o_full = [&](int w) { return bars + 8u * (18 + w); };
  auto o_empty = [&](int w) { return bars + 8u * (20 + w); };
  const uint32_t tmem_slot = bars + 8u * 22;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    prefetch_tmap(&tmap_qkv);
    prefetch_tmap(&tmap_vt);
    for (int i = 0; i < 4; ++i) { mbar_init(q_full(i >> 1, i & 1), 1); mbar_init(q_empty(i >> 1, i & 1), 1); }
    for (int s = 0; s < kKvStages; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int w = 0; w < 2; ++w) {
      mbar_init(s_full(w), 1); mbar_init(p_full(w), 4); mbar_init(o_full(w), 1); mbar_init(o_empty(w), 4);
    }
    fence_barrier_init();
  }
  if (warp == 9) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 8) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t kv_phase = 0;
      int n_done = 0;
      uint32_t q_uses[2][2] = {{0, 0}, {0, 0}};   // fills of Q buffer (buf, w): parity source
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_done) {
        const Item it = decode_item(item, n_qblk, n_head, seq_len, kv_len);
        const int buf = n_done & 1;
        const int row_base = it.b * seq_len;
        for (int w = 0; w < 2; ++w) {
          if (!it.active[w]) continue;
          const uint32_t qph = (q_uses[buf][w]++) & 1;
          mbar_wait(q_empty(buf, w), qph ^ 1);
          mbar_arrive_expect_tx(q_full(buf, w), kTile * kHd * 2);
          tma_load_2d(sbase + kOffQ + (buf * 2 + w) * 16384, &tmap_qkv, q_full(buf, w), it.h * kHd,
                      row_base + it.q0 + w * kTile);
        }
        for (int j = 0; j < it.n_kt; ++j) {
          mbar_wait(kv_empty(stage), kv_phase ^ 1);
          const uint32_t kdst = sbase + kOffKV + stage * 32768;
          mbar_arrive_expect_tx(kv_full(stage), 32768);
          tma_load_2d(kdst, &tmap_qkv, kv_full(stage), d_model + it.h * kHd, row_base + j * kTile);
          tma_load_3d(kdst + 16384, &tmap_vt, kv_full(stage), j * kTile, 0, it.b * n_head + it.h);
          tma_load_3d(kdst + 16384 + 8192, &tmap_vt, kv_full(stage), j * kTile + 64, 0, it.b * n_head + it.h);
          if (++stage == kKvStages) { stage = 0; kv_phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ============================ MMA issuer ============================
    // Flat sequence of steps (work item, key tile).  The S cursor runs one step ahead of the P.V cursor and
    // the two groups are interleaved:  PV_0(n), S_0(n+1), PV_1(n), S_1(n+1), ...  so that a group's next
    // score tile is in flight as soon as its P tile has been consumed, while the other group is still in its
    // softmax (ping-pong), also across work-item boundaries.
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64);
      struct Cursor {
        int item, j, stage, n_done;
        uint32_t kv_phase;
        Item it;
        bool valid;
      };
      auto advance = [&](Cursor& c) {
        if (++c.stage == kKvStages) { c.stage = 0; c.kv_phase ^= 1; }
        if (++c.j == c.it.n_kt) {
          c.j = 0;
          c.item += gridDim.x;
          ++c.n_done;
          c.valid = c.item < n_items;
          if (c.valid) c.it = decode_item(c.item, n_qblk, n_head, seq_len, kv_len);
        }
      };
      Cursor sc;                      // S cursor
      sc.item = blockIdx.x; sc.j = 0; sc.stage = 0; sc.n_done = 0; sc.kv_phase = 0;
      sc.valid = sc.item < n_items;
      if (sc.valid) sc.it = decode_item(sc.item, n_qblk, n_head, seq_len, kv_len);
      Cursor pc = sc;                 // P.V cursor
      uint32_t it_cnt[2] = {0, 0};    // steps completed per group (parity of p_full)
      uint32_t items_w[2] = {0, 0};   // work items started per group (parity of o_empty)
      uint32_t q_uses[2][2] = {{0, 0}, {0, 0}};
      uint32_t q_par[2] = {0, 0};

      auto issue_s = [&](const Cursor& c, int w) {
        // caller has waited kv_full(c.stage); S_w is free (its previous tile was consumed before p_full[w])
        const int buf = c.n_done & 1;
        if (c.j == 0) {
          q_par[w] = (q_uses[buf][w]++) & 1;
          mbar_wait(q_full(buf, w), q_par[w]);
        }
        tc_fence_after();
        const uint64_t qd = make_smem_desc_sw128(sbase + kOffQ + (buf * 2 + w) * 16384);
        const uint64_t kd = make_smem_desc_sw128(sbase + kOffKV + c.stage * 32768);
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k) umma_bf16(tmem_base + w * 128, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
        umma_commit(s_full(w));
        if (c.j == c.it.n_kt - 1) umma_commit(q_empty(buf, w));
      };

      if (sc.valid) {
        mbar_wait(kv_full(sc.stage), sc.kv_phase);
        for (int w = 0; w < 2; ++w)
          if (sc.it.active[w]) issue_s(sc, w);
        advance(sc);
      }
      while (pc.valid) {
        bool next_kv_ready = false;
        const uint32_t vaddr = sbase + kOffKV + pc.stage * 32768 + 16384;
        for (int w = 0; w < 2; ++w) {
          if (pc.it.active[w]) {
            const uint32_t ph = it_cnt[w] & 1;
            mbar_wait(p_full(w), ph);            // P_w in smem, S_w fully read (and O_w rescaled if needed)
            if (pc.j == 0) mbar_wait(o_empty(w), ((items_w[w]++) & 1) ^ 1);   // previous item's O_w was read out
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < kTile / 16; ++k) {
              const uint64_t pd = make_smem_desc_sw128(sbase + kOffP + w * 32768 + (k >> 2) * 16384) + 2 * (k & 3);
              const uint64_t vd = make_smem_desc_sw128(vaddr + (k >> 2) * 8192) + 2 * (k & 3);
              umma_bf16(tmem_base + 256 + w * 64, pd, vd, idesc_o, (k | pc.j) != 0);
            }
            umma_commit(o_full(w));
            ++it_cnt[w];
          }
          if (sc.valid && sc.it.active[w]) {
            if (!next_kv_ready) { mbar_wait(kv_full(sc.stage), sc.kv_phase); next_kv_ready = true; }
            issue_s(sc, w);
          }
        }
        umma_commit(kv_empty(pc.stage));         // K/V of this step are free once every MMA above has retired
        advance(pc);
        if (sc.valid) advance(sc);
      }
    }
    __syncwarp();
  } else {
    // ============================ softmax groups ============================
    const int w = warp >> 2;                       // group / query tile
    const int r = tid & 127;                       // row inside the tile
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tmem_s = tmem_base + w * 128 + lane_off;
    const uint32_t tmem_o = tmem_base + 256 + w * 64 + lane_off;
    unsigned char* prow = sptr + kOffP + w * 32768 + r * 128;
    const int sw = r & 7;
    uint32_t it_cnt = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const Item it = decode_item(item, n_qblk, n_head, seq_len, kv_len);
      if (!it.active[w]) continue;
      // m_ref: the maximum the exponents are taken against.  It is only raised (and O / l rescaled) when the
      // running maximum exceeds it by more than 8 in the log2 domain, so P <= 2^8 and the O accumulator in
      // TMEM is touched by the CUDA cores only on those rare steps and once at the end.
      float m_ref = -INFINITY, l_run = 0.f;
      for (int j = 0; j < it.n_kt; ++j, ++it_cnt) {
        const uint32_t ph = it_cnt & 1;
        mbar_wait(s_full(w), ph);
        tc_fence_after();
        const int valid = it.n_keys - j * kTile;     // columns < valid are real keys
        const bool full = valid >= kTile;
        // Pass 1 (row maximum): columns 64..127 are read, reduced and dropped, columns 0..63 stay in
        // registers for pass 2; columns 64..127 are re-read from TMEM for their exponentials.  This keeps the
        // live set at 64 score registers (the kernel is capped at 168 registers per thread).
        uint32_t v0[32], v1[32];
        float tile_max = -INFINITY, tile_max_b = -INFINITY;
        {
          uint32_t v2[32], v3[32];
          tmem_ld32(tmem_s + 64, v2);
          tmem_ld32(tmem_s + 96, v3);
          tmem_ld_wait();
          tmem_ld32(tmem_s, v0);
          tmem_ld32(tmem_s + 32, v1);
          if (full) {
            tile_max = chunk_max<true>(v2, 2, valid, tile_max);
            tile_max_b = chunk_max<true>(v3, 3, valid, tile_max_b);
          } else {
            tile_max = chunk_max<false>(v2, 2, valid, tile_max);
            tile_max_b = chunk_max<false>(v3, 3, valid, tile_max_b);
          }
          tmem_ld_wait();
        }
        if (full) {
          tile_max = chunk_max<true>(v0, 0, valid, tile_max);
          tile_max_b = chunk_max<true>(v1, 1, valid, tile_max_b);
        } else {
          tile_max = chunk_max<false>(v0, 0, valid, tile_max);
          tile_max_b = chunk_max<false>(v1, 1, valid, tile_max_b);
        }
        tile_max = fmaxf(tile_max, tile_max_b);
        if (j == 0) {
          m_ref = tile_max;                          // O is overwritten by the first P.V of the item
        } else {
          const bool raise = (tile_max - m_ref) * kLog2e > 8.0f;
          if (__any_sync(0xffffffffu, raise)) {
            const float factor = raise ? ex2_approx((m_ref - tile_max) * kLog2e) : 1.0f;
            if (raise) m_ref = tile_max;
            l_run *= factor;
            mbar_wait(o_full(w), ph ^ 1);            // P.V of the previous step has landed in TMEM
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < 2; ++ch) {
              uint32_t o[32];
              tmem_ld32(tmem_o + ch * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
              tmem_st32(tmem_o + ch * 32, o);
              tmem_st_wait();
            }
          }
        }
        const float m_scaled = m_ref * kLog2e;
        float l_tile;
        {
          uint32_t v2[32];
          tmem_ld32(tmem_s + 64, v2);                // in flight while columns 0..63 are exponentiated
          if (full) {
            l_tile = exp_chunk<true>(v0, 0, valid, m_scaled, prow, sw);
            l_tile += exp_chunk<true>(v1, 1, valid, m_scaled, prow, sw);
          } else {
            l_tile = exp_chunk<false>(v0, 0, valid, m_scaled, prow, sw);
            l_tile += exp_chunk<false>(v1, 1, valid, m_scaled, prow, sw);
          }
          tmem_ld_wait();
          tmem_ld32(tmem_s + 96, v0);
          if (full) l_tile += exp_chunk<true>(v2, 2, valid, m_scaled, prow, sw);
          else l_tile += exp_chunk<false>(v2, 2, valid, m_scaled, prow, sw);
          tmem_ld_wait();
          if (full) l_tile += exp_chunk<true>(v0, 3, valid, m_scaled, prow, sw);
          else l_tile += exp_chunk<false>(v0, 3, valid, m_scaled, prow, sw);
        }
        l_run += l_tile;
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(w));
      }
      // epilogue of the item: O (TMEM) / l -> ctx
      mbar_wait(o_full(w), (it_cnt - 1) & 1);
      tc_fence_after();
      const int q = it.q0 + w * kTile + r;
      const float inv = 1.0f / l_run;
      __nv_bfloat16* dst = ctx + ((int64_t)it.b * seq_len + q) * d_model + it.h * kHd;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t o[32];
        tmem_ld32(tmem_o + ch * 32, o);
        tmem_ld_wait();
        if (q < seq_len) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            *reinterpret_cast<uint4*>(dst + ch * 32 + i) = make_uint4(
                pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv),
                pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv),
                pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv),
                pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty(w));        // the next item's first P.V may overwrite O_w
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int num_sms_attn() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace

extern "C" int stac_mha_bf16(const uint16_t* qkv, const uint16_t* v_t, const int32_t* kv_len, int64_t batch,
                             int64_t seq_len, int64_t t_pad, int64_t d_model, int64_t n_head, uint16_t* ctx,
                             void* stream) {
  STAC_REQUIRE(qkv && v_t && kv_len && ctx && batch > 0 && batch < 65536 && seq_len > 0);
  STAC_REQUIRE(t_pad >= seq_len && t_pad % 8 == 0);
  if (d_model != n_head * kHd || n_head > 65535 || batch * seq_len >= (1ll << 31)) return STAC_ERR_UNSUPPORTED_SHAPE;
  const int64_t n_qblk = ceil_div64(seq_len, 2 * kTile);
  const int64_t n_items = batch * n_head * n_qblk;
  if (n_items >= (1ll << 31)) return STAC_ERR_UNSUPPORTED_SHAPE;
  CUtensorMap tq, tv;
  {
    const uint64_t dims[2] = {(uint64_t)(3 * d_model), (uint64_t)(batch * seq_len)};
    const uint64_t str[1] = {(uint64_t)(3 * d_model) * 2};
    const uint32_t box[2] = {kHd, 128};
    int r = encode_map(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  {
    // V^T [B*H][64][t_pad], innermost = keys
    const uint64_t dims[3] = {(uint64_t)t_pad, kHd, (uint64_t)(batch * n_head)};
    const uint64_t str[2] = {(uint64_t)t_pad * 2, (uint64_t)t_pad * kHd * 2};
    const uint32_t box[3] = {64, kHd, 1};
    int r = encode_map(&tv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, v_t, 3, dims, str, box);
    if (r != STAC_OK) return r;
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mha_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  const int grid = (int)std::min<int64_t>(n_items, num_sms_attn());
  mha_bf16_kernel<<<grid, kThreads, kSmemBytes, as_stream(stream)>>>(tq, tv, kv_len, (int)seq_len, (int)d_model,
                                                                    (int)n_head, (int)n_qblk, (int)n_items,
                                                                    reinterpret_cast<__nv_bfloat16*>(ctx));
  STAC_LAUNCH_CHECK();
}
