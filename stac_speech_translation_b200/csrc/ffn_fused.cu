// a7, position-wise feed-forward of one encoder layer as ONE kernel (bf16 mode, d_model = 256):
//     x += W2 . GELU(W1 . h + b1) + b2          h = LayerNorm2(x) in bf16, x = fp32 residual stream
// Reference behaviour replaced: SpeechBrain PositionalwiseFeedForward (Linear, GELU, Dropout(id), Linear) + the
// residual add of TransformerEncoderLayer, reached from
//   /root/reference/stac-st/modules/TransformerMultiTask.py:304-308.
//
// The two GEMMs of the block are chained through tensor memory and shared memory the way flash attention chains
// Q.K^T and P.V, so the [rows, d_ffn] hidden activation (98 MB per layer at the benchmark shape: written once and
// read once by the unfused path) never exists in HBM and one launch replaces two:
//   per 128-row tile, for every chunk c of 128 hidden units
//     S_c = h . W1[c]^T            tcgen05.mma 128x128xK=256  -> TMEM (two S buffers, S runs two chunks ahead)
//     P_c = GELU(S_c + b1[c])      16 epilogue warps: TMEM -> registers -> bf16 -> smem (two P buffers, K-major UMMA layout)
//     O  += P_c . W2[:, c]^T       tcgen05.mma 128x256xK=128 (N = 256) accumulating in TMEM
//   then x += O + b2 as TMA reduce-add from a staging tile (the residual stream is never loaded into the SM).
// W1 / W2 stream from L2 through a ring of three 32 KB units (W1: [128 hidden x 128 k] as two k-blocks; W2: [256
// outputs x 64 hidden]; bf16, 128-byte swizzle) in exactly the order the MMA warp consumes them.  Roles: warps 0-15 epilogue (warp & 3 = TMEM lane quarter, warp >> 2 = column group),
// warp 16 TMA producer, warp 17 MMA issuer (warp-uniform loops, one elected lane issues).
#include <algorithm>
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int kD = 256;            // d_model this kernel is specialised for
constexpr int kHC = 128;           // hidden units per chunk
constexpr int kKB = 16384;         // one [128 rows x 64 k] bf16 k-block (h, P, and the halves of a W1 unit)
constexpr int kUnit = 32768;       // ring unit: W1 [128 hidden x 128 k] (two k-blocks) or W2 [256 outputs x 64 hidden]
constexpr int kRing = 3;
constexpr int kEpiWarps = 16;
constexpr int kThreads = (kEpiWarps + 2) * 32;

constexpr int kOffH = 0;                       // 4 k-blocks x 16 KB: the h tile (A operand of GEMM 1)
constexpr int kOffP = 65536;                   // 2 buffers x (2 k-blocks x 16 KB): P (A operand of GEMM 2) / output staging
constexpr int kOffW = 131072;                  // weight ring
constexpr int kOffBias = kOffW + kRing * kUnit;   // b2 [256] fp32 (b1 is read through L1: no room for it here)
constexpr int kMaxFfn = 4096;
constexpr int kOffBar = kOffBias + kD * 4;
constexpr int kNumBars = 2 + 2 * kRing + 8 + 2;
constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16 + 1024;

// the same tanh-form minimax fit of the exact-erf GELU as the GEMM epilogue (gemm_tc.cu)
__device__ __forceinline__ float gelu_fast(float x) {
#ifdef FFN_NOGELU     // timing experiment only
  return x;
#endif
  const float u = fminf(x * x, 64.0f);
  float p = fmaf(-3.51516785e-04f, u, 3.70056460e-02f);
  p = fmaf(p, u, 7.97507884e-01f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * p));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
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
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

#ifdef FFN_TRACE     // timing experiment only (tools/trace_ffn.py): CTA 0 logs clock32 per (role, chunk, event) to global memory
__device__ unsigned int* g_ffn_trace = nullptr;
#define FTRACE(role, ev, n) do { if (blockIdx.x == 0 && g_ffn_trace != nullptr && (n) < 32) g_ffn_trace[((role) * 32 + (n)) * 8 + (ev)] = (unsigned int)clock64(); } while (0)
#else
#define FTRACE(role, ev, n) do {} while (0)
#endif

__global__ void __launch_bounds__(kThreads, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_w1,
                 const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_x,
                 const float* __restrict__ b1, const float* __restrict__ b2, int m_rows, int n_chunks, int num_tiles) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bars = sbase + kOffBar;
  auto h_full = [&]() { return bars; };
  auto h_free = [&]() { return bars + 8u; };
  auto w_full = [&](int s) { return bars + 8u * (2 + s); };
  auto w_empty = [&](int s) { return bars + 8u * (2 + kRing + s); };
  const uint32_t gb = bars + 8u * (2 + 2 * kRing);
  auto s_full = [&](int i) { return gb + 8u * i; };
  auto s_free = [&](int i) { return gb + 8u * (2 + i); };
  auto p_full = [&](int i) { return gb + 8u * (4 + i); };
  auto p_free = [&](int i) { return gb + 8u * (6 + i); };
  auto o_full = [&]() { return gb + 8u * 8; };
  auto o_free = [&]() { return gb + 8u * 9; };
  const uint32_t tmem_slot = bars + 8u * kNumBars;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* b2_s = reinterpret_cast<float*>(sptr + kOffBias);

  if (tid == 0) {
    prefetch_tmap(&tmap_h); prefetch_tmap(&tmap_w1); prefetch_tmap(&tmap_w2); prefetch_tmap(&tmap_x);
    mbar_init(h_full(), 1); mbar_init(h_free(), 1);
    for (int s = 0; s < kRing; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_full(i), 1); mbar_init(s_free(i), kEpiWarps); mbar_init(p_full(i), kEpiWarps); mbar_init(p_free(i), 1);
    }
    mbar_init(o_full(), 1); mbar_init(o_free(), kEpiWarps);
    fence_barrier_init();
  }
  if (warp == 17) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  for (int i = tid; i < kD; i += kThreads) b2_s[i] = __ldg(b2 + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 16) {
    // ============================ TMA producer ============================
    // unit order per tile = MMA consumption order: W1[0] W1[1] | W2[0] W1[2] | W2[1] W1[3] | ... | W2[n-2] | W2[n-1],
    // two 32 KB units each
    int stage = 0;
    uint32_t phase = 0;
    int n_done = 0;
    auto next_stage = [&]() { if (++stage == kRing) { stage = 0; phase ^= 1; } };
    auto load_w1 = [&](int c) {
      for (int u = 0; u < 2; ++u) {            // k-blocks 2u, 2u+1 of the 128 hidden rows of chunk c
        mbar_wait(w_empty(stage), phase ^ 1);
        if (elect_one()) {
#ifdef FFN_NOLOAD    // timing experiment only (tools/bench_ffn.py): the weight stream is not loaded, results are wrong
          mbar_arrive(w_full(stage));
#else
          mbar_arrive_expect_tx(w_full(stage), kUnit);
          tma_load_2d(sbase + kOffW + stage * kUnit, &tmap_w1, w_full(stage), (2 * u) * 64, c * kHC);
          tma_load_2d(sbase + kOffW + stage * kUnit + kKB, &tmap_w1, w_full(stage), (2 * u + 1) * 64, c * kHC);
#endif
        }
        __syncwarp();
        next_stage();
      }
    };
    auto load_w2 = [&](int c) {
      for (int kb = 0; kb < 2; ++kb) {         // all 256 output rows x 64 hidden units of chunk c
        mbar_wait(w_empty(stage), phase ^ 1);
        if (elect_one()) {
#ifdef FFN_NOLOAD
          mbar_arrive(w_full(stage));
#else
          mbar_arrive_expect_tx(w_full(stage), kUnit);
          tma_load_2d(sbase + kOffW + stage * kUnit, &tmap_w2, w_full(stage), c * kHC + kb * 64, 0);
#endif
        }
        __syncwarp();
        next_stage();
      }
    };
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n_done) {
      mbar_wait(h_free(), (uint32_t)(n_done & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(h_full(), 4 * kKB);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(sbase + kOffH + kb * kKB, &tmap_h, h_full(), kb * 64, tile * 128);
      }
      __syncwarp();
      load_w1(0);
      if (n_chunks > 1) load_w1(1);
      for (int c = 0; c < n_chunks; ++c) {
        load_w2(c);
        if (c + 2 < n_chunks) load_w1(c + 2);
      }
    }
  } else if (warp == 17) {
    // ============================ MMA issuer ============================
    // ONE warp consumes the weight ring, in the producer's order.  (Two issuer warps - one for S, one for O, each
    // skipping the other's units - were 8 % faster but are not safe with parity-tracked mbarriers: the warp that runs
    // ahead can test a slot's w_full barrier while the slot's PREVIOUS use, owned by the other warp, has not even been
    // filled; the parity then aliases to "complete" and it reads a slot that is still being written.  It showed up as a
    // barrier time-out on rank 0 of an 8-GPU run, where NVLink ingress slows the TMA loads.)
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 256);
    int stage = 0;
    uint32_t phase = 0;
    int n_done = 0;
    int g = 0;                              // running chunk index over all tiles: selects S / P buffer and barrier parity
    auto issue_s = [&](int gi) {
      // S(gi) = h . W1[chunk]^T : two ring units of two 64-wide k-blocks each
      const int i = gi & 1;
      if (lane == 0) FTRACE(0, 3, gi);
      mbar_wait(s_free(i), ((uint32_t)(gi >> 1) & 1) ^ 1);
      if (lane == 0) FTRACE(0, 4, gi);
      for (int u = 0; u < 2; ++u) {
        mbar_wait(w_full(stage), phase);
        if (lane == 0 && u == 0) FTRACE(0, 5, gi);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t ad = make_smem_desc_sw128(sbase + kOffH + (2 * u + h) * kKB);
            const uint64_t bd = make_smem_desc_sw128(sbase + kOffW + stage * kUnit + h * kKB);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + i * 128, ad + 2 * k, bd + 2 * k, idesc_s, (u | h | k) != 0);
          }
          umma_commit(w_empty(stage));
          if (u == 1) umma_commit(s_full(i));
        }
        __syncwarp();
        if (++stage == kRing) { stage = 0; phase ^= 1; }
      }
    };
    auto issue_o = [&](int gi, bool first) {
      // O += P(gi) . W2[:, chunk]^T : two 64-wide k-blocks, each one ring unit holding all 256 output rows (N = 256)
      const int i = gi & 1;
      if (lane == 0) FTRACE(0, 0, gi);
      mbar_wait(p_full(i), (uint32_t)(gi >> 1) & 1);
      if (lane == 0) FTRACE(0, 1, gi);
      for (int kb = 0; kb < 2; ++kb) {
        mbar_wait(w_full(stage), phase);
        if (lane == 0 && kb == 0) FTRACE(0, 2, gi);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = make_smem_desc_sw128(sbase + kOffP + i * 32768 + kb * kKB);
          const uint64_t bd = make_smem_desc_sw128(sbase + kOffW + stage * kUnit);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + 256, ad + 2 * k, bd + 2 * k, idesc_o, !(first && kb == 0 && k == 0));
          umma_commit(w_empty(stage));
          if (kb == 1) umma_commit(p_free(i));
        }
        __syncwarp();
        if (++stage == kRing) { stage = 0; phase ^= 1; }
      }
    };
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n_done) {
      auto release_h = [&]() {       // the last S of the tile has been issued: once it retires the h tile may be reloaded
        if (elect_one()) umma_commit(h_free());
        __syncwarp();
      };
      mbar_wait(h_full(), (uint32_t)n_done & 1);
      issue_s(g);
      if (n_chunks > 1) issue_s(g + 1);
      if (n_chunks <= 2) release_h();
      for (int c = 0; c < n_chunks; ++c) {
        if (c == 0) mbar_wait(o_free(), ((uint32_t)n_done & 1) ^ 1);     // the previous tile's O has been read out
        issue_o(g + c, c == 0);
        if (c + 2 < n_chunks) {
          issue_s(g + c + 2);
          if (c + 2 == n_chunks - 1) release_h();
        }
      }
      if (elect_one()) umma_commit(o_full());
      __syncwarp();
      g += n_chunks;
    }
  } else {
    // ============================ epilogue warps ============================
    const int quarter = warp & 3, cgrp = warp >> 2;
    const int r = quarter * 32 + lane;                 // row of the tile = TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int sw = r & 7;
    const uint32_t stage_buf = sbase + kOffP + warp * 4096;    // output staging (after the last P of a tile was consumed)
    int n_done = 0, g = 0;
    bool staged = false;               // output staging tiles of the previous tile may still be being read
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n_done) {
      for (int c = 0; c < n_chunks; ++c, ++g) {
        const int i = g & 1;
        const uint32_t u = (uint32_t)(g >> 1) & 1;
        if (warp == 0 && lane == 0) FTRACE(1, 0, g);
        mbar_wait(s_full(i), u);
        if (warp == 0 && lane == 0) FTRACE(1, 1, g);
        tc_fence_after();
        float x[32];
        {
          uint32_t v[32];
          tmem_ld32(tmem_base + i * 128 + cgrp * 32 + lane_off, v);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) x[k] = __uint_as_float(v[k]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free(i));
        const float* bc = b1 + c * kHC + cgrp * 32;      // warp-uniform addresses: broadcast loads served by L1
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(bc + k));
          x[k] = gelu_fast(x[k] + bv.x); x[k + 1] = gelu_fast(x[k + 1] + bv.y);
          x[k + 2] = gelu_fast(x[k + 2] + bv.z); x[k + 3] = gelu_fast(x[k + 3] + bv.w);
        }
        if (warp == 0 && lane == 0 && x[0] != 123.f) FTRACE(1, 2, g);
        mbar_wait(p_free(i), u ^ 1);                     // P.W2 of chunk g - 2 has retired
        if (staged) {
          // the previous tile's output staging tiles live in the P buffers: nobody may write P before every store has
          // read them.  Waiting here, after the first chunk's GELU, hides most of the ~3500 clk a reduce-add takes to
          // read its source behind useful work.
          bulk_wait_read0();
          epi_bar_sync();
          staged = false;
        }
        if (warp == 0 && lane == 0) FTRACE(1, 3, g);
        // my 32 hidden columns = k-block cgrp >> 1, 16-byte chunks (cgrp & 1) * 4 .. +3 of row r
        const uint32_t prow = sbase + kOffP + i * 32768 + (cgrp >> 1) * kKB + r * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_shared_v4(prow + ((((cgrp & 1) * 4 + j) ^ sw) << 4), pack_bf16x2(x[8 * j], x[8 * j + 1]),
                       pack_bf16x2(x[8 * j + 2], x[8 * j + 3]), pack_bf16x2(x[8 * j + 4], x[8 * j + 5]),
                       pack_bf16x2(x[8 * j + 6], x[8 * j + 7]));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(i));
        if (warp == 0 && lane == 0) FTRACE(1, 4, g);
      }
      // ---- x += O + b2 : 64 columns per warp, two 32-column fp32 halves through this warp's staging tile ----
      mbar_wait(o_full(), (uint32_t)n_done & 1);          // every MMA of the tile has retired (P buffers are free too)
      tc_fence_after();
      const int row0 = tile * 128 + quarter * 32;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const int col0 = cgrp * 64 + half * 32;
        uint32_t v[32];
        tmem_ld32(tmem_base + 256 + col0 + lane_off, v);
        tmem_ld_wait();
        if (half == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_free());
        }
        bulk_wait_read0();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bv = *reinterpret_cast<const float4*>(b2_s + col0 + 4 * j);
          st_shared_v4(stage_buf + lane * 128 + ((j ^ sw) << 4), __float_as_uint(__uint_as_float(v[4 * j]) + bv.x),
                       __float_as_uint(__uint_as_float(v[4 * j + 1]) + bv.y),
                       __float_as_uint(__uint_as_float(v[4 * j + 2]) + bv.z),
                       __float_as_uint(__uint_as_float(v[4 * j + 3]) + bv.w));
        }
        fence_proxy_async_smem();
        __syncwarp();
#if defined(FFN_NOOUT)       // timing experiments only (tools/bench_ffn.py): results are wrong
        if (false) {}
#elif defined(FFN_PLAINSTORE)
        if (row0 < m_rows && elect_one()) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmap_x)), "r"(stage_buf), "r"(col0), "r"(row0) : "memory");
          bulk_commit();
        }
#else
        if (row0 < m_rows && elect_one()) { tma_reduce_add_2d(&tmap_x, stage_buf, col0, row0); bulk_commit(); }
#endif
        __syncwarp();
      }
      staged = true;
    }
    bulk_wait0();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int num_sms_ffn() { return stac_grid_limit(); }

}  // namespace

#ifdef FFN_TRACE
extern "C" int stac_ffn_trace(unsigned int* buf) { cudaMemcpyToSymbol(g_ffn_trace, &buf, sizeof(buf)); return 0; }
#endif

extern "C" int stac_ffn_fused_bf16(const uint16_t* h, const uint16_t* w1, const float* b1, const uint16_t* w2,
                                   const float* b2, float* x, int64_t m, int64_t d_model, int64_t d_ffn, void* stream) {
  STAC_REQUIRE(h && w1 && b1 && w2 && b2 && x && m > 0);
  if (d_model != kD || d_ffn % kHC != 0 || d_ffn > kMaxFfn || d_ffn < kHC || m >= (1ll << 31) - 256)
    return STAC_ERR_UNSUPPORTED_SHAPE;
  CUtensorMap th, tw1, tw2, tx;
  {
    const uint64_t dims[2] = {(uint64_t)kD, (uint64_t)m};
    const uint64_t str[1] = {(uint64_t)kD * 2};
    const uint32_t box[2] = {64, 128};
    int r = encode_map(&th, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, h, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)kD, (uint64_t)d_ffn};
    const uint64_t str[1] = {(uint64_t)kD * 2};
    const uint32_t box[2] = {64, 128};
    int r = encode_map(&tw1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w1, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)d_ffn, (uint64_t)kD};
    const uint64_t str[1] = {(uint64_t)d_ffn * 2};
    const uint32_t box[2] = {64, 256};
    int r = encode_map(&tw2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w2, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)kD, (uint64_t)m};
    const uint64_t str[1] = {(uint64_t)kD * 4};
    const uint32_t box[2] = {32, 32};
    int r = encode_map(&tx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, x, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  const int num_tiles = (int)ceil_div64(m, 128);
  const int grid = std::min(num_tiles, num_sms_ffn());
  ffn_fused_kernel<<<grid, kThreads, kSmemBytes, as_stream(stream)>>>(th, tw1, tw2, tx, b1, b2, (int)m,
                                                                     (int)(d_ffn / kHC), num_tiles);
  STAC_LAUNCH_CHECK();
}
