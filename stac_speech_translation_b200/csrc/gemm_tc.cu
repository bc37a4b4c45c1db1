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
// Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2-5 = epilogue (TMEM -> registers -> fused bias/GELU/residual -> global).  Two accumulator
// stages in TMEM (2 x BLOCK_N columns) let the epilogue of tile i overlap the MMAs of tile i+1.
#include <algorithm>
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int BLOCK_M = 128, BLOCK_K = 64, UMMA_K = 16;
constexpr int kThreads = 192;
constexpr int kConvRows = 120, kConvT = 6, kConvF = 20;  // conv A tile: 6 time steps x 20 freq bins

template <int BLOCK_N>
struct Cfg {
  static constexpr int kStages = BLOCK_N == 256 ? 4 : 6;
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BLOCK_N;  // 512 or 256 (power of two)
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct EpiParams {
  const float* bias;
  const float* resid;
  int64_t resid_period;
  int act;
  void* c;
  int c_bf16;
  int64_t m, n;
  // V^T side output of a packed QKV projection
  __nv_bfloat16* vt;
  int64_t vt_cols, seq_len, t_pad;
  // conv mode
  int conv;           // 0 = linear, 1 = conv1 implicit GEMM
  int t2_len;         // output time steps per utterance (conv)
  int tiles_per_utt;  // ceil(T2 / 6)
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const EpiParams ep, const int num_m_tiles, const int num_n_tiles, const int num_k_blocks) {
  using C = Cfg<BLOCK_N>;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + C::kStages * C::kStageBytes;
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
    for (int s = 0; s < C::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_tile = tile / num_n_tiles, n_tile = tile - m_tile * num_n_tiles;
        const int conv_b = ep.conv ? m_tile / ep.tiles_per_utt : 0;
        const int conv_t0 = ep.conv ? (m_tile - conv_b * ep.tiles_per_utt) * kConvT : 0;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t a_dst = smem_base + stage * C::kStageBytes;
          const uint32_t b_dst = a_dst + C::kABytes;
          if (ep.conv) {
            mbar_arrive_expect_tx(full_bar(stage), kConvRows * BLOCK_K * 2 + C::kBBytes);
            const int tap = kb >> 2, c0 = (kb & 3) * BLOCK_K;
            const int kf = tap / 3, kt = tap - 3 * kf;
            tma_load_5d(a_dst, &tmap_a, full_bar(stage), c0, kf >> 1, conv_t0 + (kt >> 1),
                        (kt & 1) * 2 + (kf & 1), conv_b);
            tma_load_2d(b_dst, &tmap_b, full_bar(stage), c0, tap * 256 + n_tile * BLOCK_N);
          } else {
            mbar_arrive_expect_tx(full_bar(stage), C::kStageBytes);
            tma_load_2d(a_dst, &tmap_a, full_bar(stage), kb * BLOCK_K, m_tile * BLOCK_M);
            tma_load_2d(b_dst, &tmap_b, full_bar(stage), kb * BLOCK_K, n_tile * BLOCK_N);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
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
          const uint32_t a_addr = smem_base + stage * C::kStageBytes;
          const uint64_t a_desc = make_smem_desc_sw128(a_addr);
          const uint64_t b_desc = make_smem_desc_sw128(a_addr + C::kABytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // +32 B per UMMA_K step inside the 128-B swizzle atom (descriptor address unit = 16 B)
            umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(empty_bar(stage));   // frees the smem slot when these MMAs retire
          if (kb == num_k_blocks - 1) umma_commit(tfull_bar(acc));
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int quarter = warp & 3;          // TMEM lane quarter this warp may read
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_tile = tile / num_n_tiles, n_tile = tile - m_tile * num_n_tiles;
      const int r_local = quarter * 32 + lane;
      int64_t row;
      bool row_ok;
      if (ep.conv) {
        const int conv_b = m_tile / ep.tiles_per_utt;
        const int t0 = (m_tile - conv_b * ep.tiles_per_utt) * kConvT;
        row = ((int64_t)conv_b * ep.t2_len + t0) * kConvF + r_local;
        row_ok = r_local < kConvRows && (t0 + r_local / kConvF) < ep.t2_len;
      } else {
        row = (int64_t)m_tile * BLOCK_M + r_local;
        row_ok = row < ep.m;
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * BLOCK_N + ((uint32_t)(quarter * 32) << 16);
      const int64_t rrow = (ep.resid && ep.resid_period > 0) ? row % ep.resid_period : row;
#pragma unroll 1
      for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
        uint32_t v[32];
        tmem_ld32(t_addr + ch * 32, v);
        tmem_ld_wait();
        const int col0 = n_tile * BLOCK_N + ch * 32;
        if (!row_ok || col0 >= ep.n) continue;
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        const bool full = col0 + 32 <= ep.n;
        if (ep.bias) {
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + i));
              f[i] += bv.x; f[i + 1] += bv.y; f[i + 2] += bv.z; f[i + 3] += bv.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (col0 + i < ep.n) f[i] += __ldg(ep.bias + col0 + i);
          }
        }
        if (ep.act == STAC_ACT_GELU_ERF) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = gelu_erf(f[i]);
        }
        if (ep.resid) {
          const float* rp = ep.resid + rrow * ep.n + col0;
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 rv = *reinterpret_cast<const float4*>(rp + i);
              f[i] += rv.x; f[i + 1] += rv.y; f[i + 2] += rv.z; f[i + 3] += rv.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (col0 + i < ep.n) f[i] += rp[i];
          }
        }
        if (ep.vt != nullptr && col0 >= ep.n - ep.vt_cols) {
          // V columns of a packed QKV projection -> V^T [B*H][64][t_pad] (keys contiguous)
          const int64_t b = row / ep.seq_len, t = row - b * ep.seq_len;
          const int vcol = (int)(col0 - (ep.n - ep.vt_cols));
          const int64_t heads = ep.vt_cols >> 6;
          __nv_bfloat16* dst = ep.vt + ((b * heads + (vcol >> 6)) * 64 + (vcol & 63)) * ep.t_pad + t;
#pragma unroll
          for (int i = 0; i < 32; ++i) dst[(int64_t)i * ep.t_pad] = __float2bfloat16_rn(f[i]);
        } else if (ep.c_bf16) {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(ep.c) + row * ep.n + col0;
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              *reinterpret_cast<uint4*>(dst + i) =
                  make_uint4(pack_bf16x2(f[i], f[i + 1]), pack_bf16x2(f[i + 2], f[i + 3]),
                             pack_bf16x2(f[i + 4], f[i + 5]), pack_bf16x2(f[i + 6], f[i + 7]));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (col0 + i < ep.n) dst[i] = __float2bfloat16_rn(f[i]);
          }
        } else {
          float* dst = reinterpret_cast<float*>(ep.c) + row * ep.n + col0;
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(dst + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (col0 + i < ep.n) dst[i] = f[i];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BLOCK_N>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const EpiParams& ep, int m_tiles, int n_tiles,
           int k_blocks, cudaStream_t st) {
  using C = Cfg<BLOCK_N>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  const int grid = std::min(m_tiles * n_tiles, num_sms());
  gemm_bf16_kernel<BLOCK_N><<<grid, kThreads, C::kSmemBytes, st>>>(ta, tb, ep, m_tiles, n_tiles, k_blocks);
  STAC_LAUNCH_CHECK();
}

}  // namespace

extern "C" int stac_gemm_bf16(const uint16_t* a, const uint16_t* w, const float* bias, const float* resid,
                              int64_t resid_period, int act, void* c, int c_dtype, int64_t m, int64_t n,
                              int64_t k, uint16_t* vt_out, int64_t vt_cols, int64_t seq_len, int64_t t_pad,
                              void* stream) {
  STAC_REQUIRE(a && w && c && m > 0 && n > 0 && k > 0 && resid_period >= 0);
  STAC_REQUIRE(act == STAC_ACT_NONE || act == STAC_ACT_GELU_ERF);
  STAC_REQUIRE(c_dtype == STAC_DT_F32 || c_dtype == STAC_DT_BF16);
  if (k % BLOCK_K != 0 || n % 8 != 0 || m >= (1ll << 31) - 256) return STAC_ERR_UNSUPPORTED_SHAPE;
  if (vt_out) {
    STAC_REQUIRE(vt_cols > 0 && vt_cols % 64 == 0 && vt_cols <= n && (n - vt_cols) % 32 == 0);
    STAC_REQUIRE(seq_len > 0 && t_pad >= seq_len && t_pad % 8 == 0 && m % seq_len == 0);
  }
  const int block_n = (n % 256 == 0 || n > 1024) ? 256 : 128;
  CUtensorMap ta, tb;
  {
    const uint64_t dims[2] = {(uint64_t)k, (uint64_t)m};
    const uint64_t str[1] = {(uint64_t)k * 2};
    const uint32_t box[2] = {BLOCK_K, BLOCK_M};
    int r = encode_bf16_map(&ta, a, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)k, (uint64_t)n};
    const uint64_t str[1] = {(uint64_t)k * 2};
    const uint32_t box[2] = {BLOCK_K, (uint32_t)block_n};
    int r = encode_bf16_map(&tb, w, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  EpiParams ep{};
  ep.bias = bias; ep.resid = resid; ep.resid_period = resid_period; ep.act = act;
  ep.c = c; ep.c_bf16 = c_dtype == STAC_DT_BF16; ep.m = m; ep.n = n;
  ep.vt = reinterpret_cast<__nv_bfloat16*>(vt_out); ep.vt_cols = vt_cols; ep.seq_len = seq_len; ep.t_pad = t_pad;
  ep.conv = 0; ep.t2_len = 0; ep.tiles_per_utt = 1;
  const int m_tiles = (int)ceil_div64(m, BLOCK_M), k_blocks = (int)(k / BLOCK_K);
  if (block_n == 256)
    return launch<256>(ta, tb, ep, m_tiles, (int)ceil_div64(n, 256), k_blocks, as_stream(stream));
  return launch<128>(ta, tb, ep, m_tiles, (int)ceil_div64(n, 128), k_blocks, as_stream(stream));
}

extern "C" int stac_conv1_bf16(const uint16_t* xpad, const uint16_t* w1_packed, const float* b1,
                               int64_t batch, int64_t t1, float* out, void* stream) {
  STAC_REQUIRE(xpad && w1_packed && b1 && out && batch > 0 && t1 >= 2 && t1 < (1 << 30));
  const int64_t t2 = (t1 - 1) / 2 + 1, tp2 = (t1 + 3) / 2;
  const int tiles_per_utt = (int)ceil_div64(t2, kConvT);
  if (batch * tiles_per_utt >= (1ll << 31)) return STAC_ERR_UNSUPPORTED_SHAPE;
  CUtensorMap ta, tb;
  {
    // [B][4 planes][Tp2][21][256] bf16, innermost first
    const uint64_t dims[5] = {256, 21, (uint64_t)tp2, 4, (uint64_t)batch};
    const uint64_t str[4] = {256 * 2, 21 * 256 * 2, (uint64_t)tp2 * 21 * 256 * 2, (uint64_t)4 * tp2 * 21 * 256 * 2};
    const uint32_t box[5] = {BLOCK_K, kConvF, kConvT, 1, 1};
    int r = encode_bf16_map(&ta, xpad, 5, dims, str, box);
    if (r != STAC_OK) return r;
  }
  {
    const uint64_t dims[2] = {256, 9 * 256};
    const uint64_t str[1] = {256 * 2};
    const uint32_t box[2] = {BLOCK_K, 256};
    int r = encode_bf16_map(&tb, w1_packed, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  EpiParams ep{};
  ep.bias = b1; ep.c = out; ep.c_bf16 = 0; ep.m = batch * t2 * kConvF; ep.n = 256;
  ep.conv = 1; ep.t2_len = (int)t2; ep.tiles_per_utt = tiles_per_utt;
  return launch<256>(ta, tb, ep, (int)(batch * tiles_per_utt), 1, 36, as_stream(stream));
}
