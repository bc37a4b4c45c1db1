// a2 in bf16 mode: STFT as a tensor-core GEMM fused with power, mel filterbank, dB and the per-utterance maximum.
// Reference behaviour: SpeechBrain Fbank as configured at
//   /root/reference/stac-st/hparams/transformer_multitask.yaml:299-302, called at stac-st/inference.py:95.
//
// The 400-point real DFT of a hamming-windowed frame y is folded on its symmetry before it reaches the tensor core:
//   Re X[k] =  sum_{n=0}^{200} e[n] cos(2 pi k n / 400),   e[0] = y[0], e[200] = y[200], e[n] = y[n] + y[400-n]
//   Im X[k] = -sum_{n=1}^{199} o[n] sin(2 pi k n / 400),   o[n] = y[n] - y[400-n]
// (the periodic hamming window is itself symmetric, w[400-n] = w[n], so e[n] = w[n] (x[n] + x[400-n]) etc.).  One tile =
// 128 consecutive frames of one utterance:  D_cos[128 x 208] = E[128 x 208] . C^T,  D_sin = O . S^T  with fp16 operands
// and fp32 accumulation in tensor memory - 21.7 MFLOP per tile instead of 41 for the unfolded DFT, 33 GFLOP for the
// benchmark batch, i.e. ~25 us of tensor time: the stage is bound by its HBM traffic (PCM in, features out), which is the
// point.  fp16 (not bf16) operands: 11-bit mantissas keep the log-mel error at 3.6e-4 relative (bf16: 2.5e-3 and 9 dB in
// quiet bins); the fp32 mode of the library keeps the exact CUDA-core FFT kernel (fbank.cu).
//
// Pipeline (persistent, one CTA per SM, 10 warps): warps 4-7 stage the tile's PCM in smem once and build the A k-blocks
// from it (window, fold, fp16, 128-byte swizzled smem rows), warp 8 streams the matching twiddle k-blocks with TMA, warp 9 issues
// tcgen05.mma, warps 0-3 are the epilogue: thread = frame, power = re^2 + im^2 straight from TMEM, the 80 mel
// accumulators live in registers (every DFT bin feeds at most two filters; the structure is compile-time, the weights are
// runtime data), 10 log10 via lg2, row maximum -> atomicMax, rows staged in smem and written out coalesced.
#include <algorithm>
#include <utility>
#include "tc_common.cuh"
#include "fbank_mel_structure.h"

namespace {

using namespace tc;

constexpr int kNfft = 400, kHop = 160, kMel = 80;
constexpr int kRows = 128;                 // frames per tile
constexpr int kBins = 208;                 // 201 bins padded to a multiple of 16 (UMMA N)
constexpr int kKb = 8;                     // k-blocks per tile: 4 cos (e[0..207]) + 4 sin (o[1..208])
constexpr int kAStages = 2;                // A k-blocks (built by the producer warps) in flight
constexpr int kBStages = 3;                // twiddle k-blocks (TMA from L2, ~1.5 k cycles each) in flight
constexpr int kABytes = kRows * 128;       // [128 frames x 64 fp16]
constexpr int kBBytes = kBins * 128;       // [208 bins x 64 fp16]
constexpr int kProdWarps = 8;
constexpr int kThreads = (4 + kProdWarps + 2) * 32;       // 4 epilogue + producers + TMA + MMA warps
constexpr int kOutStride = 81;             // padded row of the output staging tile (bank-conflict free)

constexpr int kPcmTile = (kRows - 1) * kHop + kNfft;     // 20720 samples feed the 128 frames of a tile
constexpr int kOffA = 0;
constexpr int kOffB = kAStages * kABytes;
constexpr int kOffPcm = kOffB + kBStages * kBBytes;               // float [kPcmTile] (zero outside the utterance)
constexpr int kOffOut = kOffPcm + ((kPcmTile * 4 + 127) / 128) * 128;   // float [64][81]: output staging, half a tile at a time
constexpr int kOffTab = kOffOut + (kRows / 2) * kOutStride * 4;   // window[400] | wbin[208][2]
constexpr int kTabFloats = 400 + kBins * 2;
constexpr int kOffBar = ((kOffTab + kTabFloats * 4 + 15) / 16) * 16;
constexpr int kNumBars = 2 * kAStages + 2 * kBStages + 4;
constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16 + 1024;

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
// D[tmem] (+)= A[smem] . B[smem]^T ; fp16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
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

// one DFT bin (compile-time K) into the mel accumulators it feeds: the structure is a constant expression, so the
// accumulator indices are register names
template <int K>
__device__ __forceinline__ void mel_bin(uint32_t re, uint32_t im, const float* __restrict__ wbin, float (&acc)[kMel]) {
  constexpr int first = kMelFirst[K], cnt = kMelCnt[K];
  if constexpr (cnt > 0) {
    const float a = __uint_as_float(re), b = __uint_as_float(im);
    const float p = fmaf(a, a, b * b);
    acc[first] = fmaf(p, wbin[2 * K], acc[first]);
    if constexpr (cnt > 1) acc[first + 1] = fmaf(p, wbin[2 * K + 1], acc[first + 1]);
  }
}
// 16 consecutive bins starting at compile-time bin K0
template <int K0, int... I>
__device__ __forceinline__ void mel_chunk(const uint32_t (&re)[16], const uint32_t (&im)[16], const float* __restrict__ wbin,
                                          float (&acc)[kMel], std::integer_sequence<int, I...>) {
  (mel_bin<K0 + I>(re[I], im[I], wbin, acc), ...);
}

template <int C>
__device__ __forceinline__ void mel_all(uint32_t t_cos, uint32_t t_sin, const float* __restrict__ wbin, float (&acc)[kMel]) {
  if constexpr (C < kBins / 16) {
    uint32_t re[16], im[16];
    tmem_ld16(t_cos + C * 16, re);
    tmem_ld16(t_sin + C * 16, im);
    tmem_ld_wait();
    mel_chunk<C * 16>(re, im, wbin, acc, std::make_integer_sequence<int, 16>{});
    mel_all<C + 1>(t_cos, t_sin, wbin, acc);
  }
}

#ifdef FBANK_TRACE    // timing experiment only (tools/trace_fbank.py): CTA 0 logs clock32 per (role, tile ordinal, event)
__device__ unsigned int* g_fb_trace = nullptr;
#define BTRACE(role, ev, n) do { if (blockIdx.x == 0 && g_fb_trace != nullptr && (n) < 8) g_fb_trace[((role) * 8 + (n)) * 16 + (ev)] = (unsigned int)clock64(); } while (0)
#else
#define BTRACE(role, ev, n) do {} while (0)
#endif

__global__ void __launch_bounds__(kThreads, 1)
fbank_tc_kernel(const __grid_constant__ CUtensorMap tmap_tw, const float* __restrict__ pcm, int64_t n_samples,
                int64_t row_stride, int n_frames, int tiles_per_utt, int num_tiles, const float* __restrict__ tables,
                float* __restrict__ out, unsigned int* __restrict__ utt_max) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bars = sbase + kOffBar;
  auto a_full = [&](int s) { return bars + 8u * s; };
  auto a_empty = [&](int s) { return bars + 8u * (kAStages + s); };
  auto b_full = [&](int s) { return bars + 8u * (2 * kAStages + s); };
  auto b_empty = [&](int s) { return bars + 8u * (2 * kAStages + kBStages + s); };
  const uint32_t misc = bars + 8u * (2 * kAStages + 2 * kBStages);
  auto tfull_bar = [&]() { return misc; };
  auto tempty_bar = [&]() { return misc + 8u; };
  auto pcm_full = [&]() { return misc + 16u; };
  auto pcm_free = [&]() { return misc + 24u; };
  const uint32_t tmem_slot = bars + 8u * kNumBars;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* tab = reinterpret_cast<float*>(sptr + kOffTab);
  const float* win = tab;
  const float* wbin = tab + 400;

  if (tid == 0) {
    prefetch_tmap(&tmap_tw);
    for (int s = 0; s < kAStages; ++s) { mbar_init(a_full(s), kProdWarps); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < kBStages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    mbar_init(pcm_full(), 1);
    mbar_init(pcm_free(), kProdWarps);
    mbar_init(tfull_bar(), 1);
    mbar_init(tempty_bar(), 4);
    fence_barrier_init();
  }
  if (warp == 4 + kProdWarps + 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  for (int i = tid; i < kTabFloats; i += kThreads) tab[i] = __ldg(tables + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp >= 4 && warp < 4 + kProdWarps) {
    // ============================ A producers: window, fold, fp16 from the PCM tile in smem ============================
    // The 128 frames of a tile overlap 60 %: their 20720 samples are fetched ONCE into shared memory (one bulk async copy
    // issued by the TMA warp; samples outside the utterance are zero-filled here) and every k-block is built from there.
    // Warp pw builds rows pw, pw + 8, ...; lane l owns columns 2l, 2l + 1 of the 64-wide block.
    const int pw = warp - 4, ptid = tid - 128;
    float* pcm_s = reinterpret_cast<float*>(sptr + kOffPcm);
    int stage = 0;
    uint32_t phase = 0;
    int tn = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tn) {
      const int b = tile / tiles_per_utt;
      const int t0 = (tile - b * tiles_per_utt) * kRows;
      const int64_t g0 = (int64_t)t0 * kHop - kNfft / 2;           // global sample index of pcm_s[0]
      const int head = g0 < 0 ? (int)(-g0) : 0;                    // samples before the utterance
      const int64_t tail0 = n_samples - g0;                        // first index past the utterance
      if (ptid == 0) BTRACE(0, 0, tn);
      if (head > 0 || tail0 < kPcmTile) {
        // (the previous tile is finished by every producer warp: each of them arrived on pcm_free before the copy of
        //  this tile could start, and the copy never touches the ranges zeroed here)
        asm volatile("bar.sync 2, 256;" ::: "memory");
        for (int i = ptid; i < head; i += kProdWarps * 32) pcm_s[i] = 0.f;
        for (int i = (int)max((int64_t)0, tail0) + ptid; i < kPcmTile; i += kProdWarps * 32) pcm_s[i] = 0.f;
        asm volatile("bar.sync 2, 256;" ::: "memory");
      }
      mbar_wait(pcm_full(), (uint32_t)tn & 1);
      if (ptid == 0) BTRACE(0, 1, tn);
      for (int kb = 0; kb < kKb; ++kb) {
        const bool sin_part = kb >= 4;
        const int j0 = (kb & 3) * 64 + 2 * lane;          // column inside the part
        const int n0 = j0 + (sin_part ? 1 : 0);            // sample index inside the frame of column j0 (n0 + 1 for j0 + 1)
        // window values of my two columns (0 outside the folded range)
        const bool v0 = sin_part ? (n0 <= 199) : (n0 <= 200);
        const bool v1 = sin_part ? (n0 + 1 <= 199) : (n0 + 1 <= 200);
        const float w0 = v0 ? win[n0] : 0.f, w1 = v1 ? win[n0 + 1] : 0.f;
        // mirror partner exists for 1 <= n <= 199; clamp the indices of dead columns into the tile
        const bool m0 = n0 >= 1 && n0 <= 199, m1 = n0 + 1 >= 1 && n0 + 1 <= 199;
        const int f0 = v0 ? n0 : 0, f1 = v1 ? n0 + 1 : 0;
        const int q0 = m0 ? kNfft - n0 : 0, q1 = m1 ? kNfft - n0 - 1 : 0;
        const float sgn = sin_part ? -1.f : 1.f;
        const float wm0 = m0 ? sgn * w0 : 0.f, wm1 = m1 ? sgn * w1 : 0.f;
        mbar_wait(a_empty(stage), phase ^ 1);
        const uint32_t a_dst = sbase + kOffA + stage * kABytes;
        const uint32_t col_off = ((uint32_t)(2 * lane) >> 3), in_chunk = ((2 * lane) & 7) * 2;
        // 16 rows per warp: all shared-memory loads first (no store in between, so they are all in flight together),
        // then the 16 stores
        uint32_t pk[kRows / kProdWarps];
#pragma unroll
        for (int i = 0; i < kRows / kProdWarps; ++i) {
          const float* fr = pcm_s + (pw + i * kProdWarps) * kHop;      // first sample of frame r
          const float y0 = fmaf(wm0, fr[q0], w0 * fr[f0]);             // w (x[n] +- x[400 - n])
          const float y1 = fmaf(wm1, fr[q1], w1 * fr[f1]);
          pk[i] = pack_f16x2(y0, y1);
        }
        // row r = pw + 8 i: 4 bytes at column j0 of the 128-byte swizzled row (r & 7 = pw)
        const uint32_t row0 = a_dst + pw * 128 + ((col_off ^ (uint32_t)pw) << 4) + in_chunk;
#pragma unroll
        for (int i = 0; i < kRows / kProdWarps; ++i)
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(row0 + i * kProdWarps * 128), "r"(pk[i]) : "memory");
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full(stage));
        if (ptid == 0) BTRACE(0, 2 + kb, tn);
        if (++stage == kAStages) { stage = 0; phase ^= 1; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(pcm_free());          // this warp no longer reads the PCM tile
    }
  } else if (warp == 4 + kProdWarps) {
    // ============================ TMA: PCM tile (one bulk copy) + twiddle k-blocks ============================
    int stage = 0;
    uint32_t phase = 0;
    int tn = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tn) {
      {
        const int b = tile / tiles_per_utt;
        const int t0 = (tile - b * tiles_per_utt) * kRows;
        const int64_t g0 = (int64_t)t0 * kHop - kNfft / 2;
        const int64_t lo = g0 < 0 ? 0 : g0;
        const int64_t hi = min(g0 + (int64_t)kPcmTile, n_samples);      // > lo for every tile of the grid
        mbar_wait(pcm_free(), ((uint32_t)tn & 1) ^ 1);
        if (elect_one()) {
          const uint32_t bytes = (uint32_t)(hi - lo) * 4;                // multiple of 16: n_samples % 4 == 0 (checked on the host)
          mbar_arrive_expect_tx(pcm_full(), bytes);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(sbase + kOffPcm + (uint32_t)(lo - g0) * 4),
                         "l"(reinterpret_cast<uint64_t>(pcm + (int64_t)b * row_stride + lo)), "r"(bytes), "r"(pcm_full())
                       : "memory");
        }
        __syncwarp();
      }
      for (int kb = 0; kb < kKb; ++kb) {
        mbar_wait(b_empty(stage), phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(b_full(stage), kBBytes);
          tma_load_2d(sbase + kOffB + stage * kBBytes, &tmap_tw, b_full(stage), (kb & 3) * 64, (kb >> 2) * kBins);
        }
        __syncwarp();
        if (++stage == kBStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 4 + kProdWarps + 1) {
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc = make_idesc_f16(kRows, kBins);
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    int n_done = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n_done) {
      mbar_wait(tempty_bar(), ((uint32_t)n_done & 1) ^ 1);
      if (lane == 0) BTRACE(1, 0, n_done);
      tc_fence_after();
      for (int kb = 0; kb < kKb; ++kb) {
        mbar_wait(b_full(sb), pb);
        mbar_wait(a_full(sa), pa);
        if (lane == 0) BTRACE(1, 1 + kb, n_done);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = make_smem_desc_sw128(sbase + kOffA + sa * kABytes);
          const uint64_t bd = make_smem_desc_sw128(sbase + kOffB + sb * kBBytes);
          const uint32_t d = tmem_base + (kb >> 2) * 256;            // cos accumulator at column 0, sin at 256
          const int steps = (kb & 3) == 3 ? 1 : 4;                    // the last block of a part holds 16 real columns
          for (int k = 0; k < steps; ++k) umma_f16(d, ad + 2 * k, bd + 2 * k, idesc, ((kb & 3) | k) != 0);
          umma_commit(a_empty(sa));
          umma_commit(b_empty(sb));
          if (kb == kKb - 1) umma_commit(tfull_bar());
        }
        __syncwarp();
        if (++sa == kAStages) { sa = 0; pa ^= 1; }
        if (++sb == kBStages) { sb = 0; pb ^= 1; }
      }
    }
  } else {
    // ============================ epilogue: thread = frame ============================
    const int r = warp * 32 + lane;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    float* stage_out = reinterpret_cast<float*>(sptr + kOffOut);
    int n_done = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n_done) {
      const int b = tile / tiles_per_utt;
      const int t0 = (tile - b * tiles_per_utt) * kRows;
      const int valid = min(kRows, n_frames - t0);
      if (tid == 0) BTRACE(2, 0, n_done);
      mbar_wait(tfull_bar(), (uint32_t)n_done & 1);
      if (tid == 0) BTRACE(2, 1, n_done);
      tc_fence_after();
      float acc[kMel];
#pragma unroll
      for (int m = 0; m < kMel; ++m) acc[m] = 0.f;
      mel_all<0>(tmem_base + lane_off, tmem_base + 256 + lane_off, wbin, acc);
      // the accumulators are consumed: the next tile's MMAs may start
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar());
      if (tid == 0) BTRACE(2, 2, n_done);
      float vmax = -INFINITY;
#pragma unroll
      for (int m = 0; m < kMel; ++m) {
        acc[m] = 3.0102999566398120f * lg2_approx(fmaxf(acc[m], 1e-10f));     // 10 log10(x) = 10 log10(2) lg2(x)
        if (r < valid) vmax = fmaxf(vmax, acc[m]);
      }
      vmax = warp_max(vmax);
      if (lane == 0 && vmax > -INFINITY) atomicMax(utt_max + b, float_to_ordered(vmax));
      if (tid == 0) BTRACE(2, 3, n_done);
      // rows leave through a 64-row staging tile, half a tile at a time; [rows][80] floats are contiguous in the output
      float* dst = out + ((int64_t)b * n_frames + t0) * kMel;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        if ((r >> 6) == half) {
#pragma unroll
          for (int m = 0; m < kMel; ++m) stage_out[(r & 63) * kOutStride + m] = acc[m];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int n_out = (min(valid, half * 64 + 64) - half * 64) * kMel;      // <= 0 when the half is past the end
        for (int i = tid; i < n_out; i += 128) {
          const int rr = i / kMel, mm = i - rr * kMel;
          dst[half * 64 * kMel + i] = stage_out[rr * kOutStride + mm];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");          // the staging tile is rewritten next
      }
      if (tid == 0) BTRACE(2, 4, n_done);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4 + kProdWarps + 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace

#ifdef FBANK_TRACE
extern "C" int stac_fbank_trace(unsigned int* buf) { cudaMemcpyToSymbol(g_fb_trace, &buf, sizeof(buf)); return 0; }
#endif

extern "C" int stac_fbank_tc_tables_floats(void) { return kTabFloats; }
extern "C" int stac_fbank_tc_twiddle_halfs(void) { return 2 * kBins * 256; }

extern "C" int stac_fbank_logmel_tc(const float* pcm, int64_t batch, int64_t n_samples, int64_t pcm_row_stride,
                                    const float* tables, const uint16_t* twiddles, float* logmel_db,
                                    uint32_t* utt_max_ordered, void* stream) {
  STAC_REQUIRE(pcm && tables && twiddles && logmel_db && utt_max_ordered);
  STAC_REQUIRE(batch > 0 && batch < 65536 && n_samples > 0 && pcm_row_stride >= n_samples);
  // the PCM tile travels as one bulk async copy: 16-byte aligned rows and a sample count that is a multiple of 4
  if (n_samples % 4 != 0 || pcm_row_stride % 4 != 0 || (reinterpret_cast<uintptr_t>(pcm) & 15) != 0)
    return STAC_ERR_UNSUPPORTED_SHAPE;
  const int64_t n_frames = 1 + n_samples / kHop;
  const int64_t tiles_per_utt = ceil_div64(n_frames, kRows);
  if (n_frames >= (1ll << 30) || batch * tiles_per_utt >= (1ll << 30)) return STAC_ERR_UNSUPPORTED_SHAPE;
  CUtensorMap tw;
  {
    // twiddles fp16 [2 parts x 208 bins][256 columns]
    const uint64_t dims[2] = {256, (uint64_t)(2 * kBins)};
    const uint64_t str[1] = {256 * 2};
    const uint32_t box[2] = {64, (uint32_t)kBins};
    int r = encode_map(&tw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, twiddles, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(fbank_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  cudaError_t me = cudaMemsetAsync(utt_max_ordered, 0, (size_t)batch * sizeof(uint32_t), as_stream(stream));
  if (me != cudaSuccess) return (int)me;
  const int num_tiles = (int)(batch * tiles_per_utt);
  const int grid = std::min(num_tiles, stac_grid_limit());
  fbank_tc_kernel<<<grid, kThreads, kSmemBytes, as_stream(stream)>>>(tw, pcm, n_samples, pcm_row_stride, (int)n_frames,
                                                                    (int)tiles_per_utt, num_tiles, tables, logmel_db,
                                                                    utt_max_ordered);
  STAC_LAUNCH_CHECK();
}
