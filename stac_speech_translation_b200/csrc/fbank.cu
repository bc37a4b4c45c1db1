// a2/a3: STFT (hamming-400, hop 160) -> power -> mel-80 -> dB, top-dB clamp, global normalisation.
//
// Reference behaviour: SpeechBrain Fbank as configured at
//   /root/reference/stac-st/hparams/transformer_multitask.yaml:299-302, called at
//   /root/reference/stac-st/inference.py:95-96.
//
// The 400-point real DFT of each frame is computed in fp32 on the CUDA cores as a 200-point
// complex FFT of the even/odd-packed frame (200 = 25 x 8, Stockham autosort: one radix-25 pass
// done as 5x5 in registers, one twiddle-free radix-8 pass in place) followed by the real-FFT
// split.  A dense DFT GEMM would be 335 FLOP per HBM byte - above the tensor ridge - so the
// stage could never be HBM-bound; the FFT needs ~30 FLOP/B.  32 frames per CTA share one
// coalesced PCM tile in shared memory (frames overlap 60 %, so PCM is read once from HBM).
#include "common.cuh"

namespace {

constexpr int kNfft = 400, kHop = 160, kBins = 201, kMel = 80, kMelTaps = 16;
constexpr int kFrames = 32;                       // frames per CTA
constexpr int kThreads = 256;
constexpr int kPcmTile = kHop * (kFrames - 1) + kNfft;  // 5360 samples
constexpr int kPStride = 204;

// table layout (floats)
constexpr int kOffWindow = 0;
constexpr int kOffTw200 = kOffWindow + 400;       // float2[200]  exp(-2 pi i t/200)
constexpr int kOffTw25 = kOffTw200 + 400;         // float2[25]   exp(-2 pi i t/25)
constexpr int kOffTw400 = kOffTw25 + 50;          // float2[201]  exp(-2 pi i k/400)
constexpr int kOffMelStart = kOffTw400 + 402;     // float[80] (integers)
constexpr int kOffMelCount = kOffMelStart + kMel; // float[80]
constexpr int kOffMelW = kOffMelCount + kMel;     // float[80][16]
constexpr int kTableFloats = kOffMelW + kMel * kMelTaps;

struct Smem {
  float2 y[kFrames][200];
  float p[kFrames][kPStride];
  float pcm[kPcmTile + 8];
  float tab[kTableFloats];
  float red[40];
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// forward 5-point DFT, in place over v[0], v[s], v[2s], v[3s], v[4s]
template <int S>
__device__ __forceinline__ void dft5(float2* v) {
  constexpr float C1 = 0.30901699437494742f, C2 = -0.80901699437494742f;
  constexpr float S1 = 0.95105651629515357f, S2 = 0.58778525229247313f;
  const float2 a0 = v[0], a1 = v[S], a2 = v[2 * S], a3 = v[3 * S], a4 = v[4 * S];
  const float2 t1 = cadd(a1, a4), t2 = cadd(a2, a3), t3 = csub(a1, a4), t4 = csub(a2, a3);
  const float2 m1 = make_float2(a0.x + C1 * t1.x + C2 * t2.x, a0.y + C1 * t1.y + C2 * t2.y);
  const float2 m2 = make_float2(a0.x + C2 * t1.x + C1 * t2.x, a0.y + C2 * t1.y + C1 * t2.y);
  const float2 n1 = make_float2(S1 * t3.x + S2 * t4.x, S1 * t3.y + S2 * t4.y);
  const float2 n2 = make_float2(S2 * t3.x - S1 * t4.x, S2 * t3.y - S1 * t4.y);
  v[0] = make_float2(a0.x + t1.x + t2.x, a0.y + t1.y + t2.y);
  // b = m - i*n  ->  (m.x + n.y, m.y - n.x);   b' = m + i*n -> (m.x - n.y, m.y + n.x)
  v[S] = make_float2(m1.x + n1.y, m1.y - n1.x);
  v[4 * S] = make_float2(m1.x - n1.y, m1.y + n1.x);
  v[2 * S] = make_float2(m2.x + n2.y, m2.y - n2.x);
  v[3 * S] = make_float2(m2.x - n2.y, m2.y + n2.x);
}

__device__ __forceinline__ void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
  const float2 e0 = cadd(x0, x2), e1 = csub(x0, x2), o0 = cadd(x1, x3), d = csub(x1, x3);
  const float2 o1 = make_float2(d.y, -d.x);  // (x1-x3) * (-i)
  x0 = cadd(e0, o0); x1 = cadd(e1, o1); x2 = csub(e0, o0); x3 = csub(e1, o1);
}

__global__ void __launch_bounds__(kThreads, 2)
fbank_logmel_kernel(const float* __restrict__ pcm, int64_t n_samples, int64_t row_stride,
                    int64_t n_frames, const float* __restrict__ tables,
                    float* __restrict__ out, unsigned int* __restrict__ utt_max) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int64_t t0 = (int64_t)blockIdx.x * kFrames;
  const float* row = pcm + (int64_t)b * row_stride;

  for (int i = tid; i < kTableFloats; i += kThreads) s.tab[i] = __ldg(tables + i);
  const int64_t g0 = t0 * kHop - kNfft / 2;
  for (int i = tid; i < kPcmTile; i += kThreads) {
    const int64_t g = g0 + i;
    s.pcm[i] = (g >= 0 && g < n_samples) ? __ldg(row + g) : 0.f;
  }
  __syncthreads();

  const float* win = s.tab + kOffWindow;
  const float2* tw200 = reinterpret_cast<const float2*>(s.tab + kOffTw200);
  const float2* tw25 = reinterpret_cast<const float2*>(s.tab + kOffTw25);
  const float2* tw400 = reinterpret_cast<const float2*>(s.tab + kOffTw400);

  // ---- pass A: radix-25 (n=200, stride 1, m=8): thread = (frame f, p) ----
  {
    const int f = tid >> 3, p = tid & 7;
    float2 a[25];
    const float* x = s.pcm + f * kHop;
#pragma unroll
    for (int j = 0; j < 25; ++j) {
      const int n = p + 8 * j;
      const float2 xv = *reinterpret_cast<const float2*>(x + 2 * n);
      a[j] = make_float2(xv.x * win[2 * n], xv.y * win[2 * n + 1]);
    }
    // a[5*j1 + j2]: 5-point DFT over j1 for each j2 (stride 5), result index k1 replaces j1
#pragma unroll
    for (int j2 = 0; j2 < 5; ++j2) dft5<5>(a + j2);
#pragma unroll
    for (int k1 = 1; k1 < 5; ++k1)
#pragma unroll
      for (int j2 = 1; j2 < 5; ++j2) a[5 * k1 + j2] = cmul(a[5 * k1 + j2], tw25[j2 * k1]);
    // 5-point DFT over j2 for each k1 (stride 1): output k2 at a[5*k1 + k2] -> b[k1 + 5*k2]
#pragma unroll
    for (int k1 = 0; k1 < 5; ++k1) dft5<1>(a + 5 * k1);
#pragma unroll
    for (int k1 = 0; k1 < 5; ++k1)
#pragma unroll
      for (int k2 = 0; k2 < 5; ++k2) {
        const int k = k1 + 5 * k2;
        s.y[f][25 * p + k] = cmul(a[5 * k1 + k2], tw200[p * k]);
      }
  }
  __syncthreads();

  // ---- pass B: radix-8 (n=8, stride 25), in place, no twiddles ----
  for (int idx = tid; idx < kFrames * 25; idx += kThreads) {
    const int f = idx / 25, q = idx - 25 * f;
    float2* y = &s.y[f][q];
    float2 v0 = y[0], v1 = y[25], v2 = y[50], v3 = y[75], v4 = y[100], v5 = y[125], v6 = y[150], v7 = y[175];
    constexpr float R = 0.70710678118654752440f;
    float2 s0 = cadd(v0, v4), s1 = cadd(v1, v5), s2 = cadd(v2, v6), s3 = cadd(v3, v7);
    float2 d0 = csub(v0, v4), d1 = csub(v1, v5), d2 = csub(v2, v6), d3 = csub(v3, v7);
    d1 = make_float2(R * (d1.x + d1.y), R * (d1.y - d1.x));    // * (R - iR)
    d2 = make_float2(d2.y, -d2.x);                              // * (-i)
    d3 = make_float2(R * (d3.y - d3.x), -R * (d3.x + d3.y));   // * (-R - iR)
    dft4(s0, s1, s2, s3);
    dft4(d0, d1, d2, d3);
    y[0] = s0; y[50] = s1; y[100] = s2; y[150] = s3;
    y[25] = d0; y[75] = d1; y[125] = d2; y[175] = d3;
  }
  __syncthreads();

  // ---- real-FFT split + power: item = (frame, k in 0..100) gives bins k and 200-k ----
  for (int idx = tid; idx < kFrames * 101; idx += kThreads) {
    const int f = idx / 101, k = idx - 101 * f;
    const float2 zk = s.y[f][k == 200 ? 0 : k];
    const float2 zr = s.y[f][(200 - k) % 200];
    {
      const float2 zc = make_float2(zr.x, -zr.y);
      const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
      const float2 dd = csub(zk, zc);
      const float2 o = make_float2(0.5f * dd.y, -0.5f * dd.x);  // (zk - zc) / (2i)
      const float2 xk = cadd(e, cmul(tw400[k], o));
      s.p[f][k] = xk.x * xk.x + xk.y * xk.y;
    }
    {
      const int k2 = 200 - k;
      const float2 zc = make_float2(zk.x, -zk.y);
      const float2 e = make_float2(0.5f * (zr.x + zc.x), 0.5f * (zr.y + zc.y));
      const float2 dd = csub(zr, zc);
      const float2 o = make_float2(0.5f * dd.y, -0.5f * dd.x);
      const float2 xk = cadd(e, cmul(tw400[k2], o));
      s.p[f][k2] = xk.x * xk.x + xk.y * xk.y;
    }
  }
  __syncthreads();

  // ---- mel (sparse triangles) + dB + running max ----
  const float* mstart = s.tab + kOffMelStart;
  const float* mcount = s.tab + kOffMelCount;
  const float* mw = s.tab + kOffMelW;
  float vmax = -INFINITY;
  for (int idx = tid; idx < kFrames * kMel; idx += kThreads) {
    const int f = idx / kMel, m = idx - kMel * f;
    const int64_t t = t0 + f;
    if (t >= n_frames) continue;
    const int st = (int)mstart[m], cnt = (int)mcount[m];
    float acc = 0.f;
    for (int i = 0; i < cnt; ++i) acc = fmaf(s.p[f][st + i], mw[m * kMelTaps + i], acc);
    const float db = 10.0f * log10f(fmaxf(acc, 1e-10f));
    out[((int64_t)b * n_frames + t) * kMel + m] = db;
    vmax = fmaxf(vmax, db);
  }
  vmax = block_max(vmax, s.red);
  if (tid == 0 && vmax > -INFINITY) atomicMax(utt_max + b, float_to_ordered(vmax));
}

// (x and out may be the same buffer - the in-place call of ops.fbank - so neither is __restrict__)
__global__ void topdb_norm_kernel(const float* x, const unsigned int* __restrict__ utt_max,
                                  int per_utt, float top_db, const float* __restrict__ mean,
                                  const float* __restrict__ stdv, int64_t batch, int64_t per_row,
                                  int n_mels, float* out) {
  // one CTA-row of work per (b, chunk); per_row = frames * n_mels
  const int b = blockIdx.y;
  float mx;
  if (per_utt) {
    mx = ordered_to_float(utt_max[b]);
  } else {
    mx = -INFINITY;
    for (int i = 0; i < batch; ++i) mx = fmaxf(mx, ordered_to_float(utt_max[i]));
  }
  const float floor_db = mx - top_db;
  const float* xr = x + (int64_t)b * per_row;
  float* orow = out + (int64_t)b * per_row;
  // 16-byte path: four consecutive mel bins per thread (HBM-bound: 4-byte accesses and a 64-bit modulo per element left
  // this kernel at 2.5 TB/s); the scalar loop below takes whatever does not fit it
  const bool vec = (per_row & 3) == 0 && (n_mels & 3) == 0 && per_row < (1ll << 31) &&
                   (((uintptr_t)xr | (uintptr_t)orow | (uintptr_t)mean | (uintptr_t)stdv) & 15) == 0;
  if (vec) {
    const unsigned n4 = (unsigned)(per_row >> 2), m4n = (unsigned)n_mels >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(xr);
    float4* o4 = reinterpret_cast<float4*>(orow);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
      float4 v = x4[i];
      v.x = fmaxf(v.x, floor_db); v.y = fmaxf(v.y, floor_db); v.z = fmaxf(v.z, floor_db); v.w = fmaxf(v.w, floor_db);
      if (mean != nullptr) {
        const unsigned m4 = i % m4n;
        const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + m4);
        const float4 sd = __ldg(reinterpret_cast<const float4*>(stdv) + m4);
        v.x = (v.x - mu.x) / sd.x; v.y = (v.y - mu.y) / sd.y; v.z = (v.z - mu.z) / sd.z; v.w = (v.w - mu.w) / sd.w;
      }
      o4[i] = v;
    }
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_row;
       i += (int64_t)gridDim.x * blockDim.x) {
    float v = fmaxf(xr[i], floor_db);
    if (mean != nullptr) {
      const int m = (int)(i % n_mels);
      v = (v - mean[m]) / stdv[m];
    }
    orow[i] = v;
  }
}

__global__ void input_norm_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                  const float* __restrict__ stdv, int64_t total, int n_mels,
                                  float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i % n_mels);
    out[i] = (x[i] - mean[m]) / stdv[m];
  }
}

}  // namespace

extern "C" int stac_fbank_tables_floats(void) { return kTableFloats; }

extern "C" int stac_fbank_logmel(const float* pcm, int64_t batch, int64_t n_samples,
                                 int64_t pcm_row_stride, const float* tables, float* logmel_db,
                                 uint32_t* utt_max_ordered, void* stream) {
  STAC_REQUIRE(pcm && tables && logmel_db && utt_max_ordered);
  STAC_REQUIRE(batch > 0 && batch < 65536 && n_samples > 0 && pcm_row_stride >= n_samples);
  const int64_t n_frames = 1 + n_samples / kHop;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fbank_logmel_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  cudaError_t me = cudaMemsetAsync(utt_max_ordered, 0, (size_t)batch * sizeof(uint32_t), as_stream(stream));
  if (me != cudaSuccess) return (int)me;
  dim3 grid((unsigned)ceil_div64(n_frames, kFrames), (unsigned)batch);
  fbank_logmel_kernel<<<grid, kThreads, sizeof(Smem), as_stream(stream)>>>(
      pcm, n_samples, pcm_row_stride, n_frames, tables, logmel_db, utt_max_ordered);
  STAC_LAUNCH_CHECK();
}

extern "C" int stac_fbank_topdb_norm(const float* logmel_db, const uint32_t* utt_max_ordered,
                                     int per_utterance, float top_db, const float* mean,
                                     const float* std, int64_t batch, int64_t frames,
                                     int64_t n_mels, float* out, void* stream) {
  STAC_REQUIRE(logmel_db && utt_max_ordered && out && batch > 0 && batch < 65536 && frames > 0);
  STAC_REQUIRE((mean == nullptr) == (std == nullptr) && n_mels > 0);
  const int64_t per_row = frames * n_mels;
  dim3 grid((unsigned)std::min<int64_t>(ceil_div64(per_row, 256 * 4), 4096), (unsigned)batch);
  topdb_norm_kernel<<<grid, 256, 0, as_stream(stream)>>>(logmel_db, utt_max_ordered, per_utterance,
                                                         top_db, mean, std, batch, per_row,
                                                         (int)n_mels, out);
  STAC_LAUNCH_CHECK();
}

extern "C" int stac_input_norm(const float* x, const float* mean, const float* std, int64_t rows,
                               int64_t n_mels, float* out, void* stream) {
  STAC_REQUIRE(x && mean && std && out && rows > 0 && n_mels > 0);
  const int64_t total = rows * n_mels;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(total, 256 * 4), 148 * 16);
  input_norm_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, mean, std, total, (int)n_mels, out);
  STAC_LAUNCH_CHECK();
}
