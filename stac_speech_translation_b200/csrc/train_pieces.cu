// Training-stage pieces that sit directly on the path's tensors: SpecAugment on the normalised
// features and the CTC negative log-likelihood of the posteriors.
//
// Reference behaviour replaced:
//   * `feats = self.hparams.augmentation(feats)` (/root/reference/stac-st/train_multitask.py:63-66) with SpeechBrain's
//     SpecAugment as configured at hparams/transformer_multitask.yaml:283-293: a bicubic time warp around a random centre
//     (two torch.nn.functional.interpolate calls with align_corners=True and three in-place assignments), then
//     n_freq_mask frequency masks and n_time_mask time masks (masked_fill_).  Here: ONE elementwise pass - every output
//     element is either the fill value or a 4-tap cubic interpolation along time of the input (the frequency axis keeps
//     its size, so its interpolation weights are exactly (0, 1, 0, 0)).  The random parameters are drawn by the host in
//     SpeechBrain's order (augment.py) and arrive as integers.
//   * `self.hparams.ctc_cost(p_ctc, tokens, wav_lens, tokens_lens)` (train_multitask.py:164-170; yaml :256-258:
//     speechbrain.nnet.losses.ctc_loss = torch.nn.functional.ctc_loss with zero_infinity, blank 0, reduction batchmean):
//     the alpha recursion in log space, one CTA per utterance, the states of the extended label sequence over the
//     threads, one barrier per frame; then SpeechBrain's reductions over the batch.  Forward value only (what the
//     validation stage reports); the gradient belongs to the training loop, which is outside this path.
// Plain SIMT kernels: HBM-bound (SpecAugment: one read, one write) / latency-bound (CTC: T dependent steps).
#include "common.cuh"

namespace {

// torch's cubic convolution coefficients (A = -0.75), aten/src/ATen/native/UpSample.h
__device__ __forceinline__ float cubic1(float x, float a) { return ((a + 2.f) * x - (a + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float a) { return ((a * x - 5.f * a) * x + 8.f * a) * x - 4.f * a; }

__global__ void __launch_bounds__(256)
spec_augment_kernel(const float* __restrict__ x, int frames, int dim, int warp_c, int warp_w,
                    const int* __restrict__ freq_pos, const int* __restrict__ freq_len, int n_freq,
                    const int* __restrict__ time_pos, const int* __restrict__ time_len, int n_time, float fill,
                    float* __restrict__ out) {
  const int b = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)frames * dim) return;
  const int t = (int)(i / dim), f = (int)(i - (int64_t)t * dim);
  bool masked = false;
  for (int j = 0; j < n_freq; ++j) {
    const int p = freq_pos[b * n_freq + j];
    masked |= p <= f && f < p + freq_len[b * n_freq + j];
  }
  for (int j = 0; j < n_time; ++j) {
    const int p = time_pos[b * n_time + j];
    masked |= p <= t && t < p + time_len[b * n_time + j];
  }
  const float* xb = x + (int64_t)b * frames * dim;
  float v;
  if (masked) {
    v = fill;
  } else if (warp_w < 0) {
    v = xb[i];
  } else {
    // output rows [0, w) are rows [0, c) resampled, rows [w, T) are rows [c, T) resampled (align_corners)
    const bool left = t < warp_w;
    const int seg0 = left ? 0 : warp_c, in_len = left ? warp_c : frames - warp_c;
    const int out_len = left ? warp_w : frames - warp_w, idx = left ? t : t - warp_w;
    const float scale = out_len > 1 ? (float)(in_len - 1) / (float)(out_len - 1) : 0.f;
    const float real = __fmul_rn(scale, (float)idx);
    const int i0 = min((int)floorf(real), in_len - 1);
    const float lam = fminf(fmaxf(real - (float)i0, 0.f), 1.f);
    const float a = -0.75f;
    const float w0 = cubic2(lam + 1.f, a), w1 = cubic1(lam, a), w2 = cubic1(1.f - lam, a), w3 = cubic2(2.f - lam, a);
    const int r0 = min(max(i0 - 1, 0), in_len - 1), r1 = min(max(i0, 0), in_len - 1);
    const int r2 = min(max(i0 + 1, 0), in_len - 1), r3 = min(max(i0 + 2, 0), in_len - 1);
    const float* col = xb + (int64_t)seg0 * dim + f;
    v = w0 * col[(int64_t)r0 * dim];
    v = fmaf(w1, col[(int64_t)r1 * dim], v);
    v = fmaf(w2, col[(int64_t)r2 * dim], v);
    v = fmaf(w3, col[(int64_t)r3 * dim], v);
  }
  out[(int64_t)b * frames * dim + i] = v;
}

__device__ __forceinline__ float log_add(float a, float b) {
  // log(exp(a) + exp(b)) with -inf handled (no NaN from inf - inf)
  const float m = fmaxf(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + log1pf(expf(fminf(a, b) - m));
}

constexpr int kCtcThreads = 256;

// alpha recursion of the CTC forward pass; shared memory: labels [L] | alpha [2][2 L + 1]
__global__ void __launch_bounds__(kCtcThreads)
ctc_nll_kernel(const float* __restrict__ log_probs, const int* __restrict__ targets, const int* __restrict__ input_len,
               const int* __restrict__ target_len, int frames, int vocab, int max_targets, int blank, int zero_infinity,
               float* __restrict__ nll) {
  extern __shared__ float ctc_smem[];
  const int b = blockIdx.x;
  const int tl = min(max(target_len[b], 0), max_targets), il = min(max(input_len[b], 0), frames);
  const int n_states = 2 * tl + 1;
  int* lab = reinterpret_cast<int*>(ctc_smem);
  float* alpha = ctc_smem + max_targets;
  const int pitch = 2 * max_targets + 1;
  for (int j = threadIdx.x; j < tl; j += kCtcThreads) lab[j] = targets[(int64_t)b * max_targets + j];
  __syncthreads();
  const float* lp = log_probs + (int64_t)b * frames * vocab;
  // frame 0: only the first blank and the first label are reachable
  for (int s = threadIdx.x; s < n_states; s += kCtcThreads) {
    float v = -INFINITY;
    if (il > 0) {
      if (s == 0) v = lp[blank];
      else if (s == 1) v = lp[lab[0]];
    }
    alpha[s] = v;
  }
  __syncthreads();
  int cur = 0;
  for (int t = 1; t < il; ++t) {
    const float* row = lp + (int64_t)t * vocab;
    const float* prev = alpha + cur * pitch;
    float* next = alpha + (cur ^ 1) * pitch;
    for (int s = threadIdx.x; s < n_states; s += kCtcThreads) {
      const int sym = (s & 1) ? lab[s >> 1] : blank;
      float v = prev[s];
      if (s >= 1) v = log_add(v, prev[s - 1]);
      if ((s & 1) && s >= 3 && lab[s >> 1] != lab[(s >> 1) - 1]) v = log_add(v, prev[s - 2]);
      next[s] = v + row[sym];
    }
    __syncthreads();
    cur ^= 1;
  }
  if (threadIdx.x == 0) {
    const float* fin = alpha + cur * pitch;
    float ll = il > 0 ? fin[n_states - 1] : (tl == 0 ? 0.f : -INFINITY);
    if (il > 0 && n_states > 1) ll = log_add(ll, fin[n_states - 2]);
    float v = -ll;
    if (zero_infinity && v == INFINITY) v = 0.f;
    nll[b] = v;
  }
}

// SpeechBrain's reductions over the per-utterance values (one thread: B is a batch size)
__global__ void ctc_reduce_loss_kernel(const float* __restrict__ nll, const int* __restrict__ target_len, int batch,
                                       int mode, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (mode == 4) {                                  // "batch": per utterance / its target length
    for (int b = 0; b < batch; ++b) out[b] = nll[b] / (float)target_len[b];
    return;
  }
  double acc = 0.0;
  for (int b = 0; b < batch; ++b)
    acc += mode == 2 ? (double)(nll[b] / (float)max(target_len[b], 1)) : (double)nll[b];
  if (mode == 2 || mode == 3) acc /= (double)batch;   // "mean" (torch: per-target-length, then batch mean) / "batchmean"
  out[0] = (float)acc;
}

}  // namespace

extern "C" int stac_spec_augment(const float* x, int64_t batch, int64_t frames, int64_t dim, int warp_center,
                                 int warp_width, const int32_t* freq_pos, const int32_t* freq_len, int n_freq,
                                 const int32_t* time_pos, const int32_t* time_len, int n_time, float fill, float* out,
                                 void* stream) {
  STAC_REQUIRE(x && out && x != out && batch > 0 && batch < 65536 && frames > 0 && dim > 0);
  STAC_REQUIRE(n_freq >= 0 && n_time >= 0 && (n_freq == 0 || (freq_pos && freq_len)) && (n_time == 0 || (time_pos && time_len)));
  if (warp_width >= 0) STAC_REQUIRE(warp_center > 0 && warp_center < frames && warp_width > 0 && warp_width < frames);
  if (frames * dim >= (1ll << 40)) return STAC_ERR_UNSUPPORTED_SHAPE;
  dim3 grid((unsigned)ceil_div64(frames * dim, 256), (unsigned)batch);
  spec_augment_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, (int)frames, (int)dim, warp_center, warp_width, freq_pos,
                                                          freq_len, n_freq, time_pos, time_len, n_time, fill, out);
  STAC_LAUNCH_CHECK();
}

extern "C" int stac_ctc_loss(const float* log_probs, const int32_t* targets, const int32_t* input_len,
                             const int32_t* target_len, int64_t batch, int64_t frames, int64_t vocab, int64_t max_targets,
                             int blank, int reduction, float* nll, float* loss, void* stream) {
  STAC_REQUIRE(log_probs && targets && input_len && target_len && nll && loss);
  STAC_REQUIRE(batch > 0 && batch < 65536 && frames > 0 && vocab > 0 && max_targets > 0 && blank >= 0 && blank < vocab);
  STAC_REQUIRE(reduction >= 0 && reduction <= 4);
  const size_t smem = (size_t)(max_targets + 2 * (2 * max_targets + 1)) * sizeof(float);
  if (smem > 48 * 1024) return STAC_ERR_UNSUPPORTED_SHAPE;              // up to ~2400 target tokens
  ctc_nll_kernel<<<(unsigned)batch, kCtcThreads, smem, as_stream(stream)>>>(log_probs, targets, input_len, target_len,
                                                                           (int)frames, (int)vocab, (int)max_targets,
                                                                           blank, 1, nll);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) return (int)cudaGetLastError();
  if (reduction != 0) ctc_reduce_loss_kernel<<<1, 32, 0, as_stream(stream)>>>(nll, target_len, (int)batch, reduction, loss);
  STAC_LAUNCH_CHECK();
}
