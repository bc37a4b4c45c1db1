// Training-stage statistics of InputNormalization (first piece): per-utterance mean and unbiased
// standard deviation of every mel bin over the utterance's valid frames.
//
// Reference behaviour replaced: the Python loop of SpeechBrain's InputNormalization.forward (one torch.mean + torch.std
// per utterance, reached from /root/reference/stac-st/train_multitask.py:60-61 with epoch < update_until_epoch,
// yaml transformer_multitask.yaml:208-210): actual_size = round(lengths[b] * T) in fp32, mean / std over
// x[b, :actual_size], std floored at eps.  The running-average update of the 80 global values stays on the host.
// HBM-bound: the features are read twice (two-pass variance, the second pass from L2), 32 B written per utterance and bin.
// Parity on a B200: tests/test_gpu_ytrain_norm.py.
#include "common.cuh"

namespace {

constexpr int kGroups = 8;     // row groups per CTA (threadIdx.y); threadIdx.x = column inside a 32-wide slab

__global__ void __launch_bounds__(32 * kGroups)
utt_mean_std_kernel(const float* __restrict__ x, const float* __restrict__ wav_len, int frames, int dim, float eps,
                    float* __restrict__ mean_out, float* __restrict__ std_out) {
  __shared__ float red[kGroups][33];
  const int b = blockIdx.x;
  const int c = blockIdx.y * 32 + threadIdx.x;
  const int g = threadIdx.y;
  // torch.round(lengths[b] * T).int(): fp32 product, round half to even
  const int n = min(max((int)rintf(__fmul_rn(wav_len[b], (float)frames)), 0), frames);
  const float* xb = x + (int64_t)b * frames * dim;
  float s = 0.f;
  if (c < dim)
    for (int r = g; r < n; r += kGroups) s += xb[(int64_t)r * dim + c];
  red[g][threadIdx.x] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int k = 0; k < kGroups; ++k) tot += red[k][threadIdx.x];
  const float mean = tot / (float)n;                  // n == 0: NaN, as torch.mean of an empty slice
  __syncthreads();
  float ss = 0.f;
  if (c < dim)
    for (int r = g; r < n; r += kGroups) {
      const float d = xb[(int64_t)r * dim + c] - mean;
      ss = fmaf(d, d, ss);
    }
  red[g][threadIdx.x] = ss;
  __syncthreads();
  if (g == 0 && c < dim) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < kGroups; ++k) v += red[k][threadIdx.x];
    const float sd = sqrtf(v / (float)(n - 1));       // unbiased; n == 1: NaN, as torch.std
    mean_out[(int64_t)b * dim + c] = mean;
    std_out[(int64_t)b * dim + c] = sd != sd ? sd : fmaxf(sd, eps);
  }
}

}  // namespace

extern "C" int stac_utt_mean_std(const float* x, const float* wav_len, int64_t batch, int64_t frames, int64_t dim,
                                 float eps, float* mean, float* std, void* stream) {
  STAC_REQUIRE(x && wav_len && mean && std && batch > 0 && batch < 65536 && frames > 0 && frames < (1 << 24) && dim > 0);
  if (dim > 32 * 65535) return STAC_ERR_UNSUPPORTED_SHAPE;
  dim3 grid((unsigned)batch, (unsigned)ceil_div64(dim, 32));
  utt_mean_std_kernel<<<grid, dim3(32, kGroups), 0, as_stream(stream)>>>(x, wav_len, (int)frames, (int)dim, eps, mean,
                                                                        std);
  STAC_LAUNCH_CHECK();
}
