// Host ingest (the step in front of the path): 16-bit PCM -> the fp32 waveform batch Fbank reads.
//
// Reference behaviour replaced: the audio pipeline of /root/reference/stac-st/inference.py:250-261 (librosa.load of
// every turn, torch.cat, fp32 on the host) followed by batch.to(device) (:91) - 4 bytes per sample over PCIe.  A 16 kHz
// 16-bit file decodes to sample / 32768 exactly (what librosa / soundfile return), so shipping the int16 samples and
// scaling on the device is bit-identical and halves the host->device bytes (61 instead of 123 MB per 64 x 30 s batch).
// HBM-bound: 2 B read + 4 B written per sample.
#include <algorithm>
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
pcm_i16_to_f32_kernel(const int16_t* __restrict__ pcm, long long n, float* __restrict__ out) {
  constexpr float kScale = 1.0f / 32768.0f;
  const long long n8 = n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(pcm) + i);          // 8 samples
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    float f[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      f[2 * k] = (float)(int16_t)(w[k] & 0xffffu) * kScale;
      f[2 * k + 1] = (float)(int16_t)(w[k] >> 16) * kScale;
    }
    float4* dst = reinterpret_cast<float4*>(out) + 2 * i;
    dst[0] = make_float4(f[0], f[1], f[2], f[3]);
    dst[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  // tail (n not a multiple of 8)
  for (long long i = (n8 << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = (float)pcm[i] * kScale;
}

}  // namespace

extern "C" int stac_pcm_i16_to_f32(const int16_t* pcm, int64_t n, float* out, void* stream) {
  STAC_REQUIRE(pcm && out && n > 0);
  if ((reinterpret_cast<uintptr_t>(pcm) & 15) != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0)
    return STAC_ERR_UNSUPPORTED_SHAPE;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(ceil_div64(n, 8), 256), 148 * 16);
  pcm_i16_to_f32_kernel<<<grid, 256, 0, as_stream(stream)>>>(pcm, (long long)n, out);
  STAC_LAUNCH_CHECK();
}
