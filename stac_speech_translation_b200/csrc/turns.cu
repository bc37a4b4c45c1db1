// Speaker-turn / cross-talk spikes from the CTC-greedy ids (the consumer right behind a9).
//
// Reference behaviour replaced: append_speaker_turns, /root/reference/stac-st/inference.py:54-84 -
//   p_ctc.argmax(-1); == hparams["turn"]; == hparams["xt"]; both [B, T2] masks copied to the host and walked by a Python
//   loop over every frame of every utterance (padding frames included) that appends one RTTM line per spike.
// Here the masks never exist: the ids (already produced by the CTC head, or by stac_argmax_rows from a posterior tensor
// the caller holds) are compacted on the device into the ascending list of flat positions b * T2 + j of each class,
// which is the order the reference's loop appends in; the host reads back two counts and the spikes only.
// All of it is index work on B * T2 int32 (48 k at the benchmark shape): launch-latency bound, no roofline.
#include "common.cuh"

namespace {

constexpr int kBlock = 256;

// arg-max of every row (first index wins on ties, the rule of log_softmax_kernel and of the fused CTC head)
__global__ void __launch_bounds__(kBlock)
argmax_rows_kernel(const float* __restrict__ x, int cols, int* __restrict__ out) {
  __shared__ float red_v[kBlock / 32];
  __shared__ int red_i[kBlock / 32];
  const float* xr = x + (int64_t)blockIdx.x * cols;
  float mx = -INFINITY;
  int mi = 0x7fffffff;
  for (int i = threadIdx.x; i < cols; i += kBlock) {
    const float v = xr[i];
    if (v > mx || mi == 0x7fffffff) { mx = v; mi = i; }      // (a NaN-free row: the first element seeds the pair)
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (ov > mx || (ov == mx && oi < mi)) { mx = ov; mi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { red_v[threadIdx.x >> 5] = mx; red_i[threadIdx.x >> 5] = mi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bm = red_v[0];
    int bi = red_i[0];
    for (int w = 1; w < kBlock / 32; ++w)
      if (red_v[w] > bm || (red_v[w] == bm && red_i[w] < bi)) { bm = red_v[w]; bi = red_i[w]; }
    out[blockIdx.x] = bi;
  }
}

// block-wide sum of one int per thread; `red` holds kBlock / 32 + 1 ints
__device__ __forceinline__ int block_sum_int(int v, int* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kBlock / 32; ++w) t += red[w];
    red[kBlock / 32] = t;
  }
  __syncthreads();
  return red[kBlock / 32];
}

// pass 1: spikes of both classes per utterance row -> row_counts[b][2]
__global__ void __launch_bounds__(kBlock)
spike_count_kernel(const int* __restrict__ ids, int t2, int turn_id, int xt_id, int* __restrict__ row_counts) {
  __shared__ int red[kBlock / 32 + 1];
  const int* row = ids + (int64_t)blockIdx.x * t2;
  int n_turn = 0, n_xt = 0;
  for (int j = threadIdx.x; j < t2; j += kBlock) {
    const int id = row[j];
    n_turn += id == turn_id;
    n_xt += id == xt_id;
  }
  const int s_turn = block_sum_int(n_turn, red);
  const int s_xt = block_sum_int(n_xt, red);
  if (threadIdx.x == 0) {
    row_counts[2 * blockIdx.x] = s_turn;
    row_counts[2 * blockIdx.x + 1] = s_xt;
  }
}

// pass 2: row b writes its spikes behind those of rows 0 .. b-1, frames ascending; the last row also writes the totals
__global__ void __launch_bounds__(kBlock)
spike_compact_kernel(const int* __restrict__ ids, int batch, int t2, int turn_id, int xt_id,
                     const int* __restrict__ row_counts, int* __restrict__ spikes_turn, int* __restrict__ spikes_xt,
                     int* __restrict__ n_out) {
  __shared__ int red[kBlock / 32 + 1];
  __shared__ int warp_tot[2][kBlock / 32];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int a_turn = 0, a_xt = 0;
  for (int i = threadIdx.x; i < b; i += kBlock) { a_turn += row_counts[2 * i]; a_xt += row_counts[2 * i + 1]; }
  int off_turn = block_sum_int(a_turn, red);
  int off_xt = block_sum_int(a_xt, red);
  const int* row = ids + (int64_t)b * t2;
  for (int j0 = 0; j0 < t2; j0 += kBlock) {
    const int j = j0 + threadIdx.x;
    const int id = j < t2 ? row[j] : -1;
    const bool f_turn = j < t2 && id == turn_id;
    const bool f_xt = j < t2 && id == xt_id;
    const unsigned m_turn = __ballot_sync(0xffffffffu, f_turn);
    const unsigned m_xt = __ballot_sync(0xffffffffu, f_xt);
    __syncthreads();                        // warp_tot of the previous chunk has been read by everyone
    if (lane == 0) { warp_tot[0][warp] = __popc(m_turn); warp_tot[1][warp] = __popc(m_xt); }
    __syncthreads();
    int before_turn = 0, before_xt = 0, tot_turn = 0, tot_xt = 0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; ++w) {
      if (w < warp) { before_turn += warp_tot[0][w]; before_xt += warp_tot[1][w]; }
      tot_turn += warp_tot[0][w];
      tot_xt += warp_tot[1][w];
    }
    const unsigned lt = (1u << lane) - 1u;
    if (f_turn) spikes_turn[off_turn + before_turn + __popc(m_turn & lt)] = b * t2 + j;
    if (f_xt) spikes_xt[off_xt + before_xt + __popc(m_xt & lt)] = b * t2 + j;
    off_turn += tot_turn;
    off_xt += tot_xt;
  }
  if (b == batch - 1 && threadIdx.x == 0) { n_out[0] = off_turn; n_out[1] = off_xt; }
}

}  // namespace

extern "C" int stac_argmax_rows(const float* x, int64_t rows, int64_t cols, int32_t* out, void* stream) {
  STAC_REQUIRE(x && out && rows > 0 && rows < (1ll << 31) && cols > 0 && cols < (1ll << 31));
  argmax_rows_kernel<<<(unsigned)rows, kBlock, 0, as_stream(stream)>>>(x, (int)cols, out);
  STAC_LAUNCH_CHECK();
}

extern "C" int stac_ctc_spikes(const int32_t* ids, int64_t batch, int64_t t2, int32_t turn_id, int32_t xt_id,
                               int32_t* row_counts, int32_t* spikes_turn, int32_t* spikes_xt, int32_t* n_out,
                               void* stream) {
  STAC_REQUIRE(ids && row_counts && spikes_turn && spikes_xt && n_out && batch > 0 && t2 > 0);
  if (batch * t2 >= (1ll << 31)) return STAC_ERR_UNSUPPORTED_SHAPE;
  spike_count_kernel<<<(unsigned)batch, kBlock, 0, as_stream(stream)>>>(ids, (int)t2, turn_id, xt_id, row_counts);
  spike_compact_kernel<<<(unsigned)batch, kBlock, 0, as_stream(stream)>>>(ids, (int)batch, (int)t2, turn_id, xt_id,
                                                                        row_counts, spikes_turn, spikes_xt, n_out);
  STAC_LAUNCH_CHECK();
}
