// Flash-style multi-head self-attention on tcgen05 tensor cores (bf16 operands, fp32 softmax).
// Key-padding is by per-utterance valid length (no dense mask tensor, no B*H*T*T score matrix).
//
// Reference behaviour replaced: torch.nn.MultiheadAttention slow path (baddbmm + softmax + bmm with a
// -inf key-padding mask) reached through SpeechBrain's TransformerEncoderLayer from
//   /root/reference/stac-st/modules/TransformerMultiTask.py:304-308 (mask built at :289-294 / :225-226).
//
// CTA = (128-query tile, head, utterance), 128 threads; thread r owns query row r.
//   S = Q K^T     : tcgen05.mma 128x128x64 (4 UMMA_K steps), accumulator in TMEM columns [0,128)
//   softmax       : tcgen05.ld S -> registers, online max/sum in fp32 (exp2), P -> smem as bf16 in the
//                   K-major 128B-swizzled UMMA layout
//   O_tile = P V  : tcgen05.mma 128x64x128 with B = V^T tile (keys contiguous), TMEM columns [128,192)
//   O accumulates in registers with the usual running-max rescale.
#include <algorithm>
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int kQT = 128, kKT = 128, kHd = 64;
constexpr int kSmemQ = 0, kSmemK = 16384, kSmemV = 32768, kSmemP = 49152, kSmemBar = 81920;
constexpr int kSmemBytes = kSmemBar + 64 + 1024;
constexpr float kLog2e = 1.4426950408889634f;

__global__ void __launch_bounds__(128, 2)
mha_bf16_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_vt,
                const int* __restrict__ kv_len, int seq_len, int d_model, int n_head,
                __nv_bfloat16* __restrict__ ctx) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar_q = sbase + kSmemBar, bar_kv = bar_q + 8, bar_s = bar_q + 16, bar_o = bar_q + 24;
  const uint32_t tmem_slot = bar_q + 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kQT, h = blockIdx.y, b = blockIdx.z;
  const int n_keys = min(max(kv_len[b], 1), seq_len);
  const int n_kt = (n_keys + kKT - 1) / kKT;

  if (tid == 0) {
    prefetch_tmap(&tmap_qkv);
    prefetch_tmap(&tmap_vt);
    mbar_init(bar_q, 1); mbar_init(bar_kv, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

  const int row_base = b * seq_len;
  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, kQT * kHd * 2);
    tma_load_2d(sbase + kSmemQ, &tmap_qkv, bar_q, h * kHd, row_base + q0);
  }

  float o[kHd];
#pragma unroll
  for (int i = 0; i < kHd; ++i) o[i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;
  constexpr uint32_t idesc_s = make_idesc_bf16(128, 128);
  constexpr uint32_t idesc_o = make_idesc_bf16(128, 64);

  for (int kt = 0; kt < n_kt; ++kt) {
    const uint32_t ph = kt & 1;
    if (tid == 0) {
      mbar_arrive_expect_tx(bar_kv, 2 * kKT * kHd * 2);
      tma_load_2d(sbase + kSmemK, &tmap_qkv, bar_kv, d_model + h * kHd, row_base + kt * kKT);
      tma_load_3d(sbase + kSmemV, &tmap_vt, bar_kv, kt * kKT, 0, b * n_head + h);
      tma_load_3d(sbase + kSmemV + 8192, &tmap_vt, bar_kv, kt * kKT + 64, 0, b * n_head + h);
      if (kt == 0) mbar_wait(bar_q, 0);
      mbar_wait(bar_kv, ph);
      tc_fence_after();
      const uint64_t qd = make_smem_desc_sw128(sbase + kSmemQ), kd = make_smem_desc_sw128(sbase + kSmemK);
#pragma unroll
      for (int k = 0; k < kHd / 16; ++k) umma_bf16(tmem_s, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
      umma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_s, ph);
    tc_fence_after();

    // ---- online softmax over this key tile ----
    const int valid = n_keys - kt * kKT;  // columns < valid are real keys
    float tile_max = -INFINITY;
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t v[32];
      tmem_ld32(tmem_s + lane_off + ch * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (ch * 32 + i < valid) tile_max = fmaxf(tile_max, __uint_as_float(v[i]));
    }
    const float m_new = fmaxf(m_run, tile_max);
    const float alpha = exp2f((m_run - m_new) * kLog2e);
    const float m_scaled = m_new * kLog2e;
    float l_tile = 0.f;
    unsigned char* prow = sptr + kSmemP + tid * 128;
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t v[32];
      tmem_ld32(tmem_s + lane_off + ch * 32, v);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float p0 = (ch * 32 + i < valid) ? exp2f(fmaf(__uint_as_float(v[i]), kLog2e, -m_scaled)) : 0.f;
        const float p1 = (ch * 32 + i + 1 < valid) ? exp2f(fmaf(__uint_as_float(v[i + 1]), kLog2e, -m_scaled)) : 0.f;
        // sum what the tensor core will actually see (bf16-rounded P) so rows normalise exactly
        const __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
        l_tile += __bfloat162float(pb.x) + __bfloat162float(pb.y);
        pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&pb);
      }
      // columns ch*32 .. +31 -> k-block (ch>>1), 16-byte chunks j = (ch&1)*4 .. +3, swizzled by row
      unsigned char* blk = prow + (ch >> 1) * 16384;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int chunk = ((ch & 1) * 4 + j) ^ (tid & 7);
        *reinterpret_cast<uint4*>(blk + chunk * 16) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      }
    }
    l_run = l_run * alpha + l_tile;
    m_run = m_new;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < kKT / 16; ++k) {
        const uint64_t pd = make_smem_desc_sw128(sbase + kSmemP + (k >> 2) * 16384) + 2 * (k & 3);
        const uint64_t vd = make_smem_desc_sw128(sbase + kSmemV + (k >> 2) * 8192) + 2 * (k & 3);
        umma_bf16(tmem_o, pd, vd, idesc_o, k != 0);
      }
      umma_commit(bar_o);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kHd; ++i) o[i] *= alpha;
    mbar_wait(bar_o, ph);
    tc_fence_after();
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t v[32];
      tmem_ld32(tmem_o + lane_off + ch * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[ch * 32 + i] += __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();   // everyone is done with S/O TMEM, P smem and the K/V tiles
    tc_fence_after();
  }

  const int q = q0 + tid;
  if (q < seq_len) {
    const float inv = 1.0f / l_run;
    __nv_bfloat16* dst = ctx + ((int64_t)row_base + q) * d_model + h * kHd;
#pragma unroll
    for (int i = 0; i < kHd; i += 8) {
      *reinterpret_cast<uint4*>(dst + i) =
          make_uint4(pack_bf16x2(o[i] * inv, o[i + 1] * inv), pack_bf16x2(o[i + 2] * inv, o[i + 3] * inv),
                     pack_bf16x2(o[i + 4] * inv, o[i + 5] * inv), pack_bf16x2(o[i + 6] * inv, o[i + 7] * inv));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

}  // namespace

extern "C" int stac_mha_bf16(const uint16_t* qkv, const uint16_t* v_t, const int32_t* kv_len, int64_t batch,
                             int64_t seq_len, int64_t t_pad, int64_t d_model, int64_t n_head, uint16_t* ctx,
                             void* stream) {
  STAC_REQUIRE(qkv && v_t && kv_len && ctx && batch > 0 && batch < 65536 && seq_len > 0);
  STAC_REQUIRE(t_pad >= seq_len && t_pad % 8 == 0);
  if (d_model != n_head * kHd || n_head > 65535 || batch * seq_len >= (1ll << 31)) return STAC_ERR_UNSUPPORTED_SHAPE;
  CUtensorMap tq, tv;
  {
    const uint64_t dims[2] = {(uint64_t)(3 * d_model), (uint64_t)(batch * seq_len)};
    const uint64_t str[1] = {(uint64_t)(3 * d_model) * 2};
    const uint32_t box[2] = {kHd, 128};
    int r = encode_bf16_map(&tq, qkv, 2, dims, str, box);
    if (r != STAC_OK) return r;
  }
  {
    // V^T [B*H][64][t_pad], innermost = keys
    const uint64_t dims[3] = {(uint64_t)t_pad, kHd, (uint64_t)(batch * n_head)};
    const uint64_t str[2] = {(uint64_t)t_pad * 2, (uint64_t)t_pad * kHd * 2};
    const uint32_t box[3] = {64, kHd, 1};
    int r = encode_bf16_map(&tv, v_t, 3, dims, str, box);
    if (r != STAC_OK) return r;
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mha_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  dim3 grid((unsigned)ceil_div64(seq_len, kQT), (unsigned)n_head, (unsigned)batch);
  mha_bf16_kernel<<<grid, 128, kSmemBytes, as_stream(stream)>>>(tq, tv, kv_len, (int)seq_len, (int)d_model,
                                                               (int)n_head, reinterpret_cast<__nv_bfloat16*>(ctx));
  STAC_LAUNCH_CHECK();
}
