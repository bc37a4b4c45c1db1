/*
 * libstac_b200 - C ABI of the B200-native (sm_100a) STAC-ST encoder-side hot path.
 *
 * The reference (amazon-science/stac-speech-translation) has no FFI: its hot path is
 * six eager Python calls into SpeechBrain/PyTorch
 *   (/root/reference/stac-st/inference.py:95-107, stac-st/train_multitask.py:59-78).
 * Each entry point below replaces the arithmetic of one of those calls; the Python
 * drop-in classes in stac_speech_translation_b200/ bind them with ctypes
 * (INTEGRATION.md shows the binding and the HyperPyYAML override).
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless named host_*;
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on
 *     it: no allocation, no synchronisation, no host reads of device memory;
 *   - return value: 0 = ok, <0 = argument error (STAC_ERR_*), >0 = cudaError_t;
 *   - tensors are dense row-major, fp32 unless the name says bf16
 *     (bf16 = uint16_t storage of the upper half of an IEEE fp32);
 *   - T  = 1 + L/160 frames, T1 = (T-1)/2+1, T2 = (T1-1)/2+1 (25 Hz).
 */
#ifndef STAC_B200_H_
#define STAC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STAC_B200_VERSION 100

#define STAC_OK 0
#define STAC_ERR_INVALID_ARGUMENT (-1)
#define STAC_ERR_UNSUPPORTED_SHAPE (-2)
#define STAC_ERR_DRIVER_ENTRY (-3)   /* cuTensorMapEncodeTiled not obtainable */
#define STAC_ERR_TENSOR_MAP (-4)     /* cuTensorMapEncodeTiled failed */

/* epilogue activation of the GEMM entry points */
#define STAC_ACT_NONE 0
#define STAC_ACT_GELU_ERF 1

/* output element type selectors */
#define STAC_DT_F32 0
#define STAC_DT_BF16 1

int stac_version(void);
/* Persistent kernels launch at most (SM count - n) CTAs from now on (default 0).  Use when a communication kernel
 * (NCCL send/recv of the multi-GPU gather) runs beside the path and holds SMs. */
int stac_set_reserved_sms(int n);
/* L2 residency hint (cudaStreamAttributeAccessPolicyWindow + the device's persisting-L2 carve-out) for a buffer the
 * kernels launched - or captured - on `stream` afterwards re-read and update: the encoder's fp32 residual stream.
 * bytes == 0 clears the stream's window.  STAC_ERR_UNSUPPORTED_SHAPE: the device has no persisting L2. */
int stac_l2_persist(const void* base, int64_t bytes, float hit_ratio, void* stream);
int stac_l2_persist_limits(int64_t* max_persist_bytes, int64_t* max_window_bytes);
const char* stac_error_string(int code);

/* ---------------------------------------------------------------------------
 * a2  Fbank   -- replaces hparams.compute_features(wavs)
 *     (/root/reference/stac-st/inference.py:95; yaml transformer_multitask.yaml:299-302)
 * STFT(hamming 400, hop 160, centre zero pad) -> |.|^2 -> 80 triangular mel
 * -> 10*log10(max(.,1e-10)); also the per-utterance maximum used by the top-dB clamp.
 *
 * tables: constant block built once by the host (layout: stac_fbank_tables_floats()
 *         floats: window[400] | mel_start[80] | mel_count[80] | mel_weight[80][16]).
 * utt_max_ordered: uint32[B], order-preserving encoding of the fp32 maximum (zeroed by the
 *         call itself with a stream-ordered memset; 0 encodes "below every float").
 */
int stac_fbank_tables_floats(void);
int stac_fbank_logmel(const float* pcm, int64_t batch, int64_t n_samples, int64_t pcm_row_stride,
                      const float* tables, float* logmel_db /*[B,T,80]*/,
                      uint32_t* utt_max_ordered /*[B]*/, void* stream);

/* a2 on tensor cores (bf16 mode): the windowed 400-point DFT of 128 frames per tile is a tcgen05 GEMM against fp16
 * cos / sin matrices after folding the frame on its symmetry (K = 201 + 199 instead of 400 + 400), fused with power, mel,
 * dB and the per-utterance maximum.  Same outputs as stac_fbank_logmel (log-mel within 4e-4 relative).
 *   tables:   stac_fbank_tc_tables_floats() floats: window[400] | per-bin mel weights [208][2]
 *   twiddles: stac_fbank_tc_twiddle_halfs() fp16: [cos | sin][208 bins][256 columns], see ops.build_fbank_tc_tables */
int stac_fbank_tc_tables_floats(void);
int stac_fbank_tc_twiddle_halfs(void);
int stac_fbank_logmel_tc(const float* pcm, int64_t batch, int64_t n_samples, int64_t pcm_row_stride,
                         const float* tables, const uint16_t* twiddles, float* logmel_db /*[B,T,80]*/,
                         uint32_t* utt_max_ordered /*[B]*/, void* stream);

/* a2 on tensor cores, second design (the bf16 path's default): the same folded fp16 DFT GEMM, but the A operand is built
 * in registers by a thread that owns one frame and written straight into tensor memory (TS-form tcgen05.mma); with
 * pair != 0 two CTAs of a cluster form one cta_group::2 instance (2 x 128 frames, each CTA holds half of every twiddle
 * tile), pair == 0 runs the same kernel per single CTA.  Same outputs as stac_fbank_logmel_tc.
 *   tables:   stac_fbank_tc2_tables_floats() floats: per-bin mel weights [208][2]
 *   twiddles: stac_fbank_tc2_twiddle_halfs() fp16: [7 stages][208 bins][64 columns]; columns 0-31 of stage i are
 *             w[n] cos(2 pi k n / 400) and columns 32-63 -w[n] sin(2 pi k n / 400) for n = 32 i .. 32 i + 31 (w = the
 *             hamming window, zero where n > 200); the kernel feeds them e[n] = x[n] + x[400 - n], o[n] = x[n] - x[400 - n]
 *   n_samples must be a multiple of 32 (the PCM tile travels as tensor-map boxes of 32-sample rows): other lengths
 *   return STAC_ERR_UNSUPPORTED_SHAPE and belong to stac_fbank_logmel_tc / stac_fbank_logmel. */
int stac_fbank_tc2_tables_floats(void);
int stac_fbank_tc2_twiddle_halfs(void);
int stac_fbank_logmel_tc2(const float* pcm, int64_t batch, int64_t n_samples, int64_t pcm_row_stride,
                          const float* tables, const uint16_t* twiddles, float* logmel_db /*[B,T,80]*/,
                          uint32_t* utt_max_ordered /*[B]*/, int pair, void* stream);

/* top-dB clamp (+ optional global mean/std normalisation, a3) in one elementwise pass:
 *   y = max(x, max_b - top_db);  if (mean) y = (y - mean[m]) / std[m]
 * per_utterance != 0: max_b is utterance b's maximum, else the maximum over the batch.
 * Replaces the tail of Filterbank._amplitude_to_DB and, with mean/std, the eval-mode
 * modules.normalize(feats, wav_lens) (/root/reference/stac-st/inference.py:96).      */
int stac_fbank_topdb_norm(const float* logmel_db, const uint32_t* utt_max_ordered,
                          int per_utterance, float top_db,
                          const float* mean /*[80] or NULL*/, const float* std /*[80] or NULL*/,
                          int64_t batch, int64_t frames, int64_t n_mels, float* out, void* stream);

/* a3 alone: y = (x - mean[m]) / std[m] over [rows, n_mels] */
int stac_input_norm(const float* x, const float* mean, const float* std, int64_t rows,
                    int64_t n_mels, float* out, void* stream);

/* ---------------------------------------------------------------------------
 * a4  ConvolutionFrontEnd -- replaces modules.CNN(feats)
 *     (/root/reference/stac-st/inference.py:99; yaml :173-180)
 * Block 0: reflect-pad 1, Conv2d(1->256, 3x3, stride 2) + bias, LayerNorm over (40,256)
 *          eps 1e-5, LeakyReLU(0.01), fused in one kernel.
 *   w0 [256][3(freq)][3(time)] (torch weight [256,1,3,3]), b0 [256], ln_g/ln_b [40*256].
 *   out_mode STAC_DT_F32 : fp32 [B,T1,40,256]
 *   out_mode STAC_DT_BF16: bf16 reflect-padded, parity-split planes
 *        [B][4 = (t_par*2+f_par)][Tp2][21][256], Tp2 = (T1+3)/2, padded coords
 *        tp = t1+1, fp = f1+1, plane = (tp&1)*2 + (fp&1), indices tp>>1, fp>>1
 *        (the layout the tensor-core conv1 reads with unit-stride TMA boxes).
 */
int64_t stac_conv0_padded_elems(int64_t batch, int64_t t1);
int stac_conv0_ln_lrelu(const float* feats /*[B,T,80]*/, const float* w0, const float* b0,
                        const float* ln_g, const float* ln_b, int64_t batch, int64_t frames,
                        void* out, int out_mode, void* stream);

/* a2 tail + a3 + a4 block 0 in one kernel (bf16 mode, the fused pipeline): the same block-0 kernel reading the RAW dB
 *        features of stac_fbank_logmel(_tc) and applying, in its loader, the top-dB clamp against the utterance maximum
 *        (utt_max_ordered as written by the Fbank kernel; per_utterance = 0: one maximum for the batch) and
 *        (x - mean) * (1 / std) (mean / std [80], both NULL: clamp only; one fp32 ulp from stac_fbank_topdb_norm's
 *        quotient), so the normalised [B,T,80] tensor is never written.  Replaces the tail of compute_features (inference.py:95),
 *        modules.normalize (:96) and block 0 of modules.CNN (:99).  out: the padded bf16 layout above. */
int stac_conv0_topdb_norm_bf16(const float* logmel_db, const uint32_t* utt_max_ordered, int per_utterance,
                               float top_db, const float* mean, const float* std, const float* w0,
                               const float* b0, const float* ln_g, const float* ln_b, int64_t batch,
                               int64_t frames, uint16_t* out, void* stream);

/* Block 1 convolution only (pre-LayerNorm), fp32 CUDA-core implicit GEMM:
 *   x [B,T1,40,256] fp32, w1 packed [256(out)][9 = kf*3+kt][256(in)], out [B,T2,20,256]. */
int stac_conv1_f32(const float* x, const float* w1, const float* b1, int64_t batch, int64_t t1,
                   float* out, void* stream);

/* Block 1 on tcgen05 tensor cores (bf16 operands, fp32 accumulate) with the whole block fused:
 * conv + bias, LayerNorm over (20,256) eps 1e-5, LeakyReLU(0.01), all in the GEMM epilogue.
 *   xpad: block-0 output in the padded parity-split bf16 layout above;
 *   w1_packed: bf16 [9 = kf*3+kt][256(out)][256(in)];  ln_g/ln_b [20*256];
 *   out: bf16 [B,T2,20*256] (what the src-linear GEMM of the encoder consumes). */
int stac_conv1_bf16(const uint16_t* xpad, const uint16_t* w1_packed, const float* b1,
                    const float* ln_g, const float* ln_b, int64_t batch, int64_t t1, uint16_t* out,
                    void* stream);

/* LayerNorm over a whole row of `dim` elements (dim = F*C = 10240 / 5120) + LeakyReLU. */
int stac_group_ln_lrelu(const float* x, int64_t rows, int64_t dim, const float* gamma,
                        const float* beta, float eps, float slope, void* out, int out_dtype,
                        void* stream);

/* ---------------------------------------------------------------------------
 * a5-a7  TransformerMultiTask.encode() building blocks
 *     (/root/reference/stac-st/modules/TransformerMultiTask.py:273-309)
 */
/* y = LayerNorm(x) * gamma + beta per row of `dim` (256/512/1024); either output may be NULL */
int stac_layernorm(const float* x, int64_t rows, int64_t dim, const float* gamma, const float* beta,
                   float eps, float* out_f32, uint16_t* out_bf16, void* stream);

/* C[M,N] = act(A[M,K] . W[N,K]^T + bias[N]) + resid[(row % resid_period), N]
 *   (resid NULL = none; resid_period 0 = row-aligned residual, may alias C).
 * fp32 CUDA-core version (fp32 mode).                                                */
int stac_gemm_f32(const float* a, const float* w, const float* bias, const float* resid,
                  int64_t resid_period, int act, float* c, int64_t m, int64_t n, int64_t k,
                  void* stream);

/* tcgen05 version: bf16 A/W (K multiple of 64, 16-byte aligned rows), fp32 accumulation in
 * TMEM, fused epilogue; c_dtype selects fp32 or bf16 output.
 * vt_out/vt_cols: if vt_out != NULL the last `vt_cols` output columns (the V third of a
 * packed QKV projection) are written transposed to vt_out as bf16 [B*H][64][t_pad]
 * (t = row % seq_len, b = row / seq_len) instead of to C.
 * Plain projections with K = 256 and N a multiple of 256 (the QKV and output projections of a d_model = 256
 * layer: bias only, bf16 store or in-place fp32 residual) over enough rows run on the weight-resident kernel
 * (csrc/gemm_wres.cu); environment STAC_WRES=0 forces the general kernel for A/B timing.  Same results either way.
 * Replaces nn.Linear / MultiheadAttention in_proj / out_proj (TransformerMultiTask.py:296,304-308). */
int stac_gemm_bf16(const uint16_t* a, const uint16_t* w, const float* bias, const float* resid,
                   int64_t resid_period, int act, void* c, int c_dtype, int64_t m, int64_t n,
                   int64_t k, uint16_t* vt_out, int64_t vt_cols, int64_t seq_len, int64_t t_pad,
                   void* stream);

/* Position-wise feed-forward block of one encoder layer in ONE kernel (bf16 mode, d_model = 256):
 *   x += W2 . GELU(W1 . h + b1) + b2      h bf16 [m, 256] (= LayerNorm2(x)), x fp32 [m, 256] in place,
 *   w1 bf16 [d_ffn, 256], w2 bf16 [256, d_ffn], d_ffn a multiple of 128 (<= 4096).
 * The hidden activation lives only in tensor / shared memory.  Other widths: STAC_ERR_UNSUPPORTED_SHAPE
 * (callers fall back to two stac_gemm_bf16 calls). */
int stac_ffn_fused_bf16(const uint16_t* h, const uint16_t* w1, const float* b1, const uint16_t* w2, const float* b2,
                        float* x, int64_t m, int64_t d_model, int64_t d_ffn, void* stream);

/* Valid key count per utterance, int32 [B], from the relative lengths the reference passes around
 * (wav_lens = len / Lmax, fp32) with the reference's own fp32 arithmetic:
 *   round_rule 0: encode()  keeps j <= floor(wav_len * T2)   -> floor(.) + 1   (TransformerMultiTask.py:289-294)
 *   round_rule 1: forward() keeps j <  round(wav_len * T2)   (round half to even, make_masks :225-226)
 * clamped to [1, T2]; wav_len NULL -> T2 for every utterance. */
int stac_kv_lengths(const float* wav_len, int64_t batch, int64_t t2, int round_rule, int32_t* out, void* stream);

/* Multi-head self-attention with key-padding by valid length (head_dim 64):
 *   qkv [B*T, 3*d] packed [q|k|v] with bias already added and q pre-scaled by 1/8,
 *   kv_len int32[B] (keys j < kv_len[b] are attended), ctx [B*T, d].
 * stac_mha_bf16: v_t = NULL reads V straight from qkv (MN-major tensor-core operand, the normal path);
 *   v_t != NULL is a K-major transposed copy bf16 [B*H][64][t_pad] (t_pad % 8 == 0, zero padded).      */
int stac_mha_f32(const float* qkv, const int32_t* kv_len, int64_t batch, int64_t seq_len,
                 int64_t d_model, int64_t n_head, float* ctx, void* stream);
int stac_mha_bf16(const uint16_t* qkv, const uint16_t* v_t, const int32_t* kv_len, int64_t batch,
                  int64_t seq_len, int64_t t_pad, int64_t d_model, int64_t n_head, uint16_t* ctx,
                  void* stream);
/* The attention kernel of the bf16 path (csrc/attention_tc2.cu; STAC_MHA_V2=0 in ops.py selects stac_mha_bf16 for
 * comparisons): same contract as stac_mha_bf16 with v_t = NULL.  96-key tiles, one thread per query row, P kept in TMEM
 * as the A operand of P.V, scores double-buffered per query group, O / l epilogue in its own warpgroup.            */
int stac_mha_bf16_v2(const uint16_t* qkv, const int32_t* kv_len, int64_t batch, int64_t seq_len,
                     int64_t d_model, int64_t n_head, uint16_t* ctx, void* stream);

/* ---------------------------------------------------------------------------
 * a8/a9  CTC head -- replaces hparams.log_softmax(modules.ctc_lin(enc_out)) and
 *        p_ctc.argmax(-1) (/root/reference/stac-st/inference.py:54-56,104-107)
 */
int stac_log_softmax(const float* logits, int64_t rows, int64_t vocab, float* out,
                     int32_t* argmax /*[rows] or NULL*/, void* stream);

/* a8 + a9 fused (bf16 mode): log_softmax(enc . W^T + b) and its argmax without ever materialising the logits.
 * Two passes of the tcgen05 GEMM over the same operands: pass 1 reduces every (row, 64-column group) to
 * (max, sum exp, argmax) in its epilogue and stores no logits, a small kernel combines the groups into the row's
 * log-sum-exp and greedy id, pass 2 recomputes the tile and stores logits - lse.  HBM traffic = the fp32
 * posteriors once (instead of three times) + the 4 % statistics workspace.
 *   enc bf16 [m, d_model], w bf16 [vocab, d_model], bias fp32 [vocab] or NULL,
 *   workspace: stac_ctc_head_workspace_floats(m, vocab) floats, log_probs [m, vocab] fp32 (STAC_DT_F32, the
 *   reference's contract) or bf16 (STAC_DT_BF16: what non-root ranks ship to rank 0), argmax int32 [m] or NULL */
int64_t stac_ctc_head_workspace_floats(int64_t m, int64_t vocab);
int stac_ctc_head_bf16(const uint16_t* enc, const uint16_t* w, const float* bias, int64_t m, int64_t vocab,
                       int64_t d_model, float* workspace, void* log_probs, int out_dtype, int32_t* argmax,
                       void* stream);

/* ---------------------------------------------------------------------------
 * turn detection -- replaces append_speaker_turns
 *        (/root/reference/stac-st/inference.py:54-84): argmax, == turn / == xt, dense masks to the host, Python loop
 *        over every frame.
 * stac_argmax_rows: out[r] = first index of the row maximum (for callers that hold posteriors, not greedy ids).
 * stac_ctc_spikes:  ids int32 [B, T2] -> ascending flat positions b * T2 + j of the frames equal to turn_id
 *        (spikes_turn) / xt_id (spikes_xt), both int32 [B * T2] capacity, and their counts n_out int32 [2];
 *        row_counts int32 [B][2] is workspace.  The order is the one the reference's loop appends RTTM lines in. */
int stac_argmax_rows(const float* x, int64_t rows, int64_t cols, int32_t* out, void* stream);
int stac_ctc_spikes(const int32_t* ids, int64_t batch, int64_t t2, int32_t turn_id, int32_t xt_id,
                    int32_t* row_counts, int32_t* spikes_turn, int32_t* spikes_xt, int32_t* n_out, void* stream);

/* ---------------------------------------------------------------------------
 * decoder side -- building blocks of TransformerMultiTask.decode()
 *        (/root/reference/stac-st/modules/TransformerMultiTask.py:234-271) and of the decoder half of forward()
 *        (:185-209) that the encoder entry points above do not cover.  First correct path: fp32, CUDA cores, the whole
 *        prefix per call as the reference's forward_step asks (mutitask_decoder.py:119-128).
 * stac_embed_scale_pe: out[row] = emb[tokens[row]] * scale + pe[row % seq_len]   (NormalizedEmbedding + positional
 *        encoding, :248-256); tokens int64 [rows], emb [vocab, d_model], pe [>= seq_len, d_model].
 * stac_attention_f32: softmax(q k^T + masks) v per head (head_dim 64; q is expected pre-scaled by 1/8).
 *        q row (r, i) at q + (r * lq + i) * ldq; key / value row j of row r at
 *        k|v + (r / mem_rows_div) * kv_batch_stride + j * kv_row_stride (packed [batch][lk][ld]: strides lk*ld, ld;
 *        time-major cache [lk][rows][ld]: strides ld, rows*ld); ctx row at ctx + (r * lq + i) * ldctx;
 *        head h uses columns h*64 .. h*64+63 of each.
 *        Masks: causal != 0 hides keys j > i; kv_len int32 [rows] (or NULL) hides j >= kv_len[r]; key_tokens int64
 *        [rows, lk] (or NULL) hides keys whose token equals pad_idx.  weights (or NULL): fp32 [rows, lq, lk], the
 *        probabilities averaged over heads (what nn.MultiheadAttention returns with need_weights=True).
 * stac_attention_beam_f32: the cross-attention of ONE decoding step for all hypothesis rows of an utterance at once
 *        (lq = 1, no causal / token masks; row r reads memory block r / group, group <= 16): a key / value row is loaded
 *        once per utterance instead of once per hypothesis.  Same arguments and arithmetic as stac_attention_f32;
 *        weights (or NULL): fp32 [rows, lk].  head_scratch (or NULL; only read when weights is given): fp32
 *        [n_head, rows, lk] of work space - with it the heads run on separate CTAs (grid utterances x heads) and a second
 *        kernel adds their probabilities up in head order; without it one CTA per utterance walks the heads (the same
 *        sums, four times fewer CTAs).  STAC_ERR_UNSUPPORTED_SHAPE: group > 16 or lk too long for shared memory
 *        (use stac_attention_f32).
 * stac_attention_step_f32: the self-attention of ONE decoding step over a time-major cache (lq = 1, every row attends
 *        keys 0 .. lk-1; one warp per (row, head), any lk).  Key / value j of hypothesis row r at
 *        k|v + src * kv_row_stride + j * kv_time_stride with src = row_map[j * rows + r] (int32 [lk, rows]) or src = r
 *        when row_map is NULL: a beam search re-orders its hypotheses by permuting the small map (DecoderCache.reorder)
 *        instead of gathering the cached prefix (reference: permute_mem / the index_select of
 *        /root/reference/stac-st/modules/mutitask_decoder.py:109-112 on the token memory; the reference has no cache).
 *        Append mode (t_dev, k_new, v_new given; else all NULL / 0): the position counter t lives on the device
 *        (int32 [1]), lk is the capacity of the cache, the step attends keys 0 .. t where key t is row r of
 *        k_new / v_new (row stride ld_new: the output of the step's K|V projection) and is stored into slab t of the
 *        cache by this kernel - no argument of the launch depends on the step, so one captured launch sequence serves a
 *        whole search.
 * stac_embed_step: stac_embed_scale_pe for one position per row with the position read from the device
 *        (out[row] = emb[tokens[row]] * scale + pe[*pos_dev]). */
int stac_embed_scale_pe(const int64_t* tokens, const float* emb, const float* pe, int64_t rows, int64_t seq_len,
                        int64_t d_model, int64_t vocab, float scale, float* out, void* stream);
int stac_attention_f32(const float* q, int64_t ldq, const float* k, const float* v, int64_t kv_batch_stride,
                       int64_t kv_row_stride, int64_t rows, int64_t lq, int64_t lk, int64_t n_head,
                       int64_t mem_rows_div, int causal,
                       const int32_t* kv_len, const int64_t* key_tokens, int64_t pad_idx, float* ctx, int64_t ldctx,
                       float* weights, void* stream);
int stac_embed_step(const int64_t* tokens, const float* emb, const float* pe, int64_t rows, int64_t d_model,
                    int64_t vocab, float scale, const int32_t* pos_dev, float* out, void* stream);
int stac_attention_step_f32(const float* q, int64_t ldq, float* k, float* v, int64_t kv_row_stride,
                            int64_t kv_time_stride, int64_t rows, int64_t lk, int64_t n_head, const int32_t* row_map,
                            const float* k_new, const float* v_new, int64_t ld_new, const int32_t* t_dev, float* ctx,
                            int64_t ldctx, void* stream);
int stac_attention_beam_f32(const float* q, int64_t ldq, const float* k, const float* v, int64_t kv_batch_stride,
                            int64_t kv_row_stride, int64_t rows, int64_t group, int64_t lk, int64_t n_head,
                            const int32_t* kv_len, float* ctx, int64_t ldctx, float* weights, float* head_scratch,
                            void* stream);

/* ---------------------------------------------------------------------------
 * host ingest -- in front of a2: replaces shipping the fp32 waveform that librosa.load produced
 *        (/root/reference/stac-st/inference.py:250-261, batch.to(device) :91).  out[i] = pcm[i] / 32768 (exactly the
 *        value a 16-bit file decodes to); pcm and out 16-byte aligned. */
int stac_pcm_i16_to_f32(const int16_t* pcm, int64_t n, float* out, void* stream);

/* ---------------------------------------------------------------------------
 * training-stage statistics of InputNormalization -- replaces the per-utterance torch.mean / torch.std
 *        loop of normalize(feats, wav_lens, epoch) in train mode (/root/reference/stac-st/train_multitask.py:60-61):
 *        mean / unbiased std (floored at eps) of x[b, :round(wav_len[b] * frames), :] per utterance and bin;
 *        x [batch, frames, dim], wav_len fp32 [batch], mean / std fp32 [batch, dim]. */
int stac_utt_mean_std(const float* x, const float* wav_len, int64_t batch, int64_t frames, int64_t dim, float eps,
                      float* mean, float* std, void* stream);

/* 8f-4, training-stage pieces on the path's tensors.
 * stac_spec_augment: SpeechBrain's SpecAugment (`feats = self.hparams.augmentation(feats)`,
 *        /root/reference/stac-st/train_multitask.py:63-66, yaml transformer_multitask.yaml:283-293) in one pass: bicubic
 *        time warp (rows [0, warp_width) = rows [0, warp_center) resampled, rows [warp_width, frames) = rows
 *        [warp_center, frames) resampled, align_corners; warp_width < 0: none), then frequency / time masks
 *        [pos, pos + len) per utterance filled with `fill`.  x, out [batch, frames, dim] fp32 (out != x);
 *        freq_pos / freq_len int32 [batch, n_freq], time_pos / time_len int32 [batch, n_time].  The random parameters are
 *        the host's (augment.SpecAugment draws them in SpeechBrain's order).
 * stac_ctc_loss: speechbrain.nnet.losses.ctc_loss (`self.hparams.ctc_cost(p_ctc, tokens, wav_lens, tokens_lens)`,
 *        train_multitask.py:164-170, yaml :256-258) = torch.nn.functional.ctc_loss(zero_infinity=True): forward value.
 *        log_probs [batch, frames, vocab] fp32, targets int32 [batch, max_targets], input_len / target_len int32 [batch]
 *        (absolute), nll fp32 [batch] (always written), loss: reduction 0 none (untouched), 1 sum, 2 mean (torch:
 *        per-target-length then batch mean), 3 batchmean (sum / batch), 4 batch (fp32 [batch]: nll / target_len). */
int stac_spec_augment(const float* x, int64_t batch, int64_t frames, int64_t dim, int warp_center, int warp_width,
                      const int32_t* freq_pos, const int32_t* freq_len, int n_freq, const int32_t* time_pos,
                      const int32_t* time_len, int n_time, float fill, float* out, void* stream);
int stac_ctc_loss(const float* log_probs, const int32_t* targets, const int32_t* input_len, const int32_t* target_len,
                  int64_t batch, int64_t frames, int64_t vocab, int64_t max_targets, int blank, int reduction,
                  float* nll, float* loss, void* stream);

/* a7, bf16 mode, d_model = 256: attention output projection + residual add + the LayerNorm that follows, one kernel
 *        (north_star's "fused LayerNorm"; replaces `src = src + dropout1(self_att(...))` and `norm2(src)` of SpeechBrain's
 *        pre-LN TransformerEncoderLayer, reached from /root/reference/stac-st/modules/TransformerMultiTask.py:304-308):
 *          x[m,256] (fp32, in place) += a[m,256] (bf16) . w[256,256]^T (bf16) + bias;   h[m,256] (bf16) = LayerNorm(x) */
int stac_outproj_ln_bf16(const uint16_t* a, const uint16_t* w, const float* bias, float* x, const float* ln_g,
                         const float* ln_b, float eps, uint16_t* h, int64_t m, void* stream);

/* fp32 -> bf16 conversion (weight packing / activation hand-off) */
int stac_cast_bf16(const float* x, int64_t n, uint16_t* out, void* stream);
/* bf16 -> fp32 (exact): the staged drop-in CNN returns fp32 as the reference's does (inference.py:99) while conv1's
 *        epilogue produces bf16 */
int stac_cast_f32(const uint16_t* x, int64_t n, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STAC_B200_H_ */
