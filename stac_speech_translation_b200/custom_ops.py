"""torch custom-op registration of the C-ABI entry points that are tensor-in / tensor-out (optional layer).

The drop-in classes call ``ops.*`` directly; importing this module additionally registers the same calls under
``torch.ops.stac_b200.*`` (``torch.library.custom_op`` with fake-tensor shape functions), which is what lets
``torch.compile`` / ``torch.export`` / FakeTensor tracing see through a model that uses the drop-ins instead of breaking
the graph at a ctypes call.  Each op is a thin wrapper: argument checks and allocation in ``ops.py``, arithmetic in
``libstac_b200.so``.  No CPU implementation is registered: on CPU tensors the ops raise like the rest of the package.

    import stac_speech_translation_b200.custom_ops          # registers
    y = torch.ops.stac_b200.input_norm(feats, mean, std)

| op | replaces (reference call site) | C entry points |
|---|---|---|
| ``fbank(wavs, top_db, per_utterance)``     | ``compute_features(wavs)`` inference.py:95 (fp32 FFT kernel)   | stac_fbank_logmel, stac_fbank_topdb_norm |
| ``input_norm(x, mean, std)``               | ``normalize(feats, wav_lens)`` inference.py:96 (eval)           | stac_input_norm |
| ``linear(x, weight, bias, precision)``     | ``ctc_lin(enc_out)`` / ``seq_lin`` inference.py:105             | stac_gemm_f32 / stac_gemm_bf16 |
| ``log_softmax(logits)``                    | ``log_softmax(logits)`` inference.py:106                        | stac_log_softmax |
| ``log_softmax_greedy(logits)``             | the same + ``.argmax(-1)`` inference.py:58                      | stac_log_softmax |
| ``argmax_rows(x)``                         | ``model_ctc_outputs.argmax(-1)`` inference.py:58                | stac_argmax_rows |
| ``pcm_to_float(pcm)``                      | fp32 decode of 16-bit PCM in front of ``batch.to(device)`` :91  | stac_pcm_i16_to_f32 |
| ``conv_frontend(feats, weights, precision)`` | ``CNN(feats)`` inference.py:99                                 | stac_conv0_ln_lrelu, stac_conv1_* |
| ``encoder(src, kv_len, weights, ...)``     | ``Transformer.encode`` after the mask rule, :100               | stac_gemm_*, stac_layernorm, stac_mha_*, stac_ffn_fused_bf16 |
| ``ctc_head(enc_bf16, weight, bias)``       | ``log_softmax(ctc_lin(enc_out))`` + arg-max, :105-106 / :58    | stac_ctc_head_bf16 |

The three structured ops take the packed weights as flat tensor lists (``frontend_weight_list`` /
``encoder_weight_list`` turn the packed dataclasses of ``ops.py`` into them).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import ingest, ops, turns

_FBANK_TABLES = {}


@torch.library.custom_op("stac_b200::fbank", mutates_args=())
def fbank(wavs: torch.Tensor, top_db: float, per_utterance: bool) -> torch.Tensor:
    key = str(wavs.device)
    if key not in _FBANK_TABLES:
        _FBANK_TABLES[key] = ops.build_fbank_tables(wavs.device)
    return ops.fbank(wavs, _FBANK_TABLES[key], top_db, per_utterance)


@fbank.register_fake
def _(wavs, top_db, per_utterance):
    return wavs.new_empty(wavs.shape[0], 1 + wavs.shape[1] // ops.HOP, ops.N_MELS, dtype=torch.float32)


@torch.library.custom_op("stac_b200::input_norm", mutates_args=())
def input_norm(x: torch.Tensor, mean: torch.Tensor, std: torch.Tensor) -> torch.Tensor:
    return ops.input_norm(x, mean, std)


@input_norm.register_fake
def _(x, mean, std):
    return torch.empty_like(x, memory_format=torch.contiguous_format)


@torch.library.custom_op("stac_b200::linear", mutates_args=())
def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], precision: str) -> torch.Tensor:
    return ops.linear(x, weight, bias, precision)


@linear.register_fake
def _(x, weight, bias, precision):
    return x.new_empty(*x.shape[:-1], weight.shape[0], dtype=torch.float32)


@torch.library.custom_op("stac_b200::log_softmax", mutates_args=())
def log_softmax(logits: torch.Tensor) -> torch.Tensor:
    return ops.log_softmax(logits)


@log_softmax.register_fake
def _(logits):
    return torch.empty_like(logits, dtype=torch.float32, memory_format=torch.contiguous_format)


@torch.library.custom_op("stac_b200::log_softmax_greedy", mutates_args=())
def log_softmax_greedy(logits: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    out, ids = ops.log_softmax(logits, want_argmax=True)
    return out, ids


@log_softmax_greedy.register_fake
def _(logits):
    return (torch.empty_like(logits, dtype=torch.float32, memory_format=torch.contiguous_format),
            logits.new_empty(logits.shape[:-1], dtype=torch.int32))


@torch.library.custom_op("stac_b200::argmax_rows", mutates_args=())
def argmax_rows(x: torch.Tensor) -> torch.Tensor:
    if x.dtype != torch.float32 or x.dim() != 3:         # (greedy_ids passes int32 ids through: an op must not alias)
        raise ops._lib.StacB200Error("argmax_rows expects fp32 posteriors [B, T2, V]")
    return turns.greedy_ids(x)


@argmax_rows.register_fake
def _(x):
    return x.new_empty(x.shape[:-1], dtype=torch.int32)


@torch.library.custom_op("stac_b200::pcm_to_float", mutates_args=())
def pcm_to_float(pcm: torch.Tensor) -> torch.Tensor:
    return ingest.pcm_to_float(pcm)


@pcm_to_float.register_fake
def _(pcm):
    return torch.empty_like(pcm, dtype=torch.float32, memory_format=torch.contiguous_format)


# ---- structured stages: packed weights travel as flat tensor lists ----
def frontend_weight_list(w: ops.FrontendWeights) -> List[torch.Tensor]:
    return [w.w0, w.b0, w.g0, w.be0, w.w1, w.b1, w.g1, w.be1]


def encoder_weight_list(w: ops.EncoderWeights) -> List[torch.Tensor]:
    """[w_src, b_src, pe, lnf_g, lnf_b] + 12 tensors per layer in LayerWeights field order."""
    out = [w.w_src, w.b_src, w.pe, w.lnf_g, w.lnf_b]
    for L in w.layers:
        out += [L.ln1_g, L.ln1_b, L.w_qkv, L.b_qkv, L.w_o, L.b_o, L.ln2_g, L.ln2_b, L.w_1, L.b_1, L.w_2, L.b_2]
    return out


@torch.library.custom_op("stac_b200::conv_frontend", mutates_args=())
def conv_frontend(feats: torch.Tensor, weights: Sequence[torch.Tensor], precision: str, out_bf16: bool) -> torch.Tensor:
    w = ops.FrontendWeights(precision, *weights)
    return ops.conv_frontend(feats, w, torch.bfloat16 if out_bf16 else torch.float32)


@conv_frontend.register_fake
def _(feats, weights, precision, out_bf16):
    t2 = ops.frames_of((feats.shape[1] - 1) * ops.HOP)[2]
    return feats.new_empty(feats.shape[0], t2, ops.F2 * ops.CNN_CH, dtype=torch.bfloat16 if out_bf16 else torch.float32)


@torch.library.custom_op("stac_b200::encoder", mutates_args=())
def encoder(src: torch.Tensor, kv_len: torch.Tensor, weights: Sequence[torch.Tensor], d_model: int, nhead: int,
            precision: str) -> torch.Tensor:
    ws = list(weights)
    if (len(ws) - 5) % 12 != 0:
        raise ops._lib.StacB200Error("encoder weight list: 5 tensors + 12 per layer expected")
    w = ops.EncoderWeights(precision, d_model, nhead, ws[0], ws[1], ws[2],
                           [ops.LayerWeights(*ws[i:i + 12]) for i in range(5, len(ws), 12)], ws[3], ws[4])
    return ops.encoder_stack(src, w, kv_len)


@encoder.register_fake
def _(src, kv_len, weights, d_model, nhead, precision):
    return src.new_empty(src.shape[0], src.shape[1], d_model, dtype=torch.float32)


@torch.library.custom_op("stac_b200::ctc_head", mutates_args=())
def ctc_head(enc_bf16: torch.Tensor, weight_bf16: torch.Tensor, bias: Optional[torch.Tensor]
             ) -> Tuple[torch.Tensor, torch.Tensor]:
    p, ids = ops.ctc_head_bf16(enc_bf16, weight_bf16, bias)
    return p, ids


@ctc_head.register_fake
def _(enc_bf16, weight_bf16, bias):
    return (enc_bf16.new_empty(*enc_bf16.shape[:-1], weight_bf16.shape[0], dtype=torch.float32),
            enc_bf16.new_empty(enc_bf16.shape[:-1], dtype=torch.int32))
