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
"""
from __future__ import annotations

from typing import Optional, Tuple

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
