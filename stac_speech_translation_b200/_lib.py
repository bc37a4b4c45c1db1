"""ctypes binding of libstac_b200.so (the C ABI declared in include/stac_b200.h).

There is no CPU fallback: if the shared library is missing or a tensor is not on a CUDA
device the call raises.  Build the library with ``python -m stac_speech_translation_b200.build``
(or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_void_p
from pathlib import Path

import torch

# STAC_B200_LIB selects a variant build (python -m stac_speech_translation_b200.build --variant ...) for experiments
LIB_PATH = Path(os.environ.get("STAC_B200_LIB") or Path(__file__).resolve().parent / "libstac_b200.so")

OK = 0
ACT_NONE, ACT_GELU_ERF = 0, 1
DT_F32, DT_BF16 = 0, 1

_P = c_void_p
_SIGNATURES = {
    "stac_version": (c_int, []),
    "stac_error_string": (c_char_p, [c_int]),
    "stac_set_reserved_sms": (c_int, [c_int]),
    "stac_l2_persist": (c_int, [_P, c_int64, c_float, _P]),
    "stac_l2_persist_limits": (c_int, [_P, _P]),
    "stac_fbank_tables_floats": (c_int, []),
    "stac_fbank_logmel": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P, _P, _P]),
    "stac_fbank_tc_tables_floats": (c_int, []),
    "stac_fbank_tc_twiddle_halfs": (c_int, []),
    "stac_fbank_logmel_tc": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P]),
    "stac_fbank_tc2_tables_floats": (c_int, []),
    "stac_fbank_tc2_twiddle_halfs": (c_int, []),
    "stac_fbank_logmel_tc2": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P, _P, _P, c_int, _P]),
    "stac_spec_augment": (c_int, [_P, c_int64, c_int64, c_int64, c_int, c_int, _P, _P, c_int, _P, _P, c_int, c_float, _P, _P]),
    "stac_ctc_loss": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int64, c_int64, c_int, c_int, _P, _P, _P]),
    "stac_fbank_topdb_norm": (c_int, [_P, _P, c_int, c_float, _P, _P, c_int64, c_int64, c_int64, _P, _P]),
    "stac_input_norm": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P]),
    "stac_conv0_padded_elems": (c_int64, [c_int64, c_int64]),
    "stac_conv0_ln_lrelu": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, _P, c_int, _P]),
    "stac_conv0_topdb_norm_bf16": (c_int, [_P, _P, c_int, c_float, _P, _P, _P, _P, _P, _P, c_int64, c_int64, _P, _P]),
    "stac_conv1_f32": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P]),
    "stac_conv1_bf16": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, _P, _P]),
    "stac_group_ln_lrelu": (c_int, [_P, c_int64, c_int64, _P, _P, c_float, c_float, _P, c_int, _P]),
    "stac_layernorm": (c_int, [_P, c_int64, c_int64, _P, _P, c_float, _P, _P, _P]),
    "stac_gemm_f32": (c_int, [_P, _P, _P, _P, c_int64, c_int, _P, c_int64, c_int64, c_int64, _P]),
    "stac_gemm_bf16": (c_int, [_P, _P, _P, _P, c_int64, c_int, _P, c_int, c_int64, c_int64, c_int64,
                               _P, c_int64, c_int64, c_int64, _P]),
    "stac_ffn_fused_bf16": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int64, _P]),
    "stac_kv_lengths": (c_int, [_P, c_int64, c_int64, c_int, _P, _P]),
    "stac_mha_f32": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, _P, _P]),
    "stac_mha_bf16": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, _P, _P]),
    "stac_mha_bf16_v2": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, _P, _P]),
    "stac_log_softmax": (c_int, [_P, c_int64, c_int64, _P, _P, _P]),
    "stac_ctc_head_workspace_floats": (c_int64, [c_int64, c_int64]),
    "stac_ctc_head_bf16": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, _P, _P, c_int, _P, _P]),
    "stac_argmax_rows": (c_int, [_P, c_int64, c_int64, _P, _P]),
    "stac_ctc_spikes": (c_int, [_P, c_int64, c_int64, c_int32, c_int32, _P, _P, _P, _P, _P]),
    "stac_embed_scale_pe": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, c_int64, c_float, _P, _P]),
    "stac_attention_f32": (c_int, [_P, c_int64, _P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int,
                                   _P, _P, c_int64, _P, c_int64, _P, _P]),
    "stac_embed_step": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, c_float, _P, _P, _P]),
    "stac_attention_step_f32": (c_int, [_P, c_int64, _P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, _P, _P, _P,
                                        c_int64, _P, _P, c_int64, _P]),
    "stac_attention_beam_f32": (c_int, [_P, c_int64, _P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, _P, _P,
                                        c_int64, _P, _P, _P]),
    "stac_pcm_i16_to_f32": (c_int, [_P, c_int64, _P, _P]),
    "stac_utt_mean_std": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_float, _P, _P, _P]),
    "stac_outproj_ln_bf16": (c_int, [_P, _P, _P, _P, _P, _P, c_float, _P, c_int64, _P]),
    "stac_cast_bf16": (c_int, [_P, c_int64, _P, _P]),
    "stac_cast_f32": (c_int, [_P, c_int64, _P, _P]),
}

_lib = None


class StacB200Error(RuntimeError):
    pass


def exported_symbols():
    """Names include/stac_b200.h declares (used by the load/export test)."""
    return sorted(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise StacB200Error(
                f"{LIB_PATH} is missing: the CUDA extension has not been built and there is no "
                "CPU fallback. Run `python -m stac_speech_translation_b200.build`.")
        handle = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(code: int, what: str = "") -> None:
    if code != OK:
        msg = lib().stac_error_string(code).decode()
        raise StacB200Error(f"{what or 'libstac_b200'} failed: {msg} (code {code})")


def stream() -> c_void_p:
    """Current stream of the CURRENT device.  The library is one-device-per-process (one process per GPU, as the
    multi-GPU path launches it): ptr() refuses tensors that live on another device, so a kernel can never be enqueued
    on one GPU with pointers of another."""
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise StacB200Error("stac_b200 kernels need CUDA tensors; there is no CPU fallback")
    if t.device.index != torch.cuda.current_device():
        raise StacB200Error(f"tensor on cuda:{t.device.index} but the current device is cuda:{torch.cuda.current_device()}"
                            " (stac_b200 is one device per process: torch.cuda.set_device first)")
    if not t.is_contiguous():
        raise StacB200Error("stac_b200 kernels need contiguous tensors")
    if dtype is not None and t.dtype != dtype:
        raise StacB200Error(f"expected {dtype}, got {t.dtype}")
    return c_void_p(t.data_ptr())
