"""Host-side orchestration of the libstac_b200 kernels (one function per stage of the path).

Everything here only allocates torch tensors and enqueues kernels on the current CUDA stream
through the C ABI; there is no torch arithmetic on the hot path.  Stage numbering follows
SURVEY.md section 8(a): a2 Fbank, a3 InputNormalization, a4 ConvolutionFrontEnd, a5-a7 encoder,
a8/a9 CTC head.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import List, NamedTuple, Optional

import torch

from . import _lib
from ._lib import ACT_GELU_ERF, ACT_NONE, DT_BF16, DT_F32, check, lib, ptr, stream

N_MELS = 80
N_FFT = 400
HOP = 160
CNN_CH = 256
F1, F2 = 40, 20
PRECISIONS = ("fp32", "bf16")
FUSED_FFN = True       # tests flip this to compare against the two-GEMM path
# Top-dB clamp + normalisation inside conv block 0's loader (stac_conv0_topdb_norm_bf16) instead of their own pass.
# Measured on a B200 at the benchmark shape (profiles/r3/): the separate pass costs 0.050 ms, but block 0's producer
# warps are its critical path and the three extra operations per feature cost it 0.076 ms (0.518 -> 0.594 ms; 0.615 ms
# with a true division, 1.19 ms when the loader consumed its prefetch early).  Correct (tests/test_gpu_tc_conv.py), not
# faster: off by default.
FUSED_CONV0_NORM = os.environ.get("STAC_FUSED_CONV0_NORM", "0") == "1"
# Attention output projection + residual add + LayerNorm 2 in ONE kernel (stac_outproj_ln_bf16, d_model 256): the
# LayerNorm folded into the producer of the residual stream (north_star's "fused LayerNorm").  Correct
# (tests/test_gpu_tc_gemm.py, path tests) and 25 % less HBM traffic than GEMM + LayerNorm, but measured NOT faster on a
# B200: 43.4 us per layer against 27.5 + 15.1 us (profiles/r3/r3f_call10_ln_fusion_ab.log) - with whole rows per thread
# the epilogue is one serial chain of TMEM loads, row loads and six staged TMA stores per warp and tile, and the
# two-kernel form overlaps better.  Folding the following LayerNorm into the fused feed-forward kernel the same way was
# measured too and cost it 36 us per launch (its epilogue is the tile boundary the tensor pipe already waits for); that
# variant was removed.  Off by default.
FUSED_OUTPROJ_LN = os.environ.get("STAC_FUSED_OUTPROJ_LN", "0") == "1"
# Attention kernel of the bf16 path: csrc/attention_tc2.cu (P in TMEM, double-buffered scores; 69.7 us against 87 us at
# the benchmark shape).  STAC_MHA_V2=0 selects the first kernel (csrc/attention_tc.cu) for comparisons.
MHA_V2 = os.environ.get("STAC_MHA_V2", "1") == "1"
# L2 residency hint for the encoder's fp32 residual stream (stac_l2_persist); STAC_L2_PERSIST_RATIO = hit ratio of the window
L2_PERSIST = os.environ.get("STAC_L2_PERSIST", "0") == "1"
L2_PERSIST_RATIO = float(os.environ.get("STAC_L2_PERSIST_RATIO", "1.0"))

# Optional launch tracing: bench.py sets TRACE to a list to get (kernel, label, start_event, end_event)
# per launch; LAUNCHES counts kernel launches either way.
TRACE = None
TRACE_FILTER = None   # optional set of trace keys ("kernel" or "kernel:label"): only those launches get events
LAUNCHES = 0
_LABEL = ""


def trace_key(name, label):
    return f"{name}:{label}" if name == "stac_gemm_bf16" or name == "stac_gemm_f32" else name


def _call(name, *args):
    """Enqueue one libstac_b200 kernel on the current stream (optionally bracketed by CUDA events)."""
    global LAUNCHES
    LAUNCHES += 1
    if TRACE is None or (TRACE_FILTER is not None and trace_key(name, _LABEL) not in TRACE_FILTER):
        check(getattr(lib(), name)(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(getattr(lib(), name)(*args), name)
    e1.record()
    TRACE.append((name, _LABEL, e0, e1))


class label:
    """Context manager naming the launches inside it (for the per-kernel timing table)."""

    def __init__(self, text):
        self.text = text

    def __enter__(self):
        global _LABEL
        self.prev, _LABEL = _LABEL, self.text

    def __exit__(self, *exc):
        global _LABEL
        _LABEL = self.prev


def frames_of(n_samples: int):
    t = 1 + n_samples // HOP
    t1 = (t - 1) // 2 + 1
    t2 = (t1 - 1) // 2 + 1
    return t, t1, t2


# --------------------------------------------------------------------------
# a2 / a3
# --------------------------------------------------------------------------
def mel_filter_matrix() -> torch.Tensor:
    """[80, 201] triangular mel filterbank with the same fp32 torch expressions SpeechBrain's Filterbank uses
    (filters symmetric in Hz with the LEFT band as half-width; 201 linear bins 0..8 kHz)."""
    def to_mel(hz):
        return 2595 * math.log10(1 + hz / 700)

    mel = torch.linspace(to_mel(0), to_mel(8000.0), N_MELS + 2)
    hz = 700 * (10 ** (mel / 2595) - 1)
    band = (hz[1:] - hz[:-1])[:-1]
    f_central = hz[1:-1]
    all_freqs = torch.linspace(0, 16000 // 2, N_FFT // 2 + 1)
    slope = (all_freqs[None, :] - f_central[:, None]) / band[:, None]          # [80, 201]
    return torch.max(torch.zeros(1), torch.min(slope + 1.0, -slope + 1.0))       # [80, 201]


def build_fbank_tables(device) -> torch.Tensor:
    """Constant block of the Fbank kernel (layout documented in include/stac_b200.h).

    Window and mel weights are computed with the same fp32 torch expressions SpeechBrain's
    STFT / Filterbank use (hamming_window(400); triangular filters symmetric in Hz with the LEFT
    band as half-width; 201 linear bins 0..8 kHz), twiddles in float64 then rounded."""
    window = torch.hamming_window(N_FFT)
    k = torch.arange(200, dtype=torch.float64)
    tw200 = torch.stack([torch.cos(2 * math.pi * k / 200), -torch.sin(2 * math.pi * k / 200)], -1)
    k = torch.arange(25, dtype=torch.float64)
    tw25 = torch.stack([torch.cos(2 * math.pi * k / 25), -torch.sin(2 * math.pi * k / 25)], -1)
    k = torch.arange(201, dtype=torch.float64)
    tw400 = torch.stack([torch.cos(2 * math.pi * k / 400), -torch.sin(2 * math.pi * k / 400)], -1)

    fb = mel_filter_matrix()
    start = torch.zeros(N_MELS)
    count = torch.zeros(N_MELS)
    weights = torch.zeros(N_MELS, 16)
    for m in range(N_MELS):
        nz = torch.nonzero(fb[m] > 0).flatten()
        s, e = int(nz[0]), int(nz[-1]) + 1
        if e - s > 16:
            raise ValueError("mel filter wider than 16 taps")
        start[m], count[m] = s, e - s
        weights[m, : e - s] = fb[m, s:e]
    tab = torch.cat([window.float(), tw200.float().flatten(), tw25.float().flatten(),
                     tw400.float().flatten(), start, count, weights.flatten()])
    assert tab.numel() == lib().stac_fbank_tables_floats()
    return tab.to(device=device, dtype=torch.float32).contiguous()


_FFT_TABLES = {}


class FbankTcTables(NamedTuple):
    """Constants of the two tensor-core Fbank kernels (layouts in include/stac_b200.h)."""
    tab: torch.Tensor      # fp32: window[400] | per-bin mel weights [208][2]                    (stac_fbank_logmel_tc)
    tw: torch.Tensor       # fp16: [cos | sin][208 bins][256 columns]
    tab2: torch.Tensor     # fp32: per-bin mel weights [208][2]                                  (stac_fbank_logmel_tc2)
    tw2: torch.Tensor      # fp16: [7 stages][208 bins][64 columns: w[n] cos, n = 32 i .. + 31 | -w[n] sin, same n]


def build_fbank_tc_tables(device) -> FbankTcTables:
    """Constants of the tensor-core Fbank kernels.  Window and mel weights come from the same fp32 torch expressions
    SpeechBrain uses; the per-bin mel layout must have the sparsity structure baked into csrc/fbank_mel_structure.h
    (checked here); twiddles are computed in float64 and rounded to fp16 once."""
    window = torch.hamming_window(N_FFT)
    fb = mel_filter_matrix()                                   # [80, 201]
    wbin = torch.zeros(208, 2)
    for k in range(201):
        nz = torch.nonzero(fb[:, k] > 0).flatten().tolist()
        if len(nz) > 2 or (len(nz) == 2 and nz[1] != nz[0] + 1):
            raise ValueError("mel filterbank structure differs from csrc/fbank_mel_structure.h")
        for j, m in enumerate(nz):
            wbin[k, j] = fb[m, k]
    tab = torch.cat([window.float(), wbin.flatten()])
    assert tab.numel() == lib().stac_fbank_tc_tables_floats()
    kk = torch.arange(201, dtype=torch.float64)[:, None]
    tw = torch.zeros(2, 208, 256, dtype=torch.float64)
    n_cos = torch.arange(201, dtype=torch.float64)[None, :]
    tw[0, :201, :201] = torch.cos(2 * math.pi * kk * n_cos / 400)
    n_sin = torch.arange(1, 200, dtype=torch.float64)[None, :]
    tw[1, :201, :199] = -torch.sin(2 * math.pi * kk * n_sin / 400)
    tw = tw.to(torch.float16).reshape(2 * 208, 256).contiguous()
    assert tw.numel() == lib().stac_fbank_tc_twiddle_halfs()

    # second kernel: the frame folded on its symmetry, e[n] = x[n] + x[400 - n] (cos side), o[n] = x[n] - x[400 - n] (sin
    # side), n = 0..223 in 7 stages of 32 (n = 0 and n = 200 have no mirror partner, n > 200 does not exist: zero
    # twiddle columns); the hamming window (symmetric, w[400 - n] = w[n]) is folded into the twiddles in float64
    tab2 = wbin.flatten().clone()
    assert tab2.numel() == lib().stac_fbank_tc2_tables_floats()
    n_all = torch.arange(224, dtype=torch.float64)[None, :]
    kk208 = torch.arange(208, dtype=torch.float64)[:, None]
    w64 = torch.zeros(224, dtype=torch.float64)
    w64[:201] = window[:201].double()
    cos_t = torch.cos(2 * math.pi * kk208 * n_all / 400) * w64[None, :]        # [208, 224]
    sin_t = -torch.sin(2 * math.pi * kk208 * n_all / 400) * w64[None, :]
    cos_t[201:] = 0
    sin_t[201:] = 0
    sin_t[:, 200:] = 0
    sin_t[:, 0] = 0
    tw2 = torch.cat([cos_t.view(208, 7, 32), sin_t.view(208, 7, 32)], dim=2).permute(1, 0, 2)   # [7, 208, 64]
    tw2 = tw2.to(torch.float16).reshape(7 * 208, 64).contiguous()
    assert tw2.numel() == lib().stac_fbank_tc2_twiddle_halfs()
    return FbankTcTables(tab.to(device=device, dtype=torch.float32).contiguous(), tw.to(device),
                         tab2.to(device=device, dtype=torch.float32).contiguous(), tw2.to(device))


@dataclass
class RawFeatures:
    """Un-clamped, un-normalised dB features of the Fbank kernel + what the consumer needs to finish a2 / a3 in its own
    loader (stac_conv0_topdb_norm_bf16): the fused pipeline never writes the normalised [B, T, 80] tensor."""
    db: torch.Tensor                  # [B, T, 80] fp32: 10 log10(max(mel, 1e-10))
    utt_max: torch.Tensor             # [B] int32: order-preserving key of the utterance maximum
    top_db: float
    per_utterance: bool
    mean: Optional[torch.Tensor]
    std: Optional[torch.Tensor]


def fbank_tc(wavs: torch.Tensor, tc_tables, top_db: float = 80.0, per_utterance: bool = True,
             mean: Optional[torch.Tensor] = None, std: Optional[torch.Tensor] = None, raw: bool = False):
    """a2 (+a3) with the STFT on tensor cores (bf16 mode): [B, L] fp32 PCM -> [B, T, 80] fp32 features, or (raw=True)
    the RawFeatures hand-off for conv_frontend."""
    if wavs.dim() != 2:
        raise _lib.StacB200Error("Fbank expects [batch, samples] waveforms")
    wavs = wavs.contiguous()
    b, n = wavs.shape
    if n % 4 != 0 or wavs.data_ptr() % 16 != 0:
        # the tensor-core kernels move PCM tiles with 16-byte bulk copies; odd lengths take the exact FFT kernel
        key = str(wavs.device)
        if key not in _FFT_TABLES:
            _FFT_TABLES[key] = build_fbank_tables(wavs.device)
        return fbank(wavs, _FFT_TABLES[key], top_db, per_utterance, mean, std, raw=raw)
    t = 1 + n // HOP
    db = torch.empty(b, t, N_MELS, device=wavs.device, dtype=torch.float32)
    umax = torch.empty(b, device=wavs.device, dtype=torch.int32)      # zeroed by the call (stream-ordered memset)
    if os.environ.get("STAC_FBANK_V2", "1") != "0" and n % 32 == 0:
        # second design (A operand in tensor memory; PCM tiles as tensor-map boxes of 32-sample rows, hence the multiple
        # of 32) as two-CTA cta_group::2 instances (63.5 us at the benchmark shape); STAC_FBANK_PAIR=0: the same kernel
        # per single CTA (67.6 us)
        _call("stac_fbank_logmel_tc2", ptr(wavs, torch.float32), b, n, wavs.stride(0), ptr(tc_tables.tab2),
              ptr(tc_tables.tw2, torch.float16), ptr(db), ptr(umax),
              int(os.environ.get("STAC_FBANK_PAIR", "1") != "0"), stream())
    else:
        _call("stac_fbank_logmel_tc", ptr(wavs, torch.float32), b, n, wavs.stride(0), ptr(tc_tables.tab),
              ptr(tc_tables.tw, torch.float16), ptr(db), ptr(umax), stream())
    if raw:
        return RawFeatures(db, umax, float(top_db), bool(per_utterance), mean, std)
    _call("stac_fbank_topdb_norm", ptr(db), ptr(umax), int(per_utterance), float(top_db), ptr(mean), ptr(std),
          b, t, N_MELS, ptr(db), stream())
    return db


def fbank(wavs: torch.Tensor, tables: torch.Tensor, top_db: float = 80.0, per_utterance: bool = True,
          mean: Optional[torch.Tensor] = None, std: Optional[torch.Tensor] = None, raw: bool = False):
    """a2 (+a3 when mean/std are given): [B, L] fp32 PCM -> [B, T, 80] fp32 features."""
    if wavs.dim() != 2:
        raise _lib.StacB200Error("Fbank expects [batch, samples] waveforms")
    wavs = wavs.contiguous()
    b, n = wavs.shape
    t = 1 + n // HOP
    db = torch.empty(b, t, N_MELS, device=wavs.device, dtype=torch.float32)
    umax = torch.empty(b, device=wavs.device, dtype=torch.int32)      # zeroed by the call (stream-ordered memset)
    _call("stac_fbank_logmel", ptr(wavs, torch.float32), b, n, wavs.stride(0), ptr(tables), ptr(db),
                                  ptr(umax), stream())
    if raw:
        return RawFeatures(db, umax, float(top_db), bool(per_utterance), mean, std)
    _call("stac_fbank_topdb_norm", ptr(db), ptr(umax), int(per_utterance), float(top_db), ptr(mean), ptr(std),
                                      b, t, N_MELS, ptr(db), stream())
    return db


def input_norm(x: torch.Tensor, mean: torch.Tensor, std: torch.Tensor) -> torch.Tensor:
    x = x.contiguous()
    out = torch.empty_like(x)
    _call("stac_input_norm", ptr(x, torch.float32), ptr(mean, torch.float32), ptr(std, torch.float32),
                                x.numel() // x.shape[-1], x.shape[-1], ptr(out), stream())
    return out


# --------------------------------------------------------------------------
# a4
# --------------------------------------------------------------------------
@dataclass
class FrontendWeights:
    precision: str
    w0: torch.Tensor      # [256, 3(freq), 3(time)] fp32
    b0: torch.Tensor
    g0: torch.Tensor      # [40*256]
    be0: torch.Tensor
    w1: torch.Tensor      # fp32 [256][9][256] (fp32 mode) or bf16 [9][256][256] (bf16 mode); tap = kf*3+kt
    b1: torch.Tensor
    g1: torch.Tensor      # [20*256]
    be1: torch.Tensor


def pack_frontend(conv0_w, conv0_b, ln0_w, ln0_b, conv1_w, conv1_b, ln1_w, ln1_b, precision: str) -> FrontendWeights:
    f = lambda t: t.detach().float().contiguous()
    w1 = conv1_w.detach().float()                       # [out, in, kf, kt]
    if precision == "fp32":
        w1p = w1.permute(0, 2, 3, 1).reshape(CNN_CH, 9, CNN_CH).contiguous()
    else:
        w1p = w1.permute(2, 3, 0, 1).reshape(9, CNN_CH, CNN_CH).to(torch.bfloat16).contiguous()
    return FrontendWeights(precision, f(conv0_w).reshape(CNN_CH, 3, 3), f(conv0_b), f(ln0_w).flatten(),
                           f(ln0_b).flatten(), w1p, f(conv1_b), f(ln1_w).flatten(), f(ln1_b).flatten())


def conv_frontend(feats, w: FrontendWeights, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """a4: [B, T, 80] fp32 -> [B, T2, 5120] (fp32, or bf16 when out_dtype=torch.bfloat16).  bf16 mode also takes the
    RawFeatures of fbank(raw=True): the top-dB clamp and the normalisation then run inside block 0's loader."""
    raw = feats if isinstance(feats, RawFeatures) else None
    if raw is not None:
        if w.precision != "bf16":
            raise _lib.StacB200Error("raw feature hand-off exists in bf16 mode only")
        feats = raw.db
    feats = feats.contiguous()
    b, t, nm = feats.shape
    if nm != N_MELS:
        raise _lib.StacB200Error("ConvolutionFrontEnd kernel is specialised for 80 mel bins")
    t1 = (t - 1) // 2 + 1
    t2 = (t1 - 1) // 2 + 1
    dev = feats.device
    out_dtype = out_dtype or torch.float32
    if w.precision == "fp32":
        pre = torch.empty(b * t2, F2 * CNN_CH, device=dev, dtype=torch.float32)
        x0 = torch.empty(b, t1, F1, CNN_CH, device=dev, dtype=torch.float32)
        _call("stac_conv0_ln_lrelu", ptr(feats, torch.float32), ptr(w.w0), ptr(w.b0), ptr(w.g0), ptr(w.be0),
              b, t, ptr(x0), DT_F32, stream())
        _call("stac_conv1_f32", ptr(x0), ptr(w.w1, torch.float32), ptr(w.b1), b, t1, ptr(pre), stream())
        out = torch.empty(b, t2, F2 * CNN_CH, device=dev, dtype=out_dtype)
        _call("stac_group_ln_lrelu", ptr(pre), b * t2, F2 * CNN_CH, ptr(w.g1), ptr(w.be1), 1e-5, 0.01, ptr(out),
              DT_BF16 if out_dtype == torch.bfloat16 else DT_F32, stream())
        return out
    # (conv0 -> conv1 in rounds of a few utterances through one reused, L2-sized buffer was measured: slower at every
    # round size - 1.52 ms at 32 utterances per round to 3.9 ms at 1, against 1.56 ms for the whole batch,
    # profiles/r3/r3c_conv_chunks.log - both kernels lose more to short launches than the L2 hits give back)
    n_pad = lib().stac_conv0_padded_elems(b, t1)
    x0 = torch.empty(n_pad, device=dev, dtype=torch.bfloat16)
    if raw is not None:
        _call("stac_conv0_topdb_norm_bf16", ptr(feats, torch.float32), ptr(raw.utt_max), int(raw.per_utterance),
              raw.top_db, ptr(raw.mean), ptr(raw.std), ptr(w.w0), ptr(w.b0), ptr(w.g0), ptr(w.be0), b, t, ptr(x0),
              stream())
    else:
        _call("stac_conv0_ln_lrelu", ptr(feats, torch.float32), ptr(w.w0), ptr(w.b0), ptr(w.g0), ptr(w.be0),
              b, t, ptr(x0), DT_BF16, stream())
    out = torch.empty(b, t2, F2 * CNN_CH, device=dev, dtype=torch.bfloat16)
    _call("stac_conv1_bf16", ptr(x0), ptr(w.w1, torch.bfloat16), ptr(w.b1), ptr(w.g1), ptr(w.be1), b, t1, ptr(out),
          stream())
    if out_dtype == torch.bfloat16:
        return out
    out32 = torch.empty(out.shape, device=dev, dtype=torch.float32)
    _call("stac_cast_f32", ptr(out, torch.bfloat16), out.numel(), ptr(out32), stream())
    return out32


# --------------------------------------------------------------------------
# a5-a7
# --------------------------------------------------------------------------
@dataclass
class LayerWeights:
    ln1_g: torch.Tensor
    ln1_b: torch.Tensor
    w_qkv: torch.Tensor   # [3d, d]; q rows (and bias) pre-scaled by 1/sqrt(64)
    b_qkv: torch.Tensor
    w_o: torch.Tensor
    b_o: torch.Tensor
    ln2_g: torch.Tensor
    ln2_b: torch.Tensor
    w_1: torch.Tensor
    b_1: torch.Tensor
    w_2: torch.Tensor
    b_2: torch.Tensor


@dataclass
class EncoderWeights:
    precision: str
    d_model: int
    nhead: int
    w_src: torch.Tensor   # [d, 5120]
    b_src: torch.Tensor
    pe: torch.Tensor      # [max_len, d] fp32
    layers: List[LayerWeights] = field(default_factory=list)
    lnf_g: torch.Tensor = None
    lnf_b: torch.Tensor = None


def _wcast(t: torch.Tensor, precision: str) -> torch.Tensor:
    t = t.detach().float()
    return (t.to(torch.bfloat16) if precision == "bf16" else t).contiguous()


def pack_encoder(src_w, src_b, pe, layers, lnf_w, lnf_b, nhead: int, precision: str) -> EncoderWeights:
    """layers: iterable of dicts with in_proj_weight/in_proj_bias/out_proj_weight/out_proj_bias,
    ffn1_w/ffn1_b/ffn2_w/ffn2_b, norm1_w/norm1_b/norm2_w/norm2_b (SpeechBrain state_dict tensors)."""
    f = lambda t: t.detach().float().contiguous()
    d = src_w.shape[0]
    if d % nhead != 0 or d // nhead != 64:
        raise _lib.StacB200Error("attention kernels are specialised for head_dim 64 (all STAC-ST sizes)")
    ew = EncoderWeights(precision, d, nhead, _wcast(src_w, precision), f(src_b), f(pe).reshape(-1, d))
    scale = 1.0 / math.sqrt(64.0)   # exact power of two: folding it into W_q/b_q is lossless
    for L in layers:
        wq = L["in_proj_weight"].detach().float().clone()
        bq = L["in_proj_bias"].detach().float().clone()
        wq[:d] *= scale
        bq[:d] *= scale
        ew.layers.append(LayerWeights(
            f(L["norm1_w"]), f(L["norm1_b"]), _wcast(wq, precision), bq.contiguous(),
            _wcast(L["out_proj_weight"], precision), f(L["out_proj_bias"]),
            f(L["norm2_w"]), f(L["norm2_b"]), _wcast(L["ffn1_w"], precision), f(L["ffn1_b"]),
            _wcast(L["ffn2_w"], precision), f(L["ffn2_b"])))
    ew.lnf_g, ew.lnf_b = f(lnf_w), f(lnf_b)
    return ew


def kv_lengths(wav_len: Optional[torch.Tensor], batch: int, t2: int, device, train_mask: bool) -> torch.Tensor:
    """Valid key count per utterance, from the reference's own fp32 expressions:
    encode(): keep j <= floor(wav_len*T2)   (TransformerMultiTask.py:289-294)
    forward(): keep j <  round(wav_len*T2)  (TransformerMultiTask.py:225-226)."""
    if torch.device(device).type == "cuda":
        out = torch.empty(batch, device=device, dtype=torch.int32)
        wl = None if wav_len is None else wav_len.to(device=device, dtype=torch.float32).contiguous()
        _call("stac_kv_lengths", ptr(wl), batch, t2, int(train_mask), ptr(out), stream())
        return out
    # host-side twin of the kernel (index math on CPU tensors; used by the CPU tests of the mask rules)
    if wav_len is None:
        return torch.full((batch,), t2, device=device, dtype=torch.int32)
    wl = wav_len.to(device=device, dtype=torch.float32)
    if train_mask:
        n = torch.round(wl * t2)
    else:
        n = torch.floor(wl * t2) + 1
    return n.clamp(1, t2).to(torch.int32).contiguous()


def _gemm(a, w, bias, c, precision, resid=None, resid_period=0, act=ACT_NONE, vt=None, vt_cols=0, seq_len=0,
          t_pad=0, tag=""):
    global _LABEL
    prev, _LABEL = _LABEL, tag or _LABEL
    try:
        return _gemm_impl(a, w, bias, c, precision, resid, resid_period, act, vt, vt_cols, seq_len, t_pad)
    finally:
        _LABEL = prev


def _gemm_impl(a, w, bias, c, precision, resid, resid_period, act, vt, vt_cols, seq_len, t_pad):
    m, k = a.shape
    n = w.shape[0]
    if precision == "fp32":
        _call("stac_gemm_f32", ptr(a, torch.float32), ptr(w, torch.float32), ptr(bias), ptr(resid), resid_period,
                                  act, ptr(c, torch.float32), m, n, k, stream())
    else:
        _call("stac_gemm_bf16", ptr(a, torch.bfloat16), ptr(w, torch.bfloat16), ptr(bias), ptr(resid),
                                   resid_period, act, ptr(c), DT_BF16 if c.dtype == torch.bfloat16 else DT_F32,
                                   m, n, k, ptr(vt), vt_cols, seq_len, t_pad, stream())
    return c


def _layernorm(x, g, b, eps, out_f32=None, out_bf16=None):
    rows, dim = x.shape
    _call("stac_layernorm", ptr(x, torch.float32), rows, dim, ptr(g), ptr(b), eps, ptr(out_f32), ptr(out_bf16),
                               stream())


def encoder_stack(src: torch.Tensor, w: EncoderWeights, kv_len: torch.Tensor, want_bf16_copy: bool = False,
                  enc_out: Optional[torch.Tensor] = None):
    """a5 (src-linear + PE) and a7 (N pre-LN layers + final LN).
    src: [B, T2, 5120] fp32 (fp32 mode) or bf16 (bf16 mode).  Returns enc_out fp32 [B, T2, d]
    (and, if asked, its bf16 copy for the CTC GEMM)."""
    b, t2, k_in = src.shape
    d, h, prec = w.d_model, w.nhead, w.precision
    dev = src.device
    m = b * t2
    if t2 > w.pe.shape[0]:
        raise _lib.StacB200Error(f"sequence of {t2} frames exceeds the positional-encoding table ({w.pe.shape[0]})")
    act_dt = torch.float32 if prec == "fp32" else torch.bfloat16
    if src.dtype != act_dt:
        raise _lib.StacB200Error(f"{prec} encoder expects {act_dt} CNN features")
    x = torch.empty(m, d, device=dev, dtype=torch.float32)            # residual stream (fp32 in both modes)
    if L2_PERSIST and dev.type == "cuda":
        # keep the residual stream in L2 for the layer loop (STAC_L2_PERSIST=1; measured)
        _call("stac_l2_persist", ptr(x), x.numel() * 4, L2_PERSIST_RATIO, stream())
    _gemm(src.reshape(m, k_in), w.w_src, w.b_src, x, prec, resid=w.pe, resid_period=t2, tag="src_linear")
    hbuf = torch.empty(m, d, device=dev, dtype=act_dt)
    qkv = torch.empty(m, 3 * d, device=dev, dtype=act_dt)
    ctx = torch.empty(m, d, device=dev, dtype=act_dt)
    d_ffn = w.layers[0].w_1.shape[0] if w.layers else 4 * d
    # S model in bf16 mode: the whole feed-forward block is one kernel and the hidden activation never reaches HBM
    fused_ffn = prec == "bf16" and d == 256 and d_ffn % 128 == 0 and 128 <= d_ffn <= 4096 and FUSED_FFN
    ff = None if fused_ffn else torch.empty(m, d_ffn, device=dev, dtype=act_dt)
    t_pad = (t2 + 7) // 8 * 8
    # (enc_out: caller-provided result buffer, e.g. the slot rank 0 pulls from in the multi-GPU gather)
    enc = enc_out if enc_out is not None else torch.empty(b, t2, d, device=dev, dtype=torch.float32)
    if enc.shape != (b, t2, d) or enc.dtype != torch.float32:
        raise _lib.StacB200Error("enc_out buffer must be fp32 [batch, frames, d_model]")
    enc_bf16 = torch.empty(b, t2, d, device=dev, dtype=torch.bfloat16) if want_bf16_copy else None
    for L in w.layers:
        if prec == "fp32":
            _layernorm(x, L.ln1_g, L.ln1_b, 1e-6, out_f32=hbuf)
            _gemm(hbuf, L.w_qkv, L.b_qkv, qkv, prec, tag="qkv")
            _call("stac_mha_f32", ptr(qkv), ptr(kv_len, torch.int32), b, t2, d, h, ptr(ctx), stream())
        else:
            _layernorm(x, L.ln1_g, L.ln1_b, 1e-6, out_bf16=hbuf)
            _gemm(hbuf, L.w_qkv, L.b_qkv, qkv, prec, tag="qkv")
            # V is read straight from the packed projection (MN-major B operand): no transposed copy
            if MHA_V2:
                _call("stac_mha_bf16_v2", ptr(qkv), ptr(kv_len, torch.int32), b, t2, d, h, ptr(ctx), stream())
            else:
                _call("stac_mha_bf16", ptr(qkv), ptr(None), ptr(kv_len, torch.int32), b, t2, t_pad, d, h, ptr(ctx),
                      stream())
        if prec == "bf16" and d == 256 and FUSED_OUTPROJ_LN:
            # x += ctx W_o^T + b_o and hbuf = LayerNorm2(x) in one launch: the residual stream's producer holds whole rows
            _call("stac_outproj_ln_bf16", ptr(ctx, torch.bfloat16), ptr(L.w_o, torch.bfloat16), ptr(L.b_o), ptr(x),
                  ptr(L.ln2_g), ptr(L.ln2_b), 1e-6, ptr(hbuf, torch.bfloat16), m, stream())
        else:
            _gemm(ctx, L.w_o, L.b_o, x, prec, resid=x, tag="out_proj")
            if prec == "fp32":
                _layernorm(x, L.ln2_g, L.ln2_b, 1e-6, out_f32=hbuf)
            else:
                _layernorm(x, L.ln2_g, L.ln2_b, 1e-6, out_bf16=hbuf)
        if fused_ffn:
            _call("stac_ffn_fused_bf16", ptr(hbuf), ptr(L.w_1, torch.bfloat16), ptr(L.b_1), ptr(L.w_2, torch.bfloat16),
                  ptr(L.b_2), ptr(x), m, d, d_ffn, stream())
        else:
            _gemm(hbuf, L.w_1, L.b_1, ff, prec, act=ACT_GELU_ERF, tag="ffn1")
            _gemm(ff, L.w_2, L.b_2, x, prec, resid=x, tag="ffn2")
    _layernorm(x, w.lnf_g, w.lnf_b, 1e-6, out_f32=enc.view(m, d),
               out_bf16=None if enc_bf16 is None else enc_bf16.view(m, d))
    if L2_PERSIST and dev.type == "cuda":
        _call("stac_l2_persist", ptr(None), 0, 0.0, stream())
    return (enc, enc_bf16) if want_bf16_copy else enc


# --------------------------------------------------------------------------
# a8 / a9
# --------------------------------------------------------------------------
def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], precision: str,
           tag: str = "linear") -> torch.Tensor:
    """y = x W^T + b over the last dim; fp32 result."""
    shp = x.shape
    x2 = x.reshape(-1, shp[-1]).contiguous()
    if precision == "bf16" and x2.dtype != torch.bfloat16:
        xb = torch.empty_like(x2, dtype=torch.bfloat16)
        _call("stac_cast_bf16", ptr(x2, torch.float32), x2.numel(), ptr(xb), stream())
        x2 = xb
    out = torch.empty(x2.shape[0], weight.shape[0], device=x.device, dtype=torch.float32)
    _gemm(x2, weight, bias, out, precision, tag=tag)
    return out.view(*shp[:-1], weight.shape[0])


def ctc_head_bf16(enc_bf16: torch.Tensor, weight_bf16: torch.Tensor, bias: Optional[torch.Tensor],
                  out_dtype: torch.dtype = torch.float32, out: Optional[torch.Tensor] = None,
                  ids_out: Optional[torch.Tensor] = None):
    """a8 + a9 fused (bf16 mode): log_softmax(enc W^T + b) fp32 [.., V] and greedy ids int32 [..], logits never
    materialised (two GEMM passes, see include/stac_b200.h)."""
    shp = enc_bf16.shape
    x2 = enc_bf16.reshape(-1, shp[-1]).contiguous()
    m, d = x2.shape
    v = weight_bf16.shape[0]
    global _LABEL
    ws = torch.empty(lib().stac_ctc_head_workspace_floats(m, v), device=x2.device, dtype=torch.float32)
    if out is not None:
        if out.numel() != m * v or out.dtype != out_dtype or not out.is_contiguous():
            raise _lib.StacB200Error("posterior buffer must be contiguous [rows, vocab] of the requested dtype")
        out = out.view(m, v)
    else:
        out = torch.empty(m, v, device=x2.device, dtype=out_dtype)
    if ids_out is not None:
        if ids_out.numel() != m or ids_out.dtype != torch.int32 or not ids_out.is_contiguous():
            raise _lib.StacB200Error("greedy-id buffer must be contiguous int32 [rows]")
        ids = ids_out.view(m)
    else:
        ids = torch.empty(m, device=x2.device, dtype=torch.int32)
    prev, _LABEL = _LABEL, "ctc_head"
    try:
        _call("stac_ctc_head_bf16", ptr(x2, torch.bfloat16), ptr(weight_bf16, torch.bfloat16), ptr(bias), m, v, d,
              ptr(ws), ptr(out), DT_BF16 if out_dtype == torch.bfloat16 else DT_F32, ptr(ids), stream())
    finally:
        _LABEL = prev
    return out.view(*shp[:-1], v), ids.view(shp[:-1])


def log_softmax(logits: torch.Tensor, want_argmax: bool = False, inplace: bool = False):
    shp = logits.shape
    x = logits.reshape(-1, shp[-1]).contiguous()
    out = x if inplace else torch.empty_like(x)
    ids = torch.empty(x.shape[0], device=x.device, dtype=torch.int32) if want_argmax else None
    _call("stac_log_softmax", ptr(x, torch.float32), x.shape[0], x.shape[1], ptr(out), ptr(ids), stream())
    out = out.view(shp)
    return (out, ids.view(shp[:-1])) if want_argmax else out
