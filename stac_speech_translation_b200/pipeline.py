"""The object graph of the reference yaml, and the fused end-to-end call.

``build_modules`` instantiates what transformer_multitask.yaml:173-210,253-254,299-302 builds,
with the drop-in classes of this package.  ``compute_forward`` is the reference's call sequence
(/root/reference/stac-st/inference.py:95-107) over those objects, stage by stage, with fp32
tensors at every boundary.  ``EncoderPipeline`` is the same arithmetic without the fp32 stage
boundaries (bf16 hand-offs, fused top-dB + normalisation, greedy ids) and is what the benchmark
and the multi-GPU sharding drive.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch
from torch import nn

from . import ops
from ._lib import StacB200Error
from .convolution import ConvolutionFrontEnd
from .features import Fbank, InputNormalization
from .linear import Linear, LogSoftmax
from .transformer import TransformerMultiTask

# (d_model, nhead, num_encoder_layers, d_ffn): /root/reference/run_default.sh:73-76,
# /root/reference/ablations/run_m_and_l_size.sh:72-99
MODEL_SIZES = {
    "S": dict(d_model=256, nhead=4, num_encoder_layers=12, d_ffn=1024),
    "M": dict(d_model=512, nhead=8, num_encoder_layers=16, d_ffn=2048),
    "L": dict(d_model=1024, nhead=16, num_encoder_layers=14, d_ffn=4096),
}


@dataclass
class HParams:
    """The yaml keys of the hot path (transformer_multitask.yaml:127-170)."""
    sample_rate: int = 16000
    n_fft: int = 400
    n_mels: int = 80
    d_model: int = 256
    nhead: int = 4
    num_encoder_layers: int = 12
    num_decoder_layers: int = 6
    d_ffn: int = 1024
    transformer_dropout: float = 0.1
    output_neurons: int = 5000
    blank_index: int = 0
    turn: int = 7
    xt: int = 8
    seed: int = 8886

    @classmethod
    def for_size(cls, size: str, **kw):
        return cls(**{**MODEL_SIZES[size], **kw})


def build_modules(hp: HParams, precision: str = "bf16", device="cuda") -> Dict[str, nn.Module]:
    torch.manual_seed(hp.seed)
    mods = {
        "compute_features": Fbank(sample_rate=hp.sample_rate, n_fft=hp.n_fft, n_mels=hp.n_mels),
        "normalize": InputNormalization(norm_type="global", update_until_epoch=4),
        "CNN": ConvolutionFrontEnd(input_shape=(8, 10, hp.n_mels), num_blocks=2, num_layers_per_block=1,
                                   out_channels=(256, 256), kernel_sizes=(3, 3), strides=(2, 2),
                                   residuals=(False, False), precision=precision),
        "Transformer": TransformerMultiTask(
            input_size=5120, tgt_vocab=hp.output_neurons, d_model=hp.d_model, nhead=hp.nhead,
            num_encoder_layers=hp.num_encoder_layers, num_decoder_layers=hp.num_decoder_layers,
            d_ffn=hp.d_ffn, dropout=hp.transformer_dropout, activation=nn.GELU, encoder_module="transformer",
            attention_type="regularMHA", normalize_before=True, causal=False, precision=precision),
        "ctc_lin": Linear(input_size=hp.d_model, n_neurons=hp.output_neurons, precision=precision),
        "log_softmax": LogSoftmax(dim=-1),
    }
    for m in mods.values():
        m.to(device)
        m.eval()
    return mods


@torch.no_grad()
def compute_forward(mods: Dict[str, nn.Module], wavs: torch.Tensor, wav_lens: torch.Tensor, train_mask=False):
    """The reference's six-call sequence over the drop-in objects; returns every stage boundary."""
    out = {}
    feats = mods["compute_features"](wavs)
    out["fbank"] = feats
    feats = mods["normalize"](feats, wav_lens)
    out["feats"] = feats
    src = mods["CNN"](feats)
    out["cnn"] = src
    tr = mods["Transformer"]
    enc_out = tr.forward_encoder(src, wav_lens) if train_mask else tr.encode(src, wav_lens)
    out["enc_out"] = enc_out
    logits = mods["ctc_lin"](enc_out)
    out["logits"] = logits
    out["p_ctc"] = mods["log_softmax"](logits)
    return out


class EncoderPipeline:
    """PCM -> (enc_out, p_ctc[, greedy ids]) in one call, using the modules' packed weights.

    The same kernels as the stage-by-stage drop-ins, minus the fp32 stage boundaries:
    top-dB clamp and global normalisation are one elementwise pass, the CNN hands bf16 to the
    encoder in bf16 mode, and the CTC head also emits greedy token ids (the only thing
    ``append_speaker_turns`` reads, /root/reference/stac-st/inference.py:54-56).
    """

    def __init__(self, mods: Dict[str, nn.Module], precision: Optional[str] = None, train_mask: bool = False,
                 posterior_dtype: torch.dtype = torch.float32):
        self.mods = mods
        self.posterior_dtype = posterior_dtype      # bf16 only for ranks that ship posteriors to rank 0
        self.precision = precision or mods["Transformer"].precision
        for k in ("CNN", "Transformer", "ctc_lin"):
            if mods[k].precision != self.precision:
                raise StacB200Error("all modules of a pipeline must share one precision")
        self.train_mask = train_mask

    @torch.no_grad()
    def __call__(self, wavs: torch.Tensor, wav_lens: Optional[torch.Tensor], want_posteriors: bool = True,
                 want_greedy: bool = True, stop_after: Optional[str] = None,
                 outputs: Optional[Dict[str, torch.Tensor]] = None):
        """outputs: optional caller-owned result buffers {"enc_out", "p_ctc", "greedy"} (bf16 mode) that the last kernels
        write directly, e.g. the peer-visible slots of distributed.PeerGather."""
        m = self.mods
        outputs = outputs or {}
        dev = wavs.device
        fb, norm = m["compute_features"], m["normalize"]
        mean, std = norm.device_stats(dev, ops.N_MELS)
        bf16 = self.precision == "bf16"
        if bf16:      # STFT as a tensor-core GEMM (fp16 operands, fp32 accumulation); fp32 mode keeps the exact FFT kernel
            # (ops.FUSED_CONV0_NORM: raw hand-off, the top-dB clamp + normalisation then happen in conv block 0's loader)
            feats = ops.fbank_tc(wavs, fb.tc_tables(dev), fb.top_db, fb.top_db_per_utterance, mean, std,
                                 raw=ops.FUSED_CONV0_NORM)
        else:
            feats = ops.fbank(wavs, fb.tables(dev), fb.top_db, fb.top_db_per_utterance, mean, std)
        src = ops.conv_frontend(feats, m["CNN"].packed(), torch.bfloat16 if bf16 else torch.float32)
        if stop_after == "cnn":
            return {"cnn": src}
        tr = m["Transformer"]
        b, t2, _ = src.shape
        kv_len = ops.kv_lengths(wav_lens, b, t2, dev, self.train_mask)
        res = {"kv_len": kv_len}
        if bf16:
            enc, enc_b = ops.encoder_stack(src, tr.packed(), kv_len, want_bf16_copy=True, enc_out=outputs.get("enc_out"))
        else:
            enc = ops.encoder_stack(src, tr.packed(), kv_len)
            enc_b = enc
        res["enc_out"] = enc
        if want_posteriors or want_greedy:
            ctc = m["ctc_lin"]
            bias = None if ctc.w.bias is None else ctc.w.bias.detach().float().contiguous()
            if bf16:
                p, ids = ops.ctc_head_bf16(enc_b, ctc.packed_weight(), bias, self.posterior_dtype,
                                           out=outputs.get("p_ctc"), ids_out=outputs.get("greedy"))
                p = p.view(b, t2, -1)
                ids = ids.view(b, t2)
            else:
                logits = ops.linear(enc_b, ctc.packed_weight(), bias, self.precision, tag="ctc_lin")
                p, ids = ops.log_softmax(logits, want_argmax=True, inplace=True)
            res["p_ctc"], res["greedy"] = p, ids
        return res


class GraphedPipeline:
    """One CUDA graph of the whole path for a fixed batch shape: PCM buffer -> (enc_out, p_ctc, greedy ids).

    Every kernel of the path takes plain device pointers and a stream, so the ~80 launches of a step capture into
    one graph; a replay costs the host a few microseconds instead of a few milliseconds of Python/ctypes launch
    work, which is what the end-to-end rate depends on when the host is slow or shared.  The input lives in
    ``self.wavs`` (copy new audio there, on any stream ordered before the replay); the results are the same
    static tensors on every replay."""

    def __init__(self, pipe: "EncoderPipeline", wavs: torch.Tensor, wav_lens: Optional[torch.Tensor], warmup: int = 2,
                 pool=None, **call_kwargs):
        self.wavs = wavs
        self.wav_lens = wav_lens
        side = torch.cuda.Stream(device=wavs.device)
        side.wait_stream(torch.cuda.current_stream(wavs.device))
        with torch.cuda.stream(side):                 # first calls set kernel attributes / pack weights: not capturable
            for _ in range(max(1, warmup)):
                pipe(self.wavs, self.wav_lens, **call_kwargs)
        torch.cuda.current_stream(wavs.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        # pool: graphs that are only ever replayed one after the other on one stream may share a memory pool
        # (torch.cuda.graph_pool_handle()); results that must outlive the next replay go to caller-owned `outputs`
        with torch.cuda.graph(self.graph, pool=pool):
            self.out = pipe(self.wavs, self.wav_lens, **call_kwargs)

    def __call__(self, wavs: Optional[torch.Tensor] = None):
        if wavs is not None and wavs.data_ptr() != self.wavs.data_ptr():
            self.wavs.copy_(wavs, non_blocking=True)
        self.graph.replay()
        return self.out


def ctc_greedy_collapse(ids: torch.Tensor, lengths: Sequence[int], blank: int = 0) -> List[List[int]]:
    """CTC-greedy token sequences (merge repeats, drop blanks) over the valid frames of each utterance."""
    ids = ids.cpu()
    out = []
    for i, n in enumerate(lengths):
        seq, prev = [], None
        for tok in ids[i, : int(n)].tolist():
            if tok != prev and tok != blank:
                seq.append(tok)
            prev = tok
        out.append(seq)
    return out
