"""Decoder side of ``modules.Transformer``: parameters under SpeechBrain's names and the
orchestration of one decoder pass on the device.

Mirrors what /root/reference/stac-st/modules/TransformerMultiTask.py builds and runs:
  * ``custom_tgt_module = ModuleList(NormalizedEmbedding(d_model, tgt_vocab))`` :139 and the
    ``TransformerDecoder(num_decoder_layers, nhead, d_ffn, d_model, dropout, activation, normalize_before)`` that
    ``TransformerInterface.__init__`` creates (:64-84), with SpeechBrain's attribute names so that a reference
    checkpoint loads unchanged: ``decoder.layers.N.{self_attn,mutihead_attn}.att.*`` (sic),
    ``decoder.layers.N.pos_ffn.ffn.{0,3}.*``, ``decoder.layers.N.norm{1,2,3}.norm.*``, ``decoder.norm.norm.*``,
    ``custom_tgt_module.layers.0.emb.Embedding.weight``;
  * ``decode(tgt, encoder_out, enc_len)`` :234-271 and the decoder half of ``forward`` :185-209: embedding * sqrt(d)
    + positional encoding, N pre-LN layers (causal self-attention, cross-attention over the encoder output,
    feed-forward), final LayerNorm; the head-averaged cross-attention weights of the last layer are returned because
    the beam searcher receives them (``mutitask_decoder.py:126``).

Two drivers over the same entry points (fp32 on the CUDA cores in both precision modes: stac_gemm_f32, stac_layernorm,
stac_embed_scale_pe, stac_attention_f32):
  * ``decoder_stack``: the whole prefix per call, exactly what the reference's ``forward_step`` asks for
    (``decode`` / ``forward`` semantics, pinned by vectors the reference's own code produced);
  * ``DecoderCache``: KV-cached, one token per call (same results, tested equal), the shape a B200 wants.
None of the torch modules' ``forward`` is called; there is no CPU fallback.  Not yet run on a B200 (see
csrc/decoder_f32.cu); the host side is covered on the CPU through tests/abi_emulator.py.
"""
from __future__ import annotations

import math
from ctypes import c_void_p
from dataclasses import dataclass, field
from typing import List, Optional

import torch
from torch import nn

from . import ops
from ._lib import ACT_GELU_ERF, StacB200Error, ptr, stream
from .convolution import _Holder, _params_version

import os

# STAC_BEAM_ATTENTION=0: the cached step's cross-attention on the general kernel (one CTA per hypothesis row)
BEAM_ATTENTION = os.environ.get("STAC_BEAM_ATTENTION", "1") == "1"


class _DecoderLayerParams(nn.Module):
    def __init__(self, d_model, nhead, d_ffn, dropout, activation):
        super().__init__()
        self.self_attn = _Holder(att=nn.MultiheadAttention(d_model, nhead, dropout=dropout, bias=True))
        self.mutihead_attn = _Holder(att=nn.MultiheadAttention(d_model, nhead, dropout=dropout, bias=True))
        self.pos_ffn = _Holder(ffn=nn.Sequential(nn.Linear(d_model, d_ffn), activation(), nn.Dropout(dropout),
                                                 nn.Linear(d_ffn, d_model)))
        self.norm1 = _Holder(norm=nn.LayerNorm(d_model, eps=1e-6))
        self.norm2 = _Holder(norm=nn.LayerNorm(d_model, eps=1e-6))
        self.norm3 = _Holder(norm=nn.LayerNorm(d_model, eps=1e-6))


class DecoderParams(nn.Module):
    """``Transformer.decoder``: parameter tree of SpeechBrain's TransformerDecoder."""

    def __init__(self, num_layers, d_model, nhead, d_ffn, dropout, activation):
        super().__init__()
        self.layers = nn.ModuleList([_DecoderLayerParams(d_model, nhead, d_ffn, dropout, activation)
                                     for _ in range(num_layers)])
        self.norm = _Holder(norm=nn.LayerNorm(d_model, eps=1e-6))


class TgtModule(nn.Module):
    """``Transformer.custom_tgt_module``: ModuleList(NormalizedEmbedding) -> ``layers.0.emb.Embedding.weight``."""

    def __init__(self, d_model, vocab):
        super().__init__()
        self.layers = nn.ModuleList([_Holder(emb=_Holder(Embedding=nn.Embedding(vocab, d_model, padding_idx=0)))])

    @property
    def weight(self):
        return self.layers[0].emb.Embedding.weight


@dataclass
class DecoderLayerWeights:
    ln1_g: torch.Tensor
    ln1_b: torch.Tensor
    w_qkv: torch.Tensor      # self-attention in_proj [3d, d]; q rows (and bias) pre-scaled by 1/sqrt(64)
    b_qkv: torch.Tensor
    w_o: torch.Tensor
    b_o: torch.Tensor
    ln2_g: torch.Tensor
    ln2_b: torch.Tensor
    w_q2: torch.Tensor       # cross-attention query projection [d, d], pre-scaled
    b_q2: torch.Tensor
    w_kv2: torch.Tensor      # cross-attention key / value projection of the encoder output [2d, d]
    b_kv2: torch.Tensor
    w_o2: torch.Tensor
    b_o2: torch.Tensor
    ln3_g: torch.Tensor
    ln3_b: torch.Tensor
    w_1: torch.Tensor
    b_1: torch.Tensor
    w_2: torch.Tensor
    b_2: torch.Tensor


@dataclass
class DecoderWeights:
    d_model: int
    nhead: int
    vocab: int
    emb: torch.Tensor        # [vocab, d]
    pe: torch.Tensor         # [max_len, d]
    layers: List[DecoderLayerWeights] = field(default_factory=list)
    lnf_g: torch.Tensor = None
    lnf_b: torch.Tensor = None
    _bf16: Optional[list] = None         # per layer: the GEMM weights as bf16 (tensor-core mode of DecoderCache)

    def bf16_layers(self):
        """bf16 copies of every layer's GEMM weights (stac_cast_bf16), built on first use."""
        if self._bf16 is None:
            out = []
            for lw in self.layers:
                out.append({k: _cast_bf16(getattr(lw, k)) for k in ("w_qkv", "w_o", "w_q2", "w_kv2", "w_o2", "w_1", "w_2")})
            self._bf16 = out
        return self._bf16


def pack_decoder(decoder: DecoderParams, tgt_module: TgtModule, pe: torch.Tensor, nhead: int) -> DecoderWeights:
    f = lambda t: t.detach().float().contiguous()
    emb = f(tgt_module.weight)
    d = emb.shape[1]
    if d % nhead != 0 or d // nhead != 64:
        raise StacB200Error("attention kernels are specialised for head_dim 64 (all STAC-ST sizes)")
    dw = DecoderWeights(d, nhead, emb.shape[0], emb, f(pe).reshape(-1, d))
    scale = 1.0 / math.sqrt(64.0)   # exact power of two: folding it into W_q / b_q is lossless
    for L in decoder.layers:
        sa, ca, ffn = L.self_attn.att, L.mutihead_attn.att, L.pos_ffn.ffn
        wq = sa.in_proj_weight.detach().float().clone()
        bq = sa.in_proj_bias.detach().float().clone()
        wq[:d] *= scale
        bq[:d] *= scale
        wc = ca.in_proj_weight.detach().float()
        bc = ca.in_proj_bias.detach().float()
        dw.layers.append(DecoderLayerWeights(
            f(L.norm1.norm.weight), f(L.norm1.norm.bias), wq.contiguous(), bq.contiguous(),
            f(sa.out_proj.weight), f(sa.out_proj.bias),
            f(L.norm2.norm.weight), f(L.norm2.norm.bias),
            (wc[:d] * scale).contiguous(), (bc[:d] * scale).contiguous(), wc[d:].contiguous(), bc[d:].contiguous(),
            f(ca.out_proj.weight), f(ca.out_proj.bias),
            f(L.norm3.norm.weight), f(L.norm3.norm.bias),
            f(ffn[0].weight), f(ffn[0].bias), f(ffn[3].weight), f(ffn[3].bias)))
    dw.lnf_g, dw.lnf_b = f(decoder.norm.norm.weight), f(decoder.norm.norm.bias)
    return dw


def _cast_bf16(t: torch.Tensor) -> torch.Tensor:
    t = t.contiguous()
    out = torch.empty(t.shape, device=t.device, dtype=torch.bfloat16)
    ops._call("stac_cast_bf16", ptr(t, torch.float32), t.numel(), ptr(out), stream())
    return out


def _off(t: torch.Tensor, elems: int) -> c_void_p:
    """Device address `elems` elements into a contiguous CUDA tensor (column offset inside a packed projection)."""
    ptr(t)                       # CUDA / contiguity checks
    return c_void_p(t.data_ptr() + elems * t.element_size())


def decoder_stack(tgt: torch.Tensor, memory: torch.Tensor, w: DecoderWeights, mem_len: Optional[torch.Tensor] = None,
                  pad_idx: Optional[int] = None):
    """One pass of the decoder over the whole target prefix.

    tgt      int64 [R, L] token ids (R hypothesis rows)
    memory   fp32 [Bm, T2, d] encoder output; R must be a multiple of Bm (row r reads memory[r // (R // Bm)]; the
             reference's beam searcher passes the inflated tensor, Bm == R)
    mem_len  int32 [R] valid encoder frames per row (keys j >= mem_len[r] are masked) or None
    pad_idx  if not None, target keys equal to it are masked in the self-attention (``tgt_key_padding_mask`` of
             ``forward``; ``decode`` passes None)
    Returns (prediction fp32 [R, L, d], cross-attention weights of the last layer, head-averaged, fp32 [R, L, T2])."""
    if tgt.dim() != 2 or memory.dim() != 3:
        raise StacB200Error("decoder expects tgt [rows, length] and encoder_out [batch, frames, d_model]")
    r, L = tgt.shape
    bm, t2, d = memory.shape
    if d != w.d_model or r % bm != 0:
        raise StacB200Error("encoder_out does not match the decoder (d_model / rows not a multiple of its batch)")
    if L > w.pe.shape[0]:
        raise StacB200Error(f"prefix of {L} tokens exceeds the positional-encoding table ({w.pe.shape[0]})")
    if not w.layers:
        raise StacB200Error("the decoder has no layers")
    dev = memory.device
    h = w.nhead
    m = r * L
    tok = tgt.to(device=dev, dtype=torch.int64).contiguous()
    mem = memory.float().contiguous().view(bm * t2, d)
    f32 = dict(device=dev, dtype=torch.float32)
    x = torch.empty(m, d, **f32)
    ops._call("stac_embed_scale_pe", ptr(tok, torch.int64), ptr(w.emb), ptr(w.pe), m, L, d, w.vocab, math.sqrt(d), ptr(x),
              stream())
    hbuf = torch.empty(m, d, **f32)
    qkv = torch.empty(m, 3 * d, **f32)
    q2 = torch.empty(m, d, **f32)
    kv2 = torch.empty(bm * t2, 2 * d, **f32)
    ctx = torch.empty(m, d, **f32)
    ff = torch.empty(m, w.layers[0].w_1.shape[0], **f32)
    weights = torch.empty(r, L, t2, **f32)
    key_tok = ptr(tok, torch.int64) if pad_idx is not None else ptr(None)
    for n, lw in enumerate(w.layers):
        last = n == len(w.layers) - 1
        # causal self-attention over the prefix
        ops._layernorm(x, lw.ln1_g, lw.ln1_b, 1e-6, out_f32=hbuf)
        ops._gemm(hbuf, lw.w_qkv, lw.b_qkv, qkv, "fp32", tag="dec_qkv")
        ops._call("stac_attention_f32", ptr(qkv), 3 * d, _off(qkv, d), _off(qkv, 2 * d), L * 3 * d, 3 * d, r, L, L, h, 1, 1,
                  ptr(None), key_tok, 0 if pad_idx is None else int(pad_idx), ptr(ctx), d, ptr(None), stream())
        ops._gemm(ctx, lw.w_o, lw.b_o, x, "fp32", resid=x, tag="dec_out_proj")
        # cross-attention over the encoder output
        ops._layernorm(x, lw.ln2_g, lw.ln2_b, 1e-6, out_f32=hbuf)
        ops._gemm(hbuf, lw.w_q2, lw.b_q2, q2, "fp32", tag="dec_q")
        ops._gemm(mem, lw.w_kv2, lw.b_kv2, kv2, "fp32", tag="dec_mem_kv")
        ops._call("stac_attention_f32", ptr(q2), d, ptr(kv2), _off(kv2, d), t2 * 2 * d, 2 * d, r, L, t2, h, r // bm, 0,
                  ptr(mem_len, torch.int32) if mem_len is not None else ptr(None), ptr(None), 0, ptr(ctx), d,
                  ptr(weights) if last else ptr(None), stream())
        ops._gemm(ctx, lw.w_o2, lw.b_o2, x, "fp32", resid=x, tag="dec_out_proj2")
        # feed-forward
        ops._layernorm(x, lw.ln3_g, lw.ln3_b, 1e-6, out_f32=hbuf)
        ops._gemm(hbuf, lw.w_1, lw.b_1, ff, "fp32", act=ACT_GELU_ERF, tag="dec_ffn1")
        ops._gemm(ff, lw.w_2, lw.b_2, x, "fp32", resid=x, tag="dec_ffn2")
    out = torch.empty(r, L, d, **f32)
    ops._layernorm(x, w.lnf_g, w.lnf_b, 1e-6, out_f32=out.view(m, d))
    return out, weights


def _cross_attention_step(q, kv, d, t2, rows, group, heads, mem_len, ctx, weights, head_scratch=None):
    """Cross-attention of one decoding step: all rows of an utterance in one CTA when the beam fits (a key / value row is
    then read once per utterance, not once per hypothesis), else the general kernel."""
    gp = (group + 3) // 4 * 4
    smem = (16 * 64 + t2 * gp * (2 if weights is not None and head_scratch is None else 1) + 8 * 16 * 64) * 4
    ml = ptr(mem_len, torch.int32) if mem_len is not None else ptr(None)
    if BEAM_ATTENTION and group <= 16 and smem <= 220 * 1024:
        ops._call("stac_attention_beam_f32", ptr(q), d, ptr(kv), _off(kv, d), t2 * 2 * d, 2 * d, rows, group, t2, heads, ml,
                  ptr(ctx), d, ptr(weights), ptr(head_scratch), stream())
    else:
        ops._call("stac_attention_f32", ptr(q), d, ptr(kv), _off(kv, d), t2 * 2 * d, 2 * d, rows, 1, t2, heads, group, 0,
                  ml, ptr(None), 0, ptr(ctx), d, ptr(weights), stream())


class DecoderCache:
    """KV-cached incremental decoding: the same arithmetic as ``decoder_stack`` on the growing prefix, one token per
    call (what a beam searcher's ``forward_step`` needs, mutitask_decoder.py:119-128, without re-running the whole
    decoder over the whole prefix every step as the reference does).

    * the cross-attention keys / values of every layer are projected ONCE from the encoder output, per utterance (not
      per hypothesis row: row r reads memory[r // beam]);
    * the self-attention keys / values of the prefix live in a time-major cache [layer][max_len][rows][2 d]; the
      self-attention kernel of a step stores the new position's keys / values into slab t itself and finds the rows of
      a re-ordered beam through a [max_len, rows] row map, so a beam re-ordering moves no cache data;
    * the position counter lives on the device: no kernel argument of a step depends on t, and with ``graph=True`` a
      step is ONE CUDA-graph replay (captured at the second step) instead of ~90 launches issued from Python;
    * ``step(tokens)`` returns what ``decode(prefix)[0][:, -1]`` and ``decode(prefix)[1][:, -1]`` return.
    precision "fp32": the entry points of ``decoder_stack`` (CUDA cores).  precision "bf16": every projection and the
    feed-forward block run on the tensor cores (``stac_gemm_bf16``: bf16 operands, fp32 accumulation; the residual
    stream, the caches, LayerNorm and both attentions stay fp32), which is what a step is made of at beam-search
    sizes: 640 rows x seven small GEMMs per layer.  No CPU fallback."""

    def __init__(self, w: DecoderWeights, memory: torch.Tensor, rows: int, max_len: int,
                 mem_len: Optional[torch.Tensor] = None, precision: str = "fp32", graph: bool = False):
        bm, t2, d = memory.shape
        if d != w.d_model or rows % bm != 0:
            raise StacB200Error("encoder_out does not match the decoder (d_model / rows not a multiple of its batch)")
        if max_len > w.pe.shape[0]:
            raise StacB200Error(f"{max_len} steps exceed the positional-encoding table ({w.pe.shape[0]})")
        if precision not in ("fp32", "bf16"):
            raise StacB200Error("precision must be 'fp32' or 'bf16'")
        self.w, self.rows, self.max_len, self.t = w, rows, max_len, 0
        self.bm, self.t2, self.d = bm, t2, d
        self.mem_len = mem_len
        self.precision = precision
        dev = memory.device
        f32 = dict(device=dev, dtype=torch.float32)
        mem = memory.float().contiguous().view(bm * t2, d)
        self.wb = w.bf16_layers() if precision == "bf16" else None
        mem_a = _cast_bf16(mem) if precision == "bf16" else mem
        self.cross_kv = []
        for n, lw in enumerate(w.layers):
            kv = torch.empty(bm * t2, 2 * d, **f32)
            ops._gemm(mem_a, self.wb[n]["w_kv2"] if self.wb else lw.w_kv2, lw.b_kv2, kv, precision, tag="dec_mem_kv")
            self.cross_kv.append(kv)
        self.self_kv = torch.empty(len(w.layers), max_len, rows, 2 * d, **f32)
        self.x = torch.empty(rows, d, **f32)
        self.h = torch.empty(rows, d, **f32)
        self.q = torch.empty(rows, d, **f32)
        self.ctx = torch.empty(rows, d, **f32)
        self.qkv_new = torch.empty(rows, 3 * d, **f32)            # query | key | value of the position being decoded
        self.t_dev = torch.zeros(1, device=dev, dtype=torch.int32)   # the position counter the kernels read
        self.use_graph, self._graph = bool(graph), None
        self.tok_buf = torch.zeros(rows, device=dev, dtype=torch.int64)
        self.out = torch.empty(rows, d, **f32)
        self.weights = torch.empty(rows, t2, **f32)
        # lazy beam re-ordering: key / value j of hypothesis row r lives in cache row row_map[j, r]; reorder() permutes
        # this table (max_len x rows int32) instead of gathering the cached prefix of every layer
        self._iota = torch.arange(rows, device=dev, dtype=torch.int32)
        self.row_map = self._iota.repeat(max_len, 1).contiguous()
        # per-head probabilities of the last layer's cross-attention (its head average is what step() returns)
        self.head_scratch = torch.empty(w.nhead, rows, t2, **f32)
        self.ff = torch.empty(rows, w.layers[0].w_1.shape[0], **f32)
        if precision == "bf16":
            b16 = dict(device=dev, dtype=torch.bfloat16)
            self.h16 = torch.empty(rows, d, **b16)
            self.ctx16 = torch.empty(rows, d, **b16)
            self.ff16 = torch.empty(rows, w.layers[0].w_1.shape[0], **b16)

    def _self_attention(self, cache):
        """Self-attention of the position ``t_dev`` counts, from the fused q | k | v projection of that position: the
        kernel stores the new keys / values into slab t of the cache and attends keys 0 .. t; hypothesis rows are found
        through ``self.row_map``."""
        d, r, n = self.d, self.rows, self.qkv_new
        ops._call("stac_attention_step_f32", ptr(n), 3 * d, ptr(cache), _off(cache, d), 2 * d, r * 2 * d, r, self.max_len,
                  self.w.nhead, ptr(self.row_map, torch.int32), _off(n, d), _off(n, 2 * d), 3 * d,
                  ptr(self.t_dev, torch.int32), ptr(self.ctx), d, stream())

    def _step_bf16(self, weights):
        """One position on the tensor-core GEMMs (same sequence as the fp32 step)."""
        w, r, d, t2 = self.w, self.rows, self.d, self.t2
        h, x = w.nhead, self.x
        for n, lw in enumerate(w.layers):
            wb = self.wb[n]
            last = n == len(w.layers) - 1
            ops._layernorm(x, lw.ln1_g, lw.ln1_b, 1e-6, out_bf16=self.h16)
            ops._gemm(self.h16, wb["w_qkv"], lw.b_qkv, self.qkv_new, "bf16", tag="dec_qkv_self")
            self._self_attention(self.self_kv[n])
            ops._call("stac_cast_bf16", ptr(self.ctx), self.ctx.numel(), ptr(self.ctx16), stream())
            ops._gemm(self.ctx16, wb["w_o"], lw.b_o, x, "bf16", resid=x, tag="dec_out_proj")
            ops._layernorm(x, lw.ln2_g, lw.ln2_b, 1e-6, out_bf16=self.h16)
            ops._gemm(self.h16, wb["w_q2"], lw.b_q2, self.q, "bf16", tag="dec_q")
            _cross_attention_step(self.q, self.cross_kv[n], d, t2, r, r // self.bm, h, self.mem_len, self.ctx,
                                  weights if last else None, self.head_scratch)
            ops._call("stac_cast_bf16", ptr(self.ctx), self.ctx.numel(), ptr(self.ctx16), stream())
            ops._gemm(self.ctx16, wb["w_o2"], lw.b_o2, x, "bf16", resid=x, tag="dec_out_proj2")
            ops._layernorm(x, lw.ln3_g, lw.ln3_b, 1e-6, out_bf16=self.h16)
            ops._gemm(self.h16, wb["w_1"], lw.b_1, self.ff16, "bf16", act=ACT_GELU_ERF, tag="dec_ffn1")
            ops._gemm(self.ff16, wb["w_2"], lw.b_2, x, "bf16", resid=x, tag="dec_ffn2")

    def _step_body(self):
        """Every launch of one step; reads ``tok_buf`` and ``t_dev``, writes ``out`` / ``weights``, advances ``t_dev``."""
        w, r, d, t2 = self.w, self.rows, self.d, self.t2
        h, x, weights = w.nhead, self.x, self.weights
        ops._call("stac_embed_step", ptr(self.tok_buf, torch.int64), ptr(w.emb), ptr(w.pe), r, d, w.vocab, math.sqrt(d),
                  ptr(self.t_dev, torch.int32), ptr(x), stream())
        if self.precision == "bf16":
            self._step_bf16(weights)
        for n, lw in enumerate(w.layers if self.precision == "fp32" else []):
            last = n == len(w.layers) - 1
            ops._layernorm(x, lw.ln1_g, lw.ln1_b, 1e-6, out_f32=self.h)
            ops._gemm(self.h, lw.w_qkv, lw.b_qkv, self.qkv_new, "fp32", tag="dec_qkv_self")
            self._self_attention(self.self_kv[n])
            ops._gemm(self.ctx, lw.w_o, lw.b_o, x, "fp32", resid=x, tag="dec_out_proj")
            ops._layernorm(x, lw.ln2_g, lw.ln2_b, 1e-6, out_f32=self.h)
            ops._gemm(self.h, lw.w_q2, lw.b_q2, self.q, "fp32", tag="dec_q")
            _cross_attention_step(self.q, self.cross_kv[n], d, t2, r, r // self.bm, h, self.mem_len, self.ctx,
                                  weights if last else None, self.head_scratch)
            ops._gemm(self.ctx, lw.w_o2, lw.b_o2, x, "fp32", resid=x, tag="dec_out_proj2")
            ops._layernorm(x, lw.ln3_g, lw.ln3_b, 1e-6, out_f32=self.h)
            ops._gemm(self.h, lw.w_1, lw.b_1, self.ff, "fp32", act=ACT_GELU_ERF, tag="dec_ffn1")
            ops._gemm(self.ff, lw.w_2, lw.b_2, x, "fp32", resid=x, tag="dec_ffn2")
        ops._layernorm(x, w.lnf_g, w.lnf_b, 1e-6, out_f32=self.out)
        self.t_dev.add_(1)

    def step(self, tokens: torch.Tensor):
        """tokens int64 [rows]: the token at position t of every hypothesis.  Returns (prediction [rows, d],
        head-averaged cross-attention weights of the last layer [rows, T2]) for that position."""
        if self.t >= self.max_len:
            raise StacB200Error("decoder cache is full")
        if tokens.shape != (self.rows,):
            raise StacB200Error("one token per hypothesis row is required")
        self.tok_buf.copy_(tokens, non_blocking=True)
        if not self.use_graph or self.t == 0:
            self._step_body()                      # (the first step also sets kernel attributes: not capturable)
        else:
            if self._graph is None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):              # (capture records the launches; the replay below runs them)
                    self._step_body()
                self._graph = g
            self._graph.replay()
        self.t += 1
        return self.out.clone(), self.weights.clone()      # (the buffers are rewritten by the next step)

    def rewind(self, t: int):
        """Set the position counter (host and device) back to ``t``: the next step recomputes position t."""
        self.t = int(t)
        self.t_dev.fill_(int(t))

    def reorder(self, index: torch.Tensor):
        """Beam re-ordering (``permute_mem``, mutitask_decoder.py:109-112): hypothesis row i continues row index[i].
        No cache data moves: only the row map is permuted (the first version gathered the cached prefix of every layer,
        four passes over 2 x layers x t x rows x d_model floats per step - more than the step itself once the prefix
        passes ~150 tokens)."""
        idx = index.to(device=self.self_kv.device, dtype=torch.int64)
        if idx.shape != (self.rows,):
            raise StacB200Error("one source row per hypothesis row is required")
        # (the caches stay where they are: position j of row i is now read from cache row row_map[j, index[i]]; the
        # slab of a new position is always written at the hypothesis's own row, row_map[t] = 0 .. rows-1)
        self.row_map[:self.t] = self.row_map[:self.t].index_select(1, idx)


def decoder_params_version(decoder: nn.Module, tgt_module: nn.Module):
    return _params_version(decoder), _params_version(tgt_module)
