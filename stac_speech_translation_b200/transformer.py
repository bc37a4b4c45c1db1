"""Drop-in for ``modules.Transformer`` (the reference's ``TransformerMultiTask``), encoder side.

Mirrors /root/reference/stac-st/modules/TransformerMultiTask.py:
  * constructor arguments :90-110 (as passed by transformer_multitask.yaml:183-196),
  * ``encode(src, wav_len)`` :273-309 with its ``j > floor(wav_len*T)`` key-padding rule,
  * the encoder half of ``forward`` :144-183 with ``make_masks`` :211-232 (``round`` rule),
  * ``_init_params`` :311-314 (xavier_normal_ on every dim>1 parameter),
  * ``EncoderWrapper`` :317-349.
Parameters are held in torch modules under SpeechBrain's names so ``state_dict`` keys match a
reference checkpoint (``custom_src_module.layers.0.w.*``, ``encoder.layers.N.self_att.att.*``,
``encoder.layers.N.pos_ffn.ffn.{0,3}.*``, ``encoder.layers.N.norm{1,2}.norm.*``,
``encoder.norm.norm.*``, ``positional_encoding.pe``).  None of those modules' ``forward`` is ever
called: the arithmetic runs in libstac_b200 (ops.encoder_stack).

The decoder side (``decode`` :234-271, decoder half of ``forward`` :185-209) runs on the
device too when ``num_decoder_layers > 0`` (decoder.py: parameters under ``decoder.*`` / ``custom_tgt_module.*`` with
SpeechBrain's names, fp32 CUDA path, parity on a B200 in tests/test_gpu_decoder.py).  An external decoder can still be attached
with ``attach_decoder``; it then takes precedence.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from . import decoder as dec
from . import ops
from ._lib import StacB200Error
from .convolution import _Holder, _params_version


class PositionalEncoding(nn.Module):
    def __init__(self, input_size, max_len=2500):
        super().__init__()
        if input_size % 2 != 0:
            raise ValueError(f"Cannot use sin/cos positional encoding with odd channels (got channels={input_size})")
        self.max_len = max_len
        pe = torch.zeros(max_len, input_size)
        positions = torch.arange(0, max_len).unsqueeze(1).float()
        denominator = torch.exp(torch.arange(0, input_size, 2).float() * -(math.log(10000.0) / input_size))
        pe[:, 0::2] = torch.sin(positions * denominator)
        pe[:, 1::2] = torch.cos(positions * denominator)
        self.register_buffer("pe", pe.unsqueeze(0))

    def forward(self, x):
        return self.pe[:, : x.size(1)].clone().detach()


class _EncoderLayerParams(nn.Module):
    def __init__(self, d_model, nhead, d_ffn, dropout, activation):
        super().__init__()
        self.self_att = _Holder(att=nn.MultiheadAttention(d_model, nhead, dropout=dropout, bias=True))
        self.pos_ffn = _Holder(ffn=nn.Sequential(nn.Linear(d_model, d_ffn), activation(), nn.Dropout(dropout),
                                                 nn.Linear(d_ffn, d_model)))
        self.norm1 = _Holder(norm=nn.LayerNorm(d_model, eps=1e-6))
        self.norm2 = _Holder(norm=nn.LayerNorm(d_model, eps=1e-6))


class _EncoderParams(nn.Module):
    def __init__(self, num_layers, d_model, nhead, d_ffn, dropout, activation):
        super().__init__()
        self.layers = nn.ModuleList([_EncoderLayerParams(d_model, nhead, d_ffn, dropout, activation)
                                     for _ in range(num_layers)])
        self.norm = _Holder(norm=nn.LayerNorm(d_model, eps=1e-6))


class _SrcModule(nn.Module):
    """speechbrain ModuleList(Linear, Dropout): children under ``layers``."""

    def __init__(self, input_size, d_model, dropout):
        super().__init__()
        self.layers = nn.ModuleList([_Holder(w=nn.Linear(input_size, d_model, bias=True)), nn.Dropout(dropout)])


class TransformerMultiTask(nn.Module):
    def __init__(self, tgt_vocab, input_size, d_model=512, nhead=8, num_encoder_layers=6, num_decoder_layers=6,
                 d_ffn=2048, dropout=0.1, activation=nn.ReLU, positional_encoding="fixed_abs_sine",
                 normalize_before=False, kernel_size: Optional[int] = 31, bias: Optional[bool] = True,
                 encoder_module: Optional[str] = "transformer", conformer_activation=None,
                 attention_type: Optional[str] = "regularMHA", max_length: Optional[int] = 2500,
                 causal: Optional[bool] = True, precision="bf16"):
        super().__init__()
        if (encoder_module != "transformer" or attention_type != "regularMHA" or not normalize_before
                or positional_encoding != "fixed_abs_sine" or activation is not nn.GELU):
            raise StacB200Error(
                "stac_b200 TransformerMultiTask implements the STAC-ST encoder configuration only: "
                "encoder_module='transformer', attention_type='regularMHA', normalize_before=True, "
                "activation=torch.nn.GELU, fixed_abs_sine positions")
        if d_model % nhead != 0 or d_model // nhead != 64:
            raise StacB200Error("attention kernels are specialised for head_dim 64 (S/M/L STAC-ST sizes)")
        if precision not in ops.PRECISIONS:
            raise StacB200Error(f"precision must be one of {ops.PRECISIONS}")
        self.precision = precision
        self.d_model, self.nhead = d_model, nhead
        self.tgt_vocab = tgt_vocab
        self.num_decoder_layers = num_decoder_layers
        self.causal = causal
        self.attention_type = attention_type
        self.positional_encoding_type = positional_encoding
        self.positional_encoding = PositionalEncoding(d_model, max_length)
        self.encoder = _EncoderParams(num_encoder_layers, d_model, nhead, d_ffn, dropout, activation)
        self.custom_src_module = _SrcModule(input_size, d_model, dropout)
        self._init_params()
        # Decoder and target embedding (reference: TransformerInterface.__init__ / :139).  They are created and
        # xavier-initialised under a forked random stream, after the encoder-side modules, so that the encoder-side
        # weights drawn for a given seed do not depend on the decoder's presence (the oracle does the same).
        self.decoder = None
        self.custom_tgt_module = None
        self._external_decoder = False
        if num_decoder_layers > 0:
            with torch.random.fork_rng(devices=[]):
                torch.manual_seed(torch.initial_seed() + 1)
                self.decoder = dec.DecoderParams(num_decoder_layers, d_model, nhead, d_ffn, dropout, activation)
                self.custom_tgt_module = dec.TgtModule(d_model, tgt_vocab)
                for m in (self.decoder, self.custom_tgt_module):
                    for p in m.parameters():
                        if p.dim() > 1:
                            torch.nn.init.xavier_normal_(p)
        self._packed = None
        self._packed_key = None
        self._packed_dec = None
        self._packed_dec_key = None

    def _init_params(self):
        for p in self.parameters():
            if p.dim() > 1:
                torch.nn.init.xavier_normal_(p)

    # ---- weight packing ----
    def packed(self) -> ops.EncoderWeights:
        key = (self.precision, _params_version(self.encoder), _params_version(self.custom_src_module))
        if self._packed is None or self._packed_key != key:
            layers = []
            for L in self.encoder.layers:
                att, ffn = L.self_att.att, L.pos_ffn.ffn
                layers.append(dict(in_proj_weight=att.in_proj_weight, in_proj_bias=att.in_proj_bias,
                                   out_proj_weight=att.out_proj.weight, out_proj_bias=att.out_proj.bias,
                                   ffn1_w=ffn[0].weight, ffn1_b=ffn[0].bias, ffn2_w=ffn[3].weight, ffn2_b=ffn[3].bias,
                                   norm1_w=L.norm1.norm.weight, norm1_b=L.norm1.norm.bias,
                                   norm2_w=L.norm2.norm.weight, norm2_b=L.norm2.norm.bias))
            src = self.custom_src_module.layers[0].w
            self._packed = ops.pack_encoder(src.weight, src.bias, self.positional_encoding.pe[0], layers,
                                            self.encoder.norm.norm.weight, self.encoder.norm.norm.bias,
                                            self.nhead, self.precision)
            self._packed_key = key
        return self._packed

    # ---- encoder entry points ----
    def _run_encoder(self, src, wav_len, train_mask: bool):
        if self.training:
            raise StacB200Error("stac_b200 TransformerMultiTask encoder is inference-only: call .eval()")
        if src.dim() == 4:
            bz, t, ch1, ch2 = src.shape
            src = src.reshape(bz, t, ch1 * ch2)
        w = self.packed()
        b, t2, _ = src.shape
        kv_len = ops.kv_lengths(wav_len, b, t2, src.device, train_mask)
        src = src.contiguous()
        # the reference hands fp32 CNN features to encode(); the fused pipeline hands bf16 in bf16 mode.  Anything else
        # would need a torch cast on the product path: refuse it.
        if src.dtype not in (torch.float32, torch.bfloat16) or (self.precision == "fp32" and src.dtype != torch.float32):
            raise StacB200Error(f"{self.precision} encoder expects fp32 CNN features (got {src.dtype})")
        if self.precision == "bf16" and src.dtype != torch.bfloat16:
            srcb = torch.empty_like(src, dtype=torch.bfloat16)
            ops._call("stac_cast_bf16", ops.ptr(src), src.numel(), ops.ptr(srcb), ops.stream())
            src = srcb
        return ops.encoder_stack(src, w, kv_len)

    @torch.no_grad()
    def encode(self, src, wav_len=None):
        """Encoder forward pass (reference :273-309).  src [B,T'',20,256] or [B,T'',5120]."""
        return self._run_encoder(src, wav_len, train_mask=False)

    @torch.no_grad()
    def forward_encoder(self, src, wav_len=None):
        """Encoder half of ``forward`` (reference :144-183) with the ``make_masks`` length rule."""
        return self._run_encoder(src, wav_len, train_mask=True)

    # ---- decoder side ----
    def attach_decoder(self, decoder: nn.Module, custom_tgt_module: nn.Module):
        """Use an external decoder (e.g. SpeechBrain's TransformerDecoder / NormalizedEmbedding) for ``forward`` and
        ``decode`` instead of the built-in device path."""
        self.decoder = decoder
        self.custom_tgt_module = custom_tgt_module
        self._external_decoder = True

    def _need_decoder(self):
        if self.decoder is None:
            raise StacB200Error(
                "this TransformerMultiTask was built with num_decoder_layers=0: forward()/decode() need a decoder "
                "(construct with num_decoder_layers > 0 or attach one with attach_decoder)")

    def packed_decoder(self) -> dec.DecoderWeights:
        key = dec.decoder_params_version(self.decoder, self.custom_tgt_module)
        if self._packed_dec is None or self._packed_dec_key != key:
            self._packed_dec = dec.pack_decoder(self.decoder, self.custom_tgt_module, self.positional_encoding.pe[0],
                                                self.nhead)
            self._packed_dec_key = key
        return self._packed_dec

    def decoder_cache(self, encoder_out, rows: int, max_len: int, enc_len=None, precision: str = "fp32",
                      graph: bool = False) -> dec.DecoderCache:
        """KV-cached incremental decoding over `encoder_out` [B, T2, d] for `rows` hypothesis rows (a multiple of B;
        row r belongs to utterance r // (rows // B)): ``cache.step(tokens)`` gives what ``decode(prefix)`` gives for
        its last position, ``cache.reorder(index)`` follows a beam re-ordering.  precision "bf16": the step's GEMMs on the
        tensor cores; graph=True: a step is one CUDA-graph replay (decoder.DecoderCache)."""
        self._need_decoder()
        if self._external_decoder:
            raise StacB200Error("decoder_cache drives the built-in decoder; an external one is attached")
        mem_len = None if enc_len is None else enc_len.to(device=encoder_out.device, dtype=torch.int32).contiguous()
        return dec.DecoderCache(self.packed_decoder(), encoder_out, rows, max_len, mem_len, precision=precision,
                                graph=graph)

    def forward(self, src, tgt, wav_len=None, pad_idx=0):
        """Reference :144-209: encoder with the ``make_masks`` length rule, decoder over the whole target with the
        look-ahead mask, ``tgt == pad_idx`` key padding and the encoder's key padding on the memory."""
        self._need_decoder()
        encoder_out = self.forward_encoder(src, wav_len)
        t2 = encoder_out.shape[1]
        if not self._external_decoder:
            if self.training:
                raise StacB200Error("stac_b200 TransformerMultiTask is inference-only: call .eval()")
            mem_len = None
            if wav_len is not None:
                mem_len = ops.kv_lengths(wav_len, encoder_out.shape[0], t2, encoder_out.device, True)
            decoder_out, _ = dec.decoder_stack(tgt, encoder_out, self.packed_decoder(), mem_len=mem_len,
                                               pad_idx=pad_idx)
            return encoder_out, decoder_out
        src_key_padding_mask = None
        if wav_len is not None:
            n = ops.kv_lengths(wav_len, encoder_out.shape[0], t2, encoder_out.device, True)
            src_key_padding_mask = torch.arange(t2, device=n.device)[None, :] >= n[:, None]
        tgt_key_padding_mask = tgt == pad_idx
        sz = tgt.shape[1]
        tgt_mask = torch.triu(torch.full((sz, sz), float("-inf"), device=tgt.device), diagonal=1)
        tgt_e = self.custom_tgt_module(tgt)
        tgt_e = tgt_e + self.positional_encoding(tgt_e)
        decoder_out, _, _ = self.decoder(tgt=tgt_e, memory=encoder_out, memory_mask=None, tgt_mask=tgt_mask,
                                         tgt_key_padding_mask=tgt_key_padding_mask,
                                         memory_key_padding_mask=src_key_padding_mask)
        return encoder_out, decoder_out

    @torch.no_grad()
    def decode(self, tgt, encoder_out, enc_len=None):
        """Reference :234-271: one decoding step = the decoder over the whole prefix; returns the prediction
        [rows, length, d_model] and the last layer's head-averaged cross-attention weights [rows, length, frames]."""
        self._need_decoder()
        if not self._external_decoder:
            mem_len = None
            if enc_len is not None:       # (1 - length_to_mask(enc_len)).bool(): keys j >= enc_len are masked
                mem_len = enc_len.to(device=encoder_out.device, dtype=torch.int32).contiguous()
            return dec.decoder_stack(tgt, encoder_out, self.packed_decoder(), mem_len=mem_len, pad_idx=None)
        sz = tgt.shape[1]
        tgt_mask = torch.triu(torch.full((sz, sz), float("-inf"), device=tgt.device), diagonal=1)
        src_key_padding_mask = None
        if enc_len is not None:
            t2 = encoder_out.shape[1]
            src_key_padding_mask = torch.arange(t2, device=enc_len.device)[None, :] >= enc_len[:, None]
        tgt_e = self.custom_tgt_module(tgt)
        tgt_e = tgt_e + self.positional_encoding(tgt_e)
        prediction, self_attns, multihead_attns = self.decoder(
            tgt_e, encoder_out, tgt_mask=tgt_mask, memory_key_padding_mask=src_key_padding_mask)
        return prediction, multihead_attns[-1]


class EncoderWrapper(nn.Module):
    """reference :317-349 - ``forward`` is ``transformer.encode``."""

    def __init__(self, transformer, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.transformer = transformer

    def forward(self, x, wav_lens=None):
        return self.transformer.encode(x, wav_lens)
