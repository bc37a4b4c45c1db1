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

The autoregressive decoder (``decode`` / decoder half of ``forward``) is not on the accelerated
path.  An external decoder can be attached with ``attach_decoder`` so that
``forward``/``decode`` keep working for the beam searcher.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

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
        self.decoder = None
        self.custom_tgt_module = None
        self._init_params()
        self._packed = None
        self._packed_key = None

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
        if self.precision == "bf16" and src.dtype != torch.bfloat16:
            srcb = torch.empty_like(src, dtype=torch.bfloat16)
            src32 = src.float()
            ops._call("stac_cast_bf16", ops.ptr(src32), src.numel(), ops.ptr(srcb), ops.stream())
            src = srcb
        elif self.precision == "fp32":
            src = src.float()
        return ops.encoder_stack(src, w, kv_len)

    @torch.no_grad()
    def encode(self, src, wav_len=None):
        """Encoder forward pass (reference :273-309).  src [B,T'',20,256] or [B,T'',5120]."""
        return self._run_encoder(src, wav_len, train_mask=False)

    @torch.no_grad()
    def forward_encoder(self, src, wav_len=None):
        """Encoder half of ``forward`` (reference :144-183) with the ``make_masks`` length rule."""
        return self._run_encoder(src, wav_len, train_mask=True)

    # ---- decoder side: not accelerated, delegated if attached ----
    def attach_decoder(self, decoder: nn.Module, custom_tgt_module: nn.Module):
        """Attach SpeechBrain's TransformerDecoder / NormalizedEmbedding (or equivalents) so that
        ``forward`` and ``decode`` serve train_multitask.py and the beam searcher."""
        self.decoder = decoder
        self.custom_tgt_module = custom_tgt_module

    def _need_decoder(self):
        if self.decoder is None:
            raise StacB200Error(
                "the autoregressive decoder is outside the accelerated encoder path; attach the reference's "
                "TransformerDecoder with attach_decoder(decoder, custom_tgt_module) to use forward()/decode()")

    def forward(self, src, tgt, wav_len=None, pad_idx=0):
        self._need_decoder()
        encoder_out = self.forward_encoder(src, wav_len)
        t2 = encoder_out.shape[1]
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
        self._need_decoder()
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
