"""Drop-in for ``modules.CNN`` (SpeechBrain ``ConvolutionFrontEnd``).

Constructor arguments as in /root/reference/stac-st/hparams/transformer_multitask.yaml:173-180;
call signature as at /root/reference/stac-st/inference.py:99.  Parameters live in torch modules
with SpeechBrain's ``state_dict`` key layout (``convblock_i.convs.conv_0.conv.*``,
``convblock_i.convs.norm_0.norm.*``) so a reference ``model.ckpt`` loads unchanged; they are
repacked into kernel layouts lazily after every load.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from ._lib import StacB200Error


class _Holder(nn.Module):
    """Parameter container: gives a wrapped torch module SpeechBrain's attribute name."""

    def __init__(self, **mods):
        super().__init__()
        for k, v in mods.items():
            self.add_module(k, v)


def _params_version(module: nn.Module):
    return tuple((p.data_ptr(), p._version) for p in module.parameters())


class ConvolutionFrontEnd(nn.Module):
    def __init__(self, input_shape, num_blocks=3, num_layers_per_block=5, out_channels=[128, 256, 512],
                 kernel_sizes=[3, 3, 3], strides=[1, 2, 2], dilations=[1, 1, 1], residuals=[True, True, True],
                 conv_module=None, activation=nn.LeakyReLU, norm=None, dropout=0.1, conv_bias=True,
                 padding="same", conv_init=None, precision="bf16"):
        super().__init__()
        ok = (num_blocks == 2 and num_layers_per_block == 1 and tuple(out_channels) == (256, 256)
              and tuple(kernel_sizes) == (3, 3) and tuple(strides) == (2, 2) and not any(residuals)
              and tuple(dilations)[:2] == (1, 1) and input_shape[-1] == ops.N_MELS and conv_bias
              and padding == "same" and activation is nn.LeakyReLU and conv_module is None and norm is None)
        if not ok:
            raise StacB200Error("stac_b200 ConvolutionFrontEnd implements the STAC-ST front-end only: 2 blocks x "
                                "1 layer, 256 channels, 3x3 stride 2, LayerNorm, LeakyReLU, 80 mel inputs")
        if precision not in ops.PRECISIONS:
            raise StacB200Error(f"precision must be one of {ops.PRECISIONS}")
        self.precision = precision
        freq, chans = ops.N_MELS, 1
        for i in range(2):
            freq = (freq - 1) // 2 + 1
            conv = nn.Conv2d(chans, 256, (3, 3), stride=(2, 2), padding=0, bias=True)
            norm_i = nn.LayerNorm((freq, 256), eps=1e-5, elementwise_affine=True)
            block = _Holder(convs=_Holder(conv_0=_Holder(conv=conv), norm_0=_Holder(norm=norm_i)),
                            drop=nn.Dropout(dropout))
            self.add_module(f"convblock_{i}", block)
            chans = 256
        self._packed = None
        self._packed_key = None

    def packed(self) -> ops.FrontendWeights:
        key = (self.precision, _params_version(self))
        if self._packed is None or self._packed_key != key:
            c0, n0 = self.convblock_0.convs.conv_0.conv, self.convblock_0.convs.norm_0.norm
            c1, n1 = self.convblock_1.convs.conv_0.conv, self.convblock_1.convs.norm_0.norm
            self._packed = ops.pack_frontend(c0.weight, c0.bias, n0.weight, n0.bias, c1.weight, c1.bias,
                                             n1.weight, n1.bias, self.precision)
            self._packed_key = key
        return self._packed

    @torch.no_grad()
    def forward(self, x):
        """[B, T, 80] -> [B, T'', 20, 256] fp32 (the reference's CNN output contract)."""
        if self.training:
            raise StacB200Error("stac_b200 ConvolutionFrontEnd is inference-only: call .eval()")
        out = ops.conv_frontend(x.float(), self.packed())
        return out.view(out.shape[0], out.shape[1], ops.F2, ops.CNN_CH)
