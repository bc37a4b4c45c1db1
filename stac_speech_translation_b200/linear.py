"""Drop-ins for ``modules.ctc_lin`` (SpeechBrain ``Linear``) and ``hparams.log_softmax``.

yaml: /root/reference/stac-st/hparams/transformer_multitask.yaml:204-206,253-254;
calls: /root/reference/stac-st/inference.py:104-107.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from ._lib import StacB200Error


class Linear(nn.Module):
    """``speechbrain.nnet.linear.Linear``: parameters under ``w`` (``w.weight`` [n, in], ``w.bias``)."""

    def __init__(self, n_neurons, input_shape=None, input_size=None, bias=True, combine_dims=False,
                 precision="bf16"):
        super().__init__()
        if input_size is None:
            if input_shape is None:
                raise ValueError("Expected one of input_shape or input_size")
            input_size = input_shape[-1]
            if len(input_shape) == 4 and combine_dims:
                input_size = input_shape[2] * input_shape[3]
        if precision not in ops.PRECISIONS:
            raise StacB200Error(f"precision must be one of {ops.PRECISIONS}")
        self.combine_dims = combine_dims
        self.precision = precision
        self.w = nn.Linear(input_size, n_neurons, bias=bias)
        self._packed = None
        self._key = None

    def packed_weight(self):
        key = (self.precision, self.w.weight.data_ptr(), self.w.weight._version)
        if self._packed is None or self._key != key:
            w = self.w.weight.detach().float()
            self._packed = (w.to(torch.bfloat16) if self.precision == "bf16" else w).contiguous()
            self._key = key
        return self._packed

    @torch.no_grad()
    def forward(self, x):
        if x.ndim == 4 and self.combine_dims:
            x = x.reshape(x.shape[0], x.shape[1], x.shape[2] * x.shape[3])
        bias = None if self.w.bias is None else self.w.bias.detach().float().contiguous()
        return ops.linear(x, self.packed_weight(), bias, self.precision)


class LogSoftmax(nn.Module):
    """``torch.nn.LogSoftmax(dim=-1)`` on the GPU kernel; ``greedy`` also returns argmax ids
    (what ``append_speaker_turns`` computes, /root/reference/stac-st/inference.py:54-56)."""

    def __init__(self, dim=-1):
        super().__init__()
        if dim != -1:
            raise StacB200Error("stac_b200 LogSoftmax works over the last dimension")
        self.dim = dim

    @torch.no_grad()
    def forward(self, x):
        return ops.log_softmax(x.float())

    @torch.no_grad()
    def greedy(self, x):
        return ops.log_softmax(x.float(), want_argmax=True)
