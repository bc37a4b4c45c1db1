"""Drop-in for the training-stage augmentation hook of ``compute_forward``.

Reference: ``feats = self.hparams.augmentation(feats)`` (/root/reference/stac-st/train_multitask.py:63-66) with
``augmentation: !new:speechbrain.lobes.augment.SpecAugment`` (hparams/transformer_multitask.yaml:283-293).  Same
constructor keywords; the random parameters are drawn with torch's global CPU generator in SpeechBrain's order (warp
centre, warped centre, frequency-mask lengths and positions, time-mask lengths and positions - on a CUDA tensor
SpeechBrain draws the mask parameters from the CUDA generator; this class always uses the CPU one, so a seeded run is
reproducible whatever the device), and ONE kernel (``stac_spec_augment``) applies warp and masks.
"""
from __future__ import annotations

import torch
from torch import nn

from ._lib import StacB200Error, check, lib, ptr, stream


class SpecAugment(nn.Module):
    def __init__(self, time_warp=True, time_warp_window=5, time_warp_mode="bicubic", freq_mask=True,
                 freq_mask_width=(0, 20), n_freq_mask=2, time_mask=True, time_mask_width=(0, 100), n_time_mask=2,
                 replace_with_zero=True):
        super().__init__()
        if time_warp and time_warp_mode != "bicubic":
            raise StacB200Error("stac_b200 SpecAugment implements the bicubic time warp (the reference's setting)")
        if not replace_with_zero:
            raise StacB200Error("stac_b200 SpecAugment implements replace_with_zero=True (SpeechBrain's default)")
        self.apply_time_warp, self.time_warp_window, self.time_warp_mode = time_warp, time_warp_window, time_warp_mode
        self.freq_mask, self.time_mask = freq_mask, time_mask
        if isinstance(freq_mask_width, int):
            freq_mask_width = (0, freq_mask_width)
        if isinstance(time_mask_width, int):
            time_mask_width = (0, time_mask_width)
        self.freq_mask_width, self.time_mask_width = tuple(freq_mask_width), tuple(time_mask_width)
        self.n_freq_mask, self.n_time_mask = n_freq_mask, n_time_mask
        self.replace_with_zero = replace_with_zero

    def draw(self, batch: int, time: int, fea: int):
        """The random parameters of one call, drawn exactly as SpeechBrain's forward draws them."""
        c = w = -1
        window = self.time_warp_window
        if self.apply_time_warp and time - window > window:
            c = torch.randint(window, time - window, (1,))[0]
            w = int(torch.randint(c - window, c + window, (1,))[0] + 1)
            c = int(c)
        masks = []
        for on, d, n_mask, width in ((self.freq_mask, fea, self.n_freq_mask, self.freq_mask_width),
                                     (self.time_mask, time, self.n_time_mask, self.time_mask_width)):
            if not on:
                masks.append((None, None))
                continue
            mask_len = torch.randint(width[0], width[1], (batch, n_mask))
            mask_pos = torch.randint(0, max(1, d - int(mask_len.max())), (batch, n_mask))
            masks.append((mask_pos.to(torch.int32), mask_len.to(torch.int32)))
        return c, w, masks[0], masks[1]

    @torch.no_grad()
    def forward(self, x):
        if x.dim() != 3:
            raise StacB200Error("SpecAugment expects [batch, time, features]")
        x = x.contiguous()
        b, t, f = x.shape
        c, w, (fpos, flen), (tpos, tlen) = self.draw(b, t, f)
        if w == 0 or w == t:
            # (SpeechBrain would interpolate to an empty / from an empty tensor here and raise)
            raise StacB200Error("degenerate time warp")
        dev = x.device
        mv = lambda a: None if a is None else a.to(dev).contiguous()
        fpos, flen, tpos, tlen = mv(fpos), mv(flen), mv(tpos), mv(tlen)
        out = torch.empty_like(x)
        check(lib().stac_spec_augment(
            ptr(x, torch.float32), b, t, f, c, w, ptr(fpos), ptr(flen), 0 if fpos is None else fpos.shape[1],
            ptr(tpos), ptr(tlen), 0 if tpos is None else tpos.shape[1], 0.0, ptr(out), stream()), "stac_spec_augment")
        return out
