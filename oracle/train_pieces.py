"""CPU restatement of the two training-stage pieces of ``compute_forward`` / ``compute_objectives`` that sit right on the
path's tensors: SpecAugment on the normalised features and the CTC loss on the posteriors.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY: SpeechBrain (~v0.5.14) is not installable here, so
``SpecAugment`` follows the published ``speechbrain.lobes.augment.SpecAugment`` from memory of that release (parity
unpinned for its control flow); everything it computes is ``torch.nn.functional.interpolate`` / ``masked_fill_``, which
ARE this image's torch.  ``ctc_loss`` is SpeechBrain's thin wrapper over ``torch.nn.functional.ctc_loss``: its arithmetic
is pinned by torch itself.

Reference call sites:
  /root/reference/stac-st/train_multitask.py:63-66   feats = self.hparams.augmentation(feats)
  /root/reference/stac-st/hparams/transformer_multitask.yaml:283-293  (time_warp bicubic window 5, 2 freq masks <= 30,
                                                                        2 time masks <= 40)
  /root/reference/stac-st/train_multitask.py:164-170 self.hparams.ctc_cost(p_ctc, tokens, wav_lens, tokens_lens)
  /root/reference/stac-st/hparams/transformer_multitask.yaml:256-258  (blank_index 0, reduction batchmean)
"""
import torch
import torch.nn.functional as F


class SpecAugment(torch.nn.Module):
    """``speechbrain.lobes.augment.SpecAugment``: time warp around a random centre, then frequency and time masks.
    Input [batch, time, features]; the random draws come from torch's global generator in this order: warp centre, warped
    centre, frequency-mask lengths, frequency-mask positions, time-mask lengths, time-mask positions."""

    def __init__(self, time_warp=True, time_warp_window=5, time_warp_mode="bicubic", freq_mask=True,
                 freq_mask_width=(0, 20), n_freq_mask=2, time_mask=True, time_mask_width=(0, 100), n_time_mask=2,
                 replace_with_zero=True):
        super().__init__()
        self.apply_time_warp, self.time_warp_window, self.time_warp_mode = time_warp, time_warp_window, time_warp_mode
        self.freq_mask, self.time_mask = freq_mask, time_mask
        if isinstance(freq_mask_width, int):
            freq_mask_width = (0, freq_mask_width)
        if isinstance(time_mask_width, int):
            time_mask_width = (0, time_mask_width)
        self.freq_mask_width, self.time_mask_width = freq_mask_width, time_mask_width
        self.n_freq_mask, self.n_time_mask = n_freq_mask, n_time_mask
        self.replace_with_zero = replace_with_zero

    def forward(self, x):
        if self.apply_time_warp:
            x = self.time_warp(x)
        if self.freq_mask:
            x = self.mask_along_axis(x, dim=2)
        if self.time_mask:
            x = self.mask_along_axis(x, dim=1)
        return x

    def time_warp(self, x):
        original_size = x.shape
        window = self.time_warp_window
        if x.dim() == 3:
            x = x.unsqueeze(1)
        time = x.shape[2]
        if time - window <= window:
            return x.view(*original_size)
        c = torch.randint(window, time - window, (1,))[0]
        w = torch.randint(c - window, c + window, (1,))[0] + 1
        left = F.interpolate(x[:, :, :c], (w, x.shape[3]), mode=self.time_warp_mode, align_corners=True)
        right = F.interpolate(x[:, :, c:], (time - w, x.shape[3]), mode=self.time_warp_mode, align_corners=True)
        x[:, :, :w] = left
        x[:, :, w:] = right
        return x.view(*original_size)

    def mask_along_axis(self, x, dim):
        original_size = x.shape
        if x.dim() == 4:
            x = x.view(-1, x.shape[2], x.shape[3])
        batch, time, fea = x.shape
        if dim == 1:
            D, n_mask, width_range = time, self.n_time_mask, self.time_mask_width
        else:
            D, n_mask, width_range = fea, self.n_freq_mask, self.freq_mask_width
        mask_len = torch.randint(width_range[0], width_range[1], (batch, n_mask), device=x.device).unsqueeze(2)
        mask_pos = torch.randint(0, max(1, D - mask_len.max()), (batch, n_mask), device=x.device).unsqueeze(2)
        arange = torch.arange(D, device=x.device).view(1, 1, -1)
        mask = (mask_pos <= arange) * (arange < (mask_pos + mask_len))
        mask = mask.any(dim=1)
        mask = mask.unsqueeze(2) if dim == 1 else mask.unsqueeze(1)
        val = 0.0 if self.replace_with_zero else x.mean()
        x = x.masked_fill_(mask, val)
        return x.view(*original_size)


def ctc_loss(log_probs, targets, input_lens, target_lens, blank_index, reduction="mean"):
    """``speechbrain.nnet.losses.ctc_loss``: relative lengths to frames / tokens, ``torch.nn.functional.ctc_loss`` with
    zero_infinity, SpeechBrain's extra reductions (batchmean: sum / batch; batch: per utterance / its target length)."""
    input_lens = (input_lens * log_probs.shape[1]).round().int()
    target_lens = (target_lens * targets.shape[1]).round().int()
    log_probs = log_probs.transpose(0, 1)
    if reduction == "batchmean":
        reduction_loss = "sum"
    elif reduction == "batch":
        reduction_loss = "none"
    else:
        reduction_loss = reduction
    loss = F.ctc_loss(log_probs, targets, input_lens, target_lens, blank_index, zero_infinity=True,
                      reduction=reduction_loss)
    if reduction == "batchmean":
        return loss / targets.shape[0]
    if reduction == "batch":
        n = loss.size(0)
        return loss.view(n, -1).sum(1) / target_lens.view(n, -1).sum(1)
    return loss
