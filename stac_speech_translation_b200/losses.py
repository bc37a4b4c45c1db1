"""Drop-in for the CTC cost of ``compute_objectives``: forward value on the device.

Reference: ``self.hparams.ctc_cost(p_ctc, tokens, wav_lens, tokens_lens)``
(/root/reference/stac-st/train_multitask.py:164-170) with ``ctc_cost: !name:speechbrain.nnet.losses.ctc_loss``,
``blank_index: 0``, ``reduction: batchmean`` (hparams/transformer_multitask.yaml:256-258, :72, :138).  SpeechBrain's
function turns the relative lengths into frames / tokens (fp32 product, round half to even) and calls
``torch.nn.functional.ctc_loss(..., zero_infinity=True)``; ``stac_ctc_loss`` runs the same alpha recursion with one CTA
per utterance and SpeechBrain's reductions.  No gradient: the value is what the validation stage reports; the backward
pass belongs to the training loop, which is outside this path.
"""
from __future__ import annotations

import torch

from ._lib import StacB200Error, check, lib, ptr, stream

_REDUCTIONS = {"none": 0, "sum": 1, "mean": 2, "batchmean": 3, "batch": 4}


@torch.no_grad()
def ctc_loss(log_probs, targets, input_lens, target_lens, blank_index, reduction="mean"):
    """log_probs [batch, time, vocab] (log-softmax outputs), targets [batch, max_tokens] integer, input_lens /
    target_lens relative lengths [batch] as SpeechBrain passes them.  Returns a 0-d tensor (sum / mean / batchmean) or
    [batch] (none / batch)."""
    if reduction not in _REDUCTIONS:
        raise StacB200Error(f"unknown reduction {reduction!r}")
    if log_probs.dim() != 3 or targets.dim() != 2 or targets.shape[0] != log_probs.shape[0]:
        raise StacB200Error("ctc_loss expects log_probs [batch, time, vocab] and targets [batch, tokens]")
    dev = log_probs.device
    b, t, v = log_probs.shape
    lp = log_probs.float().contiguous() if log_probs.dtype != torch.float32 else log_probs.contiguous()
    # SpeechBrain: (input_lens * T).round().int() - a handful of scalars per batch, computed where the lengths live
    il = (input_lens.float() * t).round().to(torch.int32).to(dev).contiguous()
    tl = (target_lens.float() * targets.shape[1]).round().to(torch.int32).to(dev).contiguous()
    tg = targets.to(device=dev, dtype=torch.int32).contiguous()
    nll = torch.empty(b, device=dev, dtype=torch.float32)
    mode = _REDUCTIONS[reduction]
    out = torch.empty(b if mode == 4 else 1, device=dev, dtype=torch.float32)
    check(lib().stac_ctc_loss(ptr(lp, torch.float32), ptr(tg), ptr(il), ptr(tl), b, t, v, tg.shape[1],
                                        int(blank_index), mode, ptr(nll), ptr(out), stream()), "stac_ctc_loss")
    if mode == 0:
        return nll
    return out if mode == 4 else out[0]
