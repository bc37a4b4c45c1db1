"""Speaker-turn / cross-talk detection from the CTC posteriors.

Drop-in for ``append_speaker_turns`` (/root/reference/stac-st/inference.py:54-84).  The reference takes the
arg-max of ``p_ctc`` on the device, copies two dense [B, T2] masks to the host and walks every frame of every utterance
in Python.  Here the arg-max ids (from the fused CTC head, or ``stac_argmax_rows`` when the caller only holds the
posteriors) are compacted on the device (``stac_ctc_spikes``) and the host reads back two counts plus the spikes; the
RTTM lines come out in the reference's order (utterance-major, frames ascending) with the reference's formatting.
No CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from ._lib import StacB200Error, check, lib, ptr, stream

# frames per second at the encoder output (inference.py:46-48)
DOWNSAMPLING = 25


def greedy_ids(model_ctc_outputs: torch.Tensor) -> torch.Tensor:
    """``model_ctc_outputs.argmax(-1)`` (inference.py:58) for fp32 posteriors [B, T2, V]; int32 ids pass through."""
    if model_ctc_outputs.dtype == torch.int32 and model_ctc_outputs.dim() == 2:
        return model_ctc_outputs.contiguous()
    if model_ctc_outputs.dim() != 3 or model_ctc_outputs.dtype != torch.float32:
        raise StacB200Error("expected fp32 posteriors [B, T2, V] or int32 greedy ids [B, T2]")
    b, t2, v = model_ctc_outputs.shape
    x = model_ctc_outputs.contiguous()
    ids = torch.empty(b, t2, device=x.device, dtype=torch.int32)
    check(lib().stac_argmax_rows(ptr(x, torch.float32), b * t2, v, ptr(ids), stream()), "stac_argmax_rows")
    return ids


def ctc_spikes(ids: torch.Tensor, turn: int, xt: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Flat positions b * T2 + j (ascending) of the frames whose greedy id is `turn` / `xt`, as two CPU int32 tensors.
    One device->host read of two counts, one of the spikes."""
    if ids.dim() != 2:
        raise StacB200Error("greedy ids must be [B, T2]")
    b, t2 = ids.shape
    dev = ids.device
    row_counts = torch.empty(b, 2, device=dev, dtype=torch.int32)
    spikes = torch.empty(2, b * t2, device=dev, dtype=torch.int32)
    n_out = torch.empty(2, device=dev, dtype=torch.int32)
    check(lib().stac_ctc_spikes(ptr(ids, torch.int32), b, t2, int(turn), int(xt), ptr(row_counts), ptr(spikes[0]),
                                ptr(spikes[1]), ptr(n_out), stream()), "stac_ctc_spikes")
    n_turn, n_xt = (int(n) for n in n_out.cpu())
    return spikes[0, :n_turn].cpu(), spikes[1, :n_xt].cpu()


def rttm_line(utt_id: str, frame: int) -> str:
    """One line exactly as inference.py:66-80 formats it."""
    abs_start = int(utt_id.split("-")[2]) / 100.0
    start = frame * (1 / DOWNSAMPLING)
    return f"SPEAKER {utt_id} 1 {abs_start + start:.3f} {(1/DOWNSAMPLING)} <NA> <NA> SPK1 <NA> <NA>"


def append_speaker_turns(batch_ids: Sequence[str], model_ctc_outputs: torch.Tensor, turn: int, xt: int,
                         turn_rttm: List[str], xt_rttm: List[str]) -> None:
    """inference.py:54-84 with its module-level state made explicit: `batch_ids` is ``batch.id``, `turn` / `xt` are
    ``hparams["turn"]`` / ``hparams["xt"]``, the two lists are the module-level ``turn_rttm`` / ``xt_rttm``."""
    ids = greedy_ids(model_ctc_outputs)
    if len(batch_ids) != ids.shape[0]:
        raise StacB200Error("one utterance id per row of the posteriors is required")
    t2 = ids.shape[1]
    s_turn, s_xt = ctc_spikes(ids, turn, xt)
    for flat in s_turn.tolist():
        turn_rttm.append(rttm_line(batch_ids[flat // t2], flat % t2))
    for flat in s_xt.tolist():
        xt_rttm.append(rttm_line(batch_ids[flat // t2], flat % t2))
