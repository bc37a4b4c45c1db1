"""Host ingest in front of the path: turn files -> padded batch -> device.

Mirrors what the reference does between its JSON manifest and ``compute_forward``:
  * the audio pipeline, /root/reference/stac-st/inference.py:250-261: every path of the ``wav`` field is loaded at
    16 kHz and the turns are concatenated (``concat_turns``);
  * SpeechBrain's ``PaddedBatch`` collation behind ``batch.sig`` (:91-92): right zero-padding to the longest utterance
    and ``wav_lens = length / longest`` in fp32 (``collate``);
  * ``batch.to(device)``.
The difference is the wire format: samples stay 16-bit until they are on the device (pinned int16 staging buffer,
asynchronous copy on a copy stream, ``stac_pcm_i16_to_f32`` on the device).  A 16-bit PCM sample decodes to
sample / 32768 exactly - the value librosa returns - so the fp32 batch is bit-identical and the host->device traffic is
halved.  No CPU fallback for the device half.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import StacB200Error, check, lib, ptr, stream


def concat_turns(turns: Sequence[np.ndarray]) -> np.ndarray:
    """inference.py:256-260: the turns of one utterance, concatenated in order.  int16 in, int16 out."""
    for t in turns:
        if t.dtype != np.int16 or t.ndim != 1:
            raise StacB200Error("turns must be 1-D int16 PCM arrays (16 kHz mono)")
    return np.concatenate(list(turns)) if len(turns) else np.zeros(0, np.int16)


def collate(utterances: Sequence[np.ndarray], out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """PaddedBatch for ``sig``: (int16 [B, Lmax] zero-padded right, fp32 wav_lens = len / Lmax).  `out`, if given, is
    a flat (pinned) int16 buffer of at least B * Lmax elements; the batch is laid out densely at its start."""
    if not len(utterances):
        raise StacB200Error("empty batch")
    lens = [int(u.shape[0]) for u in utterances]
    lmax = max(lens)
    if lmax == 0:
        raise StacB200Error("every utterance of the batch is empty")
    b = len(utterances)
    if out is None:
        batch = torch.zeros(b, lmax, dtype=torch.int16)
    else:
        if out.dtype != torch.int16 or out.dim() != 1 or out.numel() < b * lmax:
            raise StacB200Error("staging buffer too small or not a flat int16 tensor")
        batch = out[:b * lmax].view(b, lmax)
        batch.zero_()
    for i, u in enumerate(utterances):
        if u.dtype != np.int16 or u.ndim != 1:
            raise StacB200Error("utterances must be 1-D int16 PCM arrays (16 kHz mono)")
        batch[i, :lens[i]] = torch.from_numpy(np.ascontiguousarray(u))
    # speechbrain.utils.data_utils.batch_pad_right: the ratio is a Python (double) division, stored as fp32
    wav_lens = torch.tensor([n / lmax for n in lens], dtype=torch.float32)
    return batch, wav_lens


def pcm_to_float(pcm: torch.Tensor) -> torch.Tensor:
    """int16 CUDA tensor -> fp32 waveform (sample / 32768), the tensor ``compute_features`` takes."""
    if pcm.dtype != torch.int16:
        raise StacB200Error("expected int16 PCM")
    out = torch.empty(pcm.shape, device=pcm.device, dtype=torch.float32)
    check(lib().stac_pcm_i16_to_f32(ptr(pcm, torch.int16), pcm.numel(), ptr(out), stream()), "stac_pcm_i16_to_f32")
    return out


class PcmStager:
    """Double-buffered pinned int16 staging: ``put`` collates a batch into the free pinned buffer and starts its copy on
    a private copy stream; ``get`` makes the current stream wait for that copy and converts on the device.  While batch
    i is being computed, batch i+1 is collated and copied."""

    def __init__(self, max_elems: int, device="cuda"):
        self.device = torch.device(device)
        self.host = [torch.zeros(max_elems, dtype=torch.int16).pin_memory() for _ in range(2)]
        self.dev = [torch.empty(max_elems, dtype=torch.int16, device=self.device) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.shape: List = [None, None]
        self.wav_lens: List = [None, None]
        self.n_put = 0
        self.n_get = 0

    def put(self, utterances: Sequence[np.ndarray]) -> None:
        if self.n_put - self.n_get >= 2:
            raise StacB200Error("both staging buffers are in flight: call get() first")
        slot = self.n_put & 1
        self.consumed[slot].synchronize()        # the previous batch of this slot has been converted on the device
        staged, wav_lens = collate(utterances, out=self.host[slot])
        n = staged.numel()
        self.copy_stream.wait_event(self.consumed[slot])
        with torch.cuda.stream(self.copy_stream):
            self.dev[slot][:n].copy_(self.host[slot][:n], non_blocking=True)
            self.ready[slot].record(self.copy_stream)
        self.shape[slot], self.wav_lens[slot] = tuple(staged.shape), wav_lens
        self.n_put += 1

    def get(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.n_get >= self.n_put:
            raise StacB200Error("nothing staged: call put() first")
        slot = self.n_get & 1
        b, lmax = self.shape[slot]
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.ready[slot])
        wavs = pcm_to_float(self.dev[slot][:b * lmax].view(b, lmax))
        self.consumed[slot].record(cur)
        self.n_get += 1
        return wavs, self.wav_lens[slot].to(self.device, non_blocking=True)
