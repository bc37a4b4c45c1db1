"""Seeded synthetic inputs for parity tests and benchmarks.

No corpus or checkpoint can be fetched, so audio is a deterministic speech-like
harmonic stack (about 80 dB of log-mel dynamic range) and batches are built the
way the reference's ``DynamicBatchSampler`` set-up would build them
(/root/reference/stac-st/dataio_and_utils.py:203-231,
/root/reference/stac-st/hparams/transformer_multitask.yaml:104-115).
Everything here is CPU torch / numpy; the tensors are then fed to the CUDA path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np
import torch

SAMPLE_RATE = 16000
HOP = 160


def frames_of(n_samples: int) -> Tuple[int, int, int]:
    """(T, T', T'') for an utterance of ``n_samples``: STFT frames, after conv block 0, after block 1."""
    t = 1 + n_samples // HOP
    t1 = (t - 1) // 2 + 1
    t2 = (t1 - 1) // 2 + 1
    return t, t1, t2


def synth_utterance(n_samples: int, gen: torch.Generator, turns: int = 1) -> torch.Tensor:
    """One speech-like utterance in [-1, 1]: sum of 29 harmonics of f0~U[90,250] Hz with 1/h
    amplitude under a slow squared-sine envelope, plus a little white noise.  ``turns`` > 1
    concatenates that many segments with 0.2-1.0 s near-silent gaps (multi-turn items)."""
    if turns <= 1:
        return _voiced(n_samples, gen)
    out = torch.empty(n_samples, dtype=torch.float32)
    gaps = (0.2 + 0.8 * torch.rand(turns - 1, generator=gen)) * SAMPLE_RATE
    gaps = gaps.long().tolist()
    speech = max(n_samples - sum(gaps), turns)
    bounds = np.linspace(0, speech, turns + 1).astype(np.int64)
    pos = 0
    for i in range(turns):
        seg = int(bounds[i + 1] - bounds[i])
        take = min(seg, n_samples - pos)
        if take > 0:
            out[pos:pos + take] = _voiced(seg, gen)[:take]
            pos += take
        if i < turns - 1:
            g = min(gaps[i], n_samples - pos)
            if g > 0:
                out[pos:pos + g] = 0.001 * torch.randn(g, generator=gen)
                pos += g
    if pos < n_samples:
        out[pos:] = 0.001 * torch.randn(n_samples - pos, generator=gen)
    return out


def _voiced(n: int, gen: torch.Generator) -> torch.Tensor:
    t = torch.arange(n, dtype=torch.float64) / SAMPLE_RATE
    f0 = 90.0 + 160.0 * float(torch.rand((), generator=gen))
    env_rate = 0.5 + 2.5 * float(torch.rand((), generator=gen))
    sig = torch.zeros(n, dtype=torch.float64)
    for h in range(1, 30):
        if f0 * h >= SAMPLE_RATE / 2:
            break
        sig += (1.0 / h) * torch.sin(2 * math.pi * f0 * h * t + h)
    env = torch.sin(math.pi * env_rate * t) ** 2
    noise = torch.randn(n, generator=gen, dtype=torch.float32)
    return (0.05 * sig * env).float() + 0.001 * noise


def synth_batch(lengths_s: Sequence[float], seed: int = 1234, turns: int = 1):
    """(wavs [B, Lmax] fp32 right-zero-padded, wav_lens [B] fp32 = len/Lmax) - the ``batch.sig``
    pair the reference's PaddedBatch hands to compute_forward
    (/root/reference/stac-st/inference.py:91-92)."""
    gen = torch.Generator().manual_seed(seed)
    lens = [int(round(s * SAMPLE_RATE)) for s in lengths_s]
    lmax = max(lens)
    wavs = torch.zeros(len(lens), lmax, dtype=torch.float32)
    for i, n in enumerate(lens):
        wavs[i, :n] = synth_utterance(n, gen, turns=turns)
    wav_lens = torch.tensor([n / lmax for n in lens], dtype=torch.float32)
    return wavs, wav_lens


def fast_synth_batch(batch: int, seconds: float, seed: int = 1234, device="cpu"):
    """Cheap full-length batch for throughput runs: one template utterance per 8 rows, each row a
    different circular shift and gain (content does not change the work the path does)."""
    gen = torch.Generator().manual_seed(seed)
    n = int(round(seconds * SAMPLE_RATE))
    n_tmpl = max(1, min(batch, 8))
    tmpl = torch.stack([synth_utterance(n, gen, turns=3) for _ in range(n_tmpl)])
    rows = []
    for i in range(batch):
        shift = int(torch.randint(0, n, (), generator=gen))
        gain = 0.5 + float(torch.rand((), generator=gen))
        rows.append(torch.roll(tmpl[i % n_tmpl], shift) * gain)
    wavs = torch.stack(rows).to(device)
    return wavs, torch.ones(batch, dtype=torch.float32, device=device)


# --------------------------------------------------------------------------
# Length-bucketed batching and rank sharding
# --------------------------------------------------------------------------
@dataclass
class Bucketed:
    batches: List[List[int]]          # utterance indices per batch
    durations: np.ndarray             # seconds per utterance


def lognormal_durations(n: int, seed: int, median_s=8.0, sigma=0.7, lo=1.0, hi=30.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    d = np.exp(rng.normal(math.log(median_s), sigma, size=n))
    return np.clip(d, lo, hi)


def uniform_durations(n: int, seed: int, lo=1.0, hi=30.0) -> np.ndarray:
    return np.random.default_rng(seed).uniform(lo, hi, size=n)


def bucket_batches(durations: np.ndarray, max_batch_len: float = 200.0, num_buckets: int = 50,
                   max_batch_ex: int = 128) -> Bucketed:
    """Length-bucketed batches in the manner of SpeechBrain's DynamicBatchSampler as the
    reference configures it: bucket boundaries at quantiles of the duration distribution,
    batch size per bucket = max(1, int(max_batch_len / bucket_upper_bound)) capped at
    ``max_batch_ex``; utterances keep their order inside a bucket."""
    durations = np.asarray(durations, dtype=np.float64)
    qs = np.quantile(durations, np.linspace(0, 1, num_buckets + 1)[1:])
    bounds = np.unique(qs)
    bucket_of = np.searchsorted(bounds, durations, side="left")
    batches: List[List[int]] = []
    for b, ub in enumerate(bounds):
        idx = np.nonzero(bucket_of == b)[0].tolist()
        if not idx:
            continue
        bs = min(max_batch_ex, max(1, int(max_batch_len / ub)))
        for i in range(0, len(idx), bs):
            batches.append(idx[i:i + bs])
    return Bucketed(batches=batches, durations=durations)


def batch_cost(durations: np.ndarray, batch: Sequence[int], c_lin: float = 1.0,
               c_quad: float = 1.0 / 1500.0) -> float:
    """Relative cost of one padded batch: B * T'' * (c_lin + c_quad * T'') with T'' at 25 Hz of
    the batch maximum (linear GEMM/conv work plus the quadratic attention term)."""
    t2 = 25.0 * float(np.max(durations[list(batch)]))
    return len(batch) * t2 * (c_lin + c_quad * t2)


def shard_batches(bucketed: Bucketed, world_size: int) -> List[List[int]]:
    """Whole batches to ranks, longest-processing-time-first (never split a batch: reference
    outputs depend on batch membership through wav_lens = len/Lmax).
    Returns, per rank, the list of batch ids (indices into ``bucketed.batches``)."""
    costs = [batch_cost(bucketed.durations, b) for b in bucketed.batches]
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    return out
