"""Multi-GPU driving of the path: whole length-bucketed batches per rank, results gathered to rank 0.

The path has no exchange step (utterances are independent at inference), so
ranks never talk while computing.  The only communication is the variable-size gather of each
batch's outputs to rank 0 (point-to-point ``isend``/``irecv`` over the process group: NCCL on
NVLink for CUDA tensors, gloo on CPU in the tests).  The reference has no counterpart (its
inference is single-GPU, /root/reference/README.md:389).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import synth

_DTYPES = [torch.float32, torch.bfloat16, torch.int32, torch.int64, torch.float16]


def plan(durations, world_size: int, max_batch_len: float = 200.0, num_buckets: int = 50,
         max_batch_ex: int = 128):
    """Bucket utterances by length and assign whole batches to ranks (longest-processing-time-first).
    Returns (bucketed, per_rank_batch_ids)."""
    bucketed = synth.bucket_batches(durations, max_batch_len, num_buckets, max_batch_ex)
    return bucketed, synth.shard_batches(bucketed, world_size)


def _meta_of(t: torch.Tensor) -> List[int]:
    shape = list(t.shape)
    return [_DTYPES.index(t.dtype), len(shape)] + shape + [0] * (4 - len(shape))


def gather_batch(outputs: Optional[Dict[str, torch.Tensor]], keys: Sequence[str], owner: int,
                 group=None, device=None) -> Optional[Dict[str, torch.Tensor]]:
    """Bring one batch's tensors from rank ``owner`` to rank 0.  Every rank calls this for every
    batch in the same order; only ``owner`` passes ``outputs``.  Shapes travel first (one small
    broadcast from the owner), then the payload as point-to-point transfers."""
    rank = dist.get_rank(group)
    if owner == 0:
        return outputs if rank == 0 else None
    if rank != 0 and rank != owner:
        return None
    meta = torch.zeros(len(keys), 6, dtype=torch.int64, device=device)
    if rank == owner:
        meta = torch.tensor([_meta_of(outputs[k]) for k in keys], dtype=torch.int64, device=device)
        dist.send(meta, dst=0, group=group)
        reqs = [dist.isend(outputs[k].contiguous(), dst=0, group=group) for k in keys]
        for r in reqs:
            r.wait()
        return None
    dist.recv(meta, src=owner, group=group)
    got, reqs = {}, []
    for k, row in zip(keys, meta.tolist()):
        shape = row[2:2 + row[1]]
        got[k] = torch.empty(shape, dtype=_DTYPES[row[0]], device=device)
        reqs.append(dist.irecv(got[k], src=owner, group=group))
    for r in reqs:
        r.wait()
    return got


def run_sharded(batches: Sequence, per_rank: Sequence[Sequence[int]], compute: Callable, keys: Sequence[str],
                group=None, device=None) -> Optional[Dict[int, Dict[str, torch.Tensor]]]:
    """Each rank runs ``compute(batches[i])`` for its own batch ids; rank 0 ends up with
    {batch_id: outputs} for every batch (the same dict a single-GPU run would produce)."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    mine = {i: compute(batches[i]) for i in per_rank[rank]}
    collected: Dict[int, Dict[str, torch.Tensor]] = {}
    owner_of = {i: r for r in range(world) for i in per_rank[r]}
    for i in sorted(owner_of):
        got = gather_batch(mine.get(i), keys, owner_of[i], group=group, device=device)
        if rank == 0:
            collected[i] = got
    return collected if rank == 0 else None
