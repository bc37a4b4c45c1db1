"""world_size-2 (gloo, CPU) test of the sharding + gather-to-rank-0 host logic."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stac_speech_translation_b200 import distributed as sd
from stac_speech_translation_b200 import synth


def _fake_compute(batch):
    """Deterministic stand-in for the GPU pipeline: ragged outputs that depend on the batch."""
    idx, t2 = batch
    base = torch.arange(len(idx) * t2, dtype=torch.float32).view(len(idx), t2, 1) + float(sum(idx))
    return {"enc_out": base.expand(-1, -1, 4).contiguous(), "greedy": (base[..., 0] % 7).to(torch.int32),
            "p_ctc": (base.expand(-1, -1, 3) * 0.5).to(torch.bfloat16)}


def _worker(rank, world, port, durations, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bucketed, per_rank = sd.plan(durations, world, max_batch_len=60.0, num_buckets=6, max_batch_ex=16)
        batches = [(b, int(25 * max(durations[i] for i in b)) + 1) for b in bucketed.batches]
        got = sd.run_sharded(batches, per_rank, _fake_compute, ["enc_out", "greedy", "p_ctc"])
        if rank == 0:
            ok = set(got) == set(range(len(batches)))
            for i, b in enumerate(batches):
                ref = _fake_compute(b)
                for k in ref:
                    ok = ok and got[i][k].dtype == ref[k].dtype and torch.equal(got[i][k], ref[k])
            ret.put((ok, len(batches), [len(p) for p in per_rank]))
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


def test_shard_and_gather_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    durations = synth.lognormal_durations(96, seed=3).tolist()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, durations, ret)) for r in range(2)]
    for p in procs:
        p.start()
    ok, n_batches, split = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and n_batches > 4 and min(split) >= 1 and sum(split) == n_batches


def _agree_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pg = sd.PeerGather.__new__(sd.PeerGather)          # the agreement step alone (the rest needs CUDA IPC)
        pg.ctl = dist.new_group(backend="gloo")
        pg._agree(True, "phase that works everywhere")     # must not raise
        raised = False
        try:
            pg._agree(rank != 1, "phase that fails on rank 1 only")
        except sd.PeerGatherUnavailable:
            raised = True
        dist.barrier(group=pg.ctl)                         # every rank is still in step after the failure
        ret.put((rank, raised))
    finally:
        dist.destroy_process_group()


def test_peer_gather_setup_failure_is_raised_on_every_rank():
    """One rank failing a set-up phase makes PeerGather raise PeerGatherUnavailable on ALL ranks together, so the caller
    (bench.py) can fall back to NCCL point-to-point without leaving a rank behind in a collective."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_agree_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(ret.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == {0: True, 1: True}


def _pipeline_worker(rank, world, port, ret):
    """Every rank runs the REAL host pipeline (fp32 mode) on its shard - the kernels behind the C ABI replaced by the
    test emulator, tests/abi_emulator.py - and rank 0 gathers; the result must equal the unsharded run."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import pytest
    import abi_emulator
    import stac_speech_translation_b200 as sb
    from util import TINY, oracle_modules, product_from_oracle
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mpatch = pytest.MonkeyPatch()
    try:
        abi_emulator.install(mpatch)
        torch.set_num_threads(1)
        mods = product_from_oracle(oracle_modules(TINY, vocab=32, num_encoder_layers=1), "fp32", device="cpu")
        pipe = sb.EncoderPipeline(mods)
        durations = [0.21, 0.5, 0.33, 0.45, 0.26, 0.38, 0.3]
        bucketed, per_rank = sd.plan(durations, world, max_batch_len=1.0, num_buckets=3, max_batch_ex=3)

        def compute(batch):
            wavs, wl = synth.synth_batch([durations[i] for i in batch], seed=100 + batch[0])
            res = pipe(wavs, wl)
            return {"enc_out": res["enc_out"], "greedy": res["greedy"], "p_ctc": res["p_ctc"]}

        got = sd.run_sharded(bucketed.batches, per_rank, compute, ["enc_out", "greedy", "p_ctc"])
        if rank == 0:
            ok = set(got) == set(range(len(bucketed.batches)))
            for i, b in enumerate(bucketed.batches):
                ref = compute(b)
                ok = ok and all(torch.equal(got[i][k], ref[k]) for k in ref)
            ret.put((ok, len(bucketed.batches), [len(p) for p in per_rank]))
    finally:
        mpatch.undo()
        dist.destroy_process_group()


def test_sharded_host_pipeline_equals_the_unsharded_run_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_pipeline_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    ok, n_batches, split = ret.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok and n_batches >= 3 and min(split) >= 1 and sum(split) == n_batches
