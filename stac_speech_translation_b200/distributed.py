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


class _IpcEvent:
    """CUDA interprocess event through the runtime API (ctypes on the libcudart torch already loaded):
    torch.cuda.Event.from_ipc_handle segfaults on wait() for an event of another device (torch 2.11)."""
    _rt = None

    class _Handle(__import__("ctypes").Structure):
        _fields_ = [("reserved", __import__("ctypes").c_char * 64)]

    @classmethod
    def rt(cls):
        import ctypes
        if cls._rt is None:
            cls._rt = ctypes.CDLL("libcudart.so.12")
            cls._rt.cudaIpcOpenEventHandle.argtypes = [ctypes.POINTER(ctypes.c_void_p), cls._Handle]
            cls._rt.cudaIpcGetEventHandle.argtypes = [ctypes.POINTER(cls._Handle), ctypes.c_void_p]
            cls._rt.cudaEventCreateWithFlags.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint]
            cls._rt.cudaEventRecord.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
            cls._rt.cudaStreamWaitEvent.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint]
            cls._rt.cudaMemcpyPeerAsync.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                                    ctypes.c_size_t, ctypes.c_void_p]
        return cls._rt

    def __init__(self, handle: Optional[bytes] = None):
        import ctypes
        rt = self.rt()
        self.ev = ctypes.c_void_p()
        if handle is None:       # create on the current device: cudaEventDisableTiming | cudaEventInterprocess
            self._check(rt.cudaEventCreateWithFlags(ctypes.byref(self.ev), 0x02 | 0x04), "cudaEventCreateWithFlags")
        else:                    # open an event exported by another process
            h = self._Handle()
            ctypes.memmove(ctypes.byref(h), handle, 64)
            self._check(rt.cudaIpcOpenEventHandle(ctypes.byref(self.ev), h), "cudaIpcOpenEventHandle")

    @staticmethod
    def _check(code, what):
        if code != 0:
            raise RuntimeError(f"{what} failed with cudaError {code}")

    def handle(self) -> bytes:
        import ctypes
        h = self._Handle()
        self._check(self.rt().cudaIpcGetEventHandle(ctypes.byref(h), self.ev), "cudaIpcGetEventHandle")
        return bytes(ctypes.string_at(ctypes.byref(h), 64))

    def record(self, stream: torch.cuda.Stream) -> None:
        import ctypes
        self._check(self.rt().cudaEventRecord(self.ev, ctypes.c_void_p(stream.cuda_stream)), "cudaEventRecord")

    def wait(self, stream: torch.cuda.Stream) -> None:
        """Make `stream` wait for the last record of this event.  The stream must live on the device the event was
        created on: cudaStreamWaitEvent on an IPC-opened event of ANOTHER device segfaults in the runtime (driver
        580.159), so rank 0 keeps one stream per peer ON the peer's device."""
        import ctypes
        self._check(self.rt().cudaStreamWaitEvent(ctypes.c_void_p(stream.cuda_stream), self.ev, 0), "cudaStreamWaitEvent")


class PeerGatherUnavailable(RuntimeError):
    """The peer-memory transport cannot be set up on this box (raised on every rank together)."""


class PeerGather:
    """Results of every rank to rank 0 WITHOUT a communication kernel: rank 0 pulls them over NVLink with
    copy-engine peer copies out of result buffers that the producing kernels wrote directly.

    Why not NCCL send/recv for the bulk: measured on 2/4/8 B200, NCCL point-to-point delivers ~75 GB/s per peer for the
    505 MB of bf16 posteriors a rank produces per step (whatever the channel settings), so the transfer outlasts the
    5 ms step it should hide under, and its kernels hold SMs the persistent compute kernels count on.  A peer copy
    runs on the copy engines at NVLink speed, uses no SM and overlaps the next step completely.
    torch.distributed stays the plumbing: a gloo group carries the per-step control messages (two tiny CPU messages
    per peer and step); CUDA IPC handles of the buffers and of the ordering events are exchanged once at set-up.

    Every non-root rank owns ``slots`` sets of result buffers (double buffering: rank 0 pulls step i while the rank
    computes step i+1 into the other set).  Per step and slot:
        rank r : begin_write(step) -> run the path with outputs=self.slot(step) -> end_write(step)
        rank 0 : collect(step)     -> pulls every peer's slot into self.gathered[r] on a side stream
    """

    def _agree(self, ok: bool, what: str):
        """All ranks learn whether a set-up phase worked everywhere; raises PeerGatherUnavailable on EVERY rank if it
        did not on any, so the caller can fall back to another transport without leaving a rank behind."""
        flag = torch.tensor([1 if ok else 0], dtype=torch.int64)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.ctl)
        if int(flag) == 0:
            raise PeerGatherUnavailable(what)

    def __init__(self, specs: Dict[str, tuple], device: torch.device, slots: int = 2):
        from torch.multiprocessing.reductions import reduce_tensor
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device, self.n_slots, self.keys = device, slots, list(specs)
        self.ctl = dist.new_group(backend="gloo")
        self._msg = torch.zeros(1, dtype=torch.int64)
        payload, ok = None, True
        try:                                    # phase 1 (local): result slots, ordering events, their IPC handles
            if self.rank != 0:
                self._slots = [{k: torch.empty(shape, dtype=dt, device=device) for k, (shape, dt) in specs.items()}
                               for _ in range(slots)]
                self._ready = [_IpcEvent() for _ in range(slots)]
                for ev in self._ready:
                    ev.record(torch.cuda.current_stream(device))
                payload = {"device": device.index,
                           "tensors": [{k: reduce_tensor(t) for k, t in s.items()} for s in self._slots],
                           "ready": [ev.handle() for ev in self._ready]}
        except Exception as e:                  # noqa: BLE001 - any failure here means "no peer transport on this box"
            ok, self._why = False, repr(e)
        self._agree(ok, "CUDA IPC export of the result buffers failed")
        gathered = [None] * self.world if self.rank == 0 else None
        dist.gather_object(payload, gathered, dst=0, group=self.ctl)
        done_handles = [None] * self.world
        try:                                    # phase 2 (rank 0): open the peers' buffers and events
            if self.rank == 0:
                self._peer, self._peer_ready, self._done, self._pstream, self._pdev, self._last = {}, {}, {}, {}, {}, {}
                self.gathered = {}
                for r in range(1, self.world):
                    p = gathered[r]
                    pd = p["device"]
                    if not torch.cuda.can_device_access_peer(device.index, pd):
                        raise RuntimeError(f"no peer access {device.index} -> {pd}")
                    self._pdev[r] = pd
                    self._peer[r] = [{k: fn(*args) for k, (fn, args) in s.items()} for s in p["tensors"]]
                    self.gathered[r] = {k: torch.empty(shape, dtype=dt, device=device)
                                        for k, (shape, dt) in specs.items()}
                    # one torch cross-device copy makes torch enable peer access between the two devices
                    self.gathered[r][self.keys[0]].view(-1)[:1].copy_(self._peer[r][0][self.keys[0]].view(-1)[:1])
                    with torch.cuda.device(pd):     # everything that touches an IPC event lives on the peer's device
                        self._pstream[r] = torch.cuda.Stream(device=pd)
                        self._peer_ready[r] = [_IpcEvent(h) for h in p["ready"]]
                        self._done[r] = [_IpcEvent() for _ in range(slots)]
                        for ev in self._done[r]:
                            ev.record(self._pstream[r])
                        done_handles[r] = {"device": pd, "done": [ev.handle() for ev in self._done[r]]}
                torch.cuda.synchronize(device)
        except Exception as e:                  # noqa: BLE001
            ok, self._why = False, repr(e)
        self._agree(ok, "rank 0 could not open the peers' buffers (devices not visible to it, or no peer access)")
        mine = [None]
        dist.scatter_object_list(mine, done_handles if self.rank == 0 else None, src=0, group=self.ctl)
        try:
            if self.rank != 0:
                d = mine[0]
                self._done_here = [_IpcEvent(h) for h in d["done"]]       # created by rank 0 on MY device
        except Exception as e:                  # noqa: BLE001
            ok, self._why = False, repr(e)
        self._agree(ok, "a rank could not open rank 0's completion events")
        dist.barrier(group=self.ctl)

    def _post(self, value: int, dst: int) -> None:
        """Non-blocking control message (a blocking gloo send waits for the matching receive: rank 0 acknowledging
        step i while the peer announces step i + 1 would deadlock)."""
        if not hasattr(self, "_inflight"):
            self._inflight = []
        t = torch.tensor([value], dtype=torch.int64)
        self._inflight.append((dist.isend(t, dst=dst, group=self.ctl), t))
        self._inflight = [(w, x) for w, x in self._inflight if not w.is_completed()]

    # ---- producing ranks ----
    def slot(self, step: int) -> Dict[str, torch.Tensor]:
        return self._slots[step % self.n_slots]

    def begin_write(self, step: int) -> None:
        """Before the kernels of `step` overwrite their slot: rank 0 must have pulled the step that used it last."""
        if self.rank == 0 or step < self.n_slots:
            return
        dist.recv(self._msg, src=0, group=self.ctl)                    # "the pull of step - n_slots is enqueued"
        self._done_here[step % self.n_slots].wait(torch.cuda.current_stream(self.device))

    def end_write(self, step: int) -> None:
        if self.rank == 0:
            return
        self._ready[step % self.n_slots].record(torch.cuda.current_stream(self.device))
        self._post(step, 0)                                            # "the results of `step` are enqueued"

    # ---- rank 0 ----
    def collect(self, step: int) -> None:
        """Enqueue the pull of every peer's results of `step` (copy engines, side stream) and acknowledge it.  `step`
        is a counter that only ever grows (the peers consume the acknowledgement of step i at step i + slots)."""
        if self.rank != 0:
            return
        import ctypes
        s = step % self.n_slots
        rt = _IpcEvent.rt()
        for r in range(1, self.world):
            dist.recv(self._msg, src=r, group=self.ctl)
            pd, st = self._pdev[r], self._pstream[r]
            with torch.cuda.device(pd):
                self._peer_ready[r][s].wait(st)
                for k in self.keys:
                    dst, src = self.gathered[r][k], self._peer[r][s][k]
                    _IpcEvent._check(rt.cudaMemcpyPeerAsync(ctypes.c_void_p(dst.data_ptr()), self.device.index,
                                                            ctypes.c_void_p(src.data_ptr()), pd,
                                                            dst.numel() * dst.element_size(),
                                                            ctypes.c_void_p(st.cuda_stream)), "cudaMemcpyPeerAsync")
                self._done[r][s].record(st)
                ev = torch.cuda.Event()
                ev.record(st)
                self._last[r] = ev
            self._post(step, r)                 # acknowledgement: the pull of `step` is enqueued

    def gathered_view(self, r: int, step: int) -> Dict[str, torch.Tensor]:
        return self.gathered[r]

    def finish(self, last_step: Optional[int] = None) -> None:
        """Rank 0: the current stream waits for every pull enqueued so far."""
        del last_step
        if self.rank == 0:
            for ev in self._last.values():
                torch.cuda.current_stream(self.device).wait_event(ev)


class PushGather:
    """Results of every rank to rank 0, PUSHED: every producing rank copies its finished result slot into buffers that
    live on rank 0's GPU (peer-mapped through CUDA IPC) with its OWN copy engines on its OWN side stream.

    Same interface as PeerGather (begin_write / slot / end_write on the producers, collect / finish on rank 0) and the
    same properties - no communication kernel, no SM, the transfer of step i overlaps the compute of step i + 1 - but
    exactly ONE working CUDA context per GPU.  PeerGather has rank 0 drive the copies from a second context on every
    peer GPU; that context's copy and event work is time-sliced against the peer's compute context, which is what a
    2-GPU step lost (4.99 ms against 4.54 ms on one GPU, and no better when only the ids travelled; round 2).  Here the
    foreign contexts that opening an IPC handle creates (every rank holds one on GPU 0) never submit work.

    Ordering.  Producer r, step i, slot s = i % slots: the compute stream waits for push(i - slots) to have read the
    local slot (plain event), the kernels write the slot, the push stream waits for them (plain event) and issues one
    cudaMemcpyPeerAsync per tensor into rank 0's gathered[r][s] followed by an 8-byte "step i + 1 has landed" mark in
    rank 0's mailbox - stream order puts the mark behind the data.  Rank 0 never waits on another device's event (that
    segfaults, see _IpcEvent.wait): `arrived(step)` reads the mailbox with a tiny device-to-host copy on a side stream.
    Flow control against rank 0's consumer is a control message: producer r may overwrite gathered[r][s] again only
    after rank 0 released it (collect(step) = wait for the arrival of `step` from every rank, hand the buffers to the
    caller, acknowledge)."""

    def _agree(self, ok: bool, what: str):
        flag = torch.tensor([1 if ok else 0], dtype=torch.int64)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.ctl)
        if int(flag) == 0:
            raise PeerGatherUnavailable(what)

    def __init__(self, specs: Dict[str, tuple], device: torch.device, slots: int = 2):
        from torch.multiprocessing.reductions import reduce_tensor
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device, self.n_slots, self.keys = device, slots, list(specs)
        self.ctl = dist.new_group(backend="gloo")
        self._msg = torch.zeros(1, dtype=torch.int64)
        ok, payload = True, [None] * self.world
        try:
            if self.rank == 0:
                # gathered[r][s][key] and the mailbox live on rank 0's GPU; every producer gets the handles of its share
                self._bufs = {r: [{k: torch.empty(shape, dtype=dt, device=device) for k, (shape, dt) in specs.items()}
                                  for _ in range(slots)] for r in range(1, self.world)}
                self._mailbox = torch.zeros(self.world, slots, dtype=torch.int64, device=device)
                self._mail_host = torch.zeros(self.world, slots, dtype=torch.int64).pin_memory()
                self._mail_stream = torch.cuda.Stream(device=device)
                torch.cuda.synchronize(device)
                for r in range(1, self.world):
                    payload[r] = {"device": device.index,
                                  "tensors": [{k: reduce_tensor(t) for k, t in s.items()} for s in self._bufs[r]],
                                  "mailbox": reduce_tensor(self._mailbox)}
        except Exception as e:                  # noqa: BLE001
            ok, self._why = False, repr(e)
        self._agree(ok, "CUDA IPC export of rank 0's gather buffers failed")
        mine = [None]
        dist.scatter_object_list(mine, payload if self.rank == 0 else None, src=0, group=self.ctl)
        try:
            if self.rank != 0:
                p = mine[0]
                self._root_dev = p["device"]
                if not torch.cuda.can_device_access_peer(device.index, self._root_dev):
                    raise RuntimeError(f"no peer access {device.index} -> {self._root_dev}")
                self._remote = [{k: fn(*args) for k, (fn, args) in s.items()} for s in p["tensors"]]
                fn, args = p["mailbox"]
                self._remote_mail = fn(*args)
                self._slots = [{k: torch.empty(shape, dtype=dt, device=device) for k, (shape, dt) in specs.items()}
                               for _ in range(slots)]
                self._tick = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(slots)]
                # one torch cross-device copy makes torch enable peer access between the two devices
                self._tick[0].copy_(self._remote_mail.view(-1)[:1])
                self._push_stream = torch.cuda.Stream(device=device)
                self._computed = [torch.cuda.Event() for _ in range(slots)]
                self._pushed = [torch.cuda.Event() for _ in range(slots)]
                for ev in self._pushed:
                    ev.record(self._push_stream)
                torch.cuda.synchronize(device)
        except Exception as e:                  # noqa: BLE001
            ok, self._why = False, repr(e)
        self._agree(ok, "a rank could not open rank 0's gather buffers (device not visible to it, or no peer access)")
        dist.barrier(group=self.ctl)

    def _post(self, value: int, dst: int) -> None:
        if not hasattr(self, "_inflight"):
            self._inflight = []
        t = torch.tensor([value], dtype=torch.int64)
        self._inflight.append((dist.isend(t, dst=dst, group=self.ctl), t))
        self._inflight = [(w, x) for w, x in self._inflight if not w.is_completed()]

    # ---- producing ranks ----
    def slot(self, step: int) -> Dict[str, torch.Tensor]:
        return self._slots[step % self.n_slots]

    def begin_write(self, step: int) -> None:
        """Before the kernels of `step` overwrite their slot: its previous push has read it (event), and rank 0 has
        released the remote buffers that push will overwrite (control message, from step `slots` on)."""
        if self.rank == 0:
            return
        s = step % self.n_slots
        torch.cuda.current_stream(self.device).wait_event(self._pushed[s])
        if step >= self.n_slots:
            dist.recv(self._msg, src=0, group=self.ctl)                # "step - slots has been consumed"

    def end_write(self, step: int) -> None:
        """Enqueue the push of `step` behind its kernels: tensors first, then the arrival mark."""
        if self.rank == 0:
            return
        import ctypes
        s = step % self.n_slots
        cur = torch.cuda.current_stream(self.device)
        self._computed[s].record(cur)
        rt = _IpcEvent.rt()
        with torch.cuda.stream(self._push_stream):
            self._push_stream.wait_event(self._computed[s])
            for k in self.keys:
                dst, src = self._remote[s][k], self._slots[s][k]
                _IpcEvent._check(rt.cudaMemcpyPeerAsync(ctypes.c_void_p(dst.data_ptr()), self._root_dev,
                                                        ctypes.c_void_p(src.data_ptr()), self.device.index,
                                                        src.numel() * src.element_size(),
                                                        ctypes.c_void_p(self._push_stream.cuda_stream)),
                                 "cudaMemcpyPeerAsync")
            self._tick[s].fill_(step + 1)
            mark = self._remote_mail[self.rank, s:s + 1]
            _IpcEvent._check(rt.cudaMemcpyPeerAsync(ctypes.c_void_p(mark.data_ptr()), self._root_dev,
                                                    ctypes.c_void_p(self._tick[s].data_ptr()), self.device.index, 8,
                                                    ctypes.c_void_p(self._push_stream.cuda_stream)),
                             "cudaMemcpyPeerAsync")
            self._pushed[s].record(self._push_stream)

    # ---- rank 0 ----
    def arrived(self, step: int) -> bool:
        """Rank 0: have the results of `step` of every producer landed?  (tiny D2H read of the mailbox, side stream)"""
        with torch.cuda.stream(self._mail_stream):
            self._mail_host.copy_(self._mailbox, non_blocking=True)
        self._mail_stream.synchronize()
        s = step % self.n_slots
        return all(int(self._mail_host[r, s]) >= step + 1 for r in range(1, self.world))

    def gathered_view(self, r: int, step: int) -> Dict[str, torch.Tensor]:
        return self._bufs[r][step % self.n_slots]

    def collect(self, step: int) -> None:
        """Rank 0, called right after its own step `step` has been enqueued: consume the PREVIOUS step of every producer
        (wait for its arrival marks; the caller may read gathered_view(r, step - 1) until the next collect) and release
        those buffers.  Lagging one step keeps rank 0's GPU queue non-empty while its host waits for the marks."""
        if self.rank != 0 or step < 1:
            return
        self._consume(step - 1)

    def _consume(self, upto: int) -> None:
        """Every step up to `upto` exactly once, in order: wait for its arrival marks, release its buffers."""
        import time
        done = getattr(self, "_consumed_upto", -1)
        for step in range(done + 1, upto + 1):
            t0 = time.perf_counter()
            while not self.arrived(step):
                if time.perf_counter() - t0 > 120.0:
                    raise RuntimeError(f"PushGather: results of step {step} did not arrive within 120 s")
            for r in range(1, self.world):
                self._post(step, r)                                      # release: gathered[r][step % slots] is free
            self._consumed_upto = step

    def finish(self, last_step: Optional[int] = None) -> None:
        """Producers: every push enqueued so far has completed.  Rank 0: everything up to `last_step` has arrived."""
        if self.rank != 0:
            self._push_stream.synchronize()
            return
        if last_step is not None and last_step >= 0:
            self._consume(last_step)
