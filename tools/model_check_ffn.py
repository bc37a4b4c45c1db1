"""Randomised model check of the synchronisation protocol of csrc/ffn_fused.cu (fused feed-forward block), in the manner
of tools/model_check_mha2.py: TMA producer (h tile + the W1 / W2 units in MMA order through a ring of three), ONE MMA
issuer (S two chunks ahead of P.W2), 16 epilogue warps (GELU: S -> P; per tile: O -> staging -> reduce-add, the staging
tiles living in the P buffers), transcribed with the kernel's loops, barrier counts and parity expressions.
`TWO_ISSUERS = True` models the earlier design (separate S and O issuer warps, each skipping the other's ring units),
which DESIGN.md section 4 records as the race that fired on an 8-GPU run: the checker finds it.
Usage:  python tools/model_check_ffn.py [runs] [seed]"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from model_check_mha2 import Hazard, MBar, wait  # noqa: E402

RING = 3
EPI = 16
TWO_ISSUERS = False


class Sim:
    def __init__(self, n_tiles, n_chunks, rng):
        self.n_tiles, self.n, self.rng = n_tiles, n_chunks, rng
        self.h_full, self.h_free = MBar("h_full", 1), MBar("h_free", 1)
        self.w_full = [MBar(f"w_full[{s}]", 1) for s in range(RING)]
        self.w_empty = [MBar(f"w_empty[{s}]", 1) for s in range(RING)]
        self.s_full = [MBar(f"s_full[{i}]", 1) for i in range(2)]
        self.s_free = [MBar(f"s_free[{i}]", EPI) for i in range(2)]
        self.p_full = [MBar(f"p_full[{i}]", EPI) for i in range(2)]
        self.p_free = [MBar(f"p_free[{i}]", 1) for i in range(2)]
        self.o_full, self.o_free = MBar("o_full", 1), MBar("o_free", EPI)
        self.async_events, self.fifo = [], [[], []]            # one MMA FIFO per issuer warp
        self.slot = [None] * RING                              # unit index held by a ring slot (None while loading)
        self.slot_reads = [0] * RING
        self.h_tile, self.h_reads = None, 0
        self.s_version, self.p_version, self.p_reads = [-1, -1], [-1, -1], [0, 0]
        self.o_state = None                                    # (tile, accumulated chunks)
        self.o_read = -1
        self.staging_busy = [0] * EPI                          # outstanding reduce-add reads of a warp's staging tile
        self.done_tiles = []

    def later(self, action):
        # mostly short, sometimes very long (a congested memory system re-orders completions across many steps)
        delay = self.rng.randint(0, 6) if self.rng.random() < 0.85 else self.rng.randint(50, 400)
        self.async_events.append([delay, action])

    def unit_order(self):
        """Global sequence of ring units: ("w1", tile, chunk, u) / ("w2", tile, chunk, kb), as the producer loads them."""
        seq = []
        for t in range(self.n_tiles):
            def w1(c):
                seq.extend(("w1", t, c, u) for u in range(2))

            def w2(c):
                seq.extend(("w2", t, c, kb) for kb in range(2))
            w1(0)
            if self.n > 1:
                w1(1)
            for c in range(self.n):
                w2(c)
                if c + 2 < self.n:
                    w1(c + 2)
        return seq

    def producer(self):
        units = self.unit_order()
        k = 0
        for t in range(self.n_tiles):
            yield from wait(self.h_free, (t & 1) ^ 1, t - 1)
            if self.h_reads:
                raise Hazard("h tile reloaded while an MMA still reads it")
            self.h_tile = None

            def h_landed(tt=t):
                self.h_tile = tt
                self.h_full.arrive()
            self.later(h_landed)
            while k < len(units) and units[k][1] == t:
                s, ph = k % RING, (k // RING) & 1
                yield from wait(self.w_empty[s], ph ^ 1, k // RING - 1)
                if self.slot_reads[s]:
                    raise Hazard("ring slot refilled while an MMA still reads it")
                self.slot[s] = None

                def landed(ss=s, kk=k):
                    self.slot[ss] = kk
                    self.w_full[ss].arrive()
                self.later(landed)
                k += 1
                yield

    def issuer(self, which):
        """which: 0 = the single issuer (or the S issuer of the two-warp design), 1 = the O issuer of that design."""
        units = self.unit_order()
        index = {u: k for k, u in enumerate(units)}
        n = self.n
        g = 0
        do_s = which == 0
        do_o = which == 1 or not TWO_ISSUERS

        def issue_s(t, c, gi):
            i = gi & 1
            yield from wait(self.s_free[i], ((gi >> 1) & 1) ^ 1, (gi >> 1) - 1)
            for u in range(2):
                k = index[("w1", t, c, u)]
                s = k % RING
                yield from wait(self.w_full[s], (k // RING) & 1, k // RING)
                if self.slot[s] != k:
                    raise Hazard(f"S: ring slot holds unit {self.slot[s]}, expected {k}")
                if self.h_tile != t:
                    raise Hazard("S: h tile of another row tile")
                self.slot_reads[s] += 1
                self.h_reads += 1
                commits = [self.w_empty[s]] + ([self.s_full[i]] if u == 1 else [])
                self.fifo[which].append(("S", gi, s, u, commits))
                yield

        def issue_o(t, c, gi):
            i = gi & 1
            yield from wait(self.p_full[i], (gi >> 1) & 1, gi >> 1)
            for kb in range(2):
                k = index[("w2", t, c, kb)]
                s = k % RING
                yield from wait(self.w_full[s], (k // RING) & 1, k // RING)
                if self.slot[s] != k:
                    raise Hazard(f"O: ring slot holds unit {self.slot[s]}, expected {k}")
                if self.p_version[i] != gi:
                    raise Hazard("O: P buffer holds another chunk")
                self.slot_reads[s] += 1
                self.p_reads[i] += 1
                commits = [self.w_empty[s]] + ([self.p_free[i]] if kb == 1 else [])
                self.fifo[which].append(("O", gi, s, (t, c, kb), commits))
                yield

        for t in range(self.n_tiles):
            if do_s:
                yield from wait(self.h_full, t & 1, t)
                yield from issue_s(t, 0, g)
                if n > 1:
                    yield from issue_s(t, 1, g + 1)
                if n <= 2:
                    self.fifo[which].append(("H", 0, None, None, [self.h_free]))
            for c in range(n):
                if do_o:
                    if c == 0:
                        yield from wait(self.o_free, (t & 1) ^ 1, t - 1)
                        if t >= 1 and self.o_read != t - 1:
                            raise Hazard("first P.W2 of a tile issued before the previous O was read out")
                    yield from issue_o(t, c, g + c)
                if do_s and c + 2 < n:
                    yield from issue_s(t, c + 2, g + c + 2)
                    if c + 2 == n - 1:
                        self.fifo[which].append(("H", 0, None, None, [self.h_free]))
            if do_o:
                self.fifo[which].append(("T", t, None, None, [self.o_full]))
            g += n
            yield

    def mma_complete(self, f):
        kind, gi, s, info, commits = self.fifo[f].pop(0)
        if kind == "S":
            self.slot_reads[s] -= 1
            self.h_reads -= 1
            if info == 1:
                self.s_version[gi & 1] = gi
        elif kind == "O":
            t, c, kb = info
            self.slot_reads[s] -= 1
            self.p_reads[gi & 1] -= 1
            if self.p_version[gi & 1] != gi:
                raise Hazard("P.W2 executed after its P buffer was overwritten")
            if kb == 1:
                self.o_state = (t, 1) if c == 0 else (t, self.o_state[1] + 1)
        for b in commits:
            b.arrive()

    def epilogue(self, warp):
        g = 0
        staged = False
        for t in range(self.n_tiles):
            for c in range(self.n):
                i, u = g & 1, (g >> 1) & 1
                yield from wait(self.s_full[i], u, g >> 1)
                if self.s_version[i] != g:
                    raise Hazard(f"GELU read S of chunk {self.s_version[i]}, expected {g}")
                yield
                if self.s_version[i] != g:
                    raise Hazard("S overwritten while it was being read")
                self.s_free[i].arrive()
                yield
                yield from wait(self.p_free[i], u ^ 1, (g >> 1) - 1)
                if staged:
                    while self.staging_busy[warp]:                 # cp.async.bulk.wait_group.read 0 (own stores)
                        yield
                    self.bar_count = getattr(self, "bar_count", 0) + 1     # bar.sync 1, 512
                    target = (self.bar_count + EPI - 1) // EPI * EPI
                    while getattr(self, "bar_count", 0) < target:
                        yield
                    staged = False
                if self.p_reads[i]:
                    raise Hazard("P buffer written while a P.W2 still reads it")
                if any(self.staging_busy[8 * i:8 * i + 8]):
                    raise Hazard("P buffer written while a reduce-add still reads a staging tile inside it")
                if warp == 0:
                    self.p_version[i] = g
                self.p_full[i].arrive()
                g += 1
                yield
            yield from wait(self.o_full, t & 1, t)
            if self.o_state != (t, self.n):
                raise Hazard(f"epilogue read O holding {self.o_state}, expected {(t, self.n)}")
            yield
            if self.o_state != (t, self.n):
                raise Hazard("O overwritten while it was being read out")
            if warp == 0:
                self.o_read = t
            self.o_free.arrive()
            for half in range(2):
                while self.staging_busy[warp]:                     # wait_group.read 0 before re-staging
                    yield
                if self.p_reads[warp >> 3]:
                    raise Hazard("staging tile written into a P buffer a P.W2 still reads")
                self.staging_busy[warp] += 1

                def read_done(w=warp):
                    self.staging_busy[w] -= 1
                self.later(read_done)
                yield
            if warp == 0:
                self.done_tiles.append(t)
            staged = True

    def run(self):
        roles = [self.producer(), self.issuer(0)] + ([self.issuer(1)] if TWO_ISSUERS else [])
        roles += [self.epilogue(k) for k in range(EPI)]
        live = list(range(len(roles)))
        idle = 0
        while live:
            progressed = False
            for ev in list(self.async_events):
                ev[0] -= 1
                if ev[0] <= 0:
                    self.async_events.remove(ev)
                    ev[1]()
                    progressed = True
            for f in self.rng.sample([0, 1], 2):
                if self.fifo[f] and self.rng.random() < 0.5:
                    self.mma_complete(f)
                    progressed = True
            k = self.rng.choice(live)
            before = self.snapshot()
            try:
                next(roles[k])
            except StopIteration:
                live.remove(k)
                progressed = True
            progressed = progressed or self.snapshot() != before
            idle = 0 if progressed or self.async_events or any(self.fifo) else idle + 1
            if idle > 4000:
                raise Hazard(f"deadlock: roles {live} are blocked with nothing in flight")
        if self.done_tiles != list(range(self.n_tiles)):
            raise Hazard("not every tile was written")

    def snapshot(self):
        bars = [self.h_full, self.h_free, self.o_full, self.o_free] + self.w_full + self.w_empty
        bars += self.s_full + self.s_free + self.p_full + self.p_free
        return tuple((b.phase, b.pending) for b in bars) + (len(self.done_tiles), getattr(self, "bar_count", 0),
                                                             tuple(self.staging_busy))


def check(runs=100, seed=0, two_issuers=False):
    global TWO_ISSUERS
    TWO_ISSUERS = two_issuers
    rng = random.Random(seed)
    for _ in range(runs):
        Sim(rng.randint(1, 4), rng.choice([1, 2, 3, 4, 8]), rng).run()
    return runs


if __name__ == "__main__":
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    two = "--two-issuers" in sys.argv
    print("ok:", check(runs, seed, two), "random schedules, no deadlock, no parity aliasing, no data hazard",
          "(two issuer warps)" if two else "(one issuer warp: the kernel as built)")
