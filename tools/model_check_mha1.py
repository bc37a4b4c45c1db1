"""Randomised model check of the synchronisation protocol of csrc/attention_tc.cu (the default attention kernel), in the
manner of tools/model_check_mha2.py: the roles (TMA producer, two MMA issuers with S running two steps ahead, 2 x 8
softmax warps, output store warp) transcribed with the kernel's own loops, barrier counts and parity expressions;
mbarrier parity semantics, asynchronous TMA and in-order MMA completions, a random scheduler; every wait states which
completion it means, and the data hazards of the K/V ring, the Q buffers (which double as output staging), the two S / P
buffers per group and the O accumulator are tracked.
Usage:  python tools/model_check_mha1.py [runs] [seed]"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from model_check_mha2 import Hazard, MBar, wait  # noqa: E402

KV_STAGES = 5
# False: one o_staged barrier per query group, as the kernel is built by default.  True: one per (Q buffer, group), the
# -DMHA_OSTAGED_PER_BUFFER build (see the finding in DESIGN.md section 9).
OSTAGED_PER_BUFFER = False


class Sim:
    def __init__(self, items, rng):
        self.items, self.rng = items, rng                      # items: (n_kt, active1)
        B = lambda name, c: MBar(name, c)
        self.q_full = [[B(f"q_full[{b}][{w}]", 1) for w in range(2)] for b in range(2)]
        self.q_empty = [[B(f"q_empty[{b}][{w}]", 2) for w in range(2)] for b in range(2)]
        self.kv_full = [B(f"kv_full[{s}]", 1) for s in range(KV_STAGES)]
        self.kv_empty = [B(f"kv_empty[{s}]", 2) for s in range(KV_STAGES)]
        self.s_full = [[B(f"s_full[{w}][{i}]", 1) for i in range(2)] for w in range(2)]
        self.s_free = [[B(f"s_free[{w}][{i}]", 8) for i in range(2)] for w in range(2)]
        self.p_full = [[B(f"p_full[{w}][{i}]", 8) for i in range(2)] for w in range(2)]
        self.p_free = [[B(f"p_free[{w}][{i}]", 1) for i in range(2)] for w in range(2)]
        self.o_free = [B(f"o_free[{w}]", 8) for w in range(2)]
        self.o_staged = [[B(f"o_staged[{b}][{w}]", 8) for w in range(2)] for b in range(2)]
        self.async_events, self.mma_fifo = [], [[], []]
        self.stage_fill, self.stage_reads = [None] * KV_STAGES, [0] * KV_STAGES
        self.q_state = [[None, None], [None, None]]            # ("q", item) | ("out", item) | None (loading)
        self.q_reads = [[0, 0], [0, 0]]
        self.s_version = [[-1, -1], [-1, -1]]
        self.p_version = [[-1, -1], [-1, -1]]
        self.p_reads = [[0, 0], [0, 0]]
        self.pv_done = [0, 0]
        self.o_item = [None, None]                             # (item use, accumulated P.V blocks)
        self.o_read = [-1, -1]                                 # last item use whose O the softmax warps have read out
        self.outputs = []

    def later(self, action):
        # mostly short, sometimes very long (a congested memory system re-orders completions across many steps)
        delay = self.rng.randint(0, 6) if self.rng.random() < 0.85 else self.rng.randint(50, 400)
        self.async_events.append([delay, action])

    def producer(self):
        stage = kv_phase = fills = tile = 0
        q_uses = [[0, 0], [0, 0]]
        for n_done, (n_kt, active1) in enumerate(self.items):
            buf = n_done & 1
            for w in range(2):
                if w == 1 and not active1:
                    continue
                k = q_uses[buf][w]
                q_uses[buf][w] += 1
                yield from wait(self.q_empty[buf][w], (k & 1) ^ 1, k - 1)
                if self.q_reads[buf][w]:
                    raise Hazard("Q buffer refilled while an MMA / the TMA store still reads it")
                self.q_state[buf][w] = None

                def landed(b=buf, ww=w, n=n_done):
                    self.q_state[b][ww] = ("q", n)
                    self.q_full[b][ww].arrive()
                self.later(landed)
            for _ in range(n_kt):
                yield from wait(self.kv_empty[stage], kv_phase ^ 1, fills // KV_STAGES - 1)
                if self.stage_reads[stage]:
                    raise Hazard("K/V stage refilled while an MMA still reads it")
                self.stage_fill[stage] = None

                def kv_landed(s=stage, t=tile):
                    self.stage_fill[s] = t
                    self.kv_full[s].arrive()
                self.later(kv_landed)
                tile += 1
                fills += 1
                stage += 1
                if stage == KV_STAGES:
                    stage, kv_phase = 0, kv_phase ^ 1
                yield

    def store_warp(self):
        cnt = [[0, 0], [0, 0]]
        for n_done, (n_kt, active1) in enumerate(self.items):
            buf = n_done & 1
            for w in range(2):
                if w == 1 and not active1:
                    continue
                sb = buf if OSTAGED_PER_BUFFER else 0
                yield from wait(self.o_staged[sb][w], cnt[sb][w] & 1, cnt[sb][w])
                cnt[sb][w] += 1
                if self.q_state[buf][w] != ("out", n_done):
                    raise Hazard("store warp found something else than this item's output tile in the Q buffer")
                self.q_reads[buf][w] += 1
                for _ in range(self.rng.randint(0, 4)):        # cp.async.bulk.wait_group.read 0
                    yield
                self.q_reads[buf][w] -= 1
                self.outputs.append((n_done, w))
                self.q_empty[buf][w].arrive()
                yield

    def issuer(self, w):
        items, n_items = self.items, len(self.items)

        class Cursor:
            pass

        def load_item(c):
            c.valid = c.n_done < n_items
            c.virt = False
            if c.valid:
                c.n_kt = items[c.n_done][0]
                c.virt = w == 1 and not items[c.n_done][1]

        def advance(c):
            c.tile += 1
            c.stage += 1
            if c.stage == KV_STAGES:
                c.stage, c.phase = 0, c.phase ^ 1
            c.j += 1
            if c.j == c.n_kt:
                c.j = 0
                c.n_done += 1
                load_item(c)

        def copy(c):
            n = Cursor()
            n.__dict__.update(c.__dict__)
            return n

        sc = Cursor()
        sc.j = sc.n_done = sc.stage = sc.phase = sc.tile = 0
        sc.n_kt = 1
        load_item(sc)
        pc = copy(sc)
        g_s = g_p = 0
        q_fill = [0, 0]
        items_started = 0
        while pc.valid:
            if w == 1 and pc.virt:
                for _ in range(pc.n_kt):
                    yield from wait(self.kv_full[pc.stage], pc.phase, pc.tile // KV_STAGES)
                    self.kv_empty[pc.stage].arrive()
                    advance(pc)
                    yield
                sc = copy(pc)
                continue
            while sc.valid and not (w == 1 and sc.virt) and g_s < g_p + 2:
                i, buf = g_s & 1, sc.n_done & 1
                yield from wait(self.kv_full[sc.stage], sc.phase, sc.tile // KV_STAGES)
                yield from wait(self.s_free[w][i], ((g_s >> 1) & 1) ^ 1, (g_s >> 1) - 1)
                if sc.j == 0:
                    yield from wait(self.q_full[buf][w], q_fill[buf] & 1, q_fill[buf])
                if self.stage_fill[sc.stage] != sc.tile:
                    raise Hazard("S: stage holds another tile")
                if self.q_state[buf][w] != ("q", sc.n_done):
                    raise Hazard("S: Q buffer does not hold this item's queries")
                last = sc.j == sc.n_kt - 1
                self.stage_reads[sc.stage] += 1
                self.q_reads[buf][w] += 1
                commits = [self.s_full[w][i]] + ([self.q_empty[buf][w]] if last else [])
                self.mma_fifo[w].append(("S", g_s, sc.stage, (buf, w), commits))
                if last:
                    q_fill[buf] += 1
                g_s += 1
                advance(sc)
                yield
            i = g_p & 1
            yield from wait(self.p_full[w][i], (g_p >> 1) & 1, g_p >> 1)
            if pc.j == 0:
                yield from wait(self.o_free[w], (items_started & 1) ^ 1, items_started - 1)
                if items_started >= 1 and self.o_read[w] != items_started - 1:
                    raise Hazard("first P.V of an item issued before the previous item's O was read out")
                items_started += 1
            if self.p_version[w][i] != g_p:
                raise Hazard(f"P.V({g_p}) issued on P of step {self.p_version[w][i]}")
            if self.stage_fill[pc.stage] != pc.tile:
                raise Hazard("P.V: stage holds another tile")
            self.stage_reads[pc.stage] += 1
            self.p_reads[w][i] += 1
            self.mma_fifo[w].append(("PV", g_p, pc.stage, (i, items_started - 1, pc.j),
                                     [self.p_free[w][i], self.kv_empty[pc.stage]]))
            g_p += 1
            advance(pc)
            yield

    def mma_complete(self, w):
        kind, g, stage, info, commits = self.mma_fifo[w].pop(0)
        self.stage_reads[stage] -= 1
        if kind == "S":
            buf, ww = info
            self.q_reads[buf][ww] -= 1
            self.s_version[w][g & 1] = g
        else:
            i, use, j = info
            if self.p_version[w][i] != g:
                raise Hazard("P.V executed after its P buffer was overwritten")
            self.p_reads[w][i] -= 1
            self.pv_done[w] += 1
            self.o_item[w] = (use, 1) if j == 0 else (use, self.o_item[w][1] + 1)
        for b in commits:
            b.arrive()

    def softmax(self, w, warp):
        g = use = 0
        for ordinal, (n_kt, active1) in enumerate(self.items):
            if w == 1 and not active1:
                continue
            for j in range(n_kt):
                i, u = g & 1, (g >> 1) & 1
                yield from wait(self.s_full[w][i], u, g >> 1)
                if self.s_version[w][i] != g:
                    raise Hazard(f"softmax read S of step {self.s_version[w][i]}, expected {g}")
                yield
                if self.s_version[w][i] != g:
                    raise Hazard("S overwritten while the softmax was reading it")
                self.s_free[w][i].arrive()
                yield
                yield from wait(self.p_free[w][i], u ^ 1, (g >> 1) - 1)
                if j > 0 and self.rng.random() < 0.3:
                    yield from wait(self.p_free[w][i ^ 1], ((g - 1) >> 1) & 1, (g - 1) >> 1)
                    if self.pv_done[w] != g:
                        raise Hazard("O rescaled while an earlier P.V was still in flight")
                if self.p_reads[w][i]:
                    raise Hazard("P buffer overwritten while a P.V still reads it")
                if warp == 0:
                    self.p_version[w][i] = g
                self.p_full[w][i].arrive()
                g += 1
                yield
            yield from wait(self.p_free[w][(g - 1) & 1], ((g - 1) >> 1) & 1, (g - 1) >> 1)
            if self.o_item[w] != (use, n_kt):
                raise Hazard(f"epilogue read O holding {self.o_item[w]}, expected {(use, n_kt)}")
            yield
            if self.o_item[w] != (use, n_kt):
                raise Hazard("O overwritten while it was being read out")
            if warp == 0:
                self.o_read[w] = use
            self.o_free[w].arrive()
            buf = ordinal & 1
            if self.q_reads[buf][w]:
                raise Hazard("output staged into a Q buffer an MMA still reads")
            if self.q_state[buf][w] not in (("q", ordinal), ("out", ordinal)):
                raise Hazard("output staged into a Q buffer that belongs to another item")
            self.q_state[buf][w] = ("out", ordinal)
            self.o_staged[buf if OSTAGED_PER_BUFFER else 0][w].arrive()
            use += 1
            yield

    def run(self):
        roles = [self.producer(), self.store_warp(), self.issuer(0), self.issuer(1)]
        roles += [self.softmax(w, k) for w in range(2) for k in range(8)]
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
            for w in self.rng.sample([0, 1], 2):
                if self.mma_fifo[w] and self.rng.random() < 0.5:
                    self.mma_complete(w)
                    progressed = True
            k = self.rng.choice(live)
            before = self.snapshot()
            try:
                next(roles[k])
            except StopIteration:
                live.remove(k)
                progressed = True
            progressed = progressed or self.snapshot() != before
            idle = 0 if progressed or self.async_events or any(self.mma_fifo) else idle + 1
            if idle > 4000:
                raise Hazard(f"deadlock: roles {live} are blocked with nothing in flight")
        want = [(n, w) for n, (_, a1) in enumerate(self.items) for w in range(2) if w == 0 or a1]
        if sorted(self.outputs) != want:
            raise Hazard("not every (item, group) tile was stored exactly once")

    def snapshot(self):
        rows = self.q_full + self.q_empty + self.s_full + self.s_free + self.p_full + self.p_free
        rows = rows + self.o_staged
        bars = [b for row in rows for b in row] + self.kv_full + self.kv_empty + self.o_free
        return tuple((b.phase, b.pending) for b in bars) + (len(self.outputs),)


def check(runs=200, seed=0, per_buffer=False):
    global OSTAGED_PER_BUFFER
    OSTAGED_PER_BUFFER = per_buffer
    rng = random.Random(seed)
    for _ in range(runs):
        items = [(rng.randint(1, 13), rng.random() < 0.7) for _ in range(rng.randint(1, 7))]
        Sim(items, rng).run()
    return runs


if __name__ == "__main__":
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    per_buffer = "--per-buffer" in sys.argv
    print("ok:", check(runs, seed, per_buffer), "random schedules, no deadlock, no parity aliasing, no data hazard",
          "(o_staged per Q buffer)" if per_buffer else "(o_staged per group: the default build)")
