"""Randomised model check of the synchronisation protocol of csrc/attention_tc2.cu (no GPU needed).

The kernel's five roles (TMA producer, two MMA issuers, two softmax groups, epilogue warpgroup) are transcribed below as
Python generators with the kernel's own control flow: same loops, same barrier, same parity expression at every wait,
same arrive / commit at every release.  mbarriers, TMA completions and tcgen05.commit arrivals are modelled with their
hardware semantics:
  * a wait on parity p passes when the barrier's current phase has parity != p (so a waiter that is two phases off
    passes WRONGLY - the aliasing race described in DESIGN.md section 4 - and the model reports it, because every wait
    also states which completion it means);
  * TMA loads and MMA blocks complete asynchronously after random delays; the MMAs of one issuer complete in issue
    order (tcgen05 pipeline), the two issuers' streams interleave arbitrarily.
A random scheduler interleaves the roles.  Checked on every run:
  * no deadlock;
  * every wait passes on exactly the completion it was written for;
  * data hazards: a K/V stage or Q buffer is refilled only after every MMA reading it has completed, and consumed only
    with the fill it expects; S is read by the softmax only when that step's Q.K^T has completed (scores are double-
    buffered per group: S runs up to two tiles ahead of P.V), P.V issues only on that step's P and executes before the
    S that reuses its buffer, the occasional O rescale (which waits on a barrier the other steps never look at) sees
    every earlier P.V completed; O / the 1/l slot are reused only after the epilogue has drained the previous item, and
    the epilogue reads a complete item.
Usage:  python tools/model_check_mha2.py [runs] [seed]      (also run by tests/test_host.py with a small budget)
"""
import random
import sys

KV_STAGES = 4


class Hazard(AssertionError):
    pass


class MBar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending < 0:
            raise Hazard(f"{self.name}: more arrivals than the barrier was initialised for")
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def passes(self, parity):
        return (self.phase & 1) != parity


def wait(bar, parity, completion):
    """Generator: blocks like mbarrier.try_wait.parity; `completion` = index (0-based) of the phase the code means."""
    while not bar.passes(parity):
        yield
    if completion >= 0 and bar.phase != completion + 1:
        raise Hazard(f"{bar.name}: wait(parity {parity}) meant completion {completion} but passed at phase {bar.phase}")


class Sim:
    def __init__(self, items, rng):
        self.items = items                     # list of (n_kt, active1)
        self.rng = rng
        self.q_full = [[MBar(f"q_full[{b}][{w}]", 1) for w in range(2)] for b in range(2)]
        self.q_empty = [[MBar(f"q_empty[{b}][{w}]", 1) for w in range(2)] for b in range(2)]
        self.kv_full = [MBar(f"kv_full[{s}]", 1) for s in range(KV_STAGES)]
        self.kv_empty = [MBar(f"kv_empty[{s}]", 2) for s in range(KV_STAGES)]
        self.s_full = [[MBar(f"s_full[{w}][{b}]", 1) for b in range(2)] for w in range(2)]
        self.p_full = [[MBar(f"p_full[{w}][{b}]", 4) for b in range(2)] for w in range(2)]
        self.pv_bar = [[MBar(f"pv_done[{w}][{b}]", 1) for b in range(2)] for w in range(2)]
        self.o_full = [MBar(f"o_full[{w}]", 1) for w in range(2)]
        self.l_full = [MBar(f"l_full[{w}]", 4) for w in range(2)]
        self.o_free = [MBar(f"o_free[{w}]", 4) for w in range(2)]
        # asynchronous completions: list of [delay, action]; the MMA streams are FIFOs (in-order per issuer)
        self.async_events = []
        self.mma_fifo = [[], []]
        # data state for the hazard checks
        self.stage_fill = [None] * KV_STAGES           # flat tile index held by the stage (None while loading / empty)
        self.stage_reads = [0] * KV_STAGES             # MMA blocks in flight that read the stage
        self.q_fill = [[None, None], [None, None]]     # item ordinal held by Q buffer (buf, w)
        self.q_reads = [[0, 0], [0, 0]]
        self.s_version = [[-1, -1], [-1, -1]]          # step whose scores are complete in S[w][buffer]
        self.p_version = [[-1, -1], [-1, -1]]          # step whose probabilities are in P[w][buffer]
        self.s_reading = [[0, 0], [0, 0]]              # softmax warps currently between "S read" and "P written"
        self.pv_done = [0, 0]                          # P.V blocks completed per group
        self.o_use = [None, None]                      # (use, accumulated P.V blocks) of the group's O
        self.o_drained = [-1, -1]                      # last use the epilogue has read out
        self.l_slot = [None, None]
        self.outputs = []

    # ---------------- roles ----------------
    def producer(self):
        stage, kv_phase, fills = 0, 0, 0
        q_par = 0
        q_count = [[0, 0], [0, 0]]
        tile = 0
        for n_done, (n_kt, active1) in enumerate(self.items):
            buf = n_done & 1
            for w in range(2):
                if w == 1 and not active1:
                    continue
                bit = buf * 2 + w
                qph = (q_par >> bit) & 1
                q_par ^= 1 << bit
                yield from wait(self.q_empty[buf][w], qph ^ 1, q_count[buf][w] - 1)
                if self.q_reads[buf][w]:
                    raise Hazard("Q buffer refilled while an MMA still reads it")
                q_count[buf][w] += 1
                self.q_fill[buf][w] = None
                self.later(lambda b=buf, ww=w, n=n_done: self.q_landed(b, ww, n))
            for j in range(n_kt):
                yield from wait(self.kv_empty[stage], kv_phase ^ 1, fills // KV_STAGES - 1)
                if self.stage_reads[stage]:
                    raise Hazard("K/V stage refilled while an MMA still reads it")
                self.stage_fill[stage] = None
                self.later(lambda s=stage, t=tile: self.kv_landed(s, t))
                tile += 1
                fills += 1
                stage += 1
                if stage == KV_STAGES:
                    stage, kv_phase = 0, kv_phase ^ 1
                yield

    def q_landed(self, buf, w, n):
        self.q_fill[buf][w] = n
        self.q_full[buf][w].arrive()

    def kv_landed(self, s, t):
        self.stage_fill[s] = t
        self.kv_full[s].arrive()

    def later(self, action):
        # mostly short, sometimes very long (a congested memory system re-orders completions across many steps)
        delay = self.rng.randint(0, 6) if self.rng.random() < 0.85 else self.rng.randint(50, 400)
        self.async_events.append([delay, action])

    def issuer(self, w):
        items = self.items
        n_items = len(items)

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
        uses = 0
        while pc.valid:
            if w == 1 and pc.virt:
                for _ in range(pc.n_kt):
                    yield from wait(self.kv_full[pc.stage], pc.phase, pc.tile // KV_STAGES)
                    if self.stage_fill[pc.stage] != pc.tile:
                        raise Hazard("virtual walk observed the wrong fill")
                    self.kv_empty[pc.stage].arrive()
                    advance(pc)
                    yield
                sc = copy(pc)
                continue
            if (sc.valid and not (w == 1 and sc.virt) and g_s < g_p + 2
                    and (g_s == g_p or self.kv_full[sc.stage].passes(sc.phase))):
                buf, sb = sc.n_done & 1, g_s & 1
                yield from wait(self.kv_full[sc.stage], sc.phase, sc.tile // KV_STAGES)
                if sc.j == 0:
                    yield from wait(self.q_full[buf][w], q_fill[buf] & 1, q_fill[buf])
                if self.stage_fill[sc.stage] != sc.tile:
                    raise Hazard(f"S: stage holds tile {self.stage_fill[sc.stage]}, expected {sc.tile}")
                if self.q_fill[buf][w] != sc.n_done:
                    raise Hazard("S: Q buffer holds another item")
                last = sc.j == sc.n_kt - 1
                self.stage_reads[sc.stage] += 1
                self.q_reads[buf][w] += 1
                commits = [self.s_full[w][sb]] + ([self.q_empty[buf][w]] if last else [])
                self.mma_fifo[w].append(("S", g_s, sc.stage, (buf, w), commits))
                if last:
                    q_fill[buf] += 1
                g_s += 1
                advance(sc)
                yield
                continue
            sb = g_p & 1
            yield from wait(self.p_full[w][sb], (g_p >> 1) & 1, g_p >> 1)
            if pc.j == 0:
                yield from wait(self.o_free[w], (uses & 1) ^ 1, uses - 1)
                if uses >= 1 and self.o_drained[w] != uses - 1:
                    raise Hazard("first P.V of an item issued into an O the epilogue has not drained")
            if self.p_version[w][sb] != g_p:
                raise Hazard(f"P.V({g_p}) issued on P of step {self.p_version[w][sb]}")
            if self.stage_fill[pc.stage] != pc.tile:
                raise Hazard("P.V: stage holds another tile")
            last = pc.j == pc.n_kt - 1
            self.stage_reads[pc.stage] += 1
            commits = [self.kv_empty[pc.stage], self.pv_bar[w][sb]] + ([self.o_full[w]] if last else [])
            self.mma_fifo[w].append(("PV", g_p, pc.stage, (uses, pc.j, pc.n_kt), commits))
            if last:
                uses += 1
            g_p += 1
            advance(pc)
            yield

    def mma_complete(self, w):
        kind, g, stage, info, commits = self.mma_fifo[w].pop(0)
        self.stage_reads[stage] -= 1
        sb = g & 1
        if kind == "S":
            buf, ww = info
            self.q_reads[buf][ww] -= 1
            if self.s_reading[w][sb]:
                raise Hazard("S overwrote a score buffer a softmax warp was still working on")
            self.s_version[w][sb] = g
            self.p_version[w][sb] = -1         # S(g) overwrites the columns P(g - 2) lived in
        else:
            use, j, n_kt = info
            if self.p_version[w][sb] != g:
                raise Hazard("P.V executed after its P was overwritten")
            self.pv_done[w] += 1
            self.o_use[w] = (use, 1) if j == 0 else (use, self.o_use[w][1] + 1)
        for b in commits:
            b.arrive()

    def softmax(self, w, warp):
        """One of the four warps of softmax group w (the barrier counts are per warp)."""
        g = uses = 0
        for n_kt, active1 in self.items:
            if w == 1 and not active1:
                continue
            for j in range(n_kt):
                sb = g & 1
                yield from wait(self.s_full[w][sb], (g >> 1) & 1, g >> 1)
                if self.s_version[w][sb] != g:
                    raise Hazard(f"softmax read S of step {self.s_version[w][sb]}, expected {g}")
                self.s_reading[w][sb] += 1
                yield
                if j > 0 and self.rng.random() < 0.3:          # lazy rescale of O: the occasional waiter of pv_done
                    yield from wait(self.pv_bar[w][sb ^ 1], ((g - 1) >> 1) & 1, (g - 1) >> 1)
                    if self.pv_done[w] != g:
                        raise Hazard("O rescaled while an earlier P.V was still in flight")
                    yield
                if self.s_version[w][sb] != g:
                    raise Hazard("S overwritten while the softmax was still reading it")
                self.s_reading[w][sb] -= 1
                if warp == 0:
                    self.p_version[w][sb] = g
                self.p_full[w][sb].arrive()
                g += 1
                yield
            yield from wait(self.o_free[w], (uses & 1) ^ 1, uses - 1)
            if uses >= 1 and self.o_drained[w] != uses - 1:
                raise Hazard("1/l slot overwritten before the epilogue read it")
            if warp == 0:
                self.l_slot[w] = uses
            self.l_full[w].arrive()
            uses += 1
            yield

    def epilogue(self, warp):
        uses = [0, 0]
        for n, (n_kt, active1) in enumerate(self.items):
            for w in range(2):
                if w == 1 and not active1:
                    continue
                use = uses[w]
                ph = uses[w] & 1
                uses[w] += 1
                yield from wait(self.l_full[w], ph, use)
                if self.l_slot[w] != use:
                    raise Hazard("epilogue read the 1/l of another item")
                yield from wait(self.o_full[w], ph, use)
                if self.o_use[w] != (use, n_kt):
                    raise Hazard(f"epilogue read O holding {self.o_use[w]}, expected {(use, n_kt)}")
                yield
                if self.o_use[w] != (use, n_kt):
                    raise Hazard("O overwritten while the epilogue was reading it")
                if warp == 0:
                    self.o_drained[w] = use
                    self.outputs.append((n, w))
                self.o_free[w].arrive()
                yield

    # ---------------- scheduler ----------------
    def run(self):
        roles = [self.producer(), self.issuer(0), self.issuer(1)]
        roles += [self.softmax(w, k) for w in range(2) for k in range(4)]
        roles += [self.epilogue(k) for k in range(4)]
        live = list(range(len(roles)))
        idle = 0
        while live:
            progressed = False
            # asynchronous machinery
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
            if self.snapshot() != before:
                progressed = True
            idle = 0 if progressed or self.async_events or any(self.mma_fifo) else idle + 1
            if idle > 2000:
                raise Hazard(f"deadlock: roles {live} are blocked with nothing in flight")
        want = [(n, w) for n, (_, a1) in enumerate(self.items) for w in range(2) if w == 0 or a1]
        if sorted(self.outputs) != want:
            raise Hazard("not every (item, group) tile was written exactly once")

    def snapshot(self):
        bars = [b for row in (self.q_full + self.q_empty + self.s_full + self.p_full + self.pv_bar) for b in row]
        bars += self.kv_full + self.kv_empty + self.o_full + self.l_full + self.o_free
        return tuple((b.phase, b.pending) for b in bars) + (len(self.outputs),)


def check(runs=200, seed=0, kv_stages=4):
    global KV_STAGES
    KV_STAGES = kv_stages
    rng = random.Random(seed)
    for r in range(runs):
        n_items = rng.randint(1, 7)
        items = [(rng.randint(1, 6), rng.random() < 0.7) for _ in range(n_items)]
        Sim(items, rng).run()
    return runs


if __name__ == "__main__":
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    stages = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    print("ok:", check(runs, seed, stages), f"random schedules ({stages} K/V stages), no deadlock, no parity aliasing, "
          "no data hazard")
