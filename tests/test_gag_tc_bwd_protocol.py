"""CPU model of the mbarrier protocol of eegan_b200/csrc/gag_tc_bwd.cu (the one-pass tcgen05 backward of GlobalAttentionGeneral).

The kernel is six roles — TMA producer, two MMA issuers, converters, pixel warps, p / output warps — that talk through ~55
mbarriers, a ring of operand slots, two p-panel buffers, one ds-panel buffer and double-buffered dP / dX accumulators.  Nothing
of that can be unit-tested on a CPU, but its PROTOCOL can: this file restates every role's loop (same barriers, same parities,
same order; keep it in step with the kernel) as a coroutine over a model of mbarrier phase parity, runs the roles under many
random interleavings with asynchronous MMA completion, and checks
  * no deadlock (a parity wait that a lapped barrier can never satisfy shows up as one: the first version of the kernel
    had ONE barrier pair for the two p-panel buffers and hung exactly like that on the GPU),
  * every consumer sees the tile it expects (ring slots, p panels, ds panels, dP, dX),
  * nobody overwrites a buffer that still has a reader (outstanding MMAs included),
  * every accumulation group is flushed before the next one starts.
It is a model, not the kernel: it holds the design, the GPU tests hold the code."""
import random

import pytest


class MBar:
    """mbarrier: `count` arrivals complete a phase; try_wait.parity(p) is true once the phase with parity p has completed,
    i.e. while the current (incomplete) phase has the other parity.  A waiter lapped by two completions waits forever."""

    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects"
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def test(self, parity):
        return (self.phase & 1) != parity


class Issuer:
    """In-order completion of one thread's MMAs: a commit arrives on its barriers when everything issued before it is done."""

    def __init__(self):
        self.queue = []  # ("mma", reads) | ("commit", bars)

    def mma(self, *reads):
        self.queue.append(("mma", reads))

    def commit(self, *bars):
        self.queue.append(("commit", bars))

    def outstanding(self, what):
        return any(kind == "mma" and what in item for kind, item in self.queue)

    def progress(self):
        """complete the oldest entry"""
        kind, item = self.queue.pop(0)
        if kind == "commit":
            for b in item:
                b.arrive()


class Model:
    own_conv = True  # one "converted" barrier array per issuer (the kernel); False: one shared array, the other's units skipped

    def __init__(self, my_tiles, NU, NS, NGR, seed):
        self.n, self.NU, self.NS, self.NGR = my_tiles, NU, NS, NGR
        self.rng = random.Random(seed)
        B = MBar
        self.full = [B(1) for _ in range(NS)]
        self.conv = {"do": [B(4) for _ in range(NS)], "x": [B(4) for _ in range(NS)]}
        if not self.own_conv:
            self.conv["x"] = self.conv["do"]
        self.sfree = [B(1) for _ in range(NS)]
        self.p_full = [B(4), B(4)]
        self.p_empty = [B(1), B(1)]
        self.ds_full, self.ds_empty = B(4), B(1)
        self.acc_full, self.acc_empty = B(2), B(4)
        self.dp_full, self.dp_empty = [B(1), B(1)], [B(4), B(4)]
        self.dx_full, self.dx_empty = [B(1), B(1)], [B(4), B(4)]
        self.A, self.Bi = Issuer(), Issuer()
        # contents
        self.slot = [None] * NS          # (kind, tile, u, "raw" | "conv")
        self.pbuf = [None, None]         # tile whose p panels the buffer holds
        self.ds = None                   # tile whose ds panels the buffer holds
        self.dp = [None, None]           # tile accumulated (complete once dp_full was waited for)
        self.dx = [None, None]
        self.pixel_read_p = -1           # last tile whose p panels the pixel warps re-read
        self.dp_read = -1                # last tile whose dP the pixel warps read
        self.dx_read = -1
        self.flushed = -1                # last accumulation group written out
        self.acc_group = -1              # group the dV / dK accumulators currently hold

    def gr0(self, g):
        return (g * self.n) // self.NGR

    def units(self):
        """the unit sequence all ring users walk: (kind, tile, u)"""
        for g in range(self.NGR):
            t0, t1 = self.gr0(g), self.gr0(g + 1)
            for i in range(t0, t1 + 1):
                if i < t1:
                    for u in range(self.NU):
                        yield ("do", i, u)
                if i > t0:
                    for u in range(self.NU):
                        yield ("x", i - 1, u)

    # ---- roles: generators yielding ("wait", bar, parity) -------------------------------------------------------
    def producer(self):
        s, ph = 0, 0
        for kind, tile, u in self.units():
            yield ("wait", self.sfree[s], ph ^ 1)
            assert not self.A.outstanding(("slot", s)) and not self.Bi.outstanding(("slot", s)), "TMA into a slot an MMA still reads"
            self.slot[s] = (kind, tile, u, "raw")
            self.full[s].arrive()  # (the TMA's completion, modelled as immediate)
            s += 1
            if s == self.NS:
                s, ph = 0, ph ^ 1

    def converters(self):
        s, ph = 0, 0
        for kind, tile, u in self.units():
            yield ("wait", self.full[s], ph)
            assert self.slot[s] == (kind, tile, u, "raw"), (self.slot[s], kind, tile, u)
            self.slot[s] = (kind, tile, u, "conv")
            for _ in range(4):
                self.conv[kind][s].arrive()
            s += 1
            if s == self.NS:
                s, ph = 0, ph ^ 1

    def issuer(self, side_a):
        me = self.A if side_a else self.Bi
        mine = self.conv["do" if side_a else "x"]
        s, ph = 0, 0
        used = 0  # bit s: parity of this issuer's uses of slot s so far (own_conv: the phase of ITS barrier of the slot)

        def adv(k):
            nonlocal s, ph
            s += k
            if s >= self.NS:
                s -= self.NS
                ph ^= 1

        def my_parity():
            nonlocal used
            if not self.own_conv:
                return ph
            par = (used >> s) & 1
            used ^= 1 << s
            return par

        for g in range(self.NGR):
            t0, t1 = self.gr0(g), self.gr0(g + 1)
            if g > 0:
                yield ("wait", self.acc_empty, (g - 1) & 1)
                assert self.flushed == g - 1
            for i in range(t0, t1 + 1):
                if i < t1:
                    if side_a:
                        a, k = i & 1, i >> 1
                        yield ("wait", self.p_full[a], k & 1)
                        yield ("wait", self.dp_empty[a], (k & 1) ^ 1)
                        assert self.pbuf[a] == i, ("p panels", self.pbuf[a], i)
                        assert self.dp_read >= i - 2, "dP buffer overwritten before the pixel warps read it"
                        for u in range(self.NU):
                            yield ("wait", mine[s], my_parity())
                            assert self.slot[s] == ("do", i, u, "conv"), (self.slot[s], i, u)
                            me.mma(("slot", s), ("p", a))
                            self.acc_group = max(self.acc_group, g)
                            me.commit(self.sfree[s])
                            adv(1)
                        self.dp[a] = i
                        me.commit(self.dp_full[a], self.p_empty[a])
                    else:
                        adv(self.NU)  # the other issuer's units
                if i > t0:
                    if not side_a:
                        j = i - 1
                        a, k = j & 1, j >> 1
                        yield ("wait", self.ds_full, j & 1)
                        yield ("wait", self.dx_empty[a], (k & 1) ^ 1)
                        assert self.ds == j, ("ds panels", self.ds, j)
                        assert self.dx_read >= j - 2, "dX buffer overwritten before the output warps read it"
                        me.mma(("ds",))
                        self.dx[a] = j
                        me.commit(self.dx_full[a])
                        for u in range(self.NU):
                            yield ("wait", mine[s], my_parity())
                            assert self.slot[s] == ("x", j, u, "conv"), (self.slot[s], j, u)
                            me.mma(("slot", s), ("ds",))
                            me.commit(self.sfree[s])
                            adv(1)
                        me.commit(self.ds_empty)
                    else:
                        adv(self.NU)
            me.commit(self.acc_full)

    def pixel(self):
        g, g_end = 0, self.gr0(1)
        for i in range(self.n):
            a, k = i & 1, i >> 1
            yield ("wait", self.dp_full[a], k & 1)
            assert self.dp[a] == i, ("dP", self.dp[a], i)
            self.dp_read = i
            for _ in range(4):
                self.dp_empty[a].arrive()
            assert self.pbuf[a] == i, ("p panels re-read by the pixel warps", self.pbuf[a], i)
            self.pixel_read_p = i
            yield ("wait", self.ds_empty, (i & 1) ^ 1)
            assert not self.Bi.outstanding(("ds",)), "ds panels overwritten under an MMA"
            self.ds = i
            for _ in range(4):
                self.ds_full.arrive()
            if i + 1 == g_end:
                yield ("wait", self.acc_full, g & 1)
                self.flushed = g
                for _ in range(4):
                    self.acc_empty.arrive()
                g += 1
                g_end = self.gr0(g + 1)

    def p_out(self):
        self.pbuf[0] = 0
        for _ in range(4):
            self.p_full[0].arrive()
        for i in range(self.n):
            a, k = i & 1, i >> 1
            if i + 1 < self.n:
                yield ("wait", self.p_empty[a ^ 1], (((i + 1) >> 1) & 1) ^ 1)
                assert not self.A.outstanding(("p", a ^ 1)), "p panels overwritten under an MMA"
                assert self.pixel_read_p >= i - 1, "p panels overwritten before the pixel warps re-read them"
                self.pbuf[a ^ 1] = i + 1
                for _ in range(4):
                    self.p_full[a ^ 1].arrive()
            yield ("wait", self.dx_full[a], k & 1)
            assert self.dx[a] == i, ("dX", self.dx[a], i)
            self.dx_read = i
            for _ in range(4):
                self.dx_empty[a].arrive()

    # ---- scheduler -------------------------------------------------------------------------------------------------
    def run(self):
        roles = {"producer": self.producer(), "conv": self.converters(), "issA": self.issuer(True), "issB": self.issuer(False),
                 "pixel": self.pixel(), "p_out": self.p_out()}
        blocked = {}
        steps = 0
        while roles:
            steps += 1
            assert steps < 2_000_000
            runnable = [r for r in roles if r not in blocked or blocked[r][0].test(blocked[r][1])]
            pend = [x for x in (self.A, self.Bi) if x.queue]
            choices = runnable + ["mma"] * len(pend)
            if not choices:
                raise AssertionError("deadlock: " + ", ".join("%s waits parity %d of a barrier in phase %d" % (r, p, b.phase)
                                                              for r, (b, p) in blocked.items()))
            c = self.rng.choice(choices)
            if c == "mma":
                self.rng.choice(pend).progress()
                continue
            blocked.pop(c, None)
            try:
                ev = next(roles[c])
                assert ev[0] == "wait"
                if not ev[1].test(ev[2]):
                    blocked[c] = (ev[1], ev[2])
                else:
                    blocked[c] = (ev[1], ev[2])  # re-tested (true) the next time the role is picked
            except StopIteration:
                del roles[c]
        while self.A.queue or self.Bi.queue:
            for x in (self.A, self.Bi):
                if x.queue:
                    x.progress()
        assert self.flushed == self.NGR - 1 and self.dx_read == self.n - 1 and self.dp_read == self.n - 1


@pytest.mark.parametrize("my_tiles,NU,NS,NGR", [(1, 1, 10, 1), (2, 2, 4, 1), (11, 2, 4, 1), (43, 1, 6, 2), (171, 1, 10, 6), (7, 2, 2, 3),
                                                 (5, 1, 2, 5), (64, 2, 5, 2)])
def test_protocol_has_no_deadlock_and_no_hazard(my_tiles, NU, NS, NGR):
    for seed in range(12 if my_tiles > 60 else 40):
        Model(my_tiles, NU, NS, NGR, seed).run()


def test_one_converted_barrier_per_slot_for_both_issuers_is_unsafe_on_a_short_ring():
    """Each issuer has its OWN "converted" barrier per slot.  With one shared barrier per slot an issuer sees only every other
    completion (the other issuer's units pass it by), and a parity wait tells completion n from n - 1 only: on a two-slot ring
    issuer A then waits for completion n while n - 1 (a unit of issuer B) has not happened, the test passes, and A contracts the
    PREVIOUS unit.  (Waiting for the other's units as well is no cure: an observer does not gate the slot's re-use, gets lapped
    and hangs — the model found that before a GPU did.)  At the 4-10 slots the launcher produces other dependencies keep the
    issuers apart, which the model also shows; the kernel does not rely on it."""

    class Shared(Model):
        own_conv = False

    bad = 0
    for seed in range(40):
        try:
            Shared(7, 2, 2, 3, seed).run()
        except AssertionError:
            bad += 1
    assert bad > 0
    for cfg in ((11, 2, 4, 1), (43, 1, 6, 2), (64, 1, 10, 2)):
        for seed in range(25):
            Shared(*cfg, seed).run()


def test_one_barrier_pair_for_two_p_buffers_deadlocks():
    """The first version of the kernel: p_full / p_empty shared by both p buffers.  A consumer can be lapped by two completions
    and then waits for a parity that never comes — the model finds the interleaving, the GPU needed ten minutes of a hung run."""

    class OnePair(Model):
        own_conv = False

        def __init__(self, *a):
            super().__init__(*a)
            self.p_full[1] = self.p_full[0]
            self.p_empty[1] = self.p_empty[0]

        def issuer(self, side_a):  # parities by tile index, as a single pair is used
            if not side_a:
                yield from super().issuer(False)
                return
            me, s, ph = self.A, 0, 0
            for i in range(self.n):
                a, k = i & 1, i >> 1
                yield ("wait", self.p_full[0], i & 1)
                yield ("wait", self.dp_empty[a], (k & 1) ^ 1)
                for u in range(self.NU):
                    yield ("wait", self.conv["do"][s], ph)
                    me.mma(("slot", s), ("p", a))
                    me.commit(self.sfree[s])
                    s += 1
                    if s == self.NS:
                        s, ph = 0, ph ^ 1
                self.dp[a] = i
                me.commit(self.dp_full[a], self.p_empty[0])
                # the x units of tile i - 1 belong to the other issuer
                if i >= 1:
                    s += self.NU
                    if s >= self.NS:
                        s, ph = s - self.NS, ph ^ 1
            me.commit(self.acc_full)

        def units(self):
            for i in range(self.n + 1):
                if i < self.n:
                    for u in range(self.NU):
                        yield ("do", i, u)
                if i >= 1:
                    for u in range(self.NU):
                        yield ("x", i - 1, u)

        def p_out(self):
            self.pbuf[0] = 0
            for _ in range(4):
                self.p_full[0].arrive()
            for i in range(self.n):
                a, k = i & 1, i >> 1
                if i + 1 < self.n:
                    yield ("wait", self.p_empty[0], (i & 1) ^ 1)
                    self.pbuf[a ^ 1] = i + 1
                    for _ in range(4):
                        self.p_full[0].arrive()
                yield ("wait", self.dx_full[a], k & 1)
                self.dx_read = i
                for _ in range(4):
                    self.dx_empty[a].arrive()

    hung = 0
    for seed in range(60):
        try:
            OnePair(12, 1, 6, 1, seed).run()
        except AssertionError as e:
            if "deadlock" in str(e) or "p panels" in str(e):
                hung += 1
    assert hung > 0
