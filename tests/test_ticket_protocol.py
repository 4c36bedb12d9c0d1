"""Model check of the SM-affine hand-out protocol of k_traceCompound (compound-ray_b200/csrc/cr_kernels.cu, "SM-affine hand-out").

The device code cannot run here, so this is a step-by-step restatement of ITS protocol -- per-SM ticket counters, the block table
filled by whoever draws the first ticket of a block, tickets drawn two ahead, "leave when both held units lie beyond the end, draw
only while the current unit is real" -- driven by a random scheduler that interleaves the warps' atomic steps in every order,
including the one the termination argument has to survive: two warps reaching the global counter out of order, so that block
numbers do not grow along an SM's sequence.  Checked: every unit is traced exactly once, every warp terminates, and no SM's block
index reaches the table size the host allocates (cr_renderer.cu attachWorkCounter: blocks + gridWarps / 16 + 2)."""
import random

import pytest

INVALID = 0xFFFFFFFF


class Warp:
    """One warp of the trace kernel as a little state machine; every yield is one atomic step the scheduler may interleave."""

    def __init__(self, sim, slot):
        self.sim, self.slot, self.done = sim, slot, False
        self.gen = self.run()

    def take_ticket(self):
        s = self.sim
        k = s.seq[self.slot]; s.seq[self.slot] += 1          # atomicAdd(smSeq + slot, 1)
        yield
        if k & 31 == 0:
            nb = s.counter; s.counter += 1                   # atomicAdd(workCounter, 1) -- possibly long after the ticket
            yield
            assert (k >> 5) < s.cap, "block table overrun"
            s.tab[(self.slot, k >> 5)] = nb                  # published with the launch's epoch
            yield
        return k

    def unit_of(self, k):
        s = self.sim
        if k == INVALID:
            return INVALID
        while (self.slot, k >> 5) not in s.tab:              # spin on the epoch tag
            yield
        nb = s.tab[(self.slot, k >> 5)]
        return (nb << 5) + (k & 31) if nb < s.n_blocks else INVALID

    def run(self):
        s = self.sim
        t0 = yield from self.take_ticket()
        t1 = yield from self.take_ticket()
        chunk = yield from self.unit_of(t0)
        nxt = t1
        while True:
            nxt = yield from self.unit_of(nxt)
            if chunk >= s.n_units and nxt >= s.n_units:
                break
            after = INVALID
            if chunk < s.n_units:
                after = yield from self.take_ticket()
                s.traced[chunk] += 1
                for _ in range(s.rng.randrange(0, 4)):       # tracing takes a while
                    yield
            chunk, nxt = nxt, after
        self.done = True


class Sim:
    def __init__(self, n_units, n_slots, warps_per_slot, seed):
        self.rng = random.Random(seed)
        self.n_units = n_units
        self.n_blocks = (n_units + 31) // 32
        grid_warps = n_slots * warps_per_slot
        self.cap = self.n_blocks + grid_warps // 16 + 2     # attachWorkCounter's table size per SM
        self.seq = [0] * n_slots
        self.counter = 0
        self.tab = {}
        self.traced = [0] * n_units
        self.warps = [Warp(self, s) for s in range(n_slots) for _ in range(warps_per_slot)]

    def run(self, bias):
        live = list(self.warps)
        steps, frozen, thaw = 0, None, 0
        while live:
            # biased scheduler: now and then one warp is frozen for a long stretch -- e.g. between drawing the first ticket of a
            # block and reaching the global counter, so that a later block of its SM gets the smaller number
            if bias and frozen is None and self.rng.random() < 0.01:
                frozen, thaw = self.rng.choice(live), steps + self.rng.randrange(50, 2000)
            if frozen is not None and (steps >= thaw or frozen not in live or len(live) == 1):
                frozen = None
            w = self.rng.choice(live)
            if w is frozen:
                steps += 1
                continue
            try:
                next(w.gen)
            except StopIteration:
                live.remove(w)
            steps += 1
            assert steps < 5_000_000, "no termination"


@pytest.mark.parametrize("n_units,n_slots,warps", [(1, 3, 2), (31, 2, 4), (32, 4, 1), (33, 4, 3), (79, 3, 2), (640, 5, 8), (1000, 7, 4),
                                                   (4096, 16, 8), (97, 1, 32), (5, 8, 8)])
def test_every_unit_exactly_once_under_random_interleavings(n_units, n_slots, warps):
    for seed in range(12):
        sim = Sim(n_units, n_slots, warps, seed)
        sim.run(bias=seed % 2 == 1)
        assert all(w.done for w in sim.warps)
        assert sim.traced == [1] * n_units, (n_units, n_slots, warps, seed)
        assert max(sim.seq) <= n_units + 2 * len(sim.warps) + 32          # tickets per SM: units + 2 per warp (+ the ragged block)


def _out_of_order_run(warp_cls):
    """One SM, 34 warps, one block of 32 units.  Warp 0 draws the first ticket of the SM's block 0 and is held before it reaches
    the global counter; the others draw 66 more tickets, so the first ticket of the SM's block 1 is drawn -- and served by the
    global counter -- first: block 1 of the SM is the real block, block 0 gets a number beyond the end."""
    sim = Sim(32, 1, 34, seed=1)
    sim.warps = [warp_cls(sim, 0) for _ in sim.warps]
    next(sim.warps[0].gen)                                  # ticket 0 drawn; the counter fetch is its next step
    for w in sim.warps[1:]:
        for _ in range(8):
            next(w.gen)                                     # both tickets (and the counter, for ticket 32), then spinning
    assert sim.tab.get((0, 1)) == 0                         # the SM's SECOND block is block number 0
    sim.run(bias=False)
    return sim


def test_out_of_order_block_numbers_lose_no_unit():
    sim = _out_of_order_run(Warp)
    assert sim.tab[(0, 0)] >= sim.n_blocks                  # the SM's first block lies beyond the end ...
    assert sim.traced == [1] * 32 and all(w.done for w in sim.warps)   # ... and still every unit is traced once


def test_the_model_notices_a_wrong_leaving_rule():
    """Mutation check of the model itself: a warp that leaves as soon as its CURRENT unit lies beyond the end (instead of both
    units it holds) abandons the real unit it already holds a ticket for -- the model must see the lost unit."""
    class Hasty(Warp):
        def run(self):
            s = self.sim
            t0 = yield from self.take_ticket()
            t1 = yield from self.take_ticket()
            chunk = yield from self.unit_of(t0)
            nxt = t1
            while True:
                nxt = yield from self.unit_of(nxt)
                if chunk >= s.n_units:
                    break
                after = yield from self.take_ticket()
                s.traced[chunk] += 1
                chunk, nxt = nxt, after
            self.done = True
    sim = _out_of_order_run(Hasty)
    assert sim.traced.count(0) > 0
