"""CPU model check of the shared-memory ring protocol of the fused orthogonalisation kernel
(arnoldimethod.jl_b200/csrc/kernels_cgs_sweep.cuh).

The kernel keeps ONE ring of `S` stages alive through three phases that walk a CTA's `ntl` tiles
forward (P1), backward (P2) and forward again (P3): local tile l always lives in slot l % S, the
last min(S, ntl) tiles of a phase are not released, and the next phase consumes them without
waiting for a refill.  Producer and consumers only exchange mbarrier phase parities (one bit per
slot on either side).  This test restates exactly that bookkeeping - the `produce` / `acquire` /
`release` lambdas and the loop bounds of the three phases - as two cooperative coroutines over
simulated mbarriers, runs them under randomly interleaved schedules, and checks for every
(ntl, S):

  * no deadlock (both sides finish), with and without the gated third phase;
  * a consumer never reads a slot that holds a different tile, or stale contents of the column
    being orthogonalised (P3 must see the v written in P2);
  * the producer never overwrites a slot the consumers have not released;
  * every tile is streamed from memory exactly once per phase except the resident ones, i.e. the
    HBM traffic saved per phase change is min(S, ntl) tiles.
"""

import random

import pytest


class MBarrier:
    """Phase-counting model of an mbarrier: `arrivals` per phase, test_wait by parity."""

    def __init__(self, count):
        self.count, self.pending, self.phase = count, 0, 0

    def arrive(self):
        self.pending += 1
        if self.pending == self.count:
            self.pending, self.phase = 0, self.phase + 1

    def done(self, parity):
        # mbarrier.try_wait.parity: true once the phase with this parity has completed, i.e. the
        # barrier's current (incomplete) phase has the other parity
        return (self.phase & 1) != parity


def simulate(ntl, S, second, seed, consumers=3, broken=False):
    rng = random.Random(seed)
    keep = min(S, ntl)
    full = [MBarrier(1) for _ in range(S)]
    empty = [MBarrier(consumers) for _ in range(S)]
    slot_tile = [None] * S        # which tile's data sits in the slot
    slot_version = [None] * S     # version of column v the slot was filled with / updated to
    slot_free = [True] * S        # released by every consumer since the last fill
    v_version = {l: 0 for l in range(ntl)}  # global memory: version of the rows of v in tile l
    loads = {1: 0, 2: 0, 3: 0}
    log = []

    def producer():
        pmask = 0
        phase_of_loop = [(1, range(0, ntl)), (2, range(ntl - 1 - keep, -1, -1))]
        if second:
            phase_of_loop.append((3, range(keep, ntl)))
        for ph, tiles in phase_of_loop:
            for l in tiles:
                s = l % S
                while not empty[s].done(((pmask >> s) & 1) ^ 1):
                    yield
                pmask ^= 1 << s
                assert slot_free[s], f"producer overwrites an unreleased slot (tile {l}, phase {ph})"
                # TMA load: the slot receives tile l with the CURRENT global version of v
                slot_tile[s], slot_version[s], slot_free[s] = l, v_version[l], False
                loads[ph] += 1
                full[s].arrive()
                yield

    released = [0] * S

    def consumer(cid):
        cmask = 0
        plan = [(1, list(range(ntl)), lambda l: False, lambda l: broken or l < ntl - keep),
                (2, list(range(ntl - 1, -1, -1)), lambda l: l >= ntl - keep, lambda l: l >= keep)]
        if second:
            plan.append((3, list(range(ntl)), lambda l: l < keep, lambda l: True))
        for ph, tiles, resident, rel in plan:
            for l in tiles:
                s = l % S
                if not resident(l):
                    while not full[s].done((cmask >> s) & 1):
                        yield
                    cmask ^= 1 << s
                assert slot_tile[s] == l, f"phase {ph}: slot {s} holds tile {slot_tile[s]}, wanted {l}"
                if ph == 3:  # must see v_new: loaded after P2 wrote it, or updated in place in the stage
                    assert slot_version[s] == 1, f"phase 3: tile {l} has stale v"
                else:        # P1 / P2 work on the v the mat-vec produced (another consumer may already have
                    assert slot_version[s] in ((0,) if ph == 1 else (0, 1))  # stored its rows of v_new in P2)
                yield
                if ph == 2 and cid == 0:
                    # the update writes v_new to global memory and back into the stage
                    v_version[l] = 1
                    slot_version[s] = 1
                if ph == 2:
                    # consumer_bar_sync(): nobody leaves the tile before the update is complete
                    log.append(("sync", l))
                    while sum(1 for e in log if e == ("sync", l)) < consumers:
                        yield
                if rel(l):
                    released[s] += 1
                    if released[s] == consumers:
                        released[s] = 0
                        slot_free[s] = True
                    empty[s].arrive()
                yield
            # grid barrier between the phases: every consumer of this CTA has finished the phase
            log.append(("phase_done", ph))
            while sum(1 for e in log if e == ("phase_done", ph)) < consumers:
                yield

    tasks = [producer()] + [consumer(c) for c in range(consumers)]
    alive = list(range(len(tasks)))
    steps = 0
    while alive:
        steps += 1
        assert steps < 200000, f"deadlock: ntl={ntl} S={S} second={second}"
        i = rng.choice(alive)
        try:
            next(tasks[i])
        except StopIteration:
            alive.remove(i)
    return loads


@pytest.mark.parametrize("S", [2, 3, 5, 8])
@pytest.mark.parametrize("ntl", [0, 1, 2, 3, 4, 5, 7, 8, 9, 16, 27, 53])
@pytest.mark.parametrize("second", [False, True])
def test_ring_protocol(ntl, S, second):
    keep = min(S, ntl)
    for seed in range(6):
        loads = simulate(ntl, S, second, seed)
        assert loads[1] == ntl
        assert loads[2] == ntl - keep       # the tiles P1 touched last are re-used
        assert loads[3] == (ntl - keep if second else 0)


def test_model_detects_a_broken_release_rule():
    """Sanity of the checker itself: if P1 released the tiles that P2 re-uses, the producer may refill their
    slots before P2 has read them - the model must notice (wrong tile in the slot, or a lost hand-shake)."""
    caught = 0
    for seed in range(20):
        try:
            simulate(9, 3, True, seed, broken=True)
        except AssertionError:
            caught += 1
    assert caught > 0
