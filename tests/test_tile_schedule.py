"""Host-side model of the tile conv kernel's producer / MMA / epilogue protocol (csrc/conv_tc.cu), all stage shapes:
one tap per stage, three kh taps per stage (haloed box or three boxes), and the 2-CTA pair mode in which each CTA
multicasts half of every weight stage into both CTAs and a stage is only free once BOTH CTAs' MMAs have read it.

The roles are replayed as coroutines over simulated mbarriers (arrival counts + transaction bytes) under random
interleavings.  Checked: no dead-lock; every (tap, chunk) unit reaches each tile's accumulator exactly once with the
operands of the right tile and K step; no shared-memory stage is overwritten while an MMA may still read it (in either
CTA of a pair); no TMEM accumulator is reused before its 8 epilogue warps drained it.
"""
import random

import pytest


class Bar:
    """mbarrier: `count` arrivals and a transaction-byte balance of zero complete a phase"""

    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _maybe_flip(self):
        if self.pending == 0 and self.tx == 0:
            self.pending = self.count
            self.phase ^= 1

    def arrive(self, expect_tx=0):
        self.tx += expect_tx
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects"
        self._maybe_flip()

    def complete_tx(self, nbytes):
        self.tx -= nbytes  # may run ahead of the expect_tx of the same phase (multicast from the peer)
        self._maybe_flip()

    def passed(self, parity):
        return self.phase != parity


def simulate(taps_per_stage, nchunks, nstages, tiles_per_cta, seed, pair=False, a_bytes=16384, w_bytes=8192):
    rng = random.Random(seed)
    ncta = 2 if pair else 1
    steps_per_tile = (27 // taps_per_stage) * nchunks
    full = [[Bar(1) for _ in range(nstages)] for _ in range(ncta)]
    empty = [[Bar(ncta) for _ in range(nstages)] for _ in range(ncta)]
    tfull = [[Bar(1) for _ in range(2)] for _ in range(ncta)]
    tempty = [[Bar(8) for _ in range(2)] for _ in range(ncta)]
    stage_a = [[None] * nstages for _ in range(ncta)]   # (tile, step) of the activation boxes held
    stage_w = [[[None, None] for _ in range(nstages)] for _ in range(ncta)]  # the two weight halves held
    acc = [[None, None] for _ in range(ncta)]           # per TMEM buffer: set of steps accumulated, None = drained
    in_flight = []                                      # asynchronous completions (TMA landings, MMA commits)
    drained = [[] for _ in range(ncta)]

    def wait(bar, parity):
        while not bar.passed(parity):
            yield

    def later(fn):
        in_flight.append(fn)

    def producer(cta):
        stage, phase = 0, 0
        for tile in range(tiles_per_cta):
            for step in range(steps_per_tile):
                yield from wait(empty[cta][stage], phase ^ 1)
                assert stage_a[cta][stage] is None, "activation stage overwritten while in use"
                full[cta][stage].arrive(expect_tx=a_bytes * taps_per_stage + w_bytes * taps_per_stage)

                def land_a(c=cta, s=stage, t=tile, k=step):
                    stage_a[c][s] = (t, k)
                    full[c][s].complete_tx(a_bytes * taps_per_stage)
                later(land_a)
                halves = (cta,) if pair else (0, 1)      # pair: my half of the rows, multicast to both CTAs
                for h in halves:
                    for dst in range(ncta):
                        def land_w(d=dst, s=stage, hh=h, t=tile, k=step):
                            assert stage_w[d][s][hh] is None, "weight half overwritten while in use"
                            stage_w[d][s][hh] = k  # weights depend on the K step only (both CTAs: same N tile)
                            full[d][s].complete_tx(w_bytes * taps_per_stage // 2)
                        later(land_w)
                stage += 1
                if stage == nstages:
                    stage, phase = 0, phase ^ 1
                yield

    def mma(cta):
        stage, phase = 0, 0
        for tile in range(tiles_per_cta):
            buf, par = tile & 1, (tile >> 1) & 1
            yield from wait(tempty[cta][buf], par ^ 1)
            assert acc[cta][buf] is None, "TMEM accumulator reused before the epilogue drained it"
            acc[cta][buf] = set()
            for step in range(steps_per_tile):
                yield from wait(full[cta][stage], phase)
                assert stage_a[cta][stage] == (tile, step), f"stage holds {stage_a[cta][stage]}, want {(tile, step)}"
                assert stage_w[cta][stage] == [step, step], f"weights {stage_w[cta][stage]} for step {step}"
                assert step not in acc[cta][buf]
                acc[cta][buf].add(step)

                def commit(c=cta, s=stage):               # tcgen05.commit: arrives once the MMAs have read the stage
                    stage_a[c][s] = None
                    stage_w[c][s] = [None, None]
                    for d in range(ncta):                 # pair: multicast arrive on both CTAs' empty barriers
                        empty[d][s].arrive()
                later(commit)
                if step == steps_per_tile - 1:
                    later(lambda c=cta, b=buf: tfull[c][b].arrive())
                stage += 1
                if stage == nstages:
                    stage, phase = 0, phase ^ 1
                yield

    def epilogue(cta, warp):
        for tile in range(tiles_per_cta):
            buf, par = tile & 1, (tile >> 1) & 1
            yield from wait(tfull[cta][buf], par)
            assert acc[cta][buf] == set(range(steps_per_tile)), f"tile {tile}: incomplete accumulator"
            yield
            if warp == 0:
                drained[cta].append(tile)
            tempty[cta][buf].arrive()
            if tempty[cta][buf].pending == tempty[cta][buf].count:  # last of the 8 warps
                acc[cta][buf] = None

    roles = []
    for cta in range(ncta):
        roles += [producer(cta), mma(cta)] + [epilogue(cta, w) for w in range(8)]
    alive = list(range(len(roles)))
    idle = 0
    while alive:
        # a stage's MMA commit may only be observed after the loads it depends on: completions retire in order per
        # kind here, at random times
        while in_flight and rng.random() < 0.7:
            in_flight.pop(0)()
        i = rng.choice(alive)
        snapshot = (len(in_flight), tuple(b.phase for c in range(ncta) for b in full[c] + empty[c] + tfull[c] + tempty[c]))
        try:
            next(roles[i])
        except StopIteration:
            alive.remove(i)
        after = (len(in_flight), tuple(b.phase for c in range(ncta) for b in full[c] + empty[c] + tfull[c] + tempty[c]))
        idle = idle + 1 if snapshot == after and not in_flight else 0
        assert idle < 50000, "dead-lock: nothing changed for 50000 scheduler steps"
    while in_flight:
        in_flight.pop(0)()
    for cta in range(ncta):
        assert drained[cta] == list(range(tiles_per_cta))


@pytest.mark.parametrize("taps_per_stage,nchunks,nstages", [
    (1, 1, 12),   # stride-2 32->64: one tap per stage, deep ring
    (1, 2, 6),    # 64->128
    (1, 8, 4),    # 1024->512 at N = 256: 4 stages
    (3, 1, 6),    # three kh taps per stage (haloed box or three boxes)
    (3, 5, 3),    # 320->320 at 8^3: the minimum of three stages the planner accepts
    (3, 2, 2),    # KHS forced with two stages
    (27, 1, 2),   # degenerate: a whole tile per stage
])
def test_tile_protocol(taps_per_stage, nchunks, nstages):
    for seed in range(5):
        simulate(taps_per_stage, nchunks, nstages, tiles_per_cta=5, seed=seed)
    simulate(taps_per_stage, nchunks, nstages, tiles_per_cta=1, seed=7)


@pytest.mark.parametrize("taps_per_stage,nchunks,nstages", [(3, 2, 3), (1, 2, 6), (1, 1, 2), (3, 4, 2)])
def test_tile_protocol_pair_mode(taps_per_stage, nchunks, nstages):
    for seed in range(5):
        simulate(taps_per_stage, nchunks, nstages, tiles_per_cta=4, seed=seed, pair=True)
