"""Host-side model of the brick conv kernel's producer / MMA / epilogue protocol (csrc/conv_brick.cu).

The four warp roles are replayed as coroutines over simulated mbarriers under random interleavings.  The model checks
what a GPU run can only show as a hang or a wrong number: no dead-lock, every activation box / weight slab is the one
the MMA issuer expects when it consumes it, every accumulator receives exactly its 27 x nchunks tap contributions
before the epilogue drains it, and no TMEM slot or shared-memory buffer is overwritten while still in use.
"""
import random

import pytest


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.pending = self.count
            self.phase ^= 1

    def passed(self, parity):
        return self.phase != parity


def simulate(P, D, nchunks, nslabbuf, nstages, units_of_cta, seed, async_commit=True, kwf=False, epi_groups=1):
    rng = random.Random(seed)
    nphases = 3 * nchunks
    resident = nslabbuf >= nphases
    assert resident or not kwf, "the kw-fused variant needs resident slabs"
    aphases = nchunks if kwf else nphases  # activation-stage phases: KWF loads one haloed box per (chunk, plane)
    full = [Bar(1) for _ in range(nstages)]
    empty = [Bar(1) for _ in range(nstages)]
    wfull = [Bar(1) for _ in range(max(nslabbuf, 1))]
    wempty = [Bar(1) for _ in range(max(nslabbuf, 1))]
    tfull = [Bar(1) for _ in range(2 * P)]
    tempty = [Bar(4) for _ in range(2 * P)]
    stage_data = [None] * nstages          # (unit, ph, p) held by each activation stage, None = free
    slab_data = [None] * max(nslabbuf, 1)  # phase id held by each slab buffer
    acc = [None] * (2 * P)                 # per TMEM slot: set of contributions, None = drained
    deferred = []                          # commits in flight: (callable)
    done_planes = []

    def wait(bar, parity):
        while not bar.passed(parity):
            yield

    def a_producer():
        stage, phase = 0, 0
        for (ui, d0) in units_of_cta:
            for ph in range(aphases):
                for p in range(P + 2):
                    yield from wait(empty[stage], phase ^ 1)
                    assert stage_data[stage] is None, "activation stage overwritten while in use"
                    stage_data[stage] = (ui, ph, p)
                    full[stage].arrive()
                    stage += 1
                    if stage == nstages:
                        stage, phase = 0, phase ^ 1
                    yield

    def w_producer():
        su = 0
        for (ui, d0) in units_of_cta:
            if resident and su >= nphases:
                break
            for ph in range(nphases):
                if resident:
                    buf = ph
                else:
                    buf = su % nslabbuf
                    yield from wait(wempty[buf], ((su // nslabbuf) & 1) ^ 1)
                    assert slab_data[buf] is None, "weight slab overwritten while in use"
                slab_data[buf] = ph
                wfull[buf].arrive()
                su += 1
                yield

    def commit(fn):
        if async_commit:
            deferred.append(fn)
        else:
            fn()

    def mma():
        stage, phase, su, tcount = 0, 0, 0, 0
        for (ui, d0) in units_of_cta:
            bb, par = tcount & 1, (tcount >> 1) & 1
            if kwf and tcount == 0:
                for b in range(nphases):
                    yield from wait(wfull[b], 0)
            for ph in range(aphases):
                if kwf:
                    buf = None
                    assert all(slab_data[ph * 3 + kw] == ph * 3 + kw for kw in range(3))
                elif resident:
                    buf = ph
                    if tcount == 0:
                        yield from wait(wfull[buf], 0)
                else:
                    buf = su % nslabbuf
                    yield from wait(wfull[buf], (su // nslabbuf) & 1)
                if not kwf:
                    assert slab_data[buf] == ph, f"slab {slab_data[buf]} != phase {ph}"
                for p in range(P + 2):
                    kd_lo, kd_hi = max(0, p - (P - 1)), min(2, p)
                    fresh = ph == 0 and p < P
                    if fresh:
                        slot = bb * P + p
                        yield from wait(tempty[slot], par ^ 1)
                        assert acc[slot] is None, "TMEM slot reused before the epilogue drained it"
                        acc[slot] = set()
                    yield from wait(full[stage], phase)
                    assert stage_data[stage] == (ui, ph, p), f"stage holds {stage_data[stage]}, want {(ui, ph, p)}"
                    for kd in range(kd_lo, kd_hi + 1):  # one wide-N MMA per (kh, k): column blocks kd_lo..kd_hi
                        slot = bb * P + (p - kd)
                        assert acc[slot] is not None, "accumulating into a slot that was never started"
                        for kh in range(3):
                            for key in ([(ph * 3 + kw, kd, kh) for kw in range(3)] if kwf else [(ph, kd, kh)]):
                                assert key not in acc[slot]
                                acc[slot].add(key)

                    def free_stage(s=stage):
                        stage_data[s] = None
                        empty[s].arrive()
                    commit(free_stage)
                    if ph == aphases - 1 and p >= 2:
                        commit(lambda s=bb * P + p - 2: tfull[s].arrive())
                    stage += 1
                    if stage == nstages:
                        stage, phase = 0, phase ^ 1
                    yield
                if not resident:
                    def free_slab(b=buf):
                        slab_data[b] = None
                        wempty[b].arrive()
                    commit(free_slab)
                su += 1
            tcount += 1

    def epilogue(widx, group):
        tcount = 0
        for (ui, d0) in units_of_cta:
            bb, par = tcount & 1, (tcount >> 1) & 1
            for q in range(group, P, epi_groups):  # CC = 16 kernels: two warp groups take alternate planes
                slot = bb * P + q
                yield from wait(tfull[slot], par)
                want = {(ph, kd, kh) for ph in range(nphases) for kd in range(3) for kh in range(3)}
                assert acc[slot] == want, f"unit {ui} plane {q}: {len(acc[slot] or ())} of {len(want)} contributions"
                if widx == 0:
                    done_planes.append((ui, q))
                yield
                tempty[slot].arrive()
                if tempty[slot].pending == tempty[slot].count:  # last of the 4 warps: slot is free again
                    acc[slot] = None
            tcount += 1

    roles = [a_producer(), w_producer(), mma()] + [epilogue(i, g) for g in range(epi_groups) for i in range(4)]
    alive = list(range(len(roles)))
    idle_rounds = 0
    while alive:
        # commits complete in issue order, some time after issue
        while deferred and rng.random() < 0.6:
            deferred.pop(0)()
        i = rng.choice(alive)
        before = (tuple(b.phase for b in full + empty + wfull + wempty + tfull + tempty), len(deferred))
        try:
            next(roles[i])
        except StopIteration:
            alive.remove(i)
        after = (tuple(b.phase for b in full + empty + wfull + wempty + tfull + tempty), len(deferred))
        idle_rounds = idle_rounds + 1 if before == after else 0
        if idle_rounds > 20000:
            raise AssertionError("dead-lock: no barrier changed phase for 20000 scheduler steps")
    while deferred:
        deferred.pop(0)()
    assert sorted(done_planes) == sorted((ui, q) for (ui, _) in units_of_cta for q in range(P))


@pytest.mark.parametrize("P,nchunks,nslabbuf,nstages", [
    (8, 1, 3, 12),   # 32->32: resident slabs
    (8, 1, 3, 6),    # 64->32
    (4, 1, 2, 4),    # 64->64: slabs stream through two buffers
    (4, 2, 2, 4),    # 128->64
    (4, 1, 3, 2),    # minimum ring depth
    (8, 2, 2, 3),
    (8, 2, 6, 4),    # 64->32 as two 32-channel chunks, resident
])
def test_brick_protocol(P, nchunks, nslabbuf, nstages):
    D = 4 * P
    for seed in range(6):
        # a CTA's share of units: bricks at the low edge, interior, high edge, and the same again
        units = [(i, (i % 4) * P) for i in range(7)]
        simulate(P, D, nchunks, nslabbuf, nstages, units, seed)
    simulate(P, P, nchunks, nslabbuf, nstages, [(0, 0), (1, 0), (2, 0)], 99)  # single-brick volume: both edges at once
    simulate(P, D, nchunks, nslabbuf, nstages, [(0, P)], 5, async_commit=False)
    if nslabbuf >= 3 * nchunks:  # resident slabs: the kernel runs the kw-fused variant
        for seed in range(4):
            simulate(P, D, nchunks, nslabbuf, nstages, [(i, (i % 4) * P) for i in range(5)], seed, kwf=True)
            simulate(P, D, nchunks, nslabbuf, nstages, [(i, (i % 4) * P) for i in range(5)], seed, kwf=True,
                     epi_groups=2)
