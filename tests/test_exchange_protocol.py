"""Host-side model of the in-kernel rank ordering of the sharded exchange (csrc/multigpu.cu, peer_finalize_kernel with a
flag table; reference call sites being sharded: run_brats2021_inference_singlethread.py:97-106, :113-128, :144-156).

R ranks are replayed as coroutines under random interleavings: every rank's stream runs, per epoch, "accumulate" (writes
its own accumulator), then the exchange kernel (several blocks), then "consume" (reads its own label volume).  The kernel
follows the device code: block 0 stores arrive[rank] = epoch into every rank's flag block; every block waits until all R
arrive flags of ITS rank's block have reached the epoch, reads all ranks' accumulators for its part of the rank's slab and
writes the labels into every rank's label volume; the last block of a rank to finish stores done[rank] = epoch
everywhere and waits for all R done flags.  Flags are only ever overwritten by later epochs (compared modulo 2^32).

Checked for every interleaving: no dead-lock; no accumulator is read before its owner finished accumulating that epoch
or after the owner started the next one; when a rank's kernel completes, its label volume holds all R slabs of the
epoch; the finished-block counter is back at zero.
"""
import random

import pytest


def reached(flag, epoch):
    return ((flag - epoch) & 0xFFFFFFFF) < 0x80000000  # static_cast<int32_t>(flag - epoch) >= 0


def simulate(R, blocks, epochs, seed, first_epoch=1, wait_done=True):
    rng = random.Random(seed)
    prev = (first_epoch - 1) & 0xFFFFFFFF  # state after the previous call (0 = the zero-initialised flag block before call 1)
    flags = [{"arrive": [prev] * R, "done": [prev] * R, "count": 0} for _ in range(R)]
    acc_epoch = [prev] * R       # epoch whose accumulation is complete in rank r's accumulator
    acc_writing = [False] * R    # rank r's stream is overwriting its accumulator (next epoch's forwards)
    labels = [[0] * (R * blocks) for _ in range(R)]  # per rank: epoch of every (slab owner, block part) of its label volume
    reads = 0

    def block(rank, b, e):
        nonlocal reads
        if b == 0:
            for r in range(R):                     # st.release.sys arrive[rank] in every rank's flag block
                flags[r]["arrive"][rank] = e
                yield
        for r in range(R):                         # threads < R spin on the local flag block
            while not reached(flags[rank]["arrive"][r], e):
                yield
        for r in range(R):                         # peer loads of every rank's accumulator (this block's part of the slab)
            assert acc_epoch[r] == e and not acc_writing[r], \
                f"rank {rank} read rank {r}'s accumulator of epoch {acc_epoch[r]} (writing: {acc_writing[r]}) in epoch {e}"
            reads += 1
            yield
        for r in range(R):                         # peer stores of the labels into every rank's label volume
            labels[r][rank * blocks + b] = e
            yield
        flags[rank]["count"] += 1                  # atomicAdd on the local finished-block counter
        last = flags[rank]["count"] == blocks
        if last:
            flags[rank]["count"] = 0
            for r in range(R):                     # st.release.sys done[rank]
                flags[r]["done"][rank] = e
                yield
            for r in range(R):
                while wait_done and not reached(flags[rank]["done"][r], e):
                    yield

    def stream(rank):
        for e in range(first_epoch, first_epoch + epochs):
            acc_writing[rank] = True               # the forwards of this epoch overwrite the accumulator
            for _ in range(rng.randint(0, 6)):
                yield
            acc_epoch[rank] = e & 0xFFFFFFFF
            acc_writing[rank] = False
            blks = [block(rank, b, e & 0xFFFFFFFF) for b in range(blocks)]
            live = list(blks)
            while live:                            # the kernel: its blocks advance in random order
                g = rng.choice(live)
                try:
                    next(g)
                except StopIteration:
                    live.remove(g)
                yield
            # launch complete on the stream: the label volume must be whole, the counter reset
            assert all(v == (e & 0xFFFFFFFF) for v in labels[rank]), f"rank {rank}: label volume not whole after epoch {e}"
            assert flags[rank]["count"] == 0

    streams = [stream(r) for r in range(R)]
    live = list(streams)
    steps = 0
    while live:
        s = rng.choice(live)
        try:
            next(s)
        except StopIteration:
            live.remove(s)
        steps += 1
        assert steps < 2_000_000, "dead-lock: the ranks stopped making progress"
    assert reads == R * R * blocks * epochs


@pytest.mark.parametrize("R,blocks", [(2, 1), (2, 3), (4, 2), (8, 2)])
def test_exchange_protocol_random_interleavings(R, blocks):
    for seed in range(40):
        simulate(R, blocks, epochs=4, seed=seed)


def test_exchange_protocol_epoch_wraparound():
    """The epoch counter is a uint32 compared modulo 2^32: crossing 0xFFFFFFFF -> 0 must not stall or let a rank run ahead."""
    for seed in range(20):
        simulate(3, 2, epochs=6, seed=seed, first_epoch=0xFFFFFFFD)


def test_model_notices_a_missing_final_wait():
    """Negative control: without the wait for the other ranks' done flags a fast rank's launch completes before the slow
    ranks' slabs have landed (and its next epoch overwrites an accumulator still being read) — the model must notice."""
    failures = 0
    for seed in range(40):
        try:
            simulate(4, 2, epochs=4, seed=seed, wait_done=False)
        except AssertionError:
            failures += 1
    assert failures >= 30
