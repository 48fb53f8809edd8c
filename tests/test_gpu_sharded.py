"""The sharded sliding window (BASELINE configs[2]) on the GPU: the exchange kernel against torch arithmetic, the
library's NCCL route on a one-rank communicator, and the whole sharded path with TWO ranks (two processes sharing the
one GPU the test box has; gloo rendezvous, CUDA-IPC peer buffers) against the single-process result and the oracle.
Reference call sites being sharded: run_brats2021_inference_singlethread.py:97-106, :113-128, :144-156."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ptr(t):
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("K,R,mode", [(1, 2, 1), (2, 4, 1), (1, 8, 0), (3, 1, 1)])
def test_finalize_peer_matches_rank_ordered_sum(K, R, mode):
    """bsg_finalize_peer over K folds x R "ranks" (all on this device here): labels and the slab partition equal the
    same arithmetic in torch (rank-ordered fp32 sums, IEEE division, fold mean, decision)."""
    from brainseg_b200 import _lib as L

    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device="cpu").manual_seed(100 + K * 10 + R)
    ncls, shape = 3, (12, 10, 16)
    nvox = int(np.prod(shape))
    accs = [[torch.rand(ncls, nvox, generator=g).to(dev) for _ in range(R)] for _ in range(K)]
    wsum = (torch.rand(nvox, generator=g) + 0.5).to(dev) * R
    segs = [torch.full((nvox,), 255, dtype=torch.uint8, device=dev) for _ in range(2)]
    table = torch.tensor([accs[k][r].data_ptr() for k in range(K) for r in range(R)], dtype=torch.int64, device=dev)
    seg_table = torch.tensor([s.data_ptr() for s in segs], dtype=torch.int64, device=dev)
    order = (C.c_int * ncls)(1, 2, 3)
    per = -(-(nvox // 4) // 3) * 4  # three slabs, as three ranks would call it
    for r in range(3):
        v0 = min(r * per, nvox)
        nv = min(per, nvox - v0)
        L.check(L.lib().bsg_finalize_peer(_ptr(table), K, R, _ptr(wsum), ncls, nvox, v0, nv, mode, order if mode else None,
                                          _ptr(seg_table), 2, L.stream_ptr()))
    probs = None
    for k in range(K):
        s = accs[k][0].clone()
        for r in range(1, R):
            s = s + accs[k][r]
        p = s / wsum
        probs = p if probs is None else probs + p
    if K > 1:
        probs = probs / K
    if mode == 0:
        ref = probs.argmax(0).to(torch.uint8)
    else:
        ref = torch.zeros(nvox, dtype=torch.uint8, device=dev)
        for i, c in enumerate((1, 2, 3)):
            ref[probs[i] > 0.5] = c
    assert torch.equal(segs[0], ref) and torch.equal(segs[1], ref)


@pytest.mark.timeout(120)
def test_finalize_peer_signal_orders_the_ranks_in_the_kernel():
    """bsg_finalize_peer_signal: R = 3 "ranks" played by three streams of this process.  Each rank's kernel announces its
    accumulators, waits for the others' announcements inside the kernel, writes its slab into every rank's label volume
    and returns only when all slabs have landed — no collective around the launches.  Three epochs reuse the flags;
    rank 2's accumulator is filled by a kernel on ITS stream right before its exchange launch, so a rank that read
    before the announcement would see the stale values."""
    from brainseg_b200 import _lib as L

    dev = torch.device("cuda", torch.cuda.current_device())
    R, ncls, nvox = 3, 3, 4 * 256 * 6
    g = torch.Generator(device="cpu").manual_seed(5)
    accs = [torch.zeros(ncls, nvox, device=dev) for _ in range(R)]
    wsum = torch.full((nvox,), float(R), device=dev)
    segs = [torch.full((nvox,), 255, dtype=torch.uint8, device=dev) for _ in range(R)]
    flags = [torch.zeros(2 * R + 2, dtype=torch.int32, device=dev) for _ in range(R)]
    table = torch.tensor([a.data_ptr() for a in accs], dtype=torch.int64, device=dev)
    seg_table = torch.tensor([s.data_ptr() for s in segs], dtype=torch.int64, device=dev)
    flag_table = torch.tensor([f.data_ptr() for f in flags], dtype=torch.int64, device=dev)
    order = (C.c_int * ncls)(1, 2, 3)
    streams = [torch.cuda.Stream(dev) for _ in range(R)]
    per = nvox // R
    # every kernel used below runs once BEFORE any rank waits inside a kernel: CUDA loads a kernel lazily at its first
    # launch, and that load waits for the device to drain — behind a kernel that is itself waiting for the launch
    torch.cuda._sleep(1000)
    accs[0].copy_(accs[1])
    torch.cuda.synchronize()
    for epoch in (1, 2, 3):
        fresh = [torch.rand(ncls, nvox, generator=g).to(dev) * 2.0 for _ in range(R)]
        torch.cuda.synchronize()
        for r in (0, 1, 2):
            with torch.cuda.stream(streams[r]):
                if r == 2:
                    torch.cuda._sleep(20_000_000)  # ~10 ms: ranks 0 and 1 are already waiting in their kernels
                accs[r].copy_(fresh[r])
                L.check(L.lib().bsg_finalize_peer_signal(_ptr(table), 1, R, _ptr(wsum), ncls, nvox, r * per, per, 1, order,
                                                         _ptr(seg_table), R, _ptr(flag_table), r, epoch, L.stream_ptr()))
        # each stream alone is enough to know the whole volume has landed in that rank's label volume
        streams[0].synchronize()
        probs = (fresh[0] + fresh[1] + fresh[2]) / wsum
        ref = torch.zeros(nvox, dtype=torch.uint8, device=dev)
        for i, c in enumerate((1, 2, 3)):
            ref[probs[i] > 0.5] = c
        assert torch.equal(segs[0], ref), f"epoch {epoch}: rank 0's label volume is not whole after its own launch"
        torch.cuda.synchronize()
        assert all(torch.equal(s, ref) for s in segs)
        assert all(int(f[2 * R]) == 0 for f in flags)  # finished-block counts reset for the next launch
        assert all(int(f[2 * R + 1]) == 0 for f in flags), "an in-kernel wait timed out"


@pytest.mark.timeout(120)
def test_finalize_peer_signal_gives_up_on_a_missing_rank():
    """A rank that never launches its exchange kernel must not wedge the others: the in-kernel waits are bounded (10 s
    each) and leave a code in word 2R + 1 of the waiting rank's flag block — here 0x100 + 1 (rank 1's slab never
    landed; the earlier code 1 + 1, rank 1 never announced, is overwritten by it)."""
    from brainseg_b200 import _lib as L

    dev = torch.device("cuda", torch.cuda.current_device())
    R, ncls, nvox = 2, 3, 4 * 256
    accs = [torch.rand(ncls, nvox, device=dev) for _ in range(R)]
    wsum = torch.full((nvox,), float(R), device=dev)
    segs = [torch.zeros(nvox, dtype=torch.uint8, device=dev) for _ in range(R)]
    flags = [torch.zeros(2 * R + 2, dtype=torch.int32, device=dev) for _ in range(R)]
    table = torch.tensor([a.data_ptr() for a in accs], dtype=torch.int64, device=dev)
    seg_table = torch.tensor([s.data_ptr() for s in segs], dtype=torch.int64, device=dev)
    flag_table = torch.tensor([f.data_ptr() for f in flags], dtype=torch.int64, device=dev)
    order = (C.c_int * ncls)(1, 2, 3)
    import time
    t0 = time.perf_counter()
    L.check(L.lib().bsg_finalize_peer_signal(_ptr(table), 1, R, _ptr(wsum), ncls, nvox, 0, nvox // 2, 1, order,
                                             _ptr(seg_table), R, _ptr(flag_table), 0, 1, L.stream_ptr()))
    torch.cuda.synchronize()
    took = time.perf_counter() - t0
    code = int(flags[0][2 * R + 1])
    print(f"kernel gave up after {took:.1f} s with status 0x{code:x}")
    assert 15.0 < took < 40.0
    assert code == 0x100 + 1
    assert int(flags[1][0]) == 1 and int(flags[1][R + 0]) == 1  # rank 0 did announce and did report its slab to rank 1


def test_nccl_route_single_rank_communicator():
    """bsg_nccl_unique_id / comm_create / reduce_accumulator / comm_destroy on a one-rank communicator: the library
    resolves libnccl at run time and the all-reduce of one rank leaves the accumulator unchanged."""
    from brainseg_b200 import _lib as L

    lib = L.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    uid = C.create_string_buffer(128)
    L.check(lib.bsg_nccl_unique_id(uid))
    comm = C.c_void_p()
    L.check(lib.bsg_nccl_comm_create(uid, 1, 0, C.byref(comm)))
    acc = torch.randn(3, 1000, device=dev)
    ref = acc.clone()
    L.check(lib.bsg_nccl_reduce_accumulator(comm, _ptr(acc), acc.numel(), -1, L.stream_ptr()))
    L.check(lib.bsg_nccl_reduce_accumulator(comm, _ptr(acc), acc.numel(), 0, L.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(acc, ref)
    L.check(lib.bsg_nccl_comm_destroy(comm))


_WORKER = r"""
import os, sys, json
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, {root!r})
rank, world, port, out_path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
torch.cuda.set_device(0)
dist.init_process_group("gloo", rank=rank, world_size=world)
from brainseg_b200 import pipeline as PL, sharded as SH
from tests.helpers import build_dropin_unet
models = [build_dropin_unet("bn", base=16, num_pool=2, seed=81), build_dropin_unet("gn", base=16, num_pool=2, groups=4, seed=82)]
vol = torch.randn(4, 40, 56, 48, generator=torch.Generator().manual_seed(12)).numpy()
patch = (32, 32, 32)
dev = torch.device("cuda", 0)
sh = SH.ShardedExchange(rank, world, dev, route="peer")
pipe = PL.BratsCasePipeline(models, patch, 0.5, (0, 1, 2), True, True, (1, 2, 3), "brats2025", batch=8, rank=rank,
                            world_size=world, shard=sh)
res = None
for rep in range(3):  # accumulators and label volumes are reused from case to case
    res = pipe.run_case(vol, features=False)
sharded = [s.cpu().numpy().copy() for s in res["model_segmentations"]] + [res["segmentation"].cpu().numpy().copy()]
n_items = [len(p.work_items(p.geometry(vol.shape[1:])[0])) for p in pipe.predictors]
sh.close()
single = PL.BratsCasePipeline(models, patch, 0.5, (0, 1, 2), True, True, (1, 2, 3), "brats2025", batch=8)
r1 = single.run_case(vol, features=False)
ref = [s.cpu().numpy() for s in r1["model_segmentations"]] + [r1["segmentation"].cpu().numpy()]
np.savez(out_path + f".rank{{rank}}.npz", s0=sharded[0], s1=sharded[1], s2=sharded[2], r0=ref[0], r1=ref[1], r2=ref[2],
         items=np.array(n_items))
dist.barrier()
dist.destroy_process_group()
"""


def test_sharded_two_ranks_peer_route(tmp_path):
    """Two ranks (processes) on one GPU: work items dealt round-robin, accumulators and label volumes mapped across the
    processes with CUDA IPC, one bsg_finalize_peer per model and rank.  Every rank ends up with the complete label
    volumes; they equal the single-process result except where a probability sits within fp32 reassociation of 0.5,
    and the oracle's on decisive voxels."""
    from oracle import sliding_window as SW
    from tests.helpers import build_dropin_unet, oracle_fns

    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    port = str(29600 + os.getpid() % 2000)
    out = str(tmp_path / "res")
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", port, out], cwd=ROOT, stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    logs = [p.communicate(timeout=600)[0] for p in procs]
    for p, lg in zip(procs, logs):
        assert p.returncode == 0, lg[-3000:]
    res = [np.load(out + f".rank{r}.npz") for r in range(2)]
    assert list(res[0]["items"]) == [48, 48] and list(res[1]["items"]) == [48, 48]  # 12 tiles x 8 mirrors / 2 ranks
    for k in ("s0", "s1", "s2"):
        assert np.array_equal(res[0][k], res[1][k]), "ranks disagree on the label volume"
    for k in range(3):
        same = float((res[0][f"s{k}"] == res[0][f"r{k}"]).mean())
        print(f"volume {k}: sharded == single-process on {same * 100:.5f}% of the voxels")
        # only the order of the fp32 additions into the accumulator differs (two partial sums instead of one running sum)
        assert same >= 0.9999
    vol = torch.randn(4, 40, 56, 48, generator=torch.Generator().manual_seed(12)).numpy()
    for m, (variant, seed, groups) in enumerate((("bn", 81, 8), ("gn", 82, 4))):
        net = build_dropin_unet(variant, base=16, num_pool=2, groups=groups, seed=seed)
        seg_ref, probs_ref = SW.predict_3d_tiled(oracle_fns(net)[0], torch.sigmoid, vol, 3, (32, 32, 32), True, (0, 1, 2), 0.5,
                                                 True, (1, 2, 3))
        decisive = np.all(np.abs(probs_ref - 0.5) > 1e-2, axis=0)
        assert np.array_equal(res[0][f"s{m}"][decisive], seg_ref[decisive])
