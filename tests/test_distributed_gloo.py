"""world_size-2 gloo test of the multi-GPU decomposition (host logic): sharding the (tile, mirror) work items of one
case over ranks and summing the per-rank accumulators with one collective reproduces the single-process result."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sliding_window as SW
from oracle import unet as OU
from tests.helpers import build_dropin_unet


def _partial_accumulator(rank, world, vol, patch, sd, arch):
    """Oracle arithmetic restricted to the rank's work items (what SlidingWindowPredictor.accumulate does on a GPU)."""
    from brainseg_b200 import sliding as S

    steps = S.compute_steps_for_sliding_window(patch, vol.shape[1:], 0.5)
    tiles = [(z, y, x) for z in steps[0] for y in steps[1] for x in steps[2]]
    gauss = torch.from_numpy(SW.get_gaussian(patch))
    acc = torch.zeros((3,) + tuple(vol.shape[1:]))
    codes = S.mirror_codes_for((0, 1, 2))
    for t, m in S.shard_work_items(len(tiles), codes, rank, world):
        z, y, x = tiles[t]
        tile = torch.from_numpy(vol[None, :, z:z + patch[0], y:y + patch[1], x:x + patch[2]].copy())
        flips = SW.MIRROR_FLIPS[m]
        xin = torch.flip(tile, flips) if flips else tile
        pred = torch.sigmoid(OU.forward(sd, arch, xin))
        if flips:
            pred = torch.flip(pred, flips)
        acc[:, z:z + patch[0], y:y + patch[1], x:x + patch[2]] += pred[0] * (1.0 / len(codes)) * gauss
    return acc


def _worker(rank, world, port, vol, patch, sd, arch, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    acc = _partial_accumulator(rank, world, vol, patch, sd, arch)
    dist.all_reduce(acc)  # the one collective of the latency mode
    if rank == 0:
        out.put(acc.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_accumulators_sum_to_the_full_prediction():
    net = build_dropin_unet("bn", base=16, num_pool=2, seed=4)
    sd = {k: v.detach().float() for k, v in net.state_dict().items()}
    arch = OU.arch_from_module(net)
    patch = (16, 16, 16)
    vol = torch.randn(4, 16, 24, 20, generator=torch.Generator().manual_seed(1)).numpy()
    fwd = lambda x: OU.forward(sd, arch, x)  # noqa: E731
    _, probs_ref = SW.predict_3d_tiled(fwd, torch.sigmoid, vol, 3, patch, True, (0, 1, 2), 0.5, True, (1, 2, 3))
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, vol, patch, sd, arch, out)) for r in range(2)]
    for p in procs:
        p.start()
    acc = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from brainseg_b200 import sliding as S
    steps = S.compute_steps_for_sliding_window(patch, vol.shape[1:], 0.5)
    wsum = np.zeros(vol.shape[1:], dtype=np.float32)
    g = SW.get_gaussian(patch)
    for z in steps[0]:
        for y in steps[1]:
            for x in steps[2]:
                wsum[z:z + 16, y:y + 16, x:x + 16] += g
    probs = acc / wsum
    assert np.abs(probs - probs_ref).max() < 1e-5  # fp32 reassociation only
