"""Exchange step of the sharded sliding window (BASELINE configs[2]): the (tile, mirror) work items of ONE case are
dealt to the ranks (sliding.shard_work_items), every rank accumulates its share into a private fp32 accumulator, and
the shares meet here.  Reference call sites being split: run_brats2021_inference_singlethread.py:97-106,113-124 (the
per-fold predict calls), :128 (fold mean), :144-156 (regions decision).

Two routes, same result up to the order of the fp32 additions:

* peer (default on one NVSwitch box): every accumulator and label volume lives in memory the other ranks have mapped
  (CUDA IPC through torch's shared-storage handles).  ONE kernel per model and rank — bsg_finalize_peer — reads the
  rank's voxel slab of ALL ranks' accumulators over NVLink, sums in rank order, divides by the weight sum, averages
  the folds, decides, and stores the uint8 labels of the slab into EVERY rank's label volume.  Two tiny NCCL
  all-reduces order the ranks around it (all accumulators complete before / all slabs written after).
* nccl: one ncclAllReduce of each accumulator (bsg_nccl_reduce_accumulator, the library's own communicator), then the
  single-GPU bsg_finalize on every rank.

torch.distributed is plumbing here (rendezvous, handle exchange, the barrier all-reduces); it must be initialised.
"""
import ctypes as C
import os

import torch

from . import _lib as L


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class ShardedExchange:
    def __init__(self, rank, world_size, device, route=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("ShardedExchange needs an initialised torch.distributed process group")
        self.dist = dist
        self.rank, self.world, self.device = int(rank), int(world_size), torch.device(device)
        route = (route or os.environ.get("BSG_SHARD_ROUTE", "peer")).lower()
        if route not in ("peer", "nccl"):
            raise ValueError(f"route {route!r}: expected 'peer' or 'nccl'")
        self.route = route
        self._bufs = {}     # key -> (local tensor, [per-rank tensors])
        self._tables = {}   # key tuple -> device pointer table
        self._token = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._comm = None
        self._side = torch.cuda.Stream(self.device)
        self.nccl_bytes = 0   # bytes handed to NCCL collectives (accumulator route) per call of reduce()
        self.peer_bytes = 0   # bytes read from / written to peer memory by bsg_finalize_peer
        self.launches = 0
        if self.route == "peer" and self.world > 1:
            # ranks are processes on one box; the kernel dereferences the peers' pointers directly, so every other
            # rank's device must be peer-accessible from this one (ranks sharing a device need nothing)
            mine = self.device.index if self.device.index is not None else torch.cuda.current_device()
            devs = [None] * self.world
            dist.all_gather_object(devs, int(mine))
            for d in sorted(set(devs)):
                if d != mine:
                    L.check(L.lib().bsg_enable_peer_access(d))

    # ------------------------------------------------------------------ shared buffers
    def shared(self, key, shape, dtype):
        """A device tensor of this rank plus every other rank's tensor of the same key, mapped into this process.
        Collective: all ranks call it with the same keys in the same order."""
        if key in self._bufs:
            return self._bufs[key]
        local = torch.zeros(shape, dtype=dtype, device=self.device)
        if self.route != "peer" or self.world == 1:
            self._bufs[key] = (local, [local])
            return self._bufs[key]
        from torch.multiprocessing.reductions import reduce_tensor
        fn, args = reduce_tensor(local)
        gathered = [None] * self.world
        self.dist.all_gather_object(gathered, (fn, args))
        views = []
        for r, (f, a) in enumerate(gathered):
            views.append(local if r == self.rank else f(*a))
        self._bufs[key] = (local, views)
        return self._bufs[key]

    def _table(self, tensors):
        key = tuple(t.data_ptr() for t in tensors)
        if key not in self._tables:
            self._tables[key] = torch.tensor(list(key), dtype=torch.int64, device=self.device)
        return self._tables[key]

    def barrier(self):
        """Stream-ordered barrier over the ranks: a 4-byte all-reduce on the current stream."""
        self.dist.all_reduce(self._token)

    # ------------------------------------------------------------------ exchange
    def side_stream(self):
        return self._side

    def finalize(self, acc_keys, wsum, ncls, regions_class_order, seg_key):
        """acc_keys: the shared-buffer keys of the K fold accumulators of one ensemble member.  Returns this rank's
        full label volume (uint8, flat [nvox]) — complete once the call's work on the current stream is."""
        lib = L.lib()
        nvox = wsum.numel()
        K = len(acc_keys)
        mode, order = 0, None
        if regions_class_order is not None:
            mode = 1
            order = (C.c_int * ncls)(*[int(c) for c in regions_class_order])
        seg_local, seg_views = self.shared(seg_key, (nvox,), torch.uint8)
        if self.route == "peer" and nvox % 4 == 0:
            accs = [self._bufs[k][1] for k in acc_keys]  # [K][R]
            table = self._table([accs[k][r] for k in range(K) for r in range(self.world)])
            segs = self._table(seg_views)
            per = -(-(nvox // 4) // self.world) * 4
            v0 = min(self.rank * per, nvox)
            nv = min(per, nvox - v0)
            self.barrier()  # every rank's accumulators are complete
            L.check(lib.bsg_finalize_peer(_ptr(table), K, self.world, _ptr(wsum), ncls, nvox, v0, nv, mode, order,
                                          _ptr(segs), self.world, L.stream_ptr()))
            self.barrier()  # every slab has landed in every label volume; accumulators may be reused
            self.launches += 1
            self.peer_bytes += K * ncls * nv * 4 * (self.world - 1) + nv * (self.world - 1)
            return seg_local
        # NCCL route: all-reduce each accumulator in place, then the single-GPU finalize
        if self._comm is None:
            self._comm = self._make_comm()
        ptrs = []
        for k in acc_keys:
            acc = self._bufs[k][0]
            L.check(lib.bsg_nccl_reduce_accumulator(self._comm, _ptr(acc), acc.numel(), -1, L.stream_ptr()))
            self.nccl_bytes += acc.numel() * 4
            ptrs.append(acc.data_ptr())
        arr = (C.c_void_p * K)(*ptrs)
        L.check(lib.bsg_finalize(arr, K, _ptr(wsum), ncls, nvox, mode, order, None, _ptr(seg_local), L.stream_ptr()))
        self.launches += 1
        return seg_local

    def _make_comm(self):
        """The library's own NCCL communicator: rank 0 creates the id, torch.distributed ships it."""
        lib = L.lib()
        buf = C.create_string_buffer(128)
        if self.rank == 0:
            L.check(lib.bsg_nccl_unique_id(buf))
        box = [bytes(buf.raw)]
        self.dist.broadcast_object_list(box, src=0)
        comm = C.c_void_p()
        L.check(lib.bsg_nccl_comm_create(C.create_string_buffer(box[0], 128), self.world, self.rank, C.byref(comm)))
        return comm

    def close(self):
        """Releases the peer mappings and the communicator (call before destroy_process_group)."""
        torch.cuda.synchronize(self.device)
        if self.dist.is_initialized():
            self.dist.barrier()
        self._tables.clear()
        self._bufs.clear()
        import gc
        gc.collect()  # the peers' mapped storages are released before the producers go away
        if self._comm is not None:
            L.lib().bsg_nccl_comm_destroy(self._comm)
            self._comm = None
