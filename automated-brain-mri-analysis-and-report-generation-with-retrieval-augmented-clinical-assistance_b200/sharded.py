"""Exchange step of the sharded sliding window (BASELINE configs[2]): the (tile, mirror) work items of ONE case are
dealt to the ranks (sliding.shard_work_items), every rank accumulates its share into a private fp32 accumulator, and
the shares meet here.  Reference call sites being split: run_brats2021_inference_singlethread.py:97-106,113-124 (the
per-fold predict calls), :128 (fold mean), :144-156 (regions decision).

Two routes, same result up to the order of the fp32 additions:

* peer (default on one NVSwitch box): every accumulator and label volume lives in a slab the other ranks have mapped
  into their own device's address space (CUDA IPC: bsg_ipc_export / bsg_ipc_open).  ONE kernel per model and rank — bsg_finalize_peer — reads the
  rank's voxel slab of ALL ranks' accumulators over NVLink, sums in rank order, divides by the weight sum, averages
  the folds, decides, and stores the uint8 labels of the slab into EVERY rank's label volume.  The ranks are ordered
  INSIDE that kernel (bsg_finalize_peer_signal: release / acquire flags in peer memory — "my accumulators are complete"
  before the reads, "my slab has landed" after the writes), so the exchange is one launch per model and rank with no
  collective around it; BSG_PEER_SYNC=nccl brings back the two 4-byte NCCL all-reduces around bsg_finalize_peer.
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
        self._bufs = {}     # key -> (local tensor, [address of every rank's tensor, as mapped into this process])
        self._slabs = []    # device allocations the buffers are carved from (one IPC handle each)
        self._opened = {}   # IPC handle bytes -> base address of the mapping in this process
        self._tables = {}   # key tuple -> device pointer table
        self._token = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._comm = None
        self._side = torch.cuda.Stream(self.device)
        self.peer_sync = os.environ.get("BSG_PEER_SYNC", "kernel").lower()  # kernel: flags inside the exchange kernel
        if self.peer_sync not in ("kernel", "nccl"):
            raise ValueError(f"BSG_PEER_SYNC {self.peer_sync!r}: expected 'kernel' or 'nccl'")
        self._epoch = 0  # exchange calls so far: the same on every rank (they call in lockstep)
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
            if len(set(devs)) < self.world:
                # ranks sharing a device are time-sliced, not concurrent: a kernel waiting for a peer's flag would spin
                # through its whole time slice — those set-ups keep the stream-ordered barriers around the kernel
                self.peer_sync = "nccl"

    # ------------------------------------------------------------------ shared buffers
    SLAB_BYTES = 384 << 20  # two 107 MB accumulators + two 9 MB label volumes of a BraTS case, with room to spare

    def _new_slab(self, min_bytes):
        """One device allocation per rank, exported with CUDA IPC and mapped by every other rank.  Collective."""
        nbytes = max(self.SLAB_BYTES, (int(min_bytes) + (1 << 21) - 1) >> 21 << 21)
        local = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        bases = [local.data_ptr()] * self.world
        if self.route == "peer" and self.world > 1:
            lib = L.lib()
            handle = C.create_string_buffer(64)
            off = C.c_size_t(0)
            L.check(lib.bsg_ipc_export(_ptr(local), handle, C.byref(off)))
            gathered = [None] * self.world
            self.dist.all_gather_object(gathered, (bytes(handle.raw), int(off.value)))
            bases = []
            for r, (hbytes, offset) in enumerate(gathered):
                if r == self.rank:
                    bases.append(local.data_ptr())
                    continue
                if hbytes not in self._opened:  # an allocation may be opened once per process
                    base = C.c_void_p()
                    L.check(lib.bsg_ipc_open(C.create_string_buffer(hbytes, 64), C.byref(base)))
                    self._opened[hbytes] = base.value
                bases.append(self._opened[hbytes] + offset)
        self._slabs.append({"local": local, "bases": bases, "used": 0})
        return self._slabs[-1]

    def shared(self, key, shape, dtype):
        """A device tensor of this rank plus the address of every rank's tensor of the same key as seen from THIS rank's
        device.  Collective: all ranks call it with the same keys in the same order."""
        if key in self._bufs:
            return self._bufs[key]
        numel = 1
        for d in shape:
            numel *= int(d)
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        slab = self._slabs[-1] if self._slabs else None
        if slab is None or slab["used"] + nbytes > slab["local"].numel():
            slab = self._new_slab(nbytes)
        off = slab["used"]
        slab["used"] = (off + nbytes + 255) // 256 * 256
        local = slab["local"][off:off + nbytes].view(dtype).view(shape)
        self._bufs[key] = (local, [b + off for b in slab["bases"]])
        return self._bufs[key]

    def _table(self, ptrs):
        key = tuple(int(p) for p in ptrs)
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
        seg_local, seg_ptrs = self.shared(seg_key, (nvox,), torch.uint8)
        if self.route == "peer" and nvox % 4 == 0:
            accs = [self._bufs[k][1] for k in acc_keys]  # [K][R] addresses
            table = self._table([accs[k][r] for k in range(K) for r in range(self.world)])
            segs = self._table(seg_ptrs)
            per = -(-(nvox // 4) // self.world) * 4
            v0 = min(self.rank * per, nvox)
            nv = min(per, nvox - v0)
            if self.peer_sync == "kernel" and self.world > 1 and nv > 0 and per * (self.world - 1) < nvox:
                # flag block of every rank: arrive[R], done[R], finished-block count (zero-initialised slab memory)
                _, flag_ptrs = self.shared(("flags",), (2 * self.world + 2,), torch.int32)
                self._epoch += 1
                L.check(lib.bsg_finalize_peer_signal(_ptr(table), K, self.world, _ptr(wsum), ncls, nvox, v0, nv, mode, order,
                                                     _ptr(segs), self.world, _ptr(self._table(flag_ptrs)), self.rank,
                                                     self._epoch, L.stream_ptr()))
            else:
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

    def check(self):
        """Raises if an in-kernel wait of the exchange timed out (a rank did not show up within 10 s): the label volumes
        produced since are invalid.  Synchronises the device."""
        torch.cuda.synchronize(self.device)
        buf = self._bufs.get(("flags",))
        if buf is not None:
            code = int(buf[0][2 * self.world + 1].item())
            if code:
                who = (code & 0xFF) - (1 if code < 0x100 else 0)
                raise L.BsgError(f"sharded exchange: rank {who} " + ("never announced its accumulators" if code < 0x100
                                 else "never delivered its slab") + " (in-kernel wait timed out)")

    def close(self):
        """Releases the peer mappings and the communicator (call before destroy_process_group).  Collective: every rank
        runs the whole sequence even when its own check() failed — the failure is raised at the end, after the barriers the
        other ranks are waiting in."""
        err = None
        try:
            self.check()
        except L.BsgError as e:
            err = e
        if self.dist.is_initialized():
            self.dist.barrier()
        self._tables.clear()
        self._bufs.clear()
        for base in self._opened.values():
            L.check(L.lib().bsg_ipc_close(C.c_void_p(base)))
        self._opened.clear()
        if self.dist.is_initialized():
            self.dist.barrier()  # nobody frees a slab a peer still has mapped
        self._slabs.clear()
        if self._comm is not None:
            L.lib().bsg_nccl_comm_destroy(self._comm)
            self._comm = None
        if err is not None:
            raise err
