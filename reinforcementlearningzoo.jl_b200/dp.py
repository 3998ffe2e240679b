"""Data-parallel plumbing (no counterpart in the reference): one process per GPU, the NCCL unique id is
created by rank 0 and shared through torch.distributed (any backend; gloo on CPU, nccl on GPUs) or passed explicitly."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def shard_bounds(global_batch, rank, world):
    """Contiguous shard [lo, hi) of a global batch for ``rank`` (B split into G contiguous shards)."""
    if global_batch % world:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per


def shard_batch(batch, rank, world):
    lo, hi = shard_bounds(len(batch["action"]), rank, world)
    return {k: (v[lo:hi] if v is not None else None) for k, v in batch.items()}


class DataParallel:
    def __init__(self, rank, world, unique_id=None):
        lib = L.lib()
        self.rank, self.world = int(rank), int(world)
        if unique_id is None:
            import torch
            import torch.distributed as dist
            buf = np.zeros(128, np.uint8)
            if rank == 0:
                L.check(lib.zoo_dp_unique_id(buf.ctypes.data))
            t = torch.from_numpy(buf)
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.broadcast(t, 0)
            unique_id = t.cpu().numpy()
        self.unique_id = np.ascontiguousarray(unique_id, np.uint8)
        h = C.c_void_p()
        L.check(lib.zoo_dp_init(self.unique_id.ctypes.data, self.rank, self.world, C.byref(h)))
        self.h = h

    def close(self):
        """ncclCommDestroy.  Collective: every rank must call it, after the learners that use this communicator are gone
        and their streams are idle.  Deliberately NOT done from __del__: at interpreter exit the ranks reach it in
        arbitrary order (and after torch's own process group is gone), where ncclCommDestroy can block forever."""
        if getattr(self, "h", None):
            L.check(L.lib().zoo_dp_destroy(self.h))
            self.h = None
