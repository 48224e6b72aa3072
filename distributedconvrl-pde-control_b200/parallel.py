"""Data parallelism over environments: one process per GPU, environments sharded by
contiguous batch-index ranges, ONE exchange step per DDPG phase -- an allreduce(sum) of the flat
gradient buffer over NCCL/NVLink (gloo in the CPU tests).  Inference-only stepping needs no collective.
"""
import numpy as np


def shard_range(n_global, rank, world_size):
    """Contiguous [lo, hi) slice of the global environment batch owned by `rank`."""
    base, rem = divmod(int(n_global), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _CudaAlias:
    """Zero-copy torch view of a device buffer owned by libpdeb200 (__cuda_array_interface__)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class Comm:
    """torch.distributed wrapper used by CustomDDPGPolicy.update."""

    def __init__(self, dist=None):
        self.dist = dist
        self.world_size = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world_size > 1 else 0
        self._views = {}

    def global_batch(self, local_batch):
        """Sum of the local batch sizes (equal on all ranks in the weak-scaling setup)."""
        return int(local_batch) * self.world_size

    def allreduce_sum_(self, tensor):
        if self.world_size > 1:
            self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM)
        return tensor

    def alias(self, env, which, dtype):
        import torch
        key = (id(env), which)
        ptr, nbytes = env.device_ptr(which)
        hit = self._views.get(key)
        if hit is not None and hit[0] == ptr and hit[1] == nbytes:
            return hit[2]
        item = {"float32": 4, "float64": 8}[dtype]
        t = torch.as_tensor(_CudaAlias(ptr, nbytes // item, {"float32": "<f4", "float64": "<f8"}[dtype]),
                            device=torch.device("cuda", env.device_index))
        self._views[key] = (ptr, nbytes, t)
        return t


def data_parallel_update(local_grads_fn, apply_fn, comm):
    """The exchange pattern in isolation (used by the gloo CPU test): grads = allreduce(sum of per-shard
    gradients already scaled by 1/global_batch); every rank applies the identical update."""
    g = local_grads_fn()
    comm.allreduce_sum_(g)
    apply_fn(g)
    return g
