"""Data parallelism over environments: one process per GPU, environments sharded by contiguous
batch-index ranges, ONE exchange per DDPG phase.  The exchange itself lives in libpdeb200 behind the C ABI
(`pdeb200_comm_init`; csrc/comm.cu, csrc/comm.cuh): the last CTA of each gradient kernel sums the reduced gradient
over the ranks through NVLink peer memory.  This module only does what a host must do around it:

  * `shard_range`  -- which environments a rank owns;
  * `Comm`         -- carries the 128-byte unique id from rank 0 to the other ranks over whatever host channel the
                      launcher provides (here: torch.distributed, gloo or nccl) and calls `pdeb200_comm_init`;
                      host-side scalar reductions for PDEhook go through `pdeb200_comm_allreduce_f64`.

The reference is single-process (no counterpart); a Julia host does the same three calls through `ccall`
(julia/B200PDE.jl `comm_unique_id`, `comm_init!`).  Inference-only stepping needs no collective.
"""
import ctypes as C

import numpy as np

from . import _lib as L


def shard_range(n_global, rank, world_size):
    """Contiguous [lo, hi) slice of the global environment batch owned by `rank`."""
    base, rem = divmod(int(n_global), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class Comm:
    """One rank's handle on the process group.  `dist` is an initialised torch.distributed module (or None for a
    single process); it is used for host-side plumbing only -- never for the gradient exchange."""

    def __init__(self, dist=None):
        self.dist = dist
        self.world_size = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world_size > 1 else 0
        self._attached = {}

    # -- host channel ------------------------------------------------------------------------------
    def broadcast_bytes(self, data, n, src=0):
        """`n` bytes from rank `src` to every rank (gloo: CPU tensor, nccl: tensor on the current device)."""
        if self.world_size == 1:
            return bytes(data)
        import torch
        dev = "cuda" if self.dist.get_backend() == "nccl" else "cpu"
        buf = np.frombuffer(bytes(data), dtype=np.uint8).copy() if self.rank == src else np.zeros(n, dtype=np.uint8)
        t = torch.from_numpy(buf).to(dev)
        self.dist.broadcast(t, src=src)
        return t.cpu().numpy().tobytes()

    def allreduce_sum_(self, tensor):
        """Host-channel sum (CPU-side bookkeeping, gloo tests)."""
        if self.world_size > 1:
            self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM)
        return tensor

    def global_batch(self, local_batch):
        """Sum of the local batch sizes when they are equal on all ranks (the weak-scaling setup)."""
        return int(local_batch) * self.world_size

    # -- the library's communicator ------------------------------------------------------------------
    def attach(self, env):
        """Collective: join `env`'s context to the process group (pdeb200_comm_unique_id on rank 0 ->
        broadcast -> pdeb200_comm_init on every rank).  Returns the transport in use
        (L.COMM_NONE / L.COMM_NCCL / L.COMM_PEER)."""
        key = id(env)
        if key in self._attached:
            return self._attached[key]
        lib = env._lib
        uid = (C.c_uint8 * L.UNIQUE_ID_BYTES)()
        if self.rank == 0:
            L.check(lib.pdeb200_comm_unique_id(uid))
        raw = self.broadcast_bytes(bytes(uid), L.UNIQUE_ID_BYTES, src=0)
        uid = (C.c_uint8 * L.UNIQUE_ID_BYTES).from_buffer_copy(raw)
        L.check(lib.pdeb200_comm_init(env._ctx, uid, int(self.rank), int(self.world_size)), env._ctx)
        tr = C.c_int32()
        L.check(lib.pdeb200_comm_info(env._ctx, None, None, C.byref(tr)), env._ctx)
        self._attached[key] = tr.value
        return tr.value

    def allreduce_f64(self, env, values):
        """In-place sum over the ranks of up to 64 host doubles through the library (NCCL on the context's stream)."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        L.check(env._lib.pdeb200_comm_allreduce_f64(env._ctx, v.ctypes.data, int(v.size)), env._ctx)
        return v


def data_parallel_update(local_grads_fn, apply_fn, comm):
    """The exchange pattern in isolation over the HOST channel (used by the gloo CPU test, which has no GPU): grads =
    sum of the per-shard gradients already scaled by 1/global_batch; every rank applies the identical update."""
    g = local_grads_fn()
    comm.allreduce_sum_(g)
    apply_fn(g)
    return g
