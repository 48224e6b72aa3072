"""CPU, world_size = 2 over gloo: the data-parallel exchange.  Each rank computes the
critic/actor gradients of ITS shard of the batch scaled by 1/global_batch (oracle stands in for the
GPU kernels: this tests the host-side sharding + allreduce logic), allreduce(sum) makes them equal to
the single-process gradients of the whole batch, and the literal-Q1 r-bar is the global mean."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import importlib
    import torch
    import torch.distributed as dist
    from oracle import agent_oracle as AO
    from test_agent_oracle import batch, make_nets
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = importlib.import_module("distributedconvrl-pde-control_b200.parallel")
    comm = par.Comm(dist)
    # the host channel that carries pdeb200_comm_unique_id's 128 bytes from rank 0 to everyone (parallel.Comm.attach)
    uid = bytes(range(128)) if rank == 0 else b""
    got = comm.broadcast_bytes(uid, 128, src=0)
    assert got == bytes(range(128)), "unique-id broadcast over the host channel"
    rng = np.random.default_rng(0)                      # same nets + global batch on both ranks
    A, Cn = make_nets(rng, 1, 1, 6, 140)
    Bg = 12
    s, a, r, t, s2 = batch(rng, 1, 1, Bg)
    lo, hi = par.shard_range(Bg, rank, world)
    d = AO.DDPG(A, Cn)
    # literal Q1 under data parallelism: r-bar is the GLOBAL batch mean -> allreduce the local sum first
    rsum = torch.tensor([float(r[lo:hi].sum())], dtype=torch.float64)
    comm.allreduce_sum_(rsum)
    rbar = np.float32(rsum.item() / Bg)
    nl = hi - lo
    # local gradient with the per-sample form on r := rbar, then rescale local-mean -> global-mean
    _, gc = d.critic_loss_and_grads(s[:, lo:hi], a[:, lo:hi], np.full(nl, rbar, np.float32), t[lo:hi], s2[:, lo:hi], False)
    g = torch.from_numpy(AO.flat_grads(gc) * np.float32(nl / Bg))
    par.data_parallel_update(lambda: g, lambda _: None, comm)
    np.save(os.path.join(out_dir, "g%d.npy" % rank), g.numpy())
    if rank == 0:
        _, gref = d.critic_loss_and_grads(s, a, r, t, s2, True)
        np.save(os.path.join(out_dir, "ref.npy"), AO.flat_grads(gref))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g0, g1, ref = (np.load(tmp_path / n) for n in ("g0.npy", "g1.npy", "ref.npy"))
    assert np.array_equal(g0, g1)                        # identical on every rank after the allreduce
    assert np.allclose(g0, ref, rtol=2e-5, atol=1e-7)


def test_shard_range_partitions_exactly(pkg):
    for n in (1, 7, 8192, 8193):
        for w in (1, 2, 3, 8):
            spans = [pkg.parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
