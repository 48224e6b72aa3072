"""torchrun-able check of the data-parallel DDPG update over NCCL.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P \
      tests/dist_ddpg_check.py

Every rank holds the same four networks and a SHARD of a seeded global batch; after each update (critic grads ->
allreduce -> ADAM; actor grads -> allreduce -> ADAM + Polyak; literal-Q1 r-bar from an allreduced sum) the weights
must be bit-identical on all ranks and equal, to fp32 summation-order tolerance, to the oracle's single-process
update on the whole batch.  Also steps a sharded KS batch (no collective) and checks it against one context
holding the whole batch.  Prints "DIST_OK" from rank 0.
"""
import importlib
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import torch
    import torch.distributed as dist
    from oracle import agent_oracle as AO
    from test_agent_oracle import batch, make_nets
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("distributedconvrl-pde-control_b200")
    A, par = pkg.agent, pkg.parallel
    comm = par.Comm(dist)
    setup = pkg.setups.KSSetup.ks22(window_size=3)
    rng = np.random.default_rng(11)                                   # same stream on every rank
    actor, critic = make_nets(rng, 3, 1, 6, 140, False)
    to_chain = lambda net: A.Chain(*[A.Dense(W, b, act) for W, b, act in net.layers])
    env = setup.make_env(n_envs=2, dtype="f64", device=local, y0=setup.y0_standard())
    # torch's NCCL ops run on torch's current stream: the context must enqueue on the same one
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    import ctypes as C
    pkg.lib.check(env._lib.pdeb200_set_stream(env._ctx, C.c_void_p(stream.cuda_stream)), env._ctx)
    pol = A.CustomDDPGPolicy(env, behavior_actor=to_chain(actor), behavior_critic=to_chain(critic), trajectory_length=4096,
                             literal_q1=True, comm=comm)
    ref = AO.DDPG(actor.copy(), critic.copy())
    Bg = 64 * world
    worst = 0.0
    for it in range(4):
        s, a, r, t, s2 = batch(rng, 3, 1, Bg)
        lo, hi = par.shard_range(Bg, rank, world)
        pol.set_batch(s[:, lo:hi], a[:, lo:hi], r[lo:hi], t[lo:hi], s2[:, lo:hi])
        pol.update()
        ref.update(s, a, r, t, s2, True)
        flat = np.concatenate([n.sync_from_device().flat() for n in
                               (pol.behavior_critic, pol.behavior_actor, pol.target_critic, pol.target_actor)])
        gathered = [torch.zeros(flat.size, dtype=torch.float32, device="cuda") for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(flat).cuda())
        for g in gathered[1:]:
            assert torch.equal(g, gathered[0]), "weights differ across ranks after update %d" % it
        want = np.concatenate([n.flat() for n in (ref.C, ref.A, ref.Ct, ref.At)])
        err = float(np.max(np.abs(flat - want)) / np.max(np.abs(want)))
        worst = max(worst, err)
        assert err < 2e-5, (it, err)
    env.close()

    # sharded stepping: rank r advances envs [lo, hi) of a global batch; equals the unsharded run bit for bit
    ks = pkg.setups.KSSetup.ks256()
    Bglob = 16 * world
    y0 = ks.generate_random_init(np.random.default_rng(5), Bglob)
    act = np.random.default_rng(6).uniform(-1, 1, (3, Bglob * 64))
    lo, hi = par.shard_range(Bglob, rank, world)
    e_loc = ks.make_env(n_envs=hi - lo, dtype="f64", device=local, y0=y0[lo:hi])
    for k in range(3):
        e_loc(act[k:k + 1, lo * 64:hi * 64])
    y_loc = torch.from_numpy(np.ascontiguousarray(e_loc.y.T)).cuda()
    ys = [torch.zeros_like(y_loc) for _ in range(world)]
    dist.all_gather(ys, y_loc)
    if rank == 0:
        e_all = ks.make_env(n_envs=Bglob, dtype="f64", device=local, y0=y0)
        for k in range(3):
            e_all(act[k:k + 1])
        assert np.array_equal(torch.cat(ys).cpu().numpy(), np.ascontiguousarray(e_all.y.T)), "sharded != unsharded"
        e_all.close()
    e_loc.close()
    dist.barrier()
    if rank == 0:
        print("DIST_OK world=%d worst_weight_err=%.2e" % (world, worst))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
