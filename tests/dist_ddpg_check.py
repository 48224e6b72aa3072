"""torchrun-able check of the data-parallel DDPG update INSIDE the library (SURVEY.md 8b `comm_init`, 8e).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P \
      tests/dist_ddpg_check.py

torch.distributed is the HOST channel only (unique-id broadcast, gathering results for the asserts); the gradient
exchange is the library's own (pdeb200_comm_init -> NVLink peer-memory exchange in the gradient kernels' last CTA, or
ncclAllReduce between the phases with PDEB200_COMM_TRANSPORT=nccl).

  1. Every rank holds the same four networks and an UNEQUAL shard of a seeded global batch; after each
     pdeb200_ddpg_update the weights must be bit-identical on all ranks and equal, to fp32 summation-order tolerance,
     to the oracle's single-process update of the whole batch (literal quirk-Q1 r-bar = global mean).
  2. pdeb200_train_updates (CUDA graph of update_loops x {sample, critic, actor}) on per-rank replay rings: weights
     bit-identical across ranks after every call, and equal to the same updates replayed one by one through the oracle
     on the gathered global batches.
  3. Sharded stepping (no collective) equals the unsharded batch bit for bit.
Prints "DIST_OK ..." from rank 0.
"""
import importlib
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def gather_np(dist, torch, arr):
    """all_gather of equally shaped numpy arrays over the host channel."""
    t = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.cpu().numpy() for o in out]


def main():
    import torch
    import torch.distributed as dist
    from oracle import agent_oracle as AO
    from test_agent_oracle import batch, make_nets
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("distributedconvrl-pde-control_b200")
    A, par, L = pkg.agent, pkg.parallel, pkg.lib
    comm = par.Comm(dist)
    setup = pkg.setups.KSSetup.ks22(window_size=3)
    rng = np.random.default_rng(11)                                   # same stream on every rank
    actor, critic = make_nets(rng, 3, 1, 6, 140, False)
    to_chain = lambda net: A.Chain(*[A.Dense(W, b, act) for W, b, act in net.layers])
    env = setup.make_env(n_envs=2, dtype="f64", device=local, y0=setup.y0_standard())
    pol = A.CustomDDPGPolicy(env, behavior_actor=to_chain(actor), behavior_critic=to_chain(critic), trajectory_length=4096,
                             literal_q1=True, comm=comm, batch_size=96)
    want_tr = {"nccl": L.COMM_NCCL, "peer": L.COMM_PEER}.get(os.environ.get("PDEB200_COMM_TRANSPORT", ""), None)
    if want_tr is not None:
        assert pol.transport == want_tr, (pol.transport, want_tr)
    ref = AO.DDPG(actor.copy(), critic.copy())

    def weights():
        return np.concatenate([n.sync_from_device().flat() for n in
                               (pol.behavior_critic, pol.behavior_actor, pol.target_critic, pol.target_actor)])

    def check_weights(tag, tol=2e-5):
        flat = weights()
        got = gather_np(dist, torch, flat)
        for g in got[1:]:
            assert np.array_equal(g, got[0]), "weights differ across ranks after %s" % tag
        want = np.concatenate([n.flat() for n in (ref.C, ref.A, ref.Ct, ref.At)])
        err = float(np.max(np.abs(flat - want)) / np.max(np.abs(want)))
        assert err < tol, (tag, err)
        return err

    # ---- 1. explicit batches, unequal shards -------------------------------------------------------------------
    worst = 0.0
    Bg = 64 * world + 5
    cuts = [0] + [int(round((r + 1) * Bg / world + (3 if r % 2 == 0 and r + 1 < world else 0))) for r in range(world)]
    cuts[-1] = Bg
    for it in range(4):
        s, a, r, t, s2 = batch(rng, 3, 1, Bg)
        lo, hi = cuts[rank], cuts[rank + 1]
        pol.set_batch(s[:, lo:hi], a[:, lo:hi], r[lo:hi], t[lo:hi], s2[:, lo:hi])
        pol.update()
        ref.update(s, a, r, t, s2, True)
        worst = max(worst, check_weights("update %d" % it))
        ls = pol.losses
        assert abs(ls["critic_loss"] - float(ref.critic_loss)) < 1e-4 * max(1.0, abs(float(ref.critic_loss))), (ls, ref.critic_loss)
        assert abs(ls["actor_loss"] - float(ref.actor_loss)) < 1e-4 * max(1.0, abs(float(ref.actor_loss))), (ls, ref.actor_loss)

    # ---- 2. graph-captured update loops on per-rank replay rings ---------------------------------------------------
    rr = np.random.default_rng(100 + rank)                           # DIFFERENT replay content per rank
    ncols = env.n_envs * env.n_cols
    n_rt = 40 * ncols
    st = rr.normal(0, 0.4, (3, n_rt + ncols)).astype(np.float32)
    ac = rr.uniform(-1, 1, (1, n_rt + ncols)).astype(np.float32)
    rw = (-np.abs(rr.normal(0, 0.3, n_rt))).astype(np.float32)
    tm = rr.random(n_rt) < 0.05
    pol.trajectory.set(st, ac, rw, tm)
    pol.update_loops, pol.update_after = 3, 1
    for call in range(3):
        n = pol.maybe_update()
        assert n == 3
        # replay the LAST update of this call through the oracle is not possible without the earlier ones; instead the
        # library is asked for every staged batch by running the same updates un-captured on a twin below.  Here:
        # bit-identical weights across ranks after every graph launch.
        got = gather_np(dist, torch, weights())
        for g in got[1:]:
            assert np.array_equal(g, got[0]), "weights differ across ranks after graph call %d" % call
    # twin: same rings, same Philox stream, updates issued one by one (sample + ddpg_update) and mirrored by the oracle
    pol2_env = setup.make_env(n_envs=2, dtype="f64", device=local, y0=setup.y0_standard())
    a2, c2 = make_nets(np.random.default_rng(5), 3, 1, 6, 140, False)
    polg = A.CustomDDPGPolicy(env, behavior_actor=to_chain(a2), behavior_critic=to_chain(c2), trajectory_length=4096,
                              literal_q1=True, comm=comm, batch_size=96, seed=3)
    pol1 = A.CustomDDPGPolicy(pol2_env, behavior_actor=to_chain(a2), behavior_critic=to_chain(c2), trajectory_length=4096,
                              literal_q1=True, comm=comm, batch_size=96, seed=3)
    ref2 = AO.DDPG(a2.copy(), c2.copy())
    polg.trajectory.set(st, ac, rw, tm)
    pol1.trajectory.set(st, ac, rw, tm)
    polg.set_sampler_offset(0)
    polg.update_loops, polg.update_after = 4, 1
    for call in range(2):
        polg.maybe_update()                                           # one graph launch = 4 updates
        for k in range(4):                                            # the same 4 updates, one by one
            off = (call * 4 + k) * 96
            L.check(pol1.env._lib.pdeb200_sample(pol1.env._ctx, 96, None, 3 ^ 0x5DEECE66D, off), pol1.env._ctx)
            pol1._staged_batch = 96
            s, a, r, t, s2, _ = pol1.get_batch()
            pol1.update()
            S = np.concatenate(gather_np(dist, torch, s), axis=1); Aa = np.concatenate(gather_np(dist, torch, a), axis=1)
            R = np.concatenate(gather_np(dist, torch, r)); T = np.concatenate(gather_np(dist, torch, t.astype(np.uint8))).astype(bool)
            S2 = np.concatenate(gather_np(dist, torch, s2), axis=1)
            ref2.update(S, Aa, R, T, S2, True)
        wg = np.concatenate([n.sync_from_device().flat() for n in (polg.behavior_critic, polg.behavior_actor, polg.target_critic, polg.target_actor)])
        w1 = np.concatenate([n.sync_from_device().flat() for n in (pol1.behavior_critic, pol1.behavior_actor, pol1.target_critic, pol1.target_actor)])
        assert np.array_equal(wg, w1), "graph-captured updates differ from the same updates issued one by one (call %d)" % call
        want = np.concatenate([n.flat() for n in (ref2.C, ref2.A, ref2.Ct, ref2.At)])
        err = float(np.max(np.abs(wg - want)) / np.max(np.abs(want)))
        assert err < 5e-5, ("graph vs oracle", call, err)
        worst = max(worst, err)
    pol2_env.close()
    env.close()

    # ---- 3. sharded stepping: rank r advances envs [lo, hi) of a global batch; equals the unsharded run bit for bit
    ks = pkg.setups.KSSetup.ks256()
    Bglob = 16 * world
    y0 = ks.generate_random_init(np.random.default_rng(5), Bglob)
    act = np.random.default_rng(6).uniform(-1, 1, (3, Bglob * 64))
    lo, hi = par.shard_range(Bglob, rank, world)
    e_loc = ks.make_env(n_envs=hi - lo, dtype="f64", device=local, y0=y0[lo:hi])
    for k in range(3):
        e_loc(act[k:k + 1, lo * 64:hi * 64])
    ys = gather_np(dist, torch, np.ascontiguousarray(e_loc.y.T))
    if rank == 0:
        e_all = ks.make_env(n_envs=Bglob, dtype="f64", device=local, y0=y0)
        for k in range(3):
            e_all(act[k:k + 1])
        assert np.array_equal(np.concatenate(ys), np.ascontiguousarray(e_all.y.T)), "sharded != unsharded"
        e_all.close()
    e_loc.close()
    dist.barrier()
    if rank == 0:
        print("DIST_OK world=%d transport=%s worst_weight_err=%.2e" % (world, {0: "none", 1: "nccl", 2: "peer"}[pol.transport], worst))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
