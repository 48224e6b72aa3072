#!/usr/bin/env python3
"""BASELINE config 5: KS full training loop (rollout + device replay + DDPG backward + gradient allreduce).

  python tools/bench_train.py --envs 8192 --batch 4096 --update-loops 1 --steps 40
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tools/bench_train.py ...

One env step of the loop = policy (actor + device Philox noise) -> PreAct push of (s, a) for every column ->
`update_loops` x {sample, critic grads, allreduce, ADAM, actor grads, allreduce, ADAM + Polyak} -> env step ->
PostAct push of (r, terminal); the stage order of RLCore's run() (scripts/Fluid/setup/FluidSetup.jl:455-519).
Prints one JSON line from rank 0: env-steps/s over all ranks (device events, max over ranks)."""
import argparse
import ctypes as C
import importlib
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=8192)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--update-loops", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=12)
    ap.add_argument("--middle", action="store_true", help="drop_middle_layer = false (wide critic 2-140-140-1 on tensor cores)")
    ap.add_argument("--updates-only", action="store_true", help="after the warm-up time only pdeb200_train_updates (us per DDPG update)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("distributedconvrl-pde-control_b200")
    A, L, par = pkg.agent, pkg.lib, pkg.parallel
    setup = pkg.setups.KSSetup.ks256()
    rng = np.random.default_rng(100 + rank)
    env = setup.make_env(n_envs=args.envs, dtype="f64", device=local, y0=setup.generate_random_init(rng, args.envs))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    L.check(env._lib.pdeb200_set_stream(env._ctx, C.c_void_p(stream.cuda_stream)), env._ctx)
    comm = par.Comm(dist if world > 1 else None)
    wrng = np.random.default_rng(7)                                     # identical initial weights on every rank
    pol = A.create_agent(env, rng=wrng, nna_scale=0.6, nna_scale_critic=7.0, drop_middle_layer=not args.middle,
                         batch_size=args.batch, start_steps=2, update_after=2, update_freq=1, update_loops=args.update_loops,
                         act_noise=1.2, trajectory_length=args.envs * env.n_cols * 8, seed=rank, comm=comm if world > 1 else None)
    traj = pol.trajectory
    env.reset()
    traj.pre_episode()

    def step():
        pol(env, learning=True)
        traj.pre_act()
        pol.maybe_update()
        env.step_device()
        traj.post_act()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if args.updates_only:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            pol.maybe_update()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record(stream)
        for _ in range(args.steps):
            pol.maybe_update()
        e1.record(stream)
        torch.cuda.synchronize()
        if os.environ.get("PDEB200_DDPG_TIMELINE"):
            buf = np.zeros(500 * 8, dtype=np.uint64)
            n = C.c_int32()
            L.check(env._lib.pdeb200_debug_timeline(env._ctx, buf.ctypes.data, 500, C.byref(n)), env._ctx)
            r = buf[:n.value * 8].reshape(-1, 8).astype(np.int64)
            if rank == 0 and len(r) > 8:
                r = r[-40:]
                t0 = r[0, 1]
                if r[0, 0] == 2:                 # one-cluster kernel: one record per update
                    for row in r[:8]:
                        print("update  entry %7.2f | critic loop %5.2f  sync %5.2f  reduce+exchange+adam %5.2f | actor loop (+sync) %5.2f  sync %5.2f  "
                              "reduce+exchange+adam+sync %5.2f | kernel %5.2f us" % ((row[1] - t0) / 1e3, (row[2] - row[1]) / 1e3, (row[3] - row[2]) / 1e3,
                              (row[4] - row[3]) / 1e3, (row[5] - row[4]) / 1e3, (row[6] - row[5]) / 1e3, (row[7] - row[6]) / 1e3, (row[7] - row[1]) / 1e3))
                    print("mean gap between kernels %.2f us; mean kernel %.2f us" % (((r[1:, 1] - r[:-1, 7]) / 1e3).mean(), ((r[:, 7] - r[:, 1]) / 1e3).mean()))
                    r = r[:0]
                for row in r[:12]:
                    print("phase %d  entry %7.2f  loop %5.2f  wait-for-last-CTA %5.2f  reduce %5.2f  exchange %5.2f  adam %5.2f | kernel %5.2f us"
                          % (row[0], (row[1] - t0) / 1e3, (row[2] - row[1]) / 1e3, (row[3] - row[2]) / 1e3, (row[4] - row[3]) / 1e3,
                             (row[5] - row[4]) / 1e3, (row[6] - row[5]) / 1e3, (row[6] - row[1]) / 1e3))
                gaps = (r[1:, 1] - r[:-1, 6]) / 1e3 if len(r) else np.zeros(1)
                if len(r):
                  print("mean gap between kernels %.2f us; mean critic %.2f us, actor %.2f us" % (
                    gaps.mean(), ((r[:, 6] - r[:, 1])[r[:, 0] == 0]).mean() / 1e3, ((r[:, 6] - r[:, 1])[r[:, 0] == 1]).mean() / 1e3))
        if rank == 0:
            print(json.dumps({"n_gpus": world, "batch": args.batch, "update_loops": args.update_loops,
                              "us_per_update": 1e3 * e0.elapsed_time(e1) / (args.steps * args.update_loops), "losses": pol.losses}))
        env.close()
        if world > 1:
            dist.destroy_process_group()
        return
    l0, u0 = env.launch_count, pol.n_updates
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        t = float(ms.item()) * 1e-3
        print(json.dumps({"workload": "KS N=256 training loop, %d envs/GPU, DDPG batch %d columns/GPU, update_loops %d, critic %s"
                          % (args.envs, args.batch, args.update_loops, pol.behavior_critic.model.sizes),
                          "n_gpus": world, "env_steps_per_s": args.envs * world * args.steps / t, "ms_per_loop_step": 1e3 * t / args.steps,
                          "updates_per_step": (pol.n_updates - u0) / args.steps, "launches_per_step": (env.launch_count - l0) / args.steps,
                          "losses": pol.losses}))
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
