#!/usr/bin/env python3
"""Navier-Stokes (BASELINE config 4: 256 x 256 vorticity, fixed-step RK4 x 81) throughput probe.

Not the headline bench (bench.py measures BASELINE's metric on KS); prints one JSON line per run with
env-steps/s, ms per rhs evaluation and the algorithmic roofline numbers used in DESIGN.md.
  python tools/bench_ns.py --envs 256 --steps 3 [--dtype f32] [--nx 256] [--oversampling 81]
"""
import argparse
import importlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--spa", type=int, default=16)
    ap.add_argument("--variance", type=float, default=0.04)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--oversampling", type=int, default=None)
    ap.add_argument("--dtype", default="f64")
    args = ap.parse_args()
    import torch
    pkg = importlib.import_module("distributedconvrl-pde-control_b200")
    setup = pkg.setups.FluidSetup(nx=args.nx, sensors_per_axis=args.spa, variance=args.variance,
                                  oversampling=args.oversampling)
    rng = np.random.default_rng(0)
    base = setup.generate_random_init(rng, 8, caseno=3)
    y0 = base[np.arange(args.envs) % 8] * (1 + 1e-4 * np.arange(args.envs))[:, None, None]
    env = setup.make_env(n_envs=args.envs, dtype=args.dtype, y0=y0)
    g = np.load(ROOT / "tests" / "golden" / "fluid16_hook.npz")
    A = pkg.agent
    if args.spa == 16:
        chain = A.Chain(A.Dense(g["best_W1"], g["best_b1"], "relu"), A.Dense(g["best_W2"], g["best_b2"], "tanh"))
        A.CustomNeuralNetworkApproximator(env, pkg.lib.NET_BEHAVIOR_ACTOR, chain)
        step = lambda: env.rollout(1)
    else:
        step = lambda: (env.put(pkg.lib.ARR_ACTION_IN, np.zeros(args.envs * env.n_actuators)), env.step_device())
    for _ in range(args.warmup):
        step()
    env.synchronize()
    l0 = env.launch_count
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    env.synchronize()
    dt = time.perf_counter() - t0
    bytes_env, flops_env = env.step_cost()
    n_rhs = 4 * setup.oversampling
    out = {"workload": "NS %dx%d, %d envs, RK4 x %d, %s" % (args.nx, args.nx, args.envs, setup.oversampling, args.dtype),
           "env_steps_per_s": args.envs * args.steps / dt, "ms_per_step": 1e3 * dt / args.steps,
           "us_per_rhs_per_env": 1e6 * dt / args.steps / n_rhs / args.envs,
           "launches_per_step": (env.launch_count - l0) / args.steps,
           "algorithmic_bytes_per_env_step": bytes_env, "algorithmic_flops_per_env_step": flops_env,
           "achieved_tflops": flops_env * args.envs * args.steps / dt / 1e12,
           "finite": bool(np.all(np.isfinite(env.reward)))}
    print(json.dumps(out))
    env.close()


if __name__ == "__main__":
    main()
