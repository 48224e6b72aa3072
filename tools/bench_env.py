#!/usr/bin/env python3
"""Throughput probe for the Keller-Segel back-ends (BASELINE config 3 and its reference-native 1-D form).
  python tools/bench_env.py --problem kseg2d --envs 2048 --steps 5 [--dtype f32]
Prints one JSON line (env-steps/s, algorithmic bytes/flops, achieved GB/s and TFLOP/s)."""
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
    ap.add_argument("--problem", default="kseg2d", choices=["kseg1d", "kseg2d"])
    ap.add_argument("--envs", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--dtype", default="f64")
    args = ap.parse_args()
    pkg = importlib.import_module("distributedconvrl-pde-control_b200")
    rng = np.random.default_rng(0)
    if args.problem == "kseg1d":
        setup = pkg.setups.KellerSegelSetup()
        y0 = setup.generate_random_init(rng, args.envs)
    else:
        setup = pkg.setups.KellerSegel2DSetup()
        x = np.arange(1, 129) * 0.1
        base = 1 + 0.3 * np.outer(np.sin(x / 2.0), np.cos(x / 3.0))
        y0 = np.stack([np.stack([base * (1 + 1e-4 * b), 1.01 * base]) for b in range(args.envs)])
    env = setup.make_env(n_envs=args.envs, dtype=args.dtype, y0=y0)
    act = rng.uniform(-1, 1, (1, args.envs * env.n_actuators))
    env.put(pkg.lib.ARR_ACTION_IN, act.T)
    for _ in range(args.warmup):
        env.step_device()
    env.synchronize()
    l0 = env.launch_count
    t0 = time.perf_counter()
    for _ in range(args.steps):
        env.step_device()
    env.synchronize()
    dt = time.perf_counter() - t0
    b, f = env.step_cost()
    print(json.dumps({"workload": "%s, %d envs, %s" % (args.problem, args.envs, args.dtype),
                      "env_steps_per_s": args.envs * args.steps / dt, "ms_per_step": 1e3 * dt / args.steps,
                      "launches_per_step": (env.launch_count - l0) / args.steps, "algorithmic_bytes_per_env_step": b,
                      "algorithmic_flops_per_env_step": f, "achieved_gbs": b * args.envs * args.steps / dt / 1e9,
                      "achieved_tflops": f * args.envs * args.steps / dt / 1e12,
                      "finite": bool(np.all(np.isfinite(env.reward)))}))
    env.close()


if __name__ == "__main__":
    main()
