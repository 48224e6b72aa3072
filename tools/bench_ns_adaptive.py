#!/usr/bin/env python3
"""Navier-Stokes adaptive-step mode (the role of FluidSetup.jl's wired-in do_step2) against the fixed-step
path: env-steps/s, integrator steps taken, and the distance between the two results.
  python tools/bench_ns_adaptive.py [--envs 256] [--nx 256] [--tols 1e-4,1e-6,1e-8]"""
import argparse
import importlib
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--tols", default="1e-4,1e-6,1e-8")
    args = ap.parse_args()
    pkg = importlib.import_module("distributedconvrl-pde-control_b200")
    setup = pkg.setups.FluidSetup(nx=args.nx, sensors_per_axis=16, variance=0.04)
    rng = np.random.default_rng(0)
    base = setup.generate_random_init(rng, 8, caseno=3)
    B = args.envs
    y0 = base[np.arange(B) % 8] * (1 + 0.5 * (np.arange(B) % 5))[:, None, None]      # five flow speeds
    a = np.zeros((1, B * setup.sensors_per_axis ** 2))

    def run(**kw):
        env = setup.make_env(n_envs=B, dtype="f64", y0=y0, **kw)
        env(a)
        first = env.substeps.copy() if kw else None
        env.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            env(a)
        env.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        y, sub = env.y, (env.substeps.copy() if kw else None)
        env.close()
        return B / dt, y, first, sub

    thr_f, y_f, _, _ = run()
    print(json.dumps({"mode": "fixed", "substeps": setup.oversampling, "env_steps_per_s": thr_f}))
    for tol in [float(x) for x in args.tols.split(",")]:
        thr, y, first, sub = run(adaptive=True, rtol=tol, atol=tol)
        d = np.abs(y - y_f).reshape(-1, B).max(0) / np.abs(y_f).reshape(-1, B).max(0)
        print(json.dumps({"mode": "adaptive", "tol": tol, "env_steps_per_s": thr,
                          "accepted_first_step_min_max": [int(first[:, 0].min()), int(first[:, 0].max())],
                          "accepted_min_max": [int(sub[:, 0].min()), int(sub[:, 0].max())],
                          "rejected_max": int(sub[:, 1].max()),
                          "max_rel_distance_to_fixed_step": float(d.max())}))


if __name__ == "__main__":
    main()
