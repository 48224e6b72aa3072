#!/usr/bin/env python3
"""Navier-Stokes back-end under compute-sanitizer (round-2 batched FFT kernels A4 / B4, adaptive mode):

  compute-sanitizer --tool memcheck  python tools/sanitize_ns.py
  compute-sanitizer --tool racecheck python tools/sanitize_ns.py

64 x 64 (96 padded, P1 = 8), 256 x 256 (384 padded, P1 = 16: the BASELINE geometry, bulk-copy staged persistent kernel B with
more jobs than warps) and the adaptive mode; batches that do not fill the last CTA."""
import importlib
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    pkg = importlib.import_module("distributedconvrl-pde-control_b200")
    rng = np.random.default_rng(0)
    for nx, spa, var, B, over, kw in ((64, 8, 0.08, 3, 2, {}), (256, 16, 0.04, 2, 1, {}),
                                      (64, 8, 0.08, 3, 4, dict(adaptive=True, rtol=1e-3, atol=1e-3))):
        ns = pkg.setups.FluidSetup(nx=nx, sensors_per_axis=spa, variance=var, oversampling=over)
        y0 = ns.generate_random_init(rng, B, caseno=3)
        env = ns.make_env(n_envs=B, dtype="f64", y0=y0, **kw)
        env(rng.uniform(-1, 1, (1, B * env.n_actuators)))
        assert np.isfinite(np.asarray(env.y)).all()
        env.close()
    print("sanitize_ns: ok")


if __name__ == "__main__":
    main()
