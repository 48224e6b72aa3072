#!/usr/bin/env python3
"""Wide-network forward and DDPG update through the tcgen05 / TMA Dense-layer kernels, for
`compute-sanitizer --tool memcheck python tools/sanitize_tc.py` (ragged column counts, tile-edge widths)."""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    pkg = importlib.import_module("distributedconvrl-pde-control_b200")
    A = pkg.agent
    rng = np.random.default_rng(7)
    g = lambda o, i: ((rng.random((o, i)) - 0.5) * np.sqrt(24.0 / (o + i))).astype(np.float32)
    setup = pkg.setups.KSSetup.ks22()
    env = setup.make_env(n_envs=1, dtype="f64", y0=setup.y0_standard())
    for sizes, n_cols in (([13, 340, 340, 1], 300), ([36, 128, 128, 1], 130)):
        layers = [A.Dense(g(o, i), (0.1 * rng.standard_normal(o)).astype(np.float32), a)
                  for i, o, a in zip(sizes[:-1], sizes[1:], ["relu", "relu", None])]
        app = A.CustomNeuralNetworkApproximator(env, pkg.lib.NET_BEHAVIOR_CRITIC, A.Chain(*layers))
        x = rng.standard_normal((sizes[0], n_cols)).astype(np.float32)
        y, used = app(x, path=0, return_info=True)
        assert used >= 1 and np.isfinite(y).all()
    env.close()
    # layer-wise DDPG update with the middle layer (critic 3-140-140-1 for the KS window-1 agent: 2 -> 140 -> 140 -> 1)
    B = 3
    env = setup.make_env(n_envs=B, dtype="f32", y0=setup.generate_random_init(rng, B))
    pol = A.create_agent(env, rng=rng, nna_scale=0.6, nna_scale_critic=7.0, drop_middle_layer=False, batch_size=200,
                         start_steps=2, update_after=30, update_loops=1, trajectory_length=2000)
    n = A.run_episode(pol, env)
    assert n == 51 and np.all(np.isfinite(pol.behavior_critic.sync_from_device().flat()))
    env.close()
    print("sanitize_tc: ok")


if __name__ == "__main__":
    main()
