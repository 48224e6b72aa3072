#!/usr/bin/env python3
"""One small step of every back-end, meant to run under compute-sanitizer:

  compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
  compute-sanitizer --tool racecheck python tools/sanitize_smoke.py

Sizes are ragged on purpose (batches that do not fill the last CTA / warp / pair) and small enough for the
sanitizer's slowdown; the oversampling is reduced where the kernels' code paths do not depend on it."""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    pkg = importlib.import_module("distributedconvrl-pde-control_b200")
    A, L = pkg.agent, pkg.lib
    rng = np.random.default_rng(0)
    g = np.load(ROOT / "tests" / "golden" / "ks200_hook.npz")
    # KS: given actions, then the fused actor (specialised actuate / observe / policy kernels), odd batch
    setup = pkg.setups.KSSetup.ks256(oversampling=3)
    B = 37
    env = setup.make_env(n_envs=B, dtype="f64", y0=setup.generate_random_init(rng, B))
    env(rng.uniform(-1, 1, (1, B * env.n_actuators)))
    chain = A.Chain(A.Dense(g["best_W1"], g["best_b1"], "relu"), A.Dense(g["best_W2"], g["best_b2"], "tanh"))
    A.CustomNeuralNetworkApproximator(env, L.NET_BEHAVIOR_ACTOR, chain)
    env.rollout(2)
    assert np.isfinite(env.y).all() and np.isfinite(env.reward).all()
    env.close()
    # Keller-Segel 1-D: 5 environments per CTA, batch not a multiple of 5
    ks = pkg.setups.KellerSegelSetup(rk4_substeps=4)
    env = ks.make_env(n_envs=7, dtype="f64", y0=ks.generate_random_init(rng, 7))
    env(rng.uniform(-1, 1, (1, 7 * env.n_actuators)))
    assert np.isfinite(env.y).all()
    env.close()
    # Keller-Segel 2-D: the compile-time 128 x 128 instantiation (clusters, DSMEM halos)
    k2 = pkg.setups.KellerSegel2DSetup(rk4_substeps=2)
    x = np.arange(1, 129) * 0.1
    base = 1 + 0.3 * np.outer(np.sin(x / 2.0), np.cos(x / 3.0))
    y0 = np.stack([np.stack([base * (1 + 1e-3 * b), 1.01 * base]) for b in range(3)])
    env = k2.make_env(n_envs=3, dtype="f64", y0=y0)
    env(rng.uniform(-1, 1, (1, 3 * env.n_actuators)))
    assert np.isfinite(env.y).all()
    env.close()
    # Navier-Stokes 64 x 64 (96 x 96 padded), two RK4 substeps
    ns = pkg.setups.FluidSetup(nx=64, sensors_per_axis=8, variance=0.08, oversampling=2)
    y0 = ns.generate_random_init(rng, 3, caseno=3)
    env = ns.make_env(n_envs=3, dtype="f64", y0=y0)
    env(rng.uniform(-1, 1, (1, 3 * env.n_actuators)))
    assert np.isfinite(np.asarray(env.y)).all()
    env.close()
    # agent: device replay rings (with wrap), sampler, fused DDPG update kernels, ADAM / Polyak
    setup = pkg.setups.KSSetup.ks22(oversampling=2)
    B = 3
    env = setup.make_env(n_envs=B, dtype="f32", y0=setup.generate_random_init(rng, B))
    pol = A.create_agent(env, rng=rng, nna_scale=0.6, nna_scale_critic=7.0, drop_middle_layer=True, batch_size=40,
                         start_steps=2, update_after=3, update_loops=2, trajectory_length=600)
    n = A.run_episode(pol, env)
    assert n == 51 and np.all(np.isfinite(pol.behavior_actor.sync_from_device().flat()))
    env.close()
    print("sanitize_smoke: ok")


if __name__ == "__main__":
    main()
