"""Helper for tests/test_glue_paths_gpu.py: run a few env steps (given actions, then the fused actor) for the KS,
Keller-Segel and KS-with-window configurations and dump the arrays the glue kernels write.  The generic / shape-
specialised actuation kernel is chosen by the PDEB200_ACTUATE_GENERIC environment variable of THIS process (the
library reads it once)."""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def main(out):
    pkg = importlib.import_module("distributedconvrl-pde-control_b200")
    A, L = pkg.agent, pkg.lib
    g = np.load(ROOT / "tests" / "golden" / "ks200_hook.npz")
    res = {}
    rng = np.random.default_rng(7)
    for name, window in (("ks_w1", 1), ("ks_w3", 3)):
        setup = pkg.setups.KSSetup.ks256(window_size=window)
        B = 37                                   # ragged: not a multiple of the environments per CTA
        env = setup.make_env(n_envs=B, dtype="f64", y0=setup.generate_random_init(np.random.default_rng(1), B))
        act = rng.uniform(-1, 1, (1, B * env.n_actuators))
        env(act)
        res[name + "_given_y"], res[name + "_given_state"] = env.y.copy(), env.state.copy()
        res[name + "_given_reward"], res[name + "_given_action"] = env.reward.copy(), env.action.copy()
        if window == 1:
            chain = A.Chain(A.Dense(g["best_W1"], g["best_b1"], "relu"), A.Dense(g["best_W2"], g["best_b2"], "tanh"))
        else:
            r2 = np.random.default_rng(3)
            chain = A.Chain(A.Dense(r2.normal(0, 0.5, (6, 3)).astype(np.float32), r2.normal(0, 0.1, 6).astype(np.float32), "relu"),
                            A.Dense(r2.normal(0, 0.5, (1, 6)).astype(np.float32), r2.normal(0, 0.1, 1).astype(np.float32), "tanh"))
        A.CustomNeuralNetworkApproximator(env, L.NET_BEHAVIOR_ACTOR, chain)
        env.rollout(3)
        res[name + "_actor_y"], res[name + "_actor_state"] = env.y.copy(), env.state.copy()
        res[name + "_actor_reward"], res[name + "_actor_action"] = env.reward.copy(), env.action.copy()
        res[name + "_actor_delta"] = env.delta_action.copy()
        env.close()
    np.savez(out, **res)


if __name__ == "__main__":
    main(sys.argv[1])
