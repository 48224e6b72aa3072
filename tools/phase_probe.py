#!/usr/bin/env python3
"""In-stream CUDA-event times of the three phases of a KS env step (actuate / core / observe), 8192 envs, fp64."""
import ctypes as C, importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
pkg = importlib.import_module("distributedconvrl-pde-control_b200"); L = pkg.lib; A = pkg.agent
setup = pkg.setups.KSSetup.ks256()
B = 8192
env = setup.make_env(n_envs=B, dtype="f64", y0=setup.generate_random_init(np.random.default_rng(0), B))
g = np.load(ROOT / "tests/golden/ks200_hook.npz")
A.CustomNeuralNetworkApproximator(env, L.NET_BEHAVIOR_ACTOR, A.Chain(A.Dense(g["best_W1"], g["best_b1"], "relu"), A.Dense(g["best_W2"], g["best_b2"], "tanh")))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
L.check(env._lib.pdeb200_set_stream(env._ctx, C.c_void_p(stream.cuda_stream)), env._ctx)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
L.check(env._lib.pdeb200_enable_step_timing(env._ctx, 1), env._ctx)
for mode in ("warm", "flush"):
    acc = []
    for i in range(30):
        if mode == "flush": flush.zero_()
        env.rollout(1)
        ph = (C.c_float * 3)(); L.check(env._lib.pdeb200_last_phase_ms(env._ctx, ph), env._ctx)
        acc.append([ph[0], ph[1], ph[2]])
    print(mode, "actuate / core / observe ms:", np.mean(np.asarray(acc)[5:], axis=0))
