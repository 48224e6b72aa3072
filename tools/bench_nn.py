#!/usr/bin/env python3
"""Wide-network operator probe: the approximator call and the DDPG update with `drop_middle_layer = false`
(critic 13 -> 340 -> 340 -> 1, the reference's NS / Keller-Segel widths) on tensor cores vs CUDA cores.
Prints one JSON line; used for profiles/ (ncu: sm__pipe_tensor_cycles_active)."""
import argparse
import ctypes as C
import importlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cols", type=int, default=524288, help="columns (= envs x actuators) of the forward call")
    ap.add_argument("--batch", type=int, default=65536, help="DDPG batch (columns)")
    ap.add_argument("--hidden", type=int, default=340)
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    import torch
    pkg = importlib.import_module("distributedconvrl-pde-control_b200")
    A, L = pkg.agent, pkg.lib
    rng = np.random.default_rng(0)
    ns, h = 12, args.hidden
    setup = pkg.setups.KSSetup.ks22(window_size=3, temporal_steps=4)
    env = setup.make_env(n_envs=1, dtype="f64", y0=setup.y0_standard())
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    L.check(env._lib.pdeb200_set_stream(env._ctx, C.c_void_p(stream.cuda_stream)), env._ctx)
    actor = A.create_chain(na=1, ns=ns, is_actor=True, rng=rng, nna_scale=2.0, drop_middle_layer=True)
    critic = A.create_chain(na=1, ns=ns, is_actor=False, rng=rng, nna_scale=h / 20.0, drop_middle_layer=False)
    pol = A.CustomDDPGPolicy(env, behavior_actor=actor, behavior_critic=critic, trajectory_length=4096)
    out = {"critic": critic.sizes, "cols": args.cols, "batch": args.batch}

    def timed(fn):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.iters):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters

    # ---- forward on device buffers -------------------------------------------------------------
    M = args.cols
    x = torch.zeros(M, 16, device="cuda", dtype=torch.float32)
    x[:, :13] = torch.randn(M, 13, device="cuda")
    y = torch.empty(M, 1, device="cuda", dtype=torch.float32)
    used = C.c_int32(0)
    for name, path in (("tensor", 0), ("cuda_core", 1)):
        ms = timed(lambda: L.check(env._lib.pdeb200_net_forward_device(env._ctx, L.NET_BEHAVIOR_CRITIC, M, C.c_void_p(x.data_ptr()), 16,
                                                                       C.c_void_p(y.data_ptr()), 1, path, C.byref(used)), env._ctx))
        flops = 2.0 * M * (13 * h + h * h + h)
        out["forward_" + name] = {"ms": ms, "tensor_layers": used.value, "tflops_fp32_equiv": flops / ms / 1e9,
                                  "columns_per_s": M / ms * 1e3}
    # ---- DDPG update ------------------------------------------------------------------------------
    B = args.batch
    s = rng.standard_normal((ns, B)).astype(np.float32); a = rng.uniform(-1, 1, (1, B)).astype(np.float32)
    r = rng.standard_normal(B).astype(np.float32); t = rng.random(B) < 0.1; s2 = rng.standard_normal((ns, B)).astype(np.float32)
    pol.set_batch(s, a, r, t, s2)
    for name, path in (("tensor", 3), ("cuda_core", 1)):
        pol.set_update_path(path)
        ms = timed(lambda: pol.update())
        # per update: 3 critic forwards + 1 more in the actor phase, 2 input-gradient and 1 weight-gradient passes of the middle layer
        flops = 2.0 * B * h * h * (4 + 2 + 1)
        out["update_" + name] = {"ms": ms, "middle_layer_tflops_fp32_equiv": flops / ms / 1e9, "columns_per_s": B / ms * 1e3}
    print(json.dumps(out))
    env.close()


if __name__ == "__main__":
    main()
