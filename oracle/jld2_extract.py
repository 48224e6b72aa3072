#!/usr/bin/env python3
"""Golden-vector extractor for the reference's shipped JLD2 checkpoints.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Run in the build container,
where /root/reference exists; the output (tests/golden/*.npz) is committed and
is what travels to the GPU box.

JLD2 is an HDF5 subset.  We do not need the group tree: every numeric array the
fixtures need is an HDF5 *object header v2* ("OHDR") carrying
  0x01 dataspace (v2)  -> dims (reversed w.r.t. Julia's column-major dims)
  0x03 datatype        -> class (1 = IEEE float, 0 = fixed point) and size
  0x08 layout (v4)     -> class 0 compact (inline bytes) / class 1 contiguous
and the arrays appear in file order:  hook.rewards, hook.rewards_compare,
bestNNA (W1, b1, W2, b2), the bestDF columns action[1..T], p[1..T], y[1..T],
reward[1..T], then currentNNA (W1, b1, W2, b2).
(PDEhook fields: /root/reference/src/PDEhook.jl:8-31; DataFrame row written at
PDEhook.jl:51-63.)
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


import importlib

_ckpt = importlib.import_module("distributedconvrl-pde-control_b200.checkpoint")
numeric_arrays = _ckpt.numeric_arrays          # the JLD2 reader lives in the product (checkpoint compatibility)


def hook_fixture(path, n_state_rows=None):
    """Split a hook.jld2 float-array stream into named pieces."""
    arrs = [a for _, a in numeric_arrays(path)]
    # rewards, rewards_compare are 1-D float64; NNA weights are float32.
    i = 0
    rewards = arrs[i]; i += 1
    rewards_compare = arrs[i]; i += 1
    best = arrs[i:i + 4]; i += 4
    assert all(a.dtype == np.float32 for a in best), [a.dtype for a in best]
    rest = arrs[i:]
    # trailing 4 float32 arrays = currentNNA
    cur = rest[-4:]
    assert all(a.dtype == np.float32 for a in cur)
    df = rest[:-4]
    assert all(a.dtype == np.float64 for a in df), set(a.dtype for a in df)
    if len(df) % 4 == 1 and df[-1].size == 1:
        df = df[:-1]                   # trailing scalar = hook.bestreward
    assert len(df) % 4 == 0, len(df)
    T = len(df) // 4
    action = np.stack([a.reshape(-1) for a in df[0:T]])
    p = np.stack([a.reshape(-1) for a in df[T:2 * T]])
    y = np.stack(df[2 * T:3 * T])
    reward = np.stack([a.reshape(-1) for a in df[3 * T:4 * T]])
    return dict(rewards=rewards, rewards_compare=rewards_compare,
                best_W1=best[0], best_b1=best[1], best_W2=best[2], best_b2=best[3],
                cur_W1=cur[0], cur_b1=cur[1], cur_W2=cur[2], cur_b2=cur[3],
                action=action, p=p, y=y, reward=reward)


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    jobs = {
        "ks22_hook": REF / "scripts/KS/KS22/saves/hook.jld2",
        "ks200_hook": REF / "scripts/KS/KS200/saves/hook.jld2",
        "ks22_global_hook": REF / "scripts/KS/KS22_global-agent/saves/hook.jld2",
        "kseg10_16_hook": REF / "scripts/Keller-Segel/Keller-Segel10_16/saves/hook.jld2",
    }
    for name, path in jobs.items():
        fx = hook_fixture(path)
        T = fx["y"].shape[0]
        if name.startswith("kseg"):
            # 1334 rows x (2,100) is 3.5 MB; keep three consecutive windows.
            keep = np.r_[0:48, 640:664, 1300:1334]
            for k in ("action", "p", "y", "reward"):
                fx[k] = fx[k][keep]
            fx["rows"] = keep.astype(np.int64)
        else:
            fx["rows"] = np.arange(T, dtype=np.int64)
        np.savez_compressed(OUT / (name + ".npz"), **fx)
        print(name, "T=%d" % T, {k: v.shape for k, v in fx.items()})
    # actor weights only (no trajectory) for the Fluid cases
    for tag in ("8", "16", "32"):
        arrs = [a for _, a in numeric_arrays(REF / f"scripts/Fluid/Fluid_{tag}/saves/hook.jld2")]
        f32 = [a for a in arrs if a.dtype == np.float32]
        f64 = [a for a in arrs if a.dtype == np.float64]
        np.savez_compressed(OUT / f"fluid{tag}_hook.npz", rewards=f64[0],
                            best_W1=f32[0], best_b1=f32[1], best_W2=f32[2], best_b2=f32[3],
                            cur_W1=f32[-4], cur_b1=f32[-3], cur_W2=f32[-2], cur_b2=f32[-1])
        print("fluid" + tag, [a.shape for a in f32])
    y0 = [a for _, a in numeric_arrays(REF / "scripts/KS/KS22_global-agent/y0.jld2")]
    np.savez_compressed(OUT / "ks22_global_y0.npz", y0=y0[0])
    print("y0", y0[0].shape)
    # KS agent.jld2 (written by save(), KSSetup.jl:378-402): the four networks, the ADAM state of the behavior networks and
    # WINDOWS of the replay rings in logical order (oldest column first) together with the rings' raw positions
    # (CircularArrayBuffer.first / nframes as stored in the file).  KS22 never wrapped (52 224 of 150 000 columns); KS200
    # wrapped 3.48 times (522 240 columns pushed), which is what pins the capacity+1 / capacity ring layout.
    for tag, n_act in (("KS22", 8), ("KS200", 80)):
        d = _ckpt.load_agent_jld2(REF / f"scripts/KS/{tag}/saves/agent.jld2")
        rp = d["replay"]
        n_sa, n_rt = rp["state"].shape[1], rp["reward"].shape[0]
        cap = rp["capacity"]
        wins = [(0, 4096)]
        if n_rt == cap:                                           # wrapped: the raw wrap point of both rings, and the newest columns
            k_wrap = cap - rp["first_rt"]
            wins += [(k_wrap - 1024, k_wrap + 1024), (n_rt - 2048, n_rt)]
        else:
            wins += [(n_rt - 2048, n_rt)]
        fx = {"n_act": np.int64(n_act), "capacity": np.int64(cap), "n_sa": np.int64(n_sa), "n_rt": np.int64(n_rt),
              "first_sa": np.int64(rp["first_sa"]), "first_rt": np.int64(rp["first_rt"]),
              "episodes": np.int64(128), "steps_per_episode": np.int64(51),
              "windows": np.asarray(wins, dtype=np.int64),
              "terminal_sum": np.int64(rp["terminal"].sum())}
        for w, (lo, hi) in enumerate(wins):
            hi_sa = min(n_sa, hi + 3 * n_act)                     # state/action columns a fetch at ind < hi can touch
            fx["w%d_state" % w] = rp["state"][:, lo:hi_sa]
            fx["w%d_action" % w] = rp["action"][:, lo:hi_sa]
            fx["w%d_reward" % w] = rp["reward"][lo:hi]
            fx["w%d_terminal" % w] = rp["terminal"][lo:hi]
        for name, chain in d["nets"].items():
            fx["net_" + name] = chain.flat()
            fx["sizes_" + name] = np.asarray(chain.sizes, dtype=np.int64)
        for name, (m, v, bp) in d["opt"].items():
            fx["opt_m_" + name], fx["opt_v_" + name], fx["opt_betap_" + name] = m, v, bp
        np.savez_compressed(OUT / f"{tag.lower()}_agent.npz", **fx)
        print(tag, "agent:", {k: (np.shape(v)) for k, v in fx.items()})


if __name__ == "__main__":
    sys.exit(main())
