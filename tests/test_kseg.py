"""Keller-Segel: oracle pinned by the reference's golden rows (CPU); CUDA kernel vs oracle and golden (GPU)."""
import numpy as np
import pytest

from conftest import relerr
from oracle import kseg_oracle as G


def _pairs(rows):
    return [t for t in range(len(rows) - 1) if rows[t + 1] == rows[t] + 1]


def test_oracle_against_golden_rows(golden):
    g = golden("kseg10_16_hook")
    cfg = G.kseg10_16_config()
    env = G.KSegEnv(cfg)
    y, p, a, r, rows = g["y"], g["p"], g["action"], g["reward"], g["rows"]
    e_ref = e_fix = 0.0
    cfg.n_sub = 40
    for n, t in enumerate(_pairs(rows)):
        assert np.array_equal(G.prepare_action(cfg, env.g_act, a[t + 1][None, :]), p[t + 1])        # exact
        rr = G.reward_function(cfg, env.g_sens, y[t + 1], a[t + 1][None, :], (a[t + 1] - a[t])[None, :])
        assert np.max(np.abs(rr - r[t + 1])) < 1e-16
        if n % 8 == 0:
            e_ref = max(e_ref, relerr(G.do_step_ref(cfg, y[t], p[t + 1]), y[t + 1]))
            e_fix = max(e_fix, relerr(G.do_step(cfg, y[t], p[t + 1]), y[t + 1]))
    # the reference integrates with an adaptive solver at 1e-8: that is the pin's resolution
    assert e_ref < 2e-8, e_ref
    assert e_fix < 2e-8, e_fix


def test_fixed_step_rk4_is_fourth_order():
    cfg = G.kseg10_16_config()
    rng = np.random.default_rng(0)
    y0 = G.generate_random_init(cfg, rng.uniform(-1, 1, 8))
    p = 3.0 * G.prepare_rectangles(cfg)[4]
    ref = G.do_step_ref(cfg, y0, p)
    errs = []
    for n in (8, 16, 32):
        cfg.n_sub = n
        errs.append(relerr(G.do_step(cfg, y0, p), ref))
    assert 10 < errs[0] / errs[1] < 24 and 10 < errs[1] / errs[2] < 24, errs


def test_adaptive_oracle_meets_its_tolerance_and_the_golden_rows(golden):
    """The adaptive controller: error vs the 1e-12 reference solution below the requested tolerance, golden
    rows at the reference solver's own 1e-8, and far fewer rhs evaluations than it would take at fixed step."""
    g = golden("kseg10_16_hook")
    cfg = G.kseg10_16_config()
    cfg.n_sub = 40
    y, p, rows = g["y"], g["p"], g["rows"]
    worst_g = worst_r = 0.0
    steps = []
    for n, t in enumerate(_pairs(rows)):
        if n % 8:
            continue
        ya, h, acc, rej = G.do_step_adaptive(cfg, y[t], p[t + 1], return_stats=True)
        worst_g = max(worst_g, relerr(ya, y[t + 1]))
        worst_r = max(worst_r, relerr(ya, G.do_step_ref(cfg, y[t], p[t + 1])))
        steps.append(acc + rej)
    assert worst_g < 1e-8 and worst_r < 2e-9, (worst_g, worst_r)
    assert max(steps) <= 20, steps          # vs 40 fixed substeps


def test_episode_length_q7():
    """te=8, dt=0.006 -> 1334 steps with Float64 clock accumulation (quirk Q7)."""
    t, n = 0.0, 0
    while not t >= 8.0:
        t += 0.006
        n += 1
    assert n == 1334


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [("f64", 1e-12), ("f32", 1e-5)])
def test_gpu_step_vs_oracle(pkg, dtype, tol):
    cfg = G.kseg10_16_config()
    cfg.n_sub = 40
    setup = pkg.setups.KellerSegelSetup()
    rng = np.random.default_rng(3)
    B = 5
    y0 = setup.generate_random_init(rng, B)
    a_prev = rng.uniform(-1, 1, (B, 16))
    a_new = rng.uniform(-1, 1, (B, 16))
    env = setup.make_env(n_envs=B, dtype=dtype, y0=y0)
    assert env.state.shape == (12, 16 * B)
    env.put(pkg.lib.ARR_ACTION, a_prev)
    env(a_new.reshape(1, -1))
    for b in range(B):
        ref = G.KSegEnv(cfg, y0=y0[b])
        ref.action = a_prev[b][None, :].copy()
        ref.step(a_new[b][None, :])
        assert relerr(env.y[:, :, b], ref.y) < tol
        assert relerr(env.p[:, b], ref.p) < tol
        assert relerr(env.state[:, b * 16:(b + 1) * 16], ref.state) < 20 * tol
        assert np.max(np.abs(env.reward[b * 16:(b + 1) * 16] - ref.reward)) < 50 * tol * max(1e-3, np.max(np.abs(ref.reward)))
    env.close()


@pytest.mark.gpu
def test_gpu_golden_rows(pkg, golden):
    """Golden (y_t, a_{t+1}) -> y_{t+1} pairs as a batch; agreement at the reference solver's own 1e-8."""
    g = golden("kseg10_16_hook")
    y, p, a, r, rows = g["y"], g["p"], g["action"], g["reward"], g["rows"]
    ts = _pairs(rows)
    B = len(ts)
    setup = pkg.setups.KellerSegelSetup()
    env = setup.make_env(n_envs=B, dtype="f64", y0=y[ts])
    env.put(pkg.lib.ARR_ACTION, a[ts])
    nxt = [t + 1 for t in ts]
    env(a[nxt].reshape(1, -1))
    assert relerr(env.y.transpose(2, 0, 1), y[nxt]) < 2e-8
    assert np.array_equal(env.p.T, p[nxt])
    assert np.max(np.abs(env.reward.reshape(B, 16) - r[nxt])) < 1e-9
    env.close()


@pytest.mark.gpu
def test_gpu_short_horizon_and_temporal_stacking(pkg):
    cfg = G.kseg10_16_config()
    cfg.n_sub = 40
    setup = pkg.setups.KellerSegelSetup()
    rng = np.random.default_rng(5)
    y0 = setup.generate_random_init(rng, 1)[0]
    env = setup.make_env(n_envs=3, dtype="f64", y0=np.stack([y0] * 3))
    ref = G.KSegEnv(cfg, y0=y0)
    for _ in range(6):
        a = rng.uniform(-1, 1, (1, 16))
        env(np.hstack([a] * 3))
        ref.step(a)
    assert relerr(env.y[:, :, 2], ref.y) < 1e-11
    assert relerr(env.state[:, 32:48], ref.state) < 1e-10        # rows 0-5 new, 6-11 previous step (temporal_steps=2)
    assert np.all(env.steps == 6)
    env.close()


@pytest.mark.gpu
def test_gpu_adaptive_mode_matches_its_oracle_the_golden_rows_and_the_fixed_step_path(pkg, golden):
    """per-environment error-controlled RK4 (the role of OrdinaryDiffEq's adaptive RK4() at 1e-8,
    KellerSegelSetup.jl:234-239).  (1) same step decisions and values as the oracle's controller, (2) golden rows at the
    reference solver's own resolution, (3) divergence from the fixed-step path (40 substeps) far below 1e-8, (4) every
    environment controls its OWN steps: a stiff environment takes more of them than a smooth one in the same launch."""
    g = golden("kseg10_16_hook")
    y, p, a, rows = g["y"], g["p"], g["action"], g["rows"]
    ts = _pairs(rows)
    B = len(ts)
    cfg = G.kseg10_16_config()
    cfg.n_sub = 40
    setup = pkg.setups.KellerSegelSetup()
    nxt = [t + 1 for t in ts]
    env_a = setup.make_env(n_envs=B, dtype="f64", y0=y[ts], adaptive=True)
    env_f = setup.make_env(n_envs=B, dtype="f64", y0=y[ts])
    for env in (env_a, env_f):
        env.put(pkg.lib.ARR_ACTION, a[ts])
        env(a[nxt].reshape(1, -1))
    ya = env_a.y.transpose(2, 0, 1)
    assert relerr(ya, y[nxt]) < 1e-8                                   # (2)
    assert relerr(ya, env_f.y.transpose(2, 0, 1)) < 1e-8               # (3) fixed vs adaptive, both inside the reference's tolerance
    sub = env_a.substeps
    assert sub.shape == (B, 2) and sub[:, 0].min() >= 1 and (sub.sum(axis=1) <= 20).all()
    for b in range(0, B, 7):                                           # (1)
        yo, h, acc, rej = G.do_step_adaptive(cfg, y[ts[b]], p[nxt[b]], return_stats=True)
        assert (acc, rej) == tuple(sub[b]), (b, acc, rej, sub[b])
        assert relerr(ya[b], yo) < 1e-12
    # (4) per-environment control: env 0 smooth (near the steady state), env 1 with a sharp front
    rng = np.random.default_rng(0)
    y0 = np.stack([np.ones((2, 100)), np.ones((2, 100))])
    y0[0] += 1e-3 * rng.standard_normal((2, 100))
    y0[1, 0, 45:55] += 6.0
    env = setup.make_env(n_envs=2, dtype="f64", y0=y0, adaptive=True, rtol=1e-10, atol=1e-10)
    env(np.zeros((1, 32)))
    s2 = env.substeps.sum(axis=1)
    assert s2[1] > s2[0], s2
    for b in range(2):
        cfg2 = G.kseg10_16_config(); cfg2.n_sub = 40
        assert relerr(env.y[:, :, b], G.do_step_adaptive(cfg2, y0[b], np.zeros(100), rtol=1e-10, atol=1e-10)) < 1e-12
    for e in (env, env_a, env_f):
        e.close()
