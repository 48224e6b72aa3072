"""Navier-Stokes back-end through the C ABI vs the oracle (oracle/ns_oracle.py; parity unpinned -- the
reference ships no NS trajectory -- so the oracle itself is held by the invariants of test_ns_oracle.py and
the same invariants are asserted on the CUDA path here)."""
import numpy as np
import pytest

from conftest import relerr
from oracle import ns_oracle as NS

pytestmark = pytest.mark.gpu


def _pair(pkg, nx, spa, variance, B, dtype, seed=0, ifpad=1, oversampling=None, **kw):
    cfg = NS.NSConfig(nx=nx, sensors_per_axis=spa, variance=variance, ifpad=ifpad, oversampling=oversampling, **kw)
    setup = pkg.setups.FluidSetup(nx=nx, sensors_per_axis=spa, variance=variance, ifpad=ifpad,
                                  oversampling=oversampling, **kw)
    rng = np.random.default_rng(seed)
    y0 = setup.generate_random_init(rng, B, caseno=3)
    env = setup.make_env(n_envs=B, dtype=dtype, y0=y0)
    refs = [NS.NSEnv(cfg, y0=y0[b]) for b in range(B)]
    return cfg, setup, env, refs, rng


def test_setup_bases_match_oracle(pkg):
    cfg = NS.NSConfig(nx=64, sensors_per_axis=8, variance=0.08)
    ops = NS.NSOperators(cfg)
    setup = pkg.setups.FluidSetup(nx=64, sensors_per_axis=8, variance=0.08)
    assert np.array_equal(setup.gaussians, NS.prepare_gaussians(cfg, ops, 1))
    assert np.array_equal(setup.gaussians_actuators, NS.prepare_gaussians(cfg, ops, 2))
    y = setup.ic(3, np.random.default_rng(5))
    assert relerr(y, NS.ic(cfg, ops, 3, np.random.default_rng(5))) < 1e-13


@pytest.mark.parametrize("nx,spa,var,dtype,tol", [
    (64, 8, 0.08, "f64", 1e-12), (64, 8, 0.08, "f32", 2e-5),
    (128, 16, 0.04, "f64", 1e-12), (128, 8, 0.08, "f32", 2e-5),
])
def test_reset_and_step_vs_oracle(pkg, nx, spa, var, dtype, tol):
    B = 3
    cfg, setup, env, refs, rng = _pair(pkg, nx, spa, var, B, dtype, oversampling=6)
    n_a = cfg.n_actuators
    # reset!: state from featurize(y0), exact window indexing
    st = env.state
    for b in range(B):
        assert relerr(st[:, b * n_a:(b + 1) * n_a], refs[b].state) < tol * 10
    for step in range(2):
        a = rng.uniform(-1, 1, (1, B * n_a))
        env(a)
        y, p, st, r = env.y, env.p, env.state, env.reward
        for b in range(B):
            refs[b].step(a[:, b * n_a:(b + 1) * n_a])
            assert relerr(p[:, :, b], refs[b].p) < tol, (step, b)
            assert relerr(y[:, :, b], refs[b].y) < tol, (step, b)
            assert relerr(st[:, b * n_a:(b + 1) * n_a], refs[b].state) < tol * 10
            assert np.allclose(r[b * n_a:(b + 1) * n_a], refs[b].reward, rtol=tol * 100, atol=tol)
    assert np.array_equal(env.steps, np.full(B, 2))
    env.close()


def test_full_oversampling_nx256_single_env(pkg):
    """Evaluation grid of the reference (nx = 256 -> 384^2 padded transforms, 81 RK4 substeps)."""
    cfg, setup, env, refs, rng = _pair(pkg, 256, 16, 0.04, 1, "f64", seed=2)
    assert cfg.oversampling == 81 and setup.oversampling == 81
    a = rng.uniform(-1, 1, (1, cfg.n_actuators))
    env(a)
    refs[0].step(a)
    assert relerr(env.y[:, :, 0], refs[0].y) < 1e-12
    assert relerr(env.state, refs[0].state) < 1e-11
    env.close()


def test_no_padding_path(pkg):
    cfg, setup, env, refs, rng = _pair(pkg, 64, 8, 0.08, 2, "f64", seed=4, ifpad=0, oversampling=5)
    a = rng.uniform(-1, 1, (1, 2 * cfg.n_actuators))
    env(a)
    for b in range(2):
        refs[b].step(a[:, b * cfg.n_actuators:(b + 1) * cfg.n_actuators])
        assert relerr(env.y[:, :, b], refs[b].y) < 1e-12
    env.close()


def test_non_hermitian_state_is_treated_like_the_reference(pkg):
    """pad()/chop() leave a non-Hermitian Nyquist row/column in omega_hat and real(ifft(.)) silently drops
    anti-Hermitian content; an arbitrary complex state must evolve exactly as in the reference."""
    cfg = NS.NSConfig(nx=64, sensors_per_axis=8, variance=0.08, oversampling=3)
    setup = pkg.setups.FluidSetup(nx=64, sensors_per_axis=8, variance=0.08, oversampling=3)
    rng = np.random.default_rng(9)
    y0 = setup.ic(3, rng)
    y0 = y0 + 0.05 * np.abs(y0).max() * (rng.standard_normal(y0.shape) + 1j * rng.standard_normal(y0.shape)) \
        * np.exp(-0.02 * (np.abs(np.fft.fftfreq(64, 1 / 64))[:, None] ** 2 + np.abs(np.fft.fftfreq(64, 1 / 64))[None, :] ** 2))
    env = setup.make_env(n_envs=1, dtype="f64", y0=y0)
    ref = NS.NSEnv(cfg, y0=y0)
    a = rng.uniform(-1, 1, (1, cfg.n_actuators))
    env(a)
    ref.step(a)
    assert relerr(env.y[:, :, 0], ref.y) < 1e-12
    assert relerr(env.state, ref.state) < 1e-11
    env.close()


def test_taylor_green_decay_on_device(pkg):
    """Exact Navier-Stokes solution: advection vanishes, each RK4 substep multiplies by R(-nu k^2 h)."""
    setup = pkg.setups.FluidSetup(nx=64, sensors_per_axis=8, variance=0.08, nu=0.01, oversampling=7)
    k = 2 * np.pi * 2
    y0 = np.fft.fft2(np.cos(k * setup.xx) * np.cos(k * setup.yy))
    env = setup.make_env(n_envs=2, dtype="f64", y0=y0)
    env(np.zeros((1, 2 * 64)))
    z = -0.01 * 2 * k * k * setup.dt / 7
    amp = (1 + z + z * z / 2 + z ** 3 / 6 + z ** 4 / 24) ** 7
    assert relerr(env.y[:, :, 1], amp * y0) < 1e-12
    env.close()


def test_fused_actor_rollout_matches_host_loop(pkg):
    """pdeb200_rollout (actor fused in actuate_kernel) == policy_act -> step_device, 2-D windows (9 -> 18 -> 1)."""
    g = np.load(__import__("conftest").GOLDEN / "fluid16_hook.npz")
    setup = pkg.setups.FluidSetup(nx=64, sensors_per_axis=16, variance=0.04, oversampling=3)
    rng = np.random.default_rng(1)
    y0 = setup.generate_random_init(rng, 2)
    envs = [setup.make_env(n_envs=2, dtype="f64", y0=y0) for _ in range(2)]
    A = pkg.agent
    for e in envs:
        chain = A.Chain(A.Dense(g["best_W1"], g["best_b1"], "relu"), A.Dense(g["best_W2"], g["best_b2"], "tanh"))
        A.CustomNeuralNetworkApproximator(e, pkg.lib.NET_BEHAVIOR_ACTOR, chain)
    envs[0].rollout(2)
    for _ in range(2):
        envs[1].policy_act()
        envs[1].step_device()
    envs[1].synchronize()
    assert np.array_equal(envs[0].y, envs[1].y)
    assert np.array_equal(envs[0].state, envs[1].state)
    for e in envs:
        e.close()


def test_adaptive_mode_matches_its_oracle_and_the_fixed_step_path(pkg):
    """SURVEY 8f row 4, Fluid: per-environment error-controlled RK4 in the role of the wired-in `do_step2`
    (FluidSetup.jl:178-186).  (1) at the shipped tolerance 1e0 the device takes the same step decisions as
    oracle/ns_oracle.py::do_step_adaptive (a handful of steps instead of `oversampling`), lands on its state to 1e-12, and
    warm-starts the next env step from the last accepted size; (2) at a tight tolerance the same holds with hundreds of
    steps and a few rejections, the faster environment takes more steps than its neighbours (per-environment control), and
    the result agrees with a finely resolved fixed-step run of the same ODE."""
    nx, spa, var, B = 64, 8, 0.08, 3
    cfg = NS.NSConfig(nx=nx, sensors_per_axis=spa, variance=var, oversampling=20)
    ops = NS.NSOperators(cfg)
    setup = pkg.setups.FluidSetup(nx=nx, sensors_per_axis=spa, variance=var, oversampling=20)
    rng = np.random.default_rng(11)
    y0 = setup.generate_random_init(rng, B, caseno=3)
    y0[1] *= 4.0                                       # a faster flow: needs smaller steps
    n_a = cfg.n_actuators
    a = rng.uniform(-1, 1, (1, B * n_a))
    g_act = NS.prepare_gaussians(cfg, ops, 2)
    p = [NS.prepare_action(cfg, g_act, a[:, b * n_a:(b + 1) * n_a]) for b in range(B)]
    # (1) the reference's tolerance, two env steps (the second one warm-started)
    env = setup.make_env(n_envs=B, dtype="f64", y0=y0, adaptive=True, rtol=1.0, atol=1.0)
    env(a)
    y_dev, nsub = env.y, env.substeps
    mid = []
    for b in range(B):
        yo, h, acc, rej = NS.do_step_adaptive(cfg, ops, y0[b], p[b], rtol=1.0, atol=1.0, return_stats=True)
        assert (nsub[b, 0], nsub[b, 1]) == (acc, rej), (b, nsub[b], acc, rej)
        assert acc < 20 and relerr(y_dev[:, :, b], yo) < 1e-12
        mid.append((yo, h))
    env(a)
    y_dev, nsub = env.y, env.substeps
    for b in range(B):
        yo, h, acc, rej = NS.do_step_adaptive(cfg, ops, mid[b][0], p[b], rtol=1.0, atol=1.0, h0=mid[b][1], return_stats=True)
        assert (nsub[b, 0], nsub[b, 1]) == (acc, rej) and acc <= 4, (b, nsub[b], acc, rej)
        assert relerr(y_dev[:, :, b], yo) < 1e-12
    env.close()
    # (2) tight tolerance; fixed-step comparison run with 4x the substeps (its own error is then ~1e-8 of the fast flow)
    tol = 1e-9
    env_a = setup.make_env(n_envs=B, dtype="f64", y0=y0, adaptive=True, rtol=tol, atol=tol)
    fine = pkg.setups.FluidSetup(nx=nx, sensors_per_axis=spa, variance=var, oversampling=80)
    env_f = fine.make_env(n_envs=B, dtype="f64", y0=y0)
    env_a(a); env_f(a)
    ya, yf, ns_ = env_a.y, env_f.y, env_a.substeps
    for b in range(B):
        yo, h, acc, rej = NS.do_step_adaptive(cfg, ops, y0[b], p[b], rtol=tol, atol=tol, return_stats=True)
        assert (ns_[b, 0], ns_[b, 1]) == (acc, rej), (b, ns_[b], acc, rej)
        assert relerr(ya[:, :, b], yo) < 1e-11
        assert relerr(ya[:, :, b], yf[:, :, b]) < 1e-7
    assert ns_[1, 0] > 2 * ns_[0, 0] and ns_[:, 1].sum() > 0       # per-environment control, some rejected attempts
    env_a.close(); env_f.close()


def test_round1_kernels_stay_in_step_with_the_batched_ones(pkg, monkeypatch):
    """PDEB200_NS_LEGACY=1 selects the one-line-per-warp kernels A / B (fft_pass.cuh); the default batched in-place kernels
    (fft_batch.cuh) must give the same states up to summation order, on a padded 16 x 12 (192-point) and a 8 x 12 grid."""
    for nx, spa, var in ((128, 16, 0.04), (64, 8, 0.08)):
        setup = pkg.setups.FluidSetup(nx=nx, sensors_per_axis=spa, variance=var)     # the scripts' substep count (a
        rng = np.random.default_rng(21)                                                 # marginal one amplifies round-off)
        y0 = setup.generate_random_init(rng, 2, caseno=3)
        a = rng.uniform(-1, 1, (1, 2 * spa * spa))
        out = []
        for legacy in ("1", "0"):
            monkeypatch.setenv("PDEB200_NS_LEGACY", legacy)
            env = setup.make_env(n_envs=2, dtype="f64", y0=y0)
            env(a); env(a)
            out.append((env.y.copy(), env.reward.copy(), env.state.copy()))
            env.close()
        assert relerr(out[0][0], out[1][0]) < 1e-12
        assert np.allclose(out[0][1], out[1][1], rtol=1e-11, atol=1e-13)
        assert relerr(out[0][2], out[1][2]) < 1e-12
