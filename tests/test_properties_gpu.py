"""Size-independent properties at BASELINE.json's full sizes (the oracle is too slow there) and ragged batches:
environments are independent, so (i) any environment of a full batch must equal the same environment stepped in a
small batch, (ii) permuting the batch permutes the outputs, (iii) spot environments match the oracle; odd /
single-environment batches exercise the two-environments-per-FFT packing of the KS kernel.

Keller-Segel and Navier-Stokes environments never share arithmetic, so (i)/(ii) hold BIT FOR BIT.  The KS kernel
advances two environments as the real and imaginary part of one complex FFT: the complex butterflies mix the two
in floating point, so an environment's round-off depends on its partner and (i)/(ii) hold to fp64 round-off
(1e-12) instead."""
import numpy as np
import pytest

from conftest import relerr
from oracle import ks_oracle as K

pytestmark = pytest.mark.gpu


def _ks_actor(pkg, golden, env):
    g = golden("ks200_hook")
    A = pkg.agent
    chain = A.Chain(A.Dense(g["best_W1"], g["best_b1"], "relu"), A.Dense(g["best_W2"], g["best_b2"], "tanh"))
    A.CustomNeuralNetworkApproximator(env, pkg.lib.NET_BEHAVIOR_ACTOR, chain)


def test_ks_full_batch_8192_independence_permutation_and_oracle_spots(pkg, golden):
    setup = pkg.setups.KSSetup.ks256()
    cfg = K.ks256_config(1)
    B = 8192
    rng = np.random.default_rng(0)
    y0 = setup.generate_random_init(rng, B)
    act = rng.uniform(-1, 1, (2, B * 64))
    env = setup.make_env(n_envs=B, dtype="f64", y0=y0)
    for k in range(2):
        env(act[k:k + 1])
    y, st, rw = env.y.copy(), env.state.copy(), env.reward.copy()
    env.close()
    # (i) + (iii): a few environments alone / against the oracle
    pick = np.array([0, 1, 4095, 4096, 8190, 8191, 1234])
    small = setup.make_env(n_envs=len(pick), dtype="f64", y0=y0[pick])
    cols = (pick[:, None] * 64 + np.arange(64)[None, :]).reshape(-1)
    for k in range(2):
        small(act[k:k + 1, cols])
    assert relerr(small.y, y[:, pick]) < 1e-12
    assert relerr(small.state, st[:, cols]) < 1e-12 and np.allclose(small.reward, rw[cols], rtol=1e-10, atol=1e-13)
    small.close()
    for b in (0, 8191):
        ref = K.KSEnv(cfg, y0=y0[b])
        for k in range(2):
            ref.step(act[k:k + 1, b * 64:(b + 1) * 64])
        assert relerr(y[:, b], ref.y) < 1e-12
        assert np.allclose(rw[b * 64:(b + 1) * 64], ref.reward, rtol=1e-10, atol=1e-13)
    # (ii) permutation
    perm = rng.permutation(B)
    pcols = (perm[:, None] * 64 + np.arange(64)[None, :]).reshape(-1)
    envp = setup.make_env(n_envs=B, dtype="f64", y0=y0[perm])
    for k in range(2):
        envp(act[k:k + 1, pcols])
    assert relerr(envp.y, y[:, perm]) < 1e-12 and np.allclose(envp.reward, rw[pcols], rtol=1e-10, atol=1e-13)
    envp.close()


@pytest.mark.parametrize("B", [1, 3, 5, 31])
def test_ks_ragged_batches(pkg, B):
    setup = pkg.setups.KSSetup.ks200()
    cfg = K.ks200_config()
    rng = np.random.default_rng(B)
    y0 = setup.generate_random_init(rng, B)
    a = rng.uniform(-1, 1, (1, B * 80))
    env = setup.make_env(n_envs=B, dtype="f64", y0=y0)
    env(a)
    for b in range(B):
        ref = K.KSEnv(cfg, y0=y0[b])
        ref.step(a[:, b * 80:(b + 1) * 80])
        assert relerr(env.y[:, b], ref.y) < 1e-12
        assert relerr(env.state[:, b * 80:(b + 1) * 80], ref.state) < 1e-12
    env.close()


def test_ks_rollout_full_batch_matches_small_batch(pkg, golden):
    """The benchmarked call (pdeb200_rollout, fused actor) at 8192 envs vs the same environments in a batch of 6."""
    setup = pkg.setups.KSSetup.ks256()
    rng = np.random.default_rng(3)
    y0 = setup.generate_random_init(rng, 8192)
    env = setup.make_env(n_envs=8192, dtype="f64", y0=y0)
    _ks_actor(pkg, golden, env)
    env.rollout(3)
    pick = np.array([0, 77, 4097, 8000, 8190, 8191])
    small = setup.make_env(n_envs=6, dtype="f64", y0=y0[pick])
    _ks_actor(pkg, golden, small)
    small.rollout(3)
    # the fp32 actor amplifies the pairing round-off of the state to ~1 ulp of fp32 in the actions
    assert relerr(small.y, env.y[:, pick]) < 1e-6
    cols = (pick[:, None] * 64 + np.arange(64)[None, :]).reshape(-1)
    assert np.max(np.abs(small.action - env.action[:, cols])) < 1e-5
    env.close(); small.close()


def test_kseg2d_full_batch_2048_independence(pkg):
    from oracle import kseg2d_oracle as K2
    setup = pkg.setups.KellerSegel2DSetup(rk4_substeps=6)
    cfg = K2.KSeg2DConfig(n_sub=6)
    rng = np.random.default_rng(1)
    base = [K2.random_init(cfg, rng) for _ in range(4)]
    B = 2048
    y0 = np.stack([base[b % 4] * (1 + 1e-3 * (b // 4) / 512) for b in range(B)])
    a = rng.uniform(-1, 1, (1, B * 256))
    env = setup.make_env(n_envs=B, dtype="f64", y0=y0)
    env(a)
    pick = np.array([0, 1023, 2047])
    small = setup.make_env(n_envs=3, dtype="f64", y0=y0[pick])
    cols = (pick[:, None] * 256 + np.arange(256)[None, :]).reshape(-1)
    small(a[:, cols])
    assert np.array_equal(small.y, env.y[..., pick])
    assert np.array_equal(small.reward, env.reward[cols])
    ref = K2.KSeg2DEnv(cfg, y0[2047], setup.gaussians)
    ref.step(a[0, 2047 * 256:])
    assert relerr(env.y[..., 2047], ref.y) < 1e-12
    env.close(); small.close()


def test_ns_batch_independence_256(pkg):
    setup = pkg.setups.FluidSetup(nx=256, sensors_per_axis=16, variance=0.04, oversampling=2)
    rng = np.random.default_rng(2)
    base = setup.generate_random_init(rng, 2)
    B = 40
    y0 = np.stack([base[b % 2] * (1 + 0.01 * b) for b in range(B)])
    a = rng.uniform(-1, 1, (1, B * 256))
    env = setup.make_env(n_envs=B, dtype="f64", y0=y0)
    env(a)
    pick = np.array([0, 17, 39])
    small = setup.make_env(n_envs=3, dtype="f64", y0=y0[pick])
    cols = (pick[:, None] * 256 + np.arange(256)[None, :]).reshape(-1)
    small(a[:, cols])
    assert np.array_equal(small.y, env.y[..., pick])
    assert np.array_equal(small.state, env.state[:, cols])
    env.close(); small.close()
