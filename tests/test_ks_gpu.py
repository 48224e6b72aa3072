"""GPU parity: fused KS step kernel (through the C ABI) vs the CPU oracle and the
reference's golden trajectories.

Tolerances (BASELINE.json north_star): <= 1e-12 relative for fp64, <= 1e-5 for
fp32, single steps and short horizons; indexing bit-exact."""
import numpy as np
import pytest

from conftest import relerr
from oracle import ks_oracle as K

pytestmark = pytest.mark.gpu

TOL = {"f64": 1e-12, "f32": 1e-5}


def _oracle_step_batch(cfg, y, act_prev, act):
    """Reference env(action) per environment. y (B,nx), act (B,n_a)."""
    env = K.KSEnv(cfg)
    ys, rs, ss, ps = [], [], [], []
    for b in range(y.shape[0]):
        env.reset()
        env.y = y[b].copy()
        env.state = K.featurize(cfg, env.g_sens, env.y)
        env.action = act_prev[b][None, :].copy()
        env.step(act[b][None, :])
        ys.append(env.y); rs.append(env.reward); ss.append(env.state); ps.append(env.p)
    return np.array(ys), np.array(rs), np.array(ss), np.array(ps)


@pytest.mark.parametrize("name,make", [("ks22", "ks22"), ("ks200", "ks200")])
def test_golden_single_steps(pkg, golden, name, make):
    """50 golden (y_t, a_{t+1}) -> y_{t+1} pairs advanced as 50 environments in ONE launch."""
    g = golden(name + "_hook")
    y, p, a, r = g["y"], g["p"], g["action"], g["reward"]
    B = 50
    setup = getattr(pkg.setups.KSSetup, make)()
    env = setup.make_env(n_envs=B, dtype="f64", y0=y[:B])
    L = pkg.lib
    env.put(L.ARR_ACTION, a[:B])                    # previous action a_t (for delta_action)
    env(a[1:B + 1].reshape(1, -1))
    assert relerr(env.y.T, y[1:B + 1]) < TOL["f64"]
    assert relerr(env.p.T, p[1:B + 1]) < TOL["f64"]
    assert relerr(env.reward.reshape(B, -1), r[1:B + 1]) < 1e-11
    assert np.allclose(env.time, 0.1)
    assert np.all(env.steps == 1)
    env.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("case", ["ks256_w1", "ks256_w3", "ks200", "ks500", "ks22", "ks256_mu", "ks256_S1"])
def test_step_vs_oracle(pkg, dtype, case):
    rng = np.random.default_rng(7)
    kw = {}
    if case == "ks256_w1":
        cfg, setup = K.ks256_config(1), pkg.setups.KSSetup.ks256(window_size=1)
    elif case == "ks256_w3":
        cfg, setup = K.ks256_config(3), pkg.setups.KSSetup.ks256(window_size=3, temporal_steps=2)
        cfg.temporal_steps = 2
    elif case == "ks256_mu":
        cfg, setup = K.ks256_config(1), pkg.setups.KSSetup.ks256(mu=0.02)
        cfg.mu = 0.02
    elif case == "ks256_S1":
        cfg, setup = K.ks256_config(1), pkg.setups.KSSetup.ks256(oversampling=1)
        cfg.oversampling = 1
    elif case == "ks200":
        cfg, setup = K.ks200_config(), pkg.setups.KSSetup.ks200()
    elif case == "ks500":
        cfg = K.KSConfig(Lx=500.0, nx=600, sensor_positions=np.arange(1, 601, 3), actuators_to_sensors=np.arange(1, 201))
        setup = pkg.setups.KSSetup.ks500()
    else:
        cfg, setup = K.ks22_config(), pkg.setups.KSSetup.ks22()
    B = 7                                            # odd: exercises the half-empty last pair
    n_a = cfg.n_actuators
    y0 = setup.generate_random_init(rng, B) * 0.3
    a_prev = rng.uniform(-1, 1, (B, n_a))
    a_new = rng.uniform(-1, 1, (B, n_a))
    env = setup.make_env(n_envs=B, dtype=dtype, y0=y0)
    L = pkg.lib
    env.put(L.ARR_ACTION, a_prev)
    s0 = env.state.copy()
    env(a_new.reshape(1, -1))
    yo, ro, so, po = _oracle_step_batch(cfg, y0, a_prev, a_new)
    tol = TOL[dtype]
    assert relerr(env.y.T, yo) < tol
    assert relerr(env.p.T, po) < tol
    assert relerr(env.reward.reshape(B, -1), ro) < 20 * tol
    st = env.state                                    # (ns, n_a*B)
    so_cols = np.concatenate([so[b] for b in range(B)], axis=1)
    if cfg.temporal_steps > 1:
        # oracle featurize(prev_state=fresh state of y0): compare the freshly produced rows and the carried rows
        assert relerr(st, so_cols) < 20 * tol
    else:
        assert relerr(st, so_cols) < 20 * tol
    assert relerr(env.delta_action, (a_new - a_prev).reshape(1, -1)) < 1e-6 if dtype == "f32" else True
    env.close()


def test_observation_indexing_bit_exact(pkg):
    """state rows must be exactly sensors[(a2s[j] - i) mod n_s] * (1/max_value): integer-exact gather."""
    setup = pkg.setups.KSSetup(Lx=100.0, nx=128, sensor_positions=np.arange(1, 129, 4),
                               actuators_to_sensors=np.array([3, 1, 32, 17, 8]), window_size=5, t_samples=228)
    rng = np.random.default_rng(3)
    B = 3
    env = setup.make_env(n_envs=B, dtype="f64", y0=setup.generate_random_init(rng, B) * 0.2)
    env(np.zeros((1, 5 * B)))
    sens = env.sensors.reshape(B, 32)
    st = env.state.T.reshape(B, 5, 5)                 # [b][actuator][row]
    for b in range(B):
        for j, m in enumerate(setup.actuators_to_sensors - 1):
            for row, i in enumerate(range(-2, 3)):
                assert st[b, j, row] == sens[b, (m - i) % 32] * (1.0 / 30.0)
    env.close()


def test_episode_done_and_divergence(pkg):
    setup = pkg.setups.KSSetup.ks22()
    B = 4
    y0 = np.tile(setup.y0_standard(), (B, 1))
    y0[2] *= 70.0                                     # env 2 starts beyond max_value=30 -> diverged flag at step 1
    env = setup.make_env(n_envs=B, dtype="f64", y0=y0)
    zeros = np.zeros((1, 8 * B))
    env(zeros)
    d = env.done
    assert list(d) == [False, False, True, False]
    n = 1
    while not env.done[0]:
        env(zeros)
        n += 1
    assert n == 51                                    # quirk Q7
    assert env.steps[0] == 51
    # reset only env 0 and 2
    env.reset(mask=[1, 0, 1, 0])
    assert list(env.steps) == [0, 51, 0, 51]
    assert np.allclose(env.y[:, 0], setup.y0_standard())
    assert list(env.done) == [False, True, False, True]
    env.close()


def test_short_horizon_fp64(pkg):
    """10 consecutive env steps with time-varying actions stay within 1e-11 of the oracle."""
    cfg, setup = K.ks200_config(), pkg.setups.KSSetup.ks200()
    rng = np.random.default_rng(11)
    y0 = setup.generate_random_init(rng, 1)
    env = setup.make_env(n_envs=2, dtype="f64", y0=np.vstack([y0, y0]))
    ref = K.KSEnv(cfg, y0=y0[0])
    for t in range(10):
        a = rng.uniform(-1, 1, (1, 80))
        env(np.hstack([a, a]))
        ref.step(a)
    assert relerr(env.y[:, 0], ref.y) < 1e-11
    # the two halves of a packed pair (Re / Im of one complex sequence) agree to round-off, not bitwise
    assert relerr(env.y[:, 0], env.y[:, 1]) < 1e-12
    assert relerr(env.reward[:80], ref.reward) < 1e-10
    env.close()


def test_global_agent_variant(pkg, golden):
    """KSglobalSetup.jl: state (n_sensors, 1), reward [mean]; pinned by the KS22 global golden rows."""
    g = golden("ks22_global_hook")
    y, a, r = g["y"], g["action"], g["reward"]
    B = 50
    setup = pkg.setups.KSSetup.ks22(mono=True)
    env = setup.make_env(n_envs=B, dtype="f64", y0=y[:B])
    env.put(pkg.lib.ARR_ACTION, a[:B])
    env(a[1:B + 1].reshape(1, -1))
    assert env.state.shape == (8, B)
    assert relerr(env.y.T, y[1:B + 1]) < 1e-12
    assert relerr(env.reward, r[1:B + 1, 0]) < 1e-11
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_memory_rows_and_temporal_stacking_over_steps(pkg, dtype):
    """memory_size > 0 (KSSetup.jl:39, 220-226): the action has 1 + memory rows, the state's last `memory` rows are the
    tail rows of the action; with window 3 x temporal 2 the older block slides down each step."""
    mem = 2
    cfg = K.ks256_config(3)
    cfg.temporal_steps, cfg.memory_size = 2, mem
    setup = pkg.setups.KSSetup.ks256(window_size=3, temporal_steps=2, memory_size=mem)
    rng = np.random.default_rng(21)
    B, n_a = 3, cfg.n_actuators
    y0 = setup.generate_random_init(rng, B) * 0.3
    env = setup.make_env(n_envs=B, dtype=dtype, y0=y0)
    refs = [K.KSEnv(cfg, y0=y0[b]) for b in range(B)]
    assert env.state.shape == (3 * 2 + mem, B * n_a) and env.action.shape == (1 + mem, B * n_a)
    tol = TOL[dtype]
    for step in range(3):
        a = rng.uniform(-1, 1, (1 + mem, B * n_a))
        env(a)
        for b in range(B):
            refs[b].step(a[:, b * n_a:(b + 1) * n_a])
            assert relerr(env.y[:, b], refs[b].y) < tol
            st = env.state[:, b * n_a:(b + 1) * n_a]
            assert relerr(st, refs[b].state) < 20 * tol, (step, b)
            assert np.allclose(st[-mem:], a[1:, b * n_a:(b + 1) * n_a], rtol=1e-6, atol=0)        # memory rows = action tail
            assert np.allclose(env.reward[b * n_a:(b + 1) * n_a], refs[b].reward, rtol=200 * tol, atol=20 * tol)
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nx", [64, 128, 320, 384, 512, 1024])
def test_other_grid_sizes_of_the_factor_table(pkg, nx):
    """Every four-step factorisation the KS back-end advertises (ks.cu kFacts) against the oracle."""
    n_s = nx // 4
    cfg = K.KSConfig(Lx=nx * 200.0 / 240, nx=nx, sensor_positions=np.arange(1, nx + 1, 4), actuators_to_sensors=np.arange(1, n_s + 1),
                     t_samples=nx + 100, oversampling=5)
    setup = pkg.setups.KSSetup(Lx=cfg.Lx, nx=nx, sensor_positions=cfg.sensor_positions, actuators_to_sensors=cfg.actuators_to_sensors,
                               t_samples=nx + 100, oversampling=5)
    rng = np.random.default_rng(nx)
    B = 3
    y0 = setup.generate_random_init(rng, B) * 0.2
    a = rng.uniform(-1, 1, (1, B * n_s))
    for dtype in ("f64", "f32"):
        env = setup.make_env(n_envs=B, dtype=dtype, y0=y0)
        env(a)
        for b in range(B):
            ref = K.KSEnv(cfg, y0=y0[b])
            ref.step(a[:, b * n_s:(b + 1) * n_s])
            assert relerr(env.y[:, b], ref.y) < TOL[dtype], (nx, dtype, b)
            assert relerr(env.state[:, b * n_s:(b + 1) * n_s], ref.state) < 20 * TOL[dtype]
        env.close()
