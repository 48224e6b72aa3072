"""CPU: pin the oracle against the golden trajectories the reference ships
(scripts/*/saves/hook.jld2 -> tests/golden/*.npz via oracle/jld2_extract.py)."""
import numpy as np
import pytest

from conftest import relerr
from oracle import ks_oracle as K


@pytest.mark.parametrize("name,cfgf,nt", [("ks22", K.ks22_config, 291), ("ks200", K.ks200_config, 340)])
def test_ks_golden_rows(golden, name, cfgf, nt):
    g = golden(name + "_hook")
    cfg = cfgf()
    env = K.KSEnv(cfg)
    dx = cfg.dx
    assert len(K.julia_float_range(dx - 50 * dx, dx, cfg.Lx + 50 * dx)) == nt      # quirk Q4
    y, p, a, r = g["y"], g["p"], g["action"], g["reward"]
    assert y.shape[0] == 51                                                         # quirk Q7
    e_step = e_p = e_r = 0.0
    for t in range(50):
        pp = K.prepare_action(cfg, env.g_act, a[t + 1][None, :])
        e_p = max(e_p, relerr(pp, p[t + 1]) if np.max(np.abs(p[t + 1])) > 0 else float(np.max(np.abs(pp))))
        yn = K.do_step(cfg, env.ops, y[t], p[t + 1])
        e_step = max(e_step, relerr(yn, y[t + 1]))
        rr = K.reward_function(cfg, env.g_sens, y[t + 1], a[t + 1][None, :], (a[t + 1] - a[t])[None, :])
        e_r = max(e_r, relerr(rr, r[t + 1]))
    assert e_step < 5e-15, e_step          # tolerance: fp64 round-off of a different (pocketfft vs FFTW) DFT
    assert e_p < 1e-14, e_p
    assert e_r < 1e-13, e_r


def test_ks_global_golden_rows(golden):
    """KS22 global-agent variant (KSglobalSetup.jl): same stepper, [mean] reward."""
    g = golden("ks22_global_hook")
    cfg = K.ks22_config()
    cfg.mono = True
    env = K.KSEnv(cfg)
    y, p, a, r = g["y"], g["p"], g["action"], g["reward"]
    e_step = e_r = 0.0
    for t in range(50):
        yn = K.do_step(cfg, env.ops, y[t], p[t + 1])
        e_step = max(e_step, relerr(yn, y[t + 1]))
        rr = K.reward_function(cfg, env.g_sens, y[t + 1], a[t + 1][None, :], (a[t + 1] - a[t])[None, :])
        e_r = max(e_r, relerr(rr, r[t + 1]))
    assert e_step < 5e-15, e_step
    assert e_r < 1e-13, e_r
    y0 = golden("ks22_global_y0")["y0"]
    assert y0.shape == (192,)


def test_ks_env_episode_length():
    """te=5, dt=0.1 gives 51 steps because time accumulates in Float64 (Q7)."""
    env = K.KSEnv(K.ks22_config())
    n = 0
    while not env.done:
        env.step(np.zeros((1, 8)))
        n += 1
    assert n == 51


def test_window_rows_order():
    """rows are circshift(sensors, i) for i = -h..h  =>  [right nbr, self, left nbr]."""
    s = np.arange(10.0)
    w = K.window_rows(s, 3, np.arange(1, 11))
    assert np.array_equal(w[:, 4], [5.0, 4.0, 3.0])
    assert np.array_equal(w[:, 0], [1.0, 0.0, 9.0])


def test_product_bases_match_oracle(pkg):
    """The product's setup code rebuilds `gaussians` independently of the oracle."""
    for make, cfgf in ((pkg.setups.KSSetup.ks22, K.ks22_config), (pkg.setups.KSSetup.ks200, K.ks200_config),
                       (pkg.setups.KSSetup.ks256, K.ks256_config)):
        s = make()
        cfg = cfgf()
        env = K.KSEnv(cfg)
        assert np.array_equal(s.gaussians, env.g_sens)
        assert np.array_equal(s.gaussians_actuators, env.g_act)
