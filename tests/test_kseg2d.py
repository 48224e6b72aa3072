"""2-D Keller-Segel (BASELINE config 3).  The reference has no 2-D model, so the pin is indirect: the 2-D
oracle / kernel on y-independent data must reproduce the 1-D oracle (itself pinned to the reference's golden
rows, tests/test_kseg.py) -- bit for bit for the oracle, to 1e-12 for the kernel."""
import numpy as np
import pytest

from conftest import relerr
from oracle import kseg2d_oracle as K2
from oracle import kseg_oracle as K1


def _pairs(rows):
    """indices t of the stored golden rows whose successor row is stored too"""
    return [t for t in range(len(rows) - 1) if rows[t + 1] == rows[t] + 1]


def _lift(y1, ny):
    """(2, nx) -> y-independent (2, nx, ny)"""
    return np.repeat(np.asarray(y1)[:, :, None], ny, axis=2)


def test_oracle_y_independent_equals_1d_oracle_bitwise(golden):
    g = golden("kseg10_16_hook")
    c1 = K1.kseg10_16_config(); c1.n_sub = 12
    c2 = K2.KSeg2DConfig(nx=100, ny=8, Lx=10.0, Ly=0.8, n_sub=12)
    t = _pairs(g["rows"])[40]
    y1, p1 = g["y"][t], g["p"][t + 1]
    # same accumulation order as the 2-D oracle for the bitwise comparison
    h = c1.dt / c1.n_sub
    ya = np.array(y1)
    for _ in range(c1.n_sub):
        k1 = K1.f(c1, ya, p1); k2 = K1.f(c1, ya + 0.5 * h * k1, p1); k3 = K1.f(c1, ya + 0.5 * h * k2, p1); k4 = K1.f(c1, ya + h * k3, p1)
        ya = ya + (h / 6) * (k1 + 2 * k2 + 2 * k3 + k4)
    yb = K2.do_step(c2, _lift(y1, 8), np.repeat(p1[:, None], 8, axis=1))
    assert np.array_equal(yb, _lift(ya, 8))
    assert relerr(ya, K1.do_step(c1, y1, p1)) < 1e-14


def test_oracle_rhs_is_second_order_accurate_in_2d():
    errs = []
    for n in (32, 64):
        c = K2.KSeg2DConfig(nx=n, ny=n, Lx=2 * np.pi, Ly=2 * np.pi)
        x = (np.arange(n) + 0.5) * c.Lx / n
        X, Y = np.meshgrid(x, x, indexing="ij")
        u, v = 1 + 0.3 * np.cos(X) * np.cos(2 * Y), 1 + 0.2 * np.cos(2 * X) * np.cos(Y)     # zero-flux compatible
        lapu, lapv = -5 * 0.3 * np.cos(X) * np.cos(2 * Y), -5 * 0.2 * np.cos(2 * X) * np.cos(Y)
        ux, uy = -0.3 * np.sin(X) * np.cos(2 * Y), -0.6 * np.cos(X) * np.sin(2 * Y)
        vx, vy = -0.4 * np.sin(2 * X) * np.cos(Y), -0.2 * np.cos(2 * X) * np.sin(Y)
        want = np.stack([lapu + u - 5.6 * (ux * vx + uy * vy) - 5.6 * u * lapv - u ** 2, lapv - v + u])
        got = K2.f(c, np.stack([u, v]), 0.0)
        errs.append(np.max(np.abs(got - want)[:, 1:-1, 1:-1]))
    assert 3.3 < errs[0] / errs[1] < 4.7, errs


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [("f64", 1e-12), ("f32", 2e-5)])
def test_gpu_y_independent_reproduces_the_1d_golden_pinned_model(pkg, golden, dtype, tol):
    g = golden("kseg10_16_hook")
    c1 = K1.kseg10_16_config(); c1.n_sub = 40
    ny = 8
    rect = K1.prepare_rectangles(c1)                                    # (20, 100) 1-D boxes
    sens2 = np.repeat(rect[:, :, None], ny, axis=2) / ny                # y-independent bases with the 1-D dot products
    pad = np.zeros((5, 100, ny)); sens2 = np.concatenate([sens2, pad])  # 25 = 5^2 sensors for the 2-D window layout
    act2 = np.repeat(rect[np.asarray(c1.actuators_to_sensors) - 1][:, :, None], ny, axis=2)
    setup = pkg.setups.KellerSegel2DSetup(nx=100, ny=ny, Lx=10.0, Ly=10.0 * ny / 100, sensors_per_axis=5,
                                          gaussians=sens2, gaussians_actuators=act2, obs_div=4.0, reward_div=800.0)
    setup.actuators_to_sensors = np.asarray(c1.actuators_to_sensors)
    pr = _pairs(g["rows"])
    rows = [pr[5], pr[len(pr) // 2], pr[-1]]
    y0 = np.stack([_lift(g["y"][t], ny) for t in rows])
    env = setup.make_env(n_envs=len(rows), dtype=dtype, y0=y0)
    a = np.concatenate([g["action"][t + 1] for t in rows])[None, :]
    env(a)
    y = env.y
    for b, t in enumerate(rows):
        want = K1.do_step(c1, g["y"][t], g["p"][t + 1])
        assert relerr(y[:, :, :, b], _lift(want, ny)) < tol
        assert np.max(np.abs(y[:, :, :, b] - y[:, :, :1, b])) == 0.0                        # stays y-independent
        assert relerr(y[:, :, 0, b], g["y"][t + 1]) < max(tol, 2e-8)                         # the reference's own row
        r = env.reward[b * 16:(b + 1) * 16]
        rw = K1.reward_function(c1, rect, want, a[:, b * 16:(b + 1) * 16], np.zeros((1, 16)))
        assert np.allclose(r, rw, rtol=tol * 100, atol=tol)
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol,nx,ny", [("f64", 1e-12, 128, 128), ("f32", 2e-5, 128, 128), ("f64", 1e-12, 96, 64)])
def test_gpu_2d_step_vs_oracle(pkg, dtype, tol, nx, ny):
    cfg = K2.KSeg2DConfig(nx=nx, ny=ny, Lx=nx * 0.1, Ly=ny * 0.1, n_sub=10, sensors_per_axis=16 if nx == 128 else 8)
    setup = pkg.setups.KellerSegel2DSetup(nx=nx, ny=ny, Lx=cfg.Lx, Ly=cfg.Ly, sensors_per_axis=cfg.sensors_per_axis,
                                          rk4_substeps=10)
    assert np.array_equal(setup.gaussians, K2.prepare_boxes(cfg))
    rng = np.random.default_rng(2)
    B = 3
    y0 = np.stack([K2.random_init(cfg, rng) for _ in range(B)])
    env = setup.make_env(n_envs=B, dtype=dtype, y0=y0)
    refs = [K2.KSeg2DEnv(cfg, y0[b], setup.gaussians) for b in range(B)]
    n_a = cfg.n_actuators
    for b in range(B):
        assert relerr(env.state[:, b * n_a:(b + 1) * n_a], refs[b].state) < tol * 10
    for step in range(2):
        a = rng.uniform(-1, 1, (1, B * n_a))
        env(a)
        y, st, r = env.y, env.state, env.reward
        for b in range(B):
            refs[b].step(a[0, b * n_a:(b + 1) * n_a])
            assert relerr(y[..., b], refs[b].y) < tol, (step, b)
            assert relerr(st[:, b * n_a:(b + 1) * n_a], refs[b].state) < tol * 10
            assert np.allclose(r[b * n_a:(b + 1) * n_a], refs[b].reward, rtol=tol * 100, atol=tol)
    assert env.state.shape == (36, B * n_a)
    env.close()
