"""NS oracle (parity unpinned: no reference trajectory exists) -- analytic invariants of the
literal restatement of src/fluid_rk4.jl + scripts/Fluid/setup/FluidSetup.jl."""
import numpy as np

from oracle import ns_oracle as NS


def _cfg(nx=32, **kw):
    return NS.NSConfig(nx=nx, sensors_per_axis=4, variance=0.08, **kw)


def test_pad_chop_round_trip_and_nyquist_side():
    cfg = _cfg(16)
    ops = NS.NSOperators(cfg)
    rng = np.random.default_rng(0)
    f = rng.standard_normal((16, 16)) + 1j * rng.standard_normal((16, 16))
    fp = NS.pad(cfg, ops, f)
    assert fp.shape == (24, 24)
    assert np.array_equal(NS.chop(cfg, ops, fp), f)
    # Nyquist row (index ny/2) sits at +ny/2 of the padded grid, nothing at -ny/2 (fluid_rk4.jl:205-208)
    assert np.array_equal(fp[8, :9], f[8, :9]) and np.all(fp[24 - 8, :] == 0)
    assert np.count_nonzero(fp) == 256


def test_taylor_green_decays_with_rk4_amplification_factor():
    # omega = cos(kx x) cos(ky y): advection vanishes, d omega_hat/dt = -nu k^2 omega_hat
    cfg = _cfg(32, nu=0.01, dt=0.02, oversampling=7)
    ops = NS.NSOperators(cfg)
    m = 2
    k = 2 * np.pi * m
    om = np.cos(k * ops.xx) * np.cos(k * ops.yy)
    y0 = np.fft.fft2(om)
    adv = NS.advection(cfg, ops, y0.copy())
    assert np.max(np.abs(adv)) < 1e-9 * np.max(np.abs(y0))
    y1 = NS.do_step(cfg, ops, y0, np.zeros_like(y0))
    z = -cfg.nu * 2 * k * k * cfg.dt / cfg.oversampling
    amp = (1 + z + z * z / 2 + z ** 3 / 6 + z ** 4 / 24) ** cfg.oversampling
    assert np.max(np.abs(y1 - amp * y0)) < 1e-12 * np.max(np.abs(y0))


def test_padded_equals_unpadded_for_band_limited_fields():
    # modes |k| < N/4 -> the quadratic term has no aliasing: 3/2-rule result == plain product
    cfg1, cfg0 = _cfg(32, ifpad=1), _cfg(32, ifpad=0)
    ops = NS.NSOperators(cfg1)
    rng = np.random.default_rng(1)
    om = np.zeros((32, 32))
    for _ in range(6):
        a, b = rng.integers(-7, 8, 2)
        om += rng.standard_normal() * np.cos(2 * np.pi * (a * ops.xx + b * ops.yy) + rng.uniform(0, 6))
    yh = np.fft.fft2(om)
    a1 = NS.advection(cfg1, ops, yh.copy())
    a0 = NS.advection(cfg0, ops, yh.copy())
    assert np.max(np.abs(a1 - a0)) < 1e-11 * np.max(np.abs(a0))


def test_advection_matches_physical_space_formula():
    # -u w_x - v w_y with u = psi_y, v = -psi_x, lap(psi) = -omega (sign convention of omg2vel: psi_hat = omega_hat / k^2)
    cfg = _cfg(32, ifpad=1)
    ops = NS.NSOperators(cfg)
    k = 2 * np.pi
    om = np.sin(k * ops.xx) + 0.5 * np.cos(2 * k * ops.yy) + 0.25 * np.sin(k * (ops.xx + ops.yy))
    psi = np.sin(k * ops.xx) / k ** 2 + 0.5 * np.cos(2 * k * ops.yy) / (2 * k) ** 2 + 0.25 * np.sin(k * (ops.xx + ops.yy)) / (2 * k ** 2)
    u = -0.5 * np.sin(2 * k * ops.yy) * 2 * k / (2 * k) ** 2 + 0.25 * np.cos(k * (ops.xx + ops.yy)) * k / (2 * k ** 2)
    v = -(np.cos(k * ops.xx) * k / k ** 2 + 0.25 * np.cos(k * (ops.xx + ops.yy)) * k / (2 * k ** 2))
    wx = k * np.cos(k * ops.xx) + 0.25 * k * np.cos(k * (ops.xx + ops.yy))
    wy = -0.5 * 2 * k * np.sin(2 * k * ops.yy) + 0.25 * k * np.cos(k * (ops.xx + ops.yy))
    want = np.fft.fft2(-u * wx - v * wy)
    got = NS.advection(cfg, ops, np.fft.fft2(om))
    assert psi.shape == om.shape
    assert np.max(np.abs(got - want)) < 1e-10 * np.max(np.abs(want))


def test_window_rows_and_sensor_layout():
    cfg = _cfg(32)
    ops = NS.NSOperators(cfg)
    g = NS.prepare_gaussians(cfg, ops, 1)
    assert g.shape == (16, 32, 32) and np.allclose(g.reshape(16, -1).sum(1), 1.0)
    # sensor i sits at x-index (i // spa) * 8, y-index (i % spa) * 8   (FluidSetup.jl:61, 142)
    for i in (0, 1, 5, 15):
        j, ii = np.unravel_index(np.argmax(g[i]), g[i].shape)
        assert (ii, j) == ((i // 4) * 8, (i % 4) * 8)
    yh = NS.ic(cfg, ops, 3, np.random.default_rng(3))
    S = NS.sensor_grid(cfg, g, yh)
    st = NS.featurize(cfg, g, yh)
    assert st.shape == (9, 16)
    r = 0
    for i in (-1, 0, 1):
        for j in (-1, 0, 1):
            for c in range(16):
                a, b = c // 4, c % 4
                assert st[r, c] == S[(a - i) % 4, (b - j) % 4]
            r += 1


def test_env_step_runs_and_julia_memory_round_trip():
    cfg = _cfg(32, dt=0.02)
    env = NS.NSEnv(cfg, y0=NS.ic(cfg, NS.NSOperators(cfg), 3, np.random.default_rng(0)))
    a = np.random.default_rng(1).uniform(-1, 1, (1, 16))
    s, r, d = env.step(a)
    assert s.shape == (9, 16) and r.shape == (16,) and not d
    assert np.all(np.isfinite(env.y.view(np.float64)))
    flat = NS.to_julia_memory(env.y)
    assert flat[1] == env.y[1, 0] and np.array_equal(NS.from_julia_memory(flat, 32, 32), env.y)


def test_adaptive_controller_meets_its_tolerance():
    """do_step_adaptive (the role of FluidSetup.jl's do_step2): the error against a finely resolved fixed-step run of the
    same ODE shrinks with the tolerance and stays below it (relative to the state), the step count grows, and the warm start
    returns the last natural step."""
    cfg = _cfg(32, dt=0.02)
    ops = NS.NSOperators(cfg)
    y0 = 3.0 * NS.ic(cfg, ops, 3, np.random.default_rng(4))
    p = np.zeros_like(y0)
    fine = y0
    for _ in range(400):
        fine = NS.rk4(cfg, ops, fine, p, cfg.dt / 400)
    prev_acc, prev_err = 0, None
    for tol in (1e-4, 1e-6, 1e-8):
        y, h, acc, rej = NS.do_step_adaptive(cfg, ops, y0, p, rtol=tol, atol=tol, return_stats=True)
        err = np.abs(y - fine).max() / np.abs(fine).max()
        assert err < tol and acc >= prev_acc and 0 < h
        assert prev_err is None or err < prev_err
        prev_acc, prev_err = acc, err
    assert prev_acc > 3
