"""CPU oracle: 2-D incompressible Navier-Stokes environment (vorticity form, pseudo-spectral RK4).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this module; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.

PARITY UNPINNED: the reference's Fluid scripts save no trajectory (`collect_bestDF = false`,
scripts/Fluid/setup/FluidSetup.jl:376), so there is no golden vector for this path.  The
restatement is literal (operation by operation, numpy's fft2/ifft2 standing in for FFTW) and is
checked by analytic invariants in tests/test_ns_oracle.py (Taylor-Green decay, RK4 amplification
factor of the viscous term, zero advection of single modes, de-aliasing of a known product).

Arrays keep the reference's orientation: a Julia matrix `A[j, i]` of size (ny, nx) -- row j <-> y /
ky, column i <-> kx / x (src/fluid_rk4.jl:28-34, FluidSetup.jl:116-118) -- is the numpy array
`A[j, i]` of shape (ny, nx).  The C ABI takes Julia's column-major memory, i.e. `A.T` flattened
C-order (`to_julia_memory`).

The env uses the FIXED-step `do_step` (FluidSetup.jl:163-172: RK4 x oversampling), which is the
path BASELINE.json's north star names; the shipped scripts wire the adaptive `do_step2`
(FluidSetup.jl:181-186, 333; OrdinaryDiffEq step control at tol = 1e0, third party, unpinned;
SURVEY.md quirk Q9).
"""
from dataclasses import dataclass, field

import numpy as np


@dataclass
class NSConfig:
    """Globals of scripts/Fluid/Fluid_*/Fluid_*.jl:11-16 + FluidSetup.jl:28-101."""
    nx: int = 128
    ny: int = None
    Lx: float = 1.0
    Ly: float = 1.0
    nu: float = 0.00005
    te: float = 6.0
    dt: float = 0.02
    oversampling: int = None                   # floor(16*nx*dt), FluidSetup.jl:48
    sensors_per_axis: int = 16
    variance: float = 0.04
    ifpad: int = 1
    window_size: int = 3
    temporal_steps: int = 1
    memory_size: int = 0
    action_punish: float = 0.002
    delta_action_punish: float = 0.002
    agent_power: float = 70.0
    max_value: float = 3.0
    check_max_value: str = "reward"

    def __post_init__(self):
        if self.ny is None:
            self.ny = self.nx
        if self.oversampling is None:
            self.oversampling = int(np.floor(16 * self.nx * self.dt))

    @property
    def dx(self):
        return self.Lx / self.nx

    @property
    def dy(self):
        return self.Ly / self.ny

    @property
    def n_sensors(self):
        return self.sensors_per_axis ** 2

    @property
    def n_actuators(self):
        return self.sensors_per_axis ** 2

    @property
    def sensor_positions(self):
        """[[i,j] for i in 1:nx/spa:nx for j in 1:ny/spa:ny] (1-based), FluidSetup.jl:61."""
        sx, sy = self.nx // self.sensors_per_axis, self.ny // self.sensors_per_axis
        return [(i, j) for i in range(1, self.nx + 1, sx) for j in range(1, self.ny + 1, sy)]


class NSOperators:
    """FluidSetup.jl:103-133: wavenumbers (Nyquist kept on the positive side), grids."""

    def __init__(self, cfg):
        nx, ny = cfg.nx, cfg.ny
        self.kx = np.concatenate([np.arange(0, nx // 2 + 1), np.arange(-nx // 2 + 1, 0)]) / cfg.Lx * 2 * np.pi
        self.ky = np.concatenate([np.arange(0, ny // 2 + 1), np.arange(-ny // 2 + 1, 0)]) / cfg.Ly * 2 * np.pi
        kx2, ky2 = self.kx ** 2, self.ky ** 2
        # kx2ky2[i, j] = ky2[i] + kx2[j]  (row <-> ky)
        self.kx2ky2 = ky2[:, None] + kx2[None, :]
        self.kx_repeat = np.tile(self.kx[None, :], (ny, 1))
        self.ky_repeat = np.tile(self.ky[:, None], (1, nx))
        # x1 = range(0, Lx, length = nx + 1)[1:nx]; meshgrid: xx[j, i] = x[i], yy[j, i] = y[j]
        x1 = np.linspace(0.0, cfg.Lx, nx + 1)[:nx]
        y1 = np.linspace(0.0, cfg.Ly, ny + 1)[:ny]
        self.xx = np.tile(x1[None, :], (ny, 1))
        self.yy = np.tile(y1[:, None], (1, nx))
        self.nxp, self.nyp = nx * 3 // 2, ny * 3 // 2


def taylorvtx_phys(cfg, ops, x0, y0, a0, U_max):
    """src/fluid_rk4.jl:54-64 (physical field before the fft)."""
    omg = np.zeros_like(ops.xx)
    for i in (-1, 0, 1):
        for j in (-1, 0, 1):
            r2 = (ops.xx - x0 - i * cfg.Lx) ** 2 + (ops.yy - y0 - j * cfg.Ly) ** 2
            omg = omg + U_max / a0 * (2 - r2 / a0 ** 2) * np.exp(0.5 * (1 - r2 / a0 ** 2))
    return omg


def taylorvtx(cfg, ops, x0, y0, a0, U_max):
    """src/fluid_rk4.jl:54-69."""
    return np.fft.fft2(taylorvtx_phys(cfg, ops, x0, y0, a0, U_max))


def ic(cfg, ops, caseno, rng):
    """src/fluid_rk4.jl:72-120 with the draws taken from a numpy Generator (Julia's StableRNG stream
    is not reproducible here: initial conditions are inputs, not oracle).  Draw order kept."""
    Lx, Ly = cfg.Lx, cfg.Ly
    if caseno == 1:
        return taylorvtx(cfg, ops, Lx / 2, Ly / 2, Lx / 8, 1.0)
    if caseno == 2:
        return taylorvtx(cfg, ops, Lx / 2, 0.4 * Ly, Lx / 10.0, 1.0) + taylorvtx(cfg, ops, Lx / 2, 0.6 * Ly, Lx / 10, 1.0)
    nv = 30 if caseno == 3 else 50
    omghat = 0
    for _ in range(nv):
        x0 = rng.random() * Lx
        y0 = rng.random() * Ly
        a0 = Lx / 20 if caseno == 3 else Lx / 20 * (0.5 + rng.random())
        um = rng.random() * 2 - 1.0
        omghat = omghat + taylorvtx(cfg, ops, x0, y0, a0, um)
    return omghat


def pad(cfg, ops, f):
    """src/fluid_rk4.jl:192-210: four-corner copy; the Nyquist row/column go to the positive side."""
    nx, ny = cfg.nx, cfg.ny
    fp = np.zeros((ops.nyp, ops.nxp), dtype=np.complex128)
    nyh, nxh = ny // 2, nx // 2
    fp[:nyh + 1, :nxh + 1] = f[:nyh + 1, :nxh + 1]
    fp[:nyh + 1, ops.nxp - nxh + 1:] = f[:nyh + 1, nxh + 1:]
    fp[ops.nyp - nyh + 1:, :nxh + 1] = f[nyh + 1:, :nxh + 1]
    fp[ops.nyp - nyh + 1:, ops.nxp - nxh + 1:] = f[nyh + 1:, nxh + 1:]
    return fp


def chop(cfg, ops, fp):
    """src/fluid_rk4.jl:212-229."""
    nx, ny = cfg.nx, cfg.ny
    f = np.zeros((ny, nx), dtype=np.complex128)
    nyh, nxh = ny // 2, nx // 2
    f[:nyh + 1, :nxh + 1] = fp[:nyh + 1, :nxh + 1]
    f[:nyh + 1, nxh + 1:] = fp[:nyh + 1, ops.nxp - nxh + 1:]
    f[nyh + 1:, :nxh + 1] = fp[ops.nyp - nyh + 1:, :nxh + 1]
    f[nyh + 1:, nxh + 1:] = fp[ops.nyp - nyh + 1:, ops.nxp - nxh + 1:]
    return f


def advection(cfg, ops, omghat):
    """src/fluid_rk4.jl:145-190."""
    with np.errstate(divide="ignore", invalid="ignore"):
        psihat = omghat / ops.kx2ky2
    psihat[0, 0] = 0.0
    domgdx = 1j * omghat * ops.kx_repeat
    domgdy = 1j * omghat * ops.ky_repeat
    vhat = -1j * psihat * ops.kx_repeat
    uhat = 1j * psihat * ops.ky_repeat
    if cfg.ifpad == 1:
        up = np.real(np.fft.ifft2(pad(cfg, ops, uhat)))
        vp = np.real(np.fft.ifft2(pad(cfg, ops, vhat)))
        domgdxp = np.real(np.fft.ifft2(pad(cfg, ops, domgdx)))
        domgdyp = np.real(np.fft.ifft2(pad(cfg, ops, domgdy)))
        temp = np.fft.fft2(-up * domgdxp - vp * domgdyp)
        return chop(cfg, ops, temp) * 1.5 * 1.5
    u = np.real(np.fft.ifft2(uhat))
    v = np.real(np.fft.ifft2(vhat))
    return np.fft.fft2(-u * np.real(np.fft.ifft2(domgdx)) - v * np.real(np.fft.ifft2(domgdy)))


def rhs(cfg, ops, omghat, p):
    """src/fluid_rk4.jl:134-143."""
    lin = -cfg.nu * (ops.kx2ky2 * omghat)
    return lin + advection(cfg, ops, omghat) + p


def rk4(cfg, ops, f, p, dt):
    """src/fluid_rk4.jl:122-132."""
    k1 = rhs(cfg, ops, f, p)
    k2 = rhs(cfg, ops, f + 0.5 * dt * k1, p)
    k3 = rhs(cfg, ops, f + 0.5 * dt * k2, p)
    k4 = rhs(cfg, ops, f + dt * k3, p)
    return f + dt / 6 * (k1 + 2 * (k2 + k3) + k4)


def do_step(cfg, ops, y, p):
    """FluidSetup.jl:163-172 (fixed-step RK4 x oversampling)."""
    dt_o = cfg.dt / cfg.oversampling
    for _ in range(cfg.oversampling):
        y = rk4(cfg, ops, y, p, dt_o)
    return y


def prepare_gaussians(cfg, ops, norm_mode=1):
    """FluidSetup.jl:139-157: thresholded Taylor-vortex profiles, dense (n_sensors, ny, nx)."""
    out = []
    for (pi, pj) in cfg.sensor_positions:
        p = np.real(np.fft.ifft2(taylorvtx(cfg, ops, pi * cfg.dx - cfg.dx, pj * cfg.dy - cfg.dy, cfg.variance, 1.0)))
        p[p < 0.1] = 0.0
        if norm_mode == 1:
            p = p / p.sum()
        else:
            p = p / p.max()
        out.append(p)
    return np.array(out)


def sensor_grid(cfg, g_sens, yhat):
    """FluidSetup.jl:205-217: sensors[floor((i-1)/spa)+1, (i-1)%spa+1] = <omega, g_i> / 70."""
    y = np.real(np.fft.ifft2(yhat))
    spa = cfg.sensors_per_axis
    sensors = np.zeros((spa, spa))
    for i in range(cfg.n_sensors):
        sensors[i // spa, i % spa] = np.sum(y * g_sens[i]) / 70
    return sensors


def featurize(cfg, g_sens, yhat, prev_state=None, action=None):
    """FluidSetup.jl:204-245.  Rows: i outer, j inner; row(i,j) = vec(circshift(sensors,[i,j])')."""
    sensors = sensor_grid(cfg, g_sens, yhat)
    h = cfg.window_size // 2
    rows = []
    for i in range(-h, h + 1):
        for j in range(-h, h + 1):
            shifted = np.roll(sensors, (i, j), axis=(0, 1))          # circshift(A,[i,j])[a,b] = A[a-i,b-j]
            rows.append(shifted.reshape(-1))                         # transpose + column-major reshape = row-major
    result = np.stack(rows)
    if cfg.temporal_steps > 1:
        if prev_state is None:
            result = np.vstack([result] * cfg.temporal_steps)
        else:
            keep = prev_state.shape[0] - result.shape[0] - cfg.memory_size
            result = np.vstack([result, prev_state[:keep]])
    if cfg.memory_size > 0:
        if prev_state is None:
            result = np.vstack([result, np.zeros((cfg.memory_size, cfg.n_actuators))])
        else:
            result = np.vstack([result, action[action.shape[0] - cfg.memory_size:]])
    return result


def reward_function(cfg, g_sens, yhat, action, delta_action):
    """FluidSetup.jl:188-202."""
    y = np.real(np.fft.ifft2(yhat))
    sensors = np.zeros(cfg.n_actuators)
    for i in range(cfg.n_actuators):
        sensors[i] = np.abs(np.sum(y * g_sens[i])) ** 1.1 / 320
    return -np.abs(sensors) - cfg.action_punish * action[0] ** 2 - cfg.delta_action_punish * delta_action[0] ** 2


def prepare_action(cfg, g_act, action):
    """FluidSetup.jl:247-261: p_hat = fft(sum_i power * a_i * g_act_i)."""
    p = np.zeros((cfg.ny, cfg.nx))
    for i in range(cfg.n_actuators):
        p = p + cfg.agent_power * action[0, i] * g_act[i]
    return np.fft.fft2(p)


def to_julia_memory(a):
    """(ny, nx) numpy array -> flat array in Julia's column-major order (j fastest)."""
    return np.ascontiguousarray(np.asarray(a).T).reshape(-1)


def from_julia_memory(flat, ny, nx):
    return np.asarray(flat).reshape(nx, ny).T


def do_step_adaptive(cfg, ops, y, p, rtol=1.0, atol=1.0, h0=None, return_stats=False):
    """Error-controlled RK4 over [t, t + dt] in the role of the WIRED-IN `do_step2` (FluidSetup.jl:178-186, :333:
    `solve(ODEProblem(f, env.y, tspan, env.p), RK4(), reltol=tol, abstol=tol)`, tol = 1e0 as shipped, 1e-8 commented).
    OrdinaryDiffEq's controller is third-party and unpinned; this is the controller of the CUDA adaptive mode
    (same rules as kseg_oracle.do_step_adaptive): step doubling with the classical tableau rk4(), e = (y2 - y1)/15, the
    extrapolated y2 + e kept, the 16x larger error of the single full step controlled, RMS norm over the complex entries
    against atol + rtol max(|y|, |y_new|) (complex moduli, like OrdinaryDiffEq's default norm), factor 0.9 err^(-1/5)
    clamped to [0.2, 5], first trial step dt / oversampling (or h0 = the previous env step's last accepted size)."""
    y = np.array(y, dtype=np.complex128)
    t, h = 0.0, (cfg.dt / cfg.oversampling if h0 is None else h0)
    acc = rej = 0
    while t < cfg.dt:
        last = t + h >= cfg.dt
        hs = cfg.dt - t if last else h
        y1 = rk4(cfg, ops, y, p, hs)
        y2 = rk4(cfg, ops, rk4(cfg, ops, y, p, 0.5 * hs), p, 0.5 * hs)
        e = (y2 - y1) / 15
        yn = y2 + e
        sc = atol + rtol * np.maximum(np.abs(y), np.abs(yn))
        err = float(np.sqrt(np.sum((16 * np.abs(e) / sc) ** 2) / e.size))
        ok = err <= 1.0
        if ok:
            y, t = yn, (cfg.dt if last else t + hs)
            acc += 1
        else:
            rej += 1
        fac = (0.9 * err ** -0.2 if err > 0 else 5.0) if err == err else 0.2
        fac = min(5.0, max(0.2, fac))
        if not (ok and hs < h):
            h = hs * fac
    return (y, h, acc, rej) if return_stats else y


@dataclass
class NSEnv:
    """PDEenv restated for the Fluid setup (src/PDEenv.jl:64-241, FluidSetup.jl:330-343)."""
    cfg: NSConfig
    y0: np.ndarray = None
    ops: NSOperators = field(init=False)

    def __post_init__(self):
        cfg = self.cfg
        self.ops = NSOperators(cfg)
        self.g_sens = prepare_gaussians(cfg, self.ops, norm_mode=1)
        self.g_act = prepare_gaussians(cfg, self.ops, norm_mode=2)
        if self.y0 is None:
            self.y0 = ic(cfg, self.ops, 2, None)
        self.action0 = np.zeros((1 + cfg.memory_size, cfg.n_actuators))
        self.reset()

    def reset(self):
        cfg = self.cfg
        self.y = np.array(self.y0, dtype=np.complex128)
        self.state = featurize(cfg, self.g_sens, self.y)
        self.action = self.action0.copy()
        self.delta_action = np.zeros_like(self.action0)
        self.p = prepare_action(cfg, self.g_act, self.action0)
        self.steps = 0
        self.time = 0.0
        self.reward = 0.0
        self.done = False

    def step(self, action):
        cfg = self.cfg
        action = np.asarray(action, dtype=np.float64).reshape(self.action0.shape)
        self.delta_action = action - self.action
        self.action = action
        self.p = prepare_action(cfg, self.g_act, self.action)
        self.y = do_step(cfg, self.ops, self.y, self.p)
        self.reward = reward_function(cfg, self.g_sens, self.y, self.action, self.delta_action)
        self.state = featurize(cfg, self.g_sens, self.y, prev_state=self.state, action=self.action)
        self.steps += 1
        self.time += cfg.dt
        if cfg.check_max_value == "y":
            self.done = bool(self.time >= cfg.te or np.max(np.abs(self.y)) > cfg.max_value)
        elif cfg.check_max_value == "reward":
            self.done = bool(self.time >= cfg.te or np.max(np.abs(self.reward)) > cfg.max_value)
        else:
            self.done = bool(self.time >= cfg.te)
        return self.state, self.reward, self.done
