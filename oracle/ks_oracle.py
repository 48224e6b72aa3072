"""CPU oracle: Kuramoto-Sivashinsky environment (literal fp64 restatement).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs use it, and only as the checker / CPU arm.

Every function restates, operation by operation, a closure of
/root/reference/scripts/KS/setup/KSSetup.jl (cited per function).  The Julia
reference cannot run in this container (no julia binary); the restatement is
PINNED by the golden trajectories the reference ships in
scripts/KS/{KS22,KS200}/saves/hook.jld2 (tests/golden/ks22_hook.npz,
ks200_hook.npz; see tests/test_oracle_golden.py).
"""
from dataclasses import dataclass, field

import numpy as np


def _julia_rat(x):
    """Base.rat (twiceprecision.jl): continued-fraction rational approximation
    with numerator/denominator bounded by maxintfloat(Float32) = 2^24."""
    y = float(x)
    a = d = 1
    b = c = 0
    m = 16777216.0
    while abs(y) <= m:
        f = int(y)                      # trunc
        y -= f
        a, c = f * a + c, a
        b, d = f * b + d, b
        if max(abs(a), abs(b)) > int(m):
            return c, d
        if b != 0 and float(a) / float(b) == x:
            break
        if y == 0:
            break
        y = 1.0 / y
    return a, b


def julia_float_range(start, step, stop):
    """Values of Julia's `collect(start:step:stop)` for Float64 arguments.

    Restates Base.(:)(start::T, step::T, stop::T) where T<:IEEEFloat
    (julia 1.9 base/twiceprecision.jl): if start/step/stop are exactly
    representable small rationals the length is computed in integers, otherwise
    len = round((stop-start)/step)+1 minus one on overshoot.  Elements are
    start+(i-1)*step: correctly rounded rationals on the first path (floatrange),
    plain fl(start + fl(i*step)) on the fallback path.  (Quirk Q4 of SURVEY.md: the range
    `dx-50dx : dx : Lx+50dx` has 291 samples for KS22 but 340 for KS200.)"""
    from fractions import Fraction
    from math import gcd
    start, step, stop = float(start), float(step), float(stop)
    step_n, step_d = _julia_rat(step)
    if step_d != 0 and float(step_n) / float(step_d) == step:
        start_n, start_d = _julia_rat(start)
        stop_n, stop_d = _julia_rat(stop)
        if (start_d != 0 and stop_d != 0 and float(start_n) / float(start_d) == start
                and float(stop_n) / float(stop_d) == stop):
            den = start_d // gcd(start_d, step_d) * step_d
            m = 2.0 ** 53
            if den != 0 and abs(start * den) <= m and abs(step * den) <= m:
                sn = round(start * den)
                tn = round(step * den)
                q = (den * stop_n - stop_d * sn)
                dd = tn * stop_d
                length = max(0, int(q / dd) if q * dd < 0 else q // dd) + 1
                last = start + (length - 1) * step
                nxt = start + length * step
                btw = lambda a, x, b: (a <= x <= b) or (b <= x <= a)
                if btw(start, last, stop + step / 2) and not btw(start, nxt, stop):
                    return np.array([float(Fraction(sn + i * tn, den)) for i in range(length)])
    lf = (stop - start) / step
    if lf < 0:
        length = 0
    elif lf == 0:
        length = 1
    else:
        length = int(round(lf)) + 1
        stop2 = start + (length - 1) * step
        length -= int(start < stop < stop2) + int(start > stop > stop2)
    # steprangelen_hp(T, start, step, nb=0, len, 1): with nb = 0 the step's high
    # word keeps all 53 bits, so u*step.hi rounds and element i is simply
    # fl(start + fl(i*step)).
    return start + step * np.arange(length, dtype=np.float64)


@dataclass
class KSConfig:
    """Globals an experiment script sets before including KSSetup.jl
    (e.g. scripts/KS/KS200/KS200.jl:10-21) plus KSSetup.jl:20-51 constants."""
    Lx: float = 200.0
    nx: int = 240
    sensor_positions: np.ndarray = None        # 1-based grid indices
    actuators_to_sensors: np.ndarray = None    # 1-based sensor indices
    sigma_sensors: float = 1.0
    sigma_actuators: float = 1.0
    mu: float = 0.0
    te: float = 5.0
    dt: float = 0.1
    oversampling: int = 30
    max_value: float = 30.0
    window_size: int = 1
    temporal_steps: int = 1
    memory_size: int = 0
    action_punish: float = 0.002
    delta_action_punish: float = 0.002
    agent_power: float = 7.5
    check_max_value: str = "y"
    mono: bool = False                         # KSglobalSetup.jl variant
    t_samples: int = None                      # override for len(t) in prepare_gaussians

    @property
    def dx(self):
        return self.Lx / self.nx

    @property
    def n_sensors(self):
        return len(self.sensor_positions)

    @property
    def n_actuators(self):
        return len(self.actuators_to_sensors)


def ks22_config():
    """scripts/KS/KS22/KS22.jl:10-21"""
    return KSConfig(Lx=22.0, nx=192, sensor_positions=np.arange(1, 193, 24),
                    actuators_to_sensors=np.arange(1, 9), sigma_sensors=0.7, sigma_actuators=0.7)


def ks200_config():
    """scripts/KS/KS200/KS200.jl:10-21"""
    return KSConfig(Lx=200.0, nx=240, sensor_positions=np.arange(1, 241, 3),
                    actuators_to_sensors=np.arange(1, 81), sigma_sensors=1.0, sigma_actuators=1.0)


def ks256_config(window_size=1):
    """BASELINE config C2 (synthetic): nx=256, same dx as KS200,
    sensors=actuators at every 4th point, explicit nx+100 basis samples."""
    return KSConfig(Lx=200.0 * 256 / 240, nx=256, sensor_positions=np.arange(1, 257, 4),
                    actuators_to_sensors=np.arange(1, 65), sigma_sensors=1.0, sigma_actuators=1.0,
                    window_size=window_size, t_samples=256 + 100)


def prepare_gaussians(cfg, sigma, norm_mode=1):
    """KSSetup.jl:82-109.  Returns (n_sensors, nx).

    Literal quirks kept: variance *multiplies* (`/ 2 * sigma^2`), prefactor
    1/sqrt(2*pi*sigma), periodic wrap of the 50-sample tails, and the length of
    the Float64 range `dx-50dx : dx : Lx+50dx` (Q4)."""
    dx, nx, Lx = cfg.dx, cfg.nx, cfg.Lx
    extra = 50
    start = dx - extra * dx
    stop = Lx + extra * dx
    if cfg.t_samples is not None:
        t = start + dx * np.arange(cfg.t_samples)
    else:
        t = julia_float_range(start, dx, stop)
    out = []
    for position in cfg.sensor_positions:
        p = (1.0 / np.sqrt(2 * np.pi * sigma)) * np.exp(-(((t - position * dx) * 1) ** 2 / 2 * sigma ** 2))
        if norm_mode == 1:
            p = p / p.sum()
        else:
            p = p / p.max()
        pleft = p[:extra]
        pright = p[extra + nx:]
        q = p[extra:extra + nx].copy()
        q[nx - len(pleft):] += pleft
        q[:len(pright)] += pright
        out.append(q)
    return np.array(out)


class KSOperators:
    """Spectral constants, KSSetup.jl:115-119."""

    def __init__(self, cfg):
        nx, Lx = cfg.nx, cfg.Lx
        h = nx // 2
        self.kx = np.concatenate([np.arange(0, h), [0], np.arange(-h + 1, 0)]).astype(np.float64)
        self.alpha = 2 * np.pi * self.kx / Lx
        self.D = 1j * self.alpha
        self.L = self.alpha ** 2 - self.alpha ** 4
        self.G = -0.5 * self.D
        self.xx = cfg.dx * np.arange(1, nx + 1)          # collect(dx:dx:Lx), KSSetup.jl:36


def do_step(cfg, ops, y, p):
    """KSSetup.jl:130-160 (CNAB2 pseudo-spectral, `oversampling` substeps).

    The literal sequence is kept, including the redundant fft(p) and fft(mu*cos)
    inside the loop (line 155), the per-call A_inv/B derivation (134-135) and
    N^{n-1} := N^n on entry (141; quirk Q2).  The global-agent variant
    (KSglobalSetup.jl:142-172) is identical minus the mu term (Q3): mu=0 gives
    fft(0)=0 which adds exact zeros, so one function serves both."""
    nx = cfg.nx
    dt_o = cfg.dt / cfg.oversampling
    dt2 = dt_o / 2
    dt32 = 3 * dt_o / 2
    A_inv = (np.ones(nx) - dt2 * ops.L) ** (-1)
    B = np.ones(nx) + dt2 * ops.L
    u = (1 + 0j) * np.asarray(y, dtype=np.float64)
    Nn = ops.G * np.fft.fft(u ** 2)
    Nn1 = Nn.copy()
    u = np.fft.fft(u)
    forcing = cfg.mu * np.cos((2 + np.pi + ops.xx / (cfg.Lx / 2)))
    for _ in range(cfg.oversampling):
        Nn1[:] = Nn
        Nn[:] = u
        Nn = np.fft.ifft(Nn)
        Nn = Nn * Nn
        Nn = np.fft.fft(Nn)
        Nn = ops.G * Nn
        u = A_inv * (B * u + dt32 * Nn - dt2 * Nn1 + dt_o * np.fft.fft(p)) + dt_o * np.fft.fft(forcing)
    u = np.fft.ifft(u)
    return np.real(u)


def prepare_action(cfg, g_act, action):
    """KSSetup.jl:231-245.  action: (1+mem, n_a); uses row 1 only."""
    p = np.zeros(cfg.nx)
    for i in range(cfg.n_actuators):
        p = p + cfg.agent_power * action[0, i] * g_act[i]
    return p


def sensor_values(cfg, g_sens, y):
    """KSSetup.jl:197-202."""
    s = np.zeros(cfg.n_sensors)
    for i in range(cfg.n_sensors):
        s[i] = np.dot(y, g_sens[i]) / cfg.max_value
    return s


def window_rows(sensors, window_size, a2s):
    """KSSetup.jl:204-207: rows circshift(sensors, i), i=-h..h; column select.
    Julia circshift(v, i)[j] = v[j-i]  <=>  np.roll(v, i)."""
    h = window_size // 2
    rows = [np.roll(sensors, i) for i in range(-h, h + 1)]
    res = np.stack(rows)
    return res[:, np.asarray(a2s) - 1]


def featurize(cfg, g_sens, y, prev_state=None, action=None):
    """KSSetup.jl:190-229.  prev_state=None <=> the `isnothing(env)` branch."""
    sens = sensor_values(cfg, g_sens, y)
    result = window_rows(sens, cfg.window_size, cfg.actuators_to_sensors)
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


def reward_function(cfg, g_sens, y, action, delta_action):
    """KSSetup.jl:162-184 (per actuator); mono => KSglobalSetup.jl:175-205 ([mean])."""
    y6 = y * 6
    a2s = np.asarray(cfg.actuators_to_sensors) - 1
    sensors = np.zeros(cfg.n_actuators)
    for i in range(cfg.n_actuators):
        sensors[i] = np.abs(np.dot(y6, g_sens[a2s[i]])) ** 1.3 / (cfg.max_value * 3)
    r = -np.abs(sensors) - cfg.action_punish * action[0] ** 2 - cfg.delta_action_punish * delta_action[0] ** 2
    if cfg.mono:
        return np.array([r.mean()])
    return r


def generate_random_init(cfg, coeffs):
    """KSSetup.jl:288-298 with the Uniform(-1,1) draws supplied by the caller
    (Julia's RNG stream is not reproducible here: inputs, not oracle)."""
    a = np.asarray(coeffs, dtype=np.float64)
    a = a / np.linalg.norm(a)
    x = cfg.dx * np.arange(1, cfg.nx + 1)
    y0 = np.zeros(cfg.nx)
    for i in range(1, len(a) + 1):
        y0 += a[i - 1] * np.sin(i * x / (2 * np.pi))
    return y0 * 30 / np.linalg.norm(y0)


@dataclass
class KSEnv:
    """One environment: PDEenv container + step functor restated
    (/root/reference/src/PDEenv.jl:64-170 ctor, :183-193 reset!, :195-241 step)."""
    cfg: KSConfig
    y0: np.ndarray = None
    ops: KSOperators = field(init=False)

    def __post_init__(self):
        cfg = self.cfg
        self.ops = KSOperators(cfg)
        self.g_sens = prepare_gaussians(cfg, cfg.sigma_sensors, norm_mode=1)
        g_act = prepare_gaussians(cfg, cfg.sigma_actuators, norm_mode=2)
        self.g_act = g_act[np.asarray(cfg.actuators_to_sensors) - 1]
        if self.y0 is None:
            # y0_1D_standard, KSSetup.jl:53
            self.y0 = np.array([0.5 if 4 <= i <= 44 else 0.0 for i in range(1, cfg.nx + 1)])
        self.action0 = np.zeros((1 + cfg.memory_size, cfg.n_actuators))
        self.reset()

    def reset(self):
        """PDEenv.jl:183-193"""
        cfg = self.cfg
        self.y = np.array(self.y0, dtype=np.float64)
        self.state = featurize(cfg, self.g_sens, self.y)
        self.action = self.action0.copy()
        self.delta_action = np.zeros_like(self.action0)
        self.p = prepare_action(cfg, self.g_act, self.action0)
        self.steps = 0
        self.time = 0.0
        self.reward = 0.0
        self.done = False

    def step(self, action):
        """PDEenv.jl:195-241"""
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
