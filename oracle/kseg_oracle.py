"""CPU oracle: Keller-Segel 1-D chemotaxis environment (fp64 restatement).

TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Restates /root/reference/scripts/Keller-Segel/setup/KellerSegelSetup.jl:
  prepare_rectangles :112-126   f (finite-difference rhs) :213-232   do_step :234-239
  reward_function :241-263      featurize :265-316                   prepare_action :318-332
  generate_random_init :373-384
The reference integrates f with OrdinaryDiffEq's adaptive RK4() at reltol=abstol=1e-8
(third-party step controller, not in the tree): the step SEQUENCE is unpinned, the RESULT is pinned
to that tolerance by the golden rows of Keller-Segel10_16/saves/hook.jld2.  `do_step` here is the
classical fixed-step RK4 with `n_sub` substeps (what the CUDA kernel implements); `do_step_ref`
integrates to 1e-12 with scipy for the golden comparison.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class KSegConfig:
    Lx: float = 10.0
    nx: int = 100
    sensor_positions: np.ndarray = None       # 1-based
    actuators_to_sensors: np.ndarray = None   # 1-based
    te: float = 8.0
    dt: float = 0.006
    n_sub: int = 8
    window_size: int = 3
    temporal_steps: int = 2
    memory_size: int = 0
    action_punish: float = 0.0
    delta_action_punish: float = 0.0
    agent_power: float = 10.0
    max_value: float = 20.0                  # PDEenv default (KellerSegelSetup passes none), PDEenv.jl:80

    @property
    def dx(self):
        return self.Lx / self.nx

    @property
    def n_sensors(self):
        return len(self.sensor_positions)

    @property
    def n_actuators(self):
        return len(self.actuators_to_sensors)


def kseg10_16_config():
    """scripts/Keller-Segel/Keller-Segel10_16/Keller-Segel10_16.jl:8-14"""
    return KSegConfig(sensor_positions=np.arange(3, 101, 5), actuators_to_sensors=np.arange(3, 19))


def prepare_rectangles(cfg, half=2):
    """KellerSegelSetup.jl:112-126: p[position-half : position+half] .= 1 (1-based, inclusive)."""
    out = np.zeros((cfg.n_sensors, cfg.nx))
    for i, pos in enumerate(cfg.sensor_positions):
        out[i, pos - half - 1:pos + half] = 1.0
    return out


def f(cfg, y, p):
    """KellerSegelSetup.jl:213-232.  y (2,nx).  circshift neighbours with the edge copies
    U[1,1]=U[1,2], U[end,3]=U[end,2] (zero-flux, quirk Q5)."""
    dx = cfg.dx
    u, v = y[0], y[1]

    def nb(w):
        left, right = np.roll(w, 1), np.roll(w, -1)
        left = left.copy(); right = right.copy()
        left[0] = w[0]; right[-1] = w[-1]
        return left, right
    ul, ur = nb(u)
    vl, vr = nb(v)
    du1 = (-0.5 / dx) * ul + 0.0 * u + (0.5 / dx) * ur
    du2 = (1.0 / dx ** 2) * ul + (-2.0 / dx ** 2) * u + (1.0 / dx ** 2) * ur
    dv1 = (-0.5 / dx) * vl + 0.0 * v + (0.5 / dx) * vr
    dv2 = (1.0 / dx ** 2) * vl + (-2.0 / dx ** 2) * v + (1.0 / dx ** 2) * vr
    vdot = dv2 - v + u + p
    udot = du2 + u - 5.6 * du1 * dv1 - 5.6 * u * dv2 - u ** 2
    return np.vstack([udot, vdot])


def do_step(cfg, y, p):
    """Classical RK4, n_sub fixed substeps over [t, t+dt] (tableau of OrdinaryDiffEq's RK4())."""
    h = cfg.dt / cfg.n_sub
    y = np.array(y, dtype=np.float64)
    for _ in range(cfg.n_sub):
        k1 = f(cfg, y, p)
        k2 = f(cfg, y + 0.5 * h * k1, p)
        k3 = f(cfg, y + 0.5 * h * k2, p)
        k4 = f(cfg, y + h * k3, p)
        y = y + (h / 6) * (k1 + 2 * (k2 + k3) + k4)
    return y


def do_step_adaptive(cfg, y, p, rtol=1e-8, atol=1e-8, h0=None, return_stats=False):
    """Error-controlled RK4 over [t, t+dt] in the role of the reference's `solve(prob, RK4(), reltol=1e-8, abstol=1e-8)`
    (KellerSegelSetup.jl:234-239).  OrdinaryDiffEq's own controller is third-party and unpinned; this is the
    controller the CUDA adaptive mode implements: step doubling with the classical tableau, e = (y2 - y1)/15, the
    extrapolated value y2 + e kept while the (16x larger) error of the single full step is what is held below the
    tolerance -- conservative, so that the result sits inside the reference solver's own 1e-8 of the golden rows --, RMS error norm over all components against atol + rtol max(|y|, |y_new|), step factor
    0.9 err^(-1/5) clamped to [0.2, 5], first trial step dt / n_sub (or h0)."""
    def rk4(y, h):
        k1 = f(cfg, y, p)
        k2 = f(cfg, y + 0.5 * h * k1, p)
        k3 = f(cfg, y + 0.5 * h * k2, p)
        k4 = f(cfg, y + h * k3, p)
        return y + (h / 6) * (k1 + 2 * (k2 + k3) + k4)
    y = np.array(y, dtype=np.float64)
    t, h = 0.0, (cfg.dt / cfg.n_sub if h0 is None else h0)
    acc = rej = 0
    while t < cfg.dt:
        last = t + h >= cfg.dt
        hs = cfg.dt - t if last else h
        y1 = rk4(y, hs)
        y2 = rk4(rk4(y, 0.5 * hs), 0.5 * hs)
        e = (y2 - y1) / 15
        yn = y2 + e
        sc = atol + rtol * np.maximum(np.abs(y), np.abs(yn))
        err = float(np.sqrt(np.sum((16 * e / sc) ** 2) / e.size))      # controlled: the error of the single full step, 16 e
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


def do_step_ref(cfg, y, p, tol=1e-12):
    """High-accuracy integration of the same ODE (stand-in for the reference's adaptive solver)."""
    from scipy.integrate import solve_ivp
    sol = solve_ivp(lambda t, z: f(cfg, z.reshape(2, -1), p).reshape(-1), (0.0, cfg.dt), np.asarray(y).reshape(-1),
                    method="DOP853", rtol=tol, atol=tol)
    return sol.y[:, -1].reshape(2, -1)


def prepare_action(cfg, g_act, action):
    """KellerSegelSetup.jl:318-332"""
    p = np.zeros(cfg.nx)
    for i in range(cfg.n_actuators):
        p = p + cfg.agent_power * action[0, i] * g_act[i]
    return p


def reward_function(cfg, g_sens, y, action, delta_action):
    """KellerSegelSetup.jl:241-263"""
    a2s = np.asarray(cfg.actuators_to_sensors) - 1
    s = np.zeros(cfg.n_actuators)
    for i in range(cfg.n_actuators):
        s[i] = np.dot(y[0] - 1.0, g_sens[a2s[i]]) ** 2 / 800
    return -np.abs(s) - cfg.action_punish * action[0] ** 2 - cfg.delta_action_punish * delta_action[0] ** 2


def featurize(cfg, g_sens, y, prev_state=None):
    """KellerSegelSetup.jl:265-316 (sees_action = false, memory_size = 0)."""
    sens = np.zeros((2, cfg.n_sensors))
    for i in range(cfg.n_sensors):
        sens[0, i] = np.dot(y[0], g_sens[i]) / 4
        sens[1, i] = np.dot(y[1], g_sens[i]) / 4
    h = cfg.window_size // 2
    a2s = np.asarray(cfg.actuators_to_sensors) - 1
    r1 = np.stack([np.roll(sens[0], i) for i in range(-h, h + 1)])[:, a2s]
    r2 = np.stack([np.roll(sens[1], i) for i in range(-h, h + 1)])[:, a2s]
    result = np.vstack([r1, r2])
    if cfg.temporal_steps > 1:
        if prev_state is None:
            result = np.vstack([result] * cfg.temporal_steps)
        else:
            result = np.vstack([result, prev_state[:prev_state.shape[0] - result.shape[0] - cfg.memory_size]])
    return result


def generate_random_init(cfg, coeffs):
    """KellerSegelSetup.jl:373-384 with the Uniform(-1,1) draws supplied by the caller."""
    n = int(np.ceil(cfg.Lx / 3))
    a = np.asarray(coeffs, dtype=np.float64)
    a = a / np.linalg.norm(a)
    x = cfg.dx * np.arange(1, cfg.nx + 1)
    y0 = np.ones((2, cfg.nx))
    for i in range(1, n + 1):
        y0[0] += a[i - 1] * np.sin(i * x / (2 * np.pi * (cfg.Lx / 22)))
        y0[1] += a[i - 1 + n] * np.sin(i * x / (2 * np.pi * (cfg.Lx / 22)))
    return y0


class KSegEnv:
    """PDEenv with the Keller-Segel closures (src/PDEenv.jl:183-241; check_max_value defaults to "y", max 20)."""

    def __init__(self, cfg, y0=None):
        self.cfg = cfg
        self.g_sens = prepare_rectangles(cfg)
        self.g_act = self.g_sens[np.asarray(cfg.actuators_to_sensors) - 1]
        self.y0 = np.vstack([np.ones(cfg.nx), 1.01 * np.ones(cfg.nx)]) if y0 is None else np.array(y0, dtype=np.float64)
        self.reset()

    def reset(self):
        cfg = self.cfg
        self.y = self.y0.copy()
        self.state = featurize(cfg, self.g_sens, self.y)
        self.action = np.zeros((1, cfg.n_actuators))
        self.delta_action = np.zeros((1, cfg.n_actuators))
        self.p = prepare_action(cfg, self.g_act, self.action)
        self.steps, self.time, self.reward, self.done = 0, 0.0, 0.0, False

    def step(self, action):
        cfg = self.cfg
        action = np.asarray(action, dtype=np.float64).reshape(1, -1)
        self.delta_action = action - self.action
        self.action = action
        self.p = prepare_action(cfg, self.g_act, action)
        self.y = do_step(cfg, self.y, self.p)
        self.reward = reward_function(cfg, self.g_sens, self.y, self.action, self.delta_action)
        self.state = featurize(cfg, self.g_sens, self.y, prev_state=self.state)
        self.steps += 1
        self.time += cfg.dt
        self.done = bool(self.time >= cfg.te or np.max(np.abs(self.y)) > cfg.max_value)
        return self.state, self.reward, self.done
