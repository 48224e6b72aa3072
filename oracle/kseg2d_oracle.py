"""CPU oracle: 2-D Keller-Segel chemotaxis (BASELINE config 3) -- TEST INFRASTRUCTURE ONLY.

The reference has NO 2-D Keller-Segel implementation (its model is 1-D, `(2, nx)`,
scripts/Keller-Segel/setup/KellerSegelSetup.jl:39, 213-239).  This restates the direct
2-D generalisation the CUDA kernel implements: the 1-D rhs `f` (KellerSegelSetup.jl:213-232) applied per axis
with the same zero-flux edge copies (quirk Q5) and the same constants.  It is PINNED to the reference through
the 1-D model: for y-independent data every y-difference is an exact zero, so the 2-D result must equal
oracle/kseg_oracle.py -- which is pinned to the golden rows of Keller-Segel10_16/saves/hook.jld2 -- bit for
bit (tests/test_kseg2d.py).  Arrays are Julia-shaped `(2, nx, ny)`.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class KSeg2DConfig:
    nx: int = 128
    ny: int = 128
    Lx: float = 12.8
    Ly: float = 12.8
    te: float = 8.0
    dt: float = 0.006
    n_sub: int = 40
    sensors_per_axis: int = 16
    window_size: int = 3
    temporal_steps: int = 2
    agent_power: float = 10.0
    max_value: float = 20.0
    obs_div: float = 20.0                     # 5x5 box sum / 20  (1-D: 5-point box sum / 4)
    reward_div: float = 20000.0               # 1-D: 800 for a 5-point box

    @property
    def n_sensors(self):
        return self.sensors_per_axis ** 2

    n_actuators = n_sensors


def _nb(w, axis):
    lo, hi = np.roll(w, 1, axis), np.roll(w, -1, axis)
    lo, hi = lo.copy(), hi.copy()
    first = [slice(None)] * w.ndim; first[axis] = 0
    last = [slice(None)] * w.ndim; last[axis] = -1
    lo[tuple(first)] = w[tuple(first)]
    hi[tuple(last)] = w[tuple(last)]
    return lo, hi


def f(cfg, y, p):
    """2-D rhs; per axis the literal 1-D expressions of KellerSegelSetup.jl:225-229."""
    dx, dy = cfg.Lx / cfg.nx, cfg.Ly / cfg.ny
    u, v = y[0], y[1]
    d = {}
    for name, w in (("u", u), ("v", v)):
        for ax, h in ((0, dx), (1, dy)):
            lo, hi = _nb(w, ax)
            d[name, 1, ax] = (-0.5 / h) * lo + 0.0 * w + (0.5 / h) * hi
            d[name, 2, ax] = (1.0 / h ** 2) * lo + (-2.0 / h ** 2) * w + (1.0 / h ** 2) * hi
    lapu = d["u", 2, 0] + d["u", 2, 1]
    lapv = d["v", 2, 0] + d["v", 2, 1]
    vdot = lapv - v + u + p
    udot = lapu + u - (5.6 * d["u", 1, 0] * d["v", 1, 0] + 5.6 * d["u", 1, 1] * d["v", 1, 1]) - 5.6 * u * lapv - u ** 2
    return np.stack([udot, vdot])


def do_step(cfg, y, p):
    """Classical RK4 with n_sub fixed substeps, accumulated like the kernel: acc = k1 + 2 k2 + 2 k3, then + k4."""
    h = cfg.dt / cfg.n_sub
    y = np.array(y, dtype=np.float64)
    for _ in range(cfg.n_sub):
        k1 = f(cfg, y, p)
        k2 = f(cfg, y + 0.5 * h * k1, p)
        k3 = f(cfg, y + 0.5 * h * k2, p)
        k4 = f(cfg, y + h * k3, p)
        y = y + (h / 6) * (k1 + 2 * k2 + 2 * k3 + k4)
    return y


def prepare_boxes(cfg, half=2):
    """5 x 5 boxes on a sensors_per_axis^2 lattice, sensor index = a * spa + b (a <-> x, b <-> y; the ordering of
    FluidSetup.jl:61).  Returns (n_sensors, nx, ny)."""
    spa = cfg.sensors_per_axis
    sx, sy = cfg.nx // spa, cfg.ny // spa
    out = np.zeros((spa * spa, cfg.nx, cfg.ny))
    for a in range(spa):
        for b in range(spa):
            cx, cy = a * sx + sx // 2, b * sy + sy // 2
            out[a * spa + b, max(cx - half, 0):cx + half + 1, max(cy - half, 0):cy + half + 1] = 1.0
    return out


def sensor_grid(cfg, g, field):
    return np.array([np.sum(field * g[i]) for i in range(cfg.n_sensors)]).reshape(cfg.sensors_per_axis, cfg.sensors_per_axis)


def featurize(cfg, g, y, prev_state=None):
    """Window rows like FluidSetup.jl:219-223 (i outer, j inner), per field like KellerSegelSetup.jl:283-290,
    temporal stacking like KellerSegelSetup.jl:292-301."""
    h = cfg.window_size // 2
    blocks = []
    for fld in (0, 1):
        S = sensor_grid(cfg, g, y[fld]) / cfg.obs_div
        blocks.append(np.stack([np.roll(S, (i, j), axis=(0, 1)).reshape(-1) for i in range(-h, h + 1) for j in range(-h, h + 1)]))
    result = np.vstack(blocks)
    if cfg.temporal_steps > 1:
        if prev_state is None:
            result = np.vstack([result] * cfg.temporal_steps)
        else:
            result = np.vstack([result, prev_state[:prev_state.shape[0] - result.shape[0]]])
    return result


def reward_function(cfg, g, y):
    return -np.abs(np.array([np.sum((y[0] - 1.0) * g[i]) ** 2 / cfg.reward_div for i in range(cfg.n_actuators)]))


def prepare_action(cfg, g, action):
    p = np.zeros((cfg.nx, cfg.ny))
    for i in range(cfg.n_actuators):
        p = p + cfg.agent_power * action[i] * g[i]
    return p


def random_init(cfg, rng):
    """1 + products of the 1-D random sine profiles of KellerSegelSetup.jl:373-384 along x and y."""
    def prof(n, L):
        m = int(np.ceil(L / 3))
        a = rng.uniform(-1, 1, m)
        a /= np.linalg.norm(a)
        x = (L / n) * np.arange(1, n + 1)
        return sum(a[i - 1] * np.sin(i * x / (2 * np.pi * (L / 22))) for i in range(1, m + 1))
    y0 = np.ones((2, cfg.nx, cfg.ny))
    y0[0] += 0.5 * np.outer(prof(cfg.nx, cfg.Lx), prof(cfg.ny, cfg.Ly))
    y0[1] += 0.5 * np.outer(prof(cfg.nx, cfg.Lx), prof(cfg.ny, cfg.Ly))
    return y0


class KSeg2DEnv:
    def __init__(self, cfg, y0, g=None):
        self.cfg, self.y0 = cfg, np.array(y0, dtype=np.float64)
        self.g = prepare_boxes(cfg) if g is None else g
        self.reset()

    def reset(self):
        self.y = self.y0.copy()
        self.state = featurize(self.cfg, self.g, self.y)
        self.steps, self.time, self.done = 0, 0.0, False

    def step(self, action):
        cfg = self.cfg
        self.p = prepare_action(cfg, self.g, np.asarray(action, dtype=np.float64).reshape(-1))
        self.y = do_step(cfg, self.y, self.p)
        self.reward = reward_function(cfg, self.g, self.y)
        self.state = featurize(cfg, self.g, self.y, prev_state=self.state)
        self.steps += 1
        self.time += cfg.dt
        self.done = bool(self.time >= cfg.te or np.max(np.abs(self.y)) > cfg.max_value)
        return self.state, self.reward, self.done
