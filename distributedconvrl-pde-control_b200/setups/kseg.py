"""Keller-Segel setup: what scripts/Keller-Segel/setup/KellerSegelSetup.jl does around the hot path."""
import numpy as np

from .. import _lib as L
from ..env import PDEenv


def prepare_rectangles(nx, sensor_positions, half_window_size=2):
    """KellerSegelSetup.jl:112-126 (1-based positions; non-wrapping 5-point boxes)."""
    out = np.zeros((len(sensor_positions), nx))
    for i, pos in enumerate(sensor_positions):
        out[i, pos - half_window_size - 1:pos + half_window_size] = 1.0
    return out


class KellerSegelSetup:
    """Globals of scripts/Keller-Segel/Keller-Segel10_16/Keller-Segel10_16.jl + KellerSegelSetup.jl:26-57."""

    def __init__(self, Lx=10.0, nx=100, sensor_positions=None, actuators_to_sensors=None, te=8.0, dt=0.006,
                 rk4_substeps=40, window_size=3, temporal_steps=2, memory_size=0, action_punish=0.0,
                 delta_action_punish=0.0, agent_power=10.0, max_value=20.0):
        self.Lx, self.nx = float(Lx), int(nx)
        self.sensor_positions = np.arange(3, nx + 1, 5) if sensor_positions is None else np.asarray(sensor_positions)
        self.actuators_to_sensors = np.arange(3, 19) if actuators_to_sensors is None else np.asarray(actuators_to_sensors)
        self.te, self.dt, self.rk4_substeps = te, dt, rk4_substeps
        self.window_size, self.temporal_steps, self.memory_size = window_size, temporal_steps, memory_size
        self.action_punish, self.delta_action_punish = action_punish, delta_action_punish
        self.agent_power, self.max_value = agent_power, max_value
        self.gaussians = prepare_rectangles(self.nx, self.sensor_positions, 2)          # :128
        self.gaussians_actuators = self.gaussians[self.actuators_to_sensors - 1]        # :129

    def y0_standard(self):
        """y0_2D_standard, KellerSegelSetup.jl:60-61"""
        return np.vstack([np.ones(self.nx), 1.01 * np.ones(self.nx)])

    def generate_random_init(self, rng, n=1):
        """KellerSegelSetup.jl:373-384, batched -> (n, 2, nx)"""
        ns = int(np.ceil(self.Lx / 3))
        a = rng.uniform(-1, 1, size=(n, 2 * ns))
        a /= np.linalg.norm(a, axis=1, keepdims=True)
        x = (self.Lx / self.nx) * np.arange(1, self.nx + 1)
        basis = np.sin(np.arange(1, ns + 1)[:, None] * x[None, :] / (2 * np.pi * (self.Lx / 22)))
        y0 = np.ones((n, 2, self.nx))
        y0[:, 0] += a[:, :ns] @ basis
        y0[:, 1] += a[:, ns:] @ basis
        return y0

    def make_env(self, n_envs=1, dtype="f64", device=0, y0=None, adaptive=False, rtol=1e-8, atol=1e-8):
        """adaptive=True: the reference's ACTIVE integrator mode -- error-controlled RK4 at (rtol, atol) with per-environment
        step control (KellerSegelSetup.jl:234-239) instead of `rk4_substeps` fixed substeps."""
        y0 = self.y0_standard() if y0 is None else np.asarray(y0, dtype=np.float64)
        if y0.ndim == 3:                      # (B, 2, nx) -> reference shape (2, nx, B)
            y0 = y0.transpose(1, 2, 0)
        return PDEenv(problem=L.KSEG1D, adaptive=int(bool(adaptive)), rtol=float(rtol), atol=float(atol), n_envs=n_envs, dtype=dtype, device=device, sensor_basis=self.gaussians,
                      actuator_basis=self.gaussians_actuators, actuators_to_sensors=self.actuators_to_sensors, y0=y0,
                      nx=self.nx, ny=1, Lx=self.Lx, dt=self.dt, te=self.te, oversampling=self.rk4_substeps,
                      max_value=self.max_value, window_size=self.window_size, temporal_steps=self.temporal_steps,
                      memory_size=self.memory_size, action_punish=self.action_punish,
                      delta_action_punish=self.delta_action_punish, agent_power=self.agent_power, check_max_value="y")


class KellerSegel2DSetup:
    """BASELINE config 3: 2-D Keller-Segel, 128 x 128, 16 x 16 distributed box sensors/actuators.

    The reference ships only the 1-D model; this is its per-axis generalisation (same rhs terms and
    constants, same zero-flux edge rule, same RK4), see csrc/kseg2d.cu and DESIGN.md.  Sensor scales
    follow the 1-D ones per box point: 25-point box / 20 (1-D: 5-point box / 4), reward / 20000 (1-D: / 800)."""

    def __init__(self, nx=128, ny=128, Lx=12.8, Ly=12.8, sensors_per_axis=16, te=8.0, dt=0.006, rk4_substeps=40,
                 window_size=3, temporal_steps=2, agent_power=10.0, max_value=20.0, obs_div=20.0, reward_div=20000.0,
                 gaussians=None, gaussians_actuators=None):
        self.nx, self.ny, self.Lx, self.Ly = int(nx), int(ny), float(Lx), float(Ly)
        self.sensors_per_axis = int(sensors_per_axis)
        self.te, self.dt, self.rk4_substeps = te, dt, rk4_substeps
        self.window_size, self.temporal_steps = window_size, temporal_steps
        self.agent_power, self.max_value, self.obs_div, self.reward_div = agent_power, max_value, obs_div, reward_div
        self.gaussians = self.prepare_boxes() if gaussians is None else np.asarray(gaussians, dtype=np.float64)
        self.gaussians_actuators = self.gaussians if gaussians_actuators is None else np.asarray(gaussians_actuators)
        self.actuators_to_sensors = np.arange(1, self.gaussians_actuators.shape[0] + 1)

    def prepare_boxes(self, half=2):
        """(n_sensors, nx, ny); sensor index = a * spa + b with a <-> x, b <-> y."""
        spa = self.sensors_per_axis
        sx, sy = self.nx // spa, self.ny // spa
        out = np.zeros((spa * spa, self.nx, self.ny))
        for a in range(spa):
            for b in range(spa):
                cx, cy = a * sx + sx // 2, b * sy + sy // 2
                out[a * spa + b, max(cx - half, 0):cx + half + 1, max(cy - half, 0):cy + half + 1] = 1.0
        return out

    @staticmethod
    def _julia_flat(a):
        """(..., nx, ny) -> (..., nx*ny) with x fastest (Julia column-major)."""
        a = np.asarray(a)
        return np.ascontiguousarray(np.swapaxes(a, -1, -2)).reshape(a.shape[:-2] + (-1,))

    def make_env(self, n_envs=1, dtype="f64", device=0, y0=None):
        y0 = np.ones((2, self.nx, self.ny)) * np.array([1.0, 1.01])[:, None, None] if y0 is None else np.asarray(y0)
        if y0.ndim == 4:                      # (B, 2, nx, ny) -> (2, nx, ny, B)
            y0 = y0.transpose(1, 2, 3, 0)
        return PDEenv(problem=L.KSEG2D, n_envs=n_envs, dtype=dtype, device=device,
                      sensor_basis=self._julia_flat(self.gaussians), actuator_basis=self._julia_flat(self.gaussians_actuators),
                      actuators_to_sensors=self.actuators_to_sensors, y0=y0, nx=self.nx, ny=self.ny, Lx=self.Lx, Ly=self.Ly,
                      dt=self.dt, te=self.te, oversampling=self.rk4_substeps, sensors_per_axis=self.sensors_per_axis,
                      max_value=self.max_value, window_size=self.window_size, temporal_steps=self.temporal_steps,
                      agent_power=self.agent_power, obs_scale=1.0 / self.obs_div, reward_div=self.reward_div,
                      check_max_value="y")
