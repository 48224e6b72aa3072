"""KS setup: what scripts/KS/setup/KSSetup.jl (and KSglobalSetup.jl) do around the hot path.

Builds the sensor / actuator bases on the host exactly like the reference
(`prepare_gaussians`, KSSetup.jl:82-113) and hands them to the library as
arrays; the spectral operators (KSSetup.jl:115-119) are derived inside
libpdeb200 from (Lx, nx, dt, oversampling).
"""
import numpy as np

from .. import _lib as L
from ..env import PDEenv
from .julia_ranges import float_range


def prepare_gaussians(Lx, nx, sensor_positions, sigma, norm_mode=1, t_samples=None):
    """KSSetup.jl:82-109, vectorised over sensors.  Returns (n_sensors, nx) float64."""
    dx = Lx / nx
    extra = 50
    start, stop = dx - extra * dx, Lx + extra * dx
    t = (start + dx * np.arange(t_samples)) if t_samples else float_range(start, dx, stop)
    pos = np.asarray(sensor_positions, dtype=np.float64)[:, None] * dx
    p = (1.0 / np.sqrt(2 * np.pi * sigma)) * np.exp(-(((t[None, :] - pos) * 1) ** 2 / 2 * sigma ** 2))
    p = p / (p.sum(axis=1, keepdims=True) if norm_mode == 1 else p.max(axis=1, keepdims=True))
    left, right = p[:, :extra], p[:, extra + nx:]
    core = p[:, extra:extra + nx].copy()
    core[:, nx - left.shape[1]:] += left
    core[:, :right.shape[1]] += right
    return core


class KSSetup:
    """Globals of an experiment script (e.g. scripts/KS/KS200/KS200.jl:10-21) + KSSetup.jl:20-51."""

    def __init__(self, Lx=200.0, nx=240, sensor_positions=None, actuators_to_sensors=None, sigma_sensors=1.0,
                 sigma_actuators=1.0, mu=0.0, te=5.0, dt=0.1, oversampling=30, max_value=30.0, window_size=1,
                 temporal_steps=1, memory_size=0, action_punish=0.002, delta_action_punish=0.002, agent_power=7.5,
                 check_max_value="y", mono=False, t_samples=None):
        self.Lx, self.nx = float(Lx), int(nx)
        self.sensor_positions = np.arange(1, nx + 1, 3) if sensor_positions is None else np.asarray(sensor_positions)
        self.actuators_to_sensors = (np.arange(1, len(self.sensor_positions) + 1) if actuators_to_sensors is None
                                     else np.asarray(actuators_to_sensors))
        self.sigma_sensors, self.sigma_actuators = sigma_sensors, sigma_actuators
        self.mu, self.te, self.dt, self.oversampling, self.max_value = mu, te, dt, oversampling, max_value
        self.window_size, self.temporal_steps, self.memory_size = window_size, temporal_steps, memory_size
        self.action_punish, self.delta_action_punish, self.agent_power = action_punish, delta_action_punish, agent_power
        self.check_max_value, self.mono, self.t_samples = check_max_value, mono, t_samples
        self.gaussians = prepare_gaussians(self.Lx, self.nx, self.sensor_positions, sigma_sensors, 1, t_samples)
        ga = prepare_gaussians(self.Lx, self.nx, self.sensor_positions, sigma_actuators, 2, t_samples)
        self.gaussians_actuators = ga[self.actuators_to_sensors - 1]

    # canned configurations -------------------------------------------------------------------
    @classmethod
    def ks22(cls, **kw):
        """scripts/KS/KS22/KS22.jl"""
        return cls(Lx=22.0, nx=192, sensor_positions=np.arange(1, 193, 24), actuators_to_sensors=np.arange(1, 9),
                   sigma_sensors=0.7, sigma_actuators=0.7, **kw)

    @classmethod
    def ks200(cls, **kw):
        """scripts/KS/KS200/KS200.jl"""
        return cls(Lx=200.0, nx=240, sensor_positions=np.arange(1, 241, 3), actuators_to_sensors=np.arange(1, 81), **kw)

    @classmethod
    def ks500(cls, **kw):
        """scripts/KS/KS500/KS500.jl (transfer of the KS200 agent)"""
        return cls(Lx=500.0, nx=600, sensor_positions=np.arange(1, 601, 3), actuators_to_sensors=np.arange(1, 201), **kw)

    @classmethod
    def ks256(cls, **kw):
        """BASELINE config C2 (synthetic): nx=256 at KS200's dx, 64 sensors/actuators."""
        kw.setdefault("t_samples", 256 + 100)
        return cls(Lx=200.0 * 256 / 240, nx=256, sensor_positions=np.arange(1, 257, 4),
                   actuators_to_sensors=np.arange(1, 65), **kw)

    def y0_standard(self):
        """y0_1D_standard, KSSetup.jl:53"""
        return np.array([0.5 if 4 <= i <= 44 else 0.0 for i in range(1, self.nx + 1)])

    def generate_random_init(self, rng, n=1):
        """KSSetup.jl:288-298, batched: n initial states from `rng` (numpy Generator)."""
        a = rng.uniform(-1, 1, size=(n, 8))
        a /= np.linalg.norm(a, axis=1, keepdims=True)
        x = (self.Lx / self.nx) * np.arange(1, self.nx + 1)
        basis = np.sin(np.arange(1, 9)[:, None] * x[None, :] / (2 * np.pi))
        y0 = a @ basis
        return y0 * 30 / np.linalg.norm(y0, axis=1, keepdims=True)

    def make_env(self, n_envs=1, dtype="f64", device=0, y0=None, drop_tol=1e-17):
        """initialize_setup(), KSSetup.jl:249-262 -- the PDEenv part.

        drop_tol: basis entries below drop_tol * max(row) are not stored in the device's banded tables.
        1e-17 is below half an ulp of the largest term of each dot product, i.e. lossless in fp64;
        pass 0.0 to keep every non-zero (including the 1e-300 Gaussian tails the reference carries)."""
        if y0 is None:
            y0 = self.y0_standard()
        y0 = np.asarray(y0, dtype=np.float64)
        if y0.ndim == 2:                       # (B, nx) -> reference shape (nx, B)
            y0 = y0.T
        return PDEenv(problem=L.KS, n_envs=n_envs, dtype=dtype, device=device, sensor_basis=self.gaussians,
                      actuator_basis=self.gaussians_actuators, actuators_to_sensors=self.actuators_to_sensors,
                      y0=y0, drop_tol=drop_tol, nx=self.nx, ny=1, Lx=self.Lx, dt=self.dt, te=self.te,
                      oversampling=self.oversampling, mu=self.mu, max_value=self.max_value,
                      window_size=self.window_size, temporal_steps=self.temporal_steps, memory_size=self.memory_size,
                      action_punish=self.action_punish, delta_action_punish=self.delta_action_punish,
                      agent_power=self.agent_power, check_max_value=self.check_max_value, mono=int(self.mono),
                      obs_scale=1.0 / self.max_value, reward_div=3.0 * self.max_value)
