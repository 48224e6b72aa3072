"""Fluid (2-D Navier-Stokes) setup: what scripts/Fluid/setup/FluidSetup.jl does around the hot path.

Host-side construction of the sensor / actuator bases (`prepare_gaussians`, FluidSetup.jl:139-157, built
from `taylorvtx`, src/fluid_rk4.jl:54-69) and of initial conditions (`ic`, fluid_rk4.jl:72-120); the bases go
to the library as arrays in Julia's column-major flattening of the (ny, nx) matrices.  The wavenumber tables
(FluidSetup.jl:103-124) are derived inside libpdeb200 from (nx, Lx, Ly).

By default the env steps with the FIXED-step `do_step` (RK4 x oversampling, FluidSetup.jl:163-172), the path named by
BASELINE.json.  The shipped scripts wire the adaptive `do_step2` (quirk Q9): `make_env(adaptive=True, rtol, atol)`
selects the error-controlled mode (per-environment step control; the reference's tolerance is 1e0, FluidSetup.jl:178-179).
"""
import numpy as np

from .. import _lib as L
from ..env import PDEenv


class FluidSetup:
    """Globals of scripts/Fluid/Fluid_*/Fluid_*.jl:11-16 + FluidSetup.jl:28-101."""

    def __init__(self, nx=128, sensors_per_axis=16, variance=0.04, Lx=1.0, Ly=1.0, nu=0.00005, te=6.0, dt=0.02,
                 oversampling=None, ifpad=1, window_size=3, temporal_steps=1, memory_size=0, action_punish=0.002,
                 delta_action_punish=0.002, agent_power=70.0, max_value=3.0, check_max_value="reward"):
        self.nx = self.ny = int(nx)
        self.Lx, self.Ly = float(Lx), float(Ly)
        self.dx, self.dy = self.Lx / self.nx, self.Ly / self.ny
        self.sensors_per_axis, self.variance = int(sensors_per_axis), float(variance)
        self.nu, self.te, self.dt, self.ifpad = nu, te, dt, int(ifpad)
        self.oversampling = int(np.floor(16 * self.nx * dt)) if oversampling is None else int(oversampling)   # :48
        self.window_size, self.temporal_steps, self.memory_size = window_size, temporal_steps, memory_size
        self.action_punish, self.delta_action_punish = action_punish, delta_action_punish
        self.agent_power, self.max_value, self.check_max_value = agent_power, max_value, check_max_value
        # x1 = range(0, Lx, length = nx + 1)[1:nx]; xx[j, i] = x[i], yy[j, i] = y[j]   (:126-133, meshgrid)
        x1 = np.linspace(0.0, self.Lx, self.nx + 1)[:self.nx]
        y1 = np.linspace(0.0, self.Ly, self.ny + 1)[:self.ny]
        self.xx = np.tile(x1[None, :], (self.ny, 1))
        self.yy = np.tile(y1[:, None], (1, self.nx))
        sx, sy = self.nx // self.sensors_per_axis, self.ny // self.sensors_per_axis
        self.sensor_positions = [(i, j) for i in range(1, self.nx + 1, sx) for j in range(1, self.ny + 1, sy)]   # :61
        self.actuators_to_sensors = np.arange(1, len(self.sensor_positions) + 1)
        self.gaussians = self.prepare_gaussians(1)
        self.gaussians_actuators = self.prepare_gaussians(2)[self.actuators_to_sensors - 1]

    @classmethod
    def fluid16(cls, evaluation=False, **kw):
        """scripts/Fluid/Fluid_16/Fluid_16.jl (nx = 256 when `evaluation`, FluidSetup.jl:33-36)"""
        return cls(nx=256 if evaluation else 128, sensors_per_axis=16, variance=0.04, **kw)

    @classmethod
    def fluid8(cls, evaluation=False, **kw):
        return cls(nx=256 if evaluation else 128, sensors_per_axis=8, variance=0.08, **kw)

    @classmethod
    def fluid32(cls, evaluation=False, **kw):
        return cls(nx=256 if evaluation else 128, sensors_per_axis=32, variance=0.022, **kw)

    def taylorvtx_phys(self, x0, y0, a0, U_max):
        """fluid_rk4.jl:54-64: periodic sum of nine Taylor vortices, physical space, (ny, nx)."""
        omg = np.zeros_like(self.xx)
        for i in (-1, 0, 1):
            for j in (-1, 0, 1):
                r2 = (self.xx - x0 - i * self.Lx) ** 2 + (self.yy - y0 - j * self.Ly) ** 2
                omg = omg + U_max / a0 * (2 - r2 / a0 ** 2) * np.exp(0.5 * (1 - r2 / a0 ** 2))
        return omg

    def taylorvtx(self, x0, y0, a0, U_max):
        return np.fft.fft2(self.taylorvtx_phys(x0, y0, a0, U_max))

    def prepare_gaussians(self, norm_mode=1):
        """FluidSetup.jl:139-157 -> (n_sensors, ny, nx) float64 (dense; the device tables are sparse)."""
        out = np.empty((len(self.sensor_positions), self.ny, self.nx))
        for n, (pi, pj) in enumerate(self.sensor_positions):
            p = np.real(np.fft.ifft2(self.taylorvtx(pi * self.dx - self.dx, pj * self.dy - self.dy, self.variance, 1.0)))
            p[p < 0.1] = 0.0
            out[n] = p / (p.sum() if norm_mode == 1 else p.max())
        return out

    def ic(self, caseno, rng=None):
        """fluid_rk4.jl:72-120 with a numpy Generator (Julia's RNG stream is an input, not reproducible)."""
        Lx, Ly = self.Lx, self.Ly
        if caseno == 1:
            return self.taylorvtx(Lx / 2, Ly / 2, Lx / 8, 1.0)
        if caseno == 2:
            return self.taylorvtx(Lx / 2, 0.4 * Ly, Lx / 10.0, 1.0) + self.taylorvtx(Lx / 2, 0.6 * Ly, Lx / 10, 1.0)
        nv = 30 if caseno == 3 else 50
        omg = np.zeros_like(self.xx)
        for _ in range(nv):
            x0, y0 = rng.random() * Lx, rng.random() * Ly
            a0 = Lx / 20 if caseno == 3 else Lx / 20 * (0.5 + rng.random())
            omg = omg + self.taylorvtx_phys(x0, y0, a0, rng.random() * 2 - 1.0)
        return np.fft.fft2(omg)

    def generate_random_init(self, rng, n=1, caseno=3):
        """FluidSetup.jl:387-395, batched -> complex (n, ny, nx)"""
        return np.stack([self.ic(caseno, rng) for _ in range(n)])

    @staticmethod
    def _julia_flat(a):
        """(..., ny, nx) -> (..., nx*ny) in Julia column-major order (row index j fastest)."""
        a = np.asarray(a)
        return np.ascontiguousarray(np.swapaxes(a, -1, -2)).reshape(a.shape[:-2] + (-1,))

    def make_env(self, n_envs=1, dtype="f64", device=0, y0=None, drop_tol=0.0, adaptive=False, rtol=1.0, atol=1.0):
        """initialize_setup(), FluidSetup.jl:330-343 -- the PDEenv part.  adaptive=True: `do_step2`'s role
        (FluidSetup.jl:181-186), error-controlled RK4 with step doubling per environment at (rtol, atol)."""
        y0 = self.ic(2) if y0 is None else np.asarray(y0)
        if y0.ndim == 3:                      # (B, ny, nx) -> reference shape (ny, nx, B)
            y0 = y0.transpose(1, 2, 0)
        return PDEenv(problem=L.NS2D, adaptive=int(bool(adaptive)), rtol=float(rtol), atol=float(atol), n_envs=n_envs,
                      dtype=dtype, device=device,
                      sensor_basis=self._julia_flat(self.gaussians),
                      actuator_basis=self._julia_flat(self.gaussians_actuators),
                      actuators_to_sensors=self.actuators_to_sensors, y0=y0, drop_tol=drop_tol,
                      nx=self.nx, ny=self.ny, Lx=self.Lx, Ly=self.Ly, nu=self.nu, dt=self.dt, te=self.te,
                      oversampling=self.oversampling, ifpad=self.ifpad, sensors_per_axis=self.sensors_per_axis,
                      max_value=self.max_value, window_size=self.window_size, temporal_steps=self.temporal_steps,
                      memory_size=self.memory_size, action_punish=self.action_punish,
                      delta_action_punish=self.delta_action_punish, agent_power=self.agent_power,
                      check_max_value=self.check_max_value)
