"""Host-side mirror of the reference's `PDEenv` (src/PDEenv.jl) over the C ABI.

Same field names, argument meaning and call order as the Julia struct
(PDEenv.jl:26-62); the four closures `prepare_action` / `do_step` /
`reward_function` / `featurize` (PDEenv.jl:31-35) are replaced by the library's
env step (csrc/: actuation -> PDE core -> observation, three sm_100a launches on one stream).  The environment batch B
is folded into the actuator (column) axis exactly as SURVEY.md 8b prescribes, so
every array has the reference's Julia shape with `n_act*B` columns:

    env.state   (ns, n_act*B)        env.action  (1+mem, n_act*B)
    env.reward  (n_act*B,)           env.y       (nx, B) | (2, nx, B) | (2, nx, ny, B) | complex (ny, nx, B)

(returned as numpy views whose memory is the Julia column-major layout).
"""
import ctypes as C

import numpy as np

from . import _lib as L


def _np_dtype(dtype):
    return np.float64 if dtype == L.F64 else np.float32


class PDEenv:
    """Batched PDE environment.  Mirrors `PDEenv(; ...)` (PDEenv.jl:64-170).

    Parameters (keyword, reference names where they exist)
      problem            L.KS | L.KSEG1D | L.KSEG2D | L.NS2D
      n_envs             B independent environments (reference: 1)
      dtype              "f64" | "f32" arithmetic of the PDE path
      sensor_basis       `gaussians`            (n_sensors, npts) float64
      actuator_basis     `gaussians_actuators`  (n_actuators, npts) float64
      actuators_to_sensors   1-BASED like the Julia scripts
      y0                 (y_shape) broadcast or (B, y_shape)
      te, t0, dt, oversampling, max_value, check_max_value  -- as PDEenv.jl:64-82
      plus the setup-file constants (window_size, temporal_steps, memory_size,
      agent_power, mu, ...) that the reference reads from globals.
    """

    def __init__(self, *, problem, n_envs=1, dtype="f64", device=0, sensor_basis, actuator_basis,
                 actuators_to_sensors, y0, drop_tol=0.0, **kw):
        lib = L.load()
        self._lib = lib
        self._ctx = C.c_void_p()
        cfg = L.Config()
        L.check(lib.pdeb200_default_config(problem, C.byref(cfg)))
        cfg.dtype = L.F64 if dtype in ("f64", np.float64, L.F64) and dtype != L.F32 else L.F32
        if dtype in ("f32", np.float32):
            cfg.dtype = L.F32
        cfg.n_envs = int(n_envs)
        sensor_basis = np.ascontiguousarray(sensor_basis, dtype=np.float64)
        actuator_basis = np.ascontiguousarray(actuator_basis, dtype=np.float64)
        cfg.n_sensors = sensor_basis.shape[0]
        cfg.n_actuators = actuator_basis.shape[0]
        names = {n for n, _ in L.Config._fields_}
        cmv = kw.pop("check_max_value", None)
        if cmv is not None:
            cfg.check_max_value = {"y": L.CHECK_Y, "reward": L.CHECK_REWARD}.get(cmv, L.CHECK_NONE) \
                if isinstance(cmv, str) else int(cmv)
        for k, v in kw.items():
            if k not in names:
                raise TypeError("PDEenv: unknown keyword %r" % k)
            setattr(cfg, k, v)
        self.cfg = cfg
        self.np_dtype = _np_dtype(cfg.dtype)
        L.check(lib.pdeb200_create(C.byref(cfg), int(device), C.byref(self._ctx)))
        self.device_index = int(device)
        self.n_envs = cfg.n_envs
        self.n_actuators = cfg.n_actuators
        self.n_sensors = cfg.n_sensors
        self.a_rows = 1 + cfg.memory_size
        self.ns = lib.pdeb200_obs_rows(self._ctx)
        self.n_cols = lib.pdeb200_obs_cols(self._ctx)
        self.n_rew = 1 if cfg.mono else cfg.n_actuators
        npts = cfg.nx * cfg.ny
        if sensor_basis.shape[1] != npts or actuator_basis.shape[1] != npts:
            raise ValueError("basis arrays must be (n, nx*ny)")
        a2s = np.ascontiguousarray(np.asarray(actuators_to_sensors, dtype=np.int64) - 1, dtype=np.int32)
        if a2s.shape != (cfg.n_actuators,):
            raise ValueError("actuators_to_sensors must have n_actuators entries")
        L.check(lib.pdeb200_set_bases(self._ctx, sensor_basis.ctypes.data, actuator_basis.ctypes.data,
                                      a2s.ctypes.data, float(drop_tol)), self._ctx)
        if problem == L.KS:
            self._y_shape = (cfg.nx,)
        elif problem == L.KSEG1D:
            self._y_shape = (2, cfg.nx)
        elif problem == L.KSEG2D:
            self._y_shape = (cfg.ny, cfg.nx, 2)      # Julia (2, nx, ny): memory [iy][ix][field]
        else:
            self._y_shape = (cfg.nx, cfg.ny, 2)      # Julia complex (ny, nx): memory [col i][row j][re,im]
        self.problem = problem
        self.set_y0(y0)
        # reference field names (PDEenv.jl:26-62)
        self.te, self.t0, self.dt = cfg.te, cfg.t0, cfg.dt
        self.oversampling = cfg.oversampling
        self.max_value = cfg.max_value
        self.check_max_value = cmv
        self.reset()

    # -- lifecycle ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.pdeb200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- raw array access ----------------------------------------------------------------------
    def _elems(self, which):
        B = self.n_envs
        ye = int(np.prod(self._y_shape))
        pe = ye if self.problem == L.NS2D else self.cfg.nx * self.cfg.ny
        return {
            L.ARR_Y: (B * ye, self.np_dtype), L.ARR_Y0: (B * ye, self.np_dtype), L.ARR_P: (B * pe, self.np_dtype),
            L.ARR_STATE: (B * self.n_cols * self.ns, self.np_dtype),
            L.ARR_ACTION: (B * self.n_actuators * self.a_rows, self.np_dtype),
            L.ARR_DELTA_ACTION: (B * self.n_actuators * self.a_rows, self.np_dtype),
            L.ARR_ACTION_IN: (B * self.n_actuators * self.a_rows, self.np_dtype),
            L.ARR_REWARD: (B * self.n_rew, self.np_dtype), L.ARR_DONE: (B, np.uint8),
            L.ARR_TIME: (B, np.float64), L.ARR_STEPS: (B, np.int32), L.ARR_NSUB: (2 * B, np.int32),
            L.ARR_SENSORS: (B * (2 if self.problem in (L.KSEG1D, L.KSEG2D) else 1) * self.n_sensors, self.np_dtype),
        }[which]

    def get(self, which):
        n, dt = self._elems(which)
        out = np.empty(n, dtype=dt)
        L.check(self._lib.pdeb200_get(self._ctx, which, out.ctypes.data, out.nbytes), self._ctx)
        return out

    def get_env(self, which, b):
        """One environment's slice of a per-environment array (no whole-batch copy)."""
        n, dt = self._elems(which)
        out = np.empty(n // self.n_envs, dtype=dt)
        L.check(self._lib.pdeb200_get_env(self._ctx, which, int(b), out.ctypes.data, out.nbytes), self._ctx)
        return out

    def y_of(self, b):
        """env.y of environment b in the reference's shape ((nx,), (2,nx), (2,nx,ny), complex (ny,nx))."""
        flat = self.get_env(L.ARR_Y, b)
        saved, self.n_envs = self.n_envs, 1
        try:
            return self._y_view(flat)[..., 0]
        finally:
            self.n_envs = saved

    def p_of(self, b):
        flat = self.get_env(L.ARR_P, b)
        if self.problem == L.NS2D:
            saved, self.n_envs = self.n_envs, 1
            try:
                return self._y_view(flat)[..., 0]
            finally:
                self.n_envs = saved
        return flat

    def put(self, which, arr):
        n, dt = self._elems(which)
        a = np.ascontiguousarray(arr, dtype=dt).reshape(-1)
        if a.size != n:
            raise ValueError("size mismatch: got %d elements, expected %d" % (a.size, n))
        L.check(self._lib.pdeb200_set(self._ctx, which, a.ctypes.data, a.nbytes), self._ctx)

    def device_ptr(self, which):
        p, n = C.c_void_p(), C.c_size_t()
        L.check(self._lib.pdeb200_device_ptr(self._ctx, which, C.byref(p), C.byref(n)), self._ctx)
        return p.value, n.value

    # -- reference-shaped views ----------------------------------------------------------------
    def _y_view(self, flat):
        B = self.n_envs
        if self.problem == L.KS:
            return flat.reshape(B, self.cfg.nx).T                              # (nx, B)
        if self.problem == L.KSEG1D:
            # Julia (2, nx) column-major => memory [x][field]
            return flat.reshape(B, self.cfg.nx, 2).transpose(2, 1, 0)          # (2, nx, B)
        if self.problem == L.KSEG2D:
            return flat.reshape(B, self.cfg.ny, self.cfg.nx, 2).transpose(3, 2, 1, 0)     # (2, nx, ny, B)
        z = flat.reshape(B, self.cfg.nx, self.cfg.ny, 2)
        return (z[..., 0] + 1j * z[..., 1]).transpose(2, 1, 0)                # (ny, nx, B)

    @property
    def y(self):
        return self._y_view(self.get(L.ARR_Y))

    @property
    def y0(self):
        return self._y_view(self.get(L.ARR_Y0))

    @property
    def p(self):
        flat = self.get(L.ARR_P)
        if self.problem == L.NS2D:
            return self._y_view(flat)
        return flat.reshape(self.n_envs, -1).T

    @property
    def state(self):
        return self.get(L.ARR_STATE).reshape(self.n_envs * self.n_cols, self.ns).T

    @property
    def action(self):
        return self.get(L.ARR_ACTION).reshape(self.n_envs * self.n_actuators, self.a_rows).T

    @property
    def delta_action(self):
        return self.get(L.ARR_DELTA_ACTION).reshape(self.n_envs * self.n_actuators, self.a_rows).T

    @property
    def reward(self):
        return self.get(L.ARR_REWARD)

    @property
    def done(self):
        """Per-environment termination flags (reference: one Bool)."""
        return self.get(L.ARR_DONE).astype(bool)

    @property
    def time(self):
        return self.get(L.ARR_TIME)

    @property
    def steps(self):
        return self.get(L.ARR_STEPS)

    @property
    def substeps(self):
        """Adaptive-step mode: (accepted, rejected) integrator steps of the last env step, per environment -> (B, 2)."""
        return self.get(L.ARR_NSUB).reshape(self.n_envs, 2)

    @property
    def sensors(self):
        return self.get(L.ARR_SENSORS)

    # -- reference methods ---------------------------------------------------------------------
    def _y_to_memory(self, y):
        """Accepts reference-shaped y (nx,), (2,nx), complex (ny,nx) [+ trailing batch] -> memory order."""
        y = np.asarray(y)
        B = self.n_envs
        if self.problem == L.KS:
            base = (self.cfg.nx,)
            if y.shape == base:
                return y.astype(np.float64), True
            if y.shape == base + (B,):
                return np.ascontiguousarray(y.T, dtype=np.float64), False
        elif self.problem == L.KSEG1D:
            base = (2, self.cfg.nx)
            if y.shape == base:
                return np.ascontiguousarray(y.T, dtype=np.float64), True
            if y.shape == base + (B,):
                return np.ascontiguousarray(y.transpose(2, 1, 0), dtype=np.float64), False
        elif self.problem == L.KSEG2D:
            base = (2, self.cfg.nx, self.cfg.ny)
            if y.shape == base:
                return np.ascontiguousarray(y.transpose(2, 1, 0), dtype=np.float64), True
            if y.shape == base + (B,):
                return np.ascontiguousarray(y.transpose(3, 2, 1, 0), dtype=np.float64), False
        else:
            base = (self.cfg.ny, self.cfg.nx)
            if y.shape == base:
                z = np.ascontiguousarray(y.T.astype(np.complex128))
                return z.view(np.float64), True
            if y.shape == base + (B,):
                z = np.ascontiguousarray(y.transpose(2, 1, 0).astype(np.complex128))
                return z.view(np.float64), False
        raise ValueError("y has shape %s; expected %s or %s + (B,)" % (y.shape, base, base))

    def set_y0(self, y0):
        mem, bcast = self._y_to_memory(y0)
        mem = np.ascontiguousarray(mem, dtype=np.float64)
        L.check(self._lib.pdeb200_set_y0(self._ctx, mem.ctypes.data, 1 if bcast else 0), self._ctx)

    def set_y(self, y):
        mem, bcast = self._y_to_memory(y)
        if bcast:
            mem = np.broadcast_to(mem.reshape(1, -1), (self.n_envs, mem.size))
        self.put(L.ARR_Y, mem)

    def reset(self, mask=None):
        """RLBase.reset!(env), PDEenv.jl:183-193; `mask` selects environments (batched extension)."""
        m = None
        if mask is not None:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            if m.shape != (self.n_envs,):
                raise ValueError("mask must have n_envs entries")
        L.check(self._lib.pdeb200_reset(self._ctx, m.ctypes.data if m is not None else None), self._ctx)

    def reset_diverged(self, sync=True):
        """Batched termination (PDEenv.jl:226-240 per environment): reset the environments that diverged before the time
        limit; returns (n_done, n_time_limit, n_diverged) -- or None with sync=False (nothing read back)."""
        if not sync:
            L.check(self._lib.pdeb200_reset_diverged(self._ctx, None), self._ctx)
            return None
        counts = np.zeros(3, dtype=np.int32)
        L.check(self._lib.pdeb200_reset_diverged(self._ctx, counts.ctypes.data), self._ctx)
        return int(counts[0]), int(counts[1]), int(counts[2])

    def _action_to_memory(self, action):
        a = np.asarray(action, dtype=self.np_dtype)
        n = self.n_envs * self.n_actuators
        if a.shape == (self.a_rows, n):
            return np.ascontiguousarray(a.T)
        if a.size == n * self.a_rows and self.a_rows == 1:
            return np.ascontiguousarray(a.reshape(n, 1))
        raise ValueError("action must have shape (%d, %d)" % (self.a_rows, n))

    def __call__(self, action):
        """env(action), PDEenv.jl:195-241 for all environments (three launches, no host synchronisation inside)."""
        a = self._action_to_memory(action)
        L.check(self._lib.pdeb200_step(self._ctx, a.ctypes.data), self._ctx)

    step = __call__

    def step_device(self):
        """env step on the action buffer a previous policy call left on the device (no host traffic)."""
        L.check(self._lib.pdeb200_step_device(self._ctx, None), self._ctx)

    def synchronize(self):
        L.check(self._lib.pdeb200_synchronize(self._ctx), self._ctx)

    def step_cost(self):
        b, f = C.c_double(), C.c_double()
        L.check(self._lib.pdeb200_step_cost(self._ctx, C.byref(b), C.byref(f)), self._ctx)
        return b.value, f.value

    def measure_fma_peak(self, dtype=None):
        """Measured CUDA-core FMA peak (TFLOP/s) for "f32" / "f64" (default: the context dtype)."""
        dt = self.cfg.dtype if dtype is None else (L.F64 if dtype in ("f64", L.F64) and dtype != L.F32 else L.F32)
        if dtype == "f32":
            dt = L.F32
        out = C.c_double()
        L.check(self._lib.pdeb200_measure_fma_peak(self._ctx, dt, C.byref(out)), self._ctx)
        return out.value

    @property
    def launch_count(self):
        return int(self._lib.pdeb200_launch_count(self._ctx))

    # -- policy entry points that keep everything on the device ----------------------------------
    def policy_act(self, noise=None, act_noise=0.0, act_limit=1.0):
        """actions = clamp(behavior_actor(state) + noise*act_noise, +-act_limit) into the device action
        staging buffer (src/PDEagent.jl:189-204); follow with step_device()."""
        ptr = None
        if noise is not None:
            noise = np.ascontiguousarray(noise, dtype=np.float64)
            n_out = self.n_actuators * self.a_rows if self.cfg.mono else self.a_rows
            want = self.n_envs * self.n_cols * (n_out - (0 if self.cfg.mono else self.cfg.memory_size))
            if noise.size != want:
                raise ValueError("noise must have %d entries ([B][n_cols][na - mem]), got %d" % (want, noise.size))
            ptr = noise.ctypes.data
        L.check(self._lib.pdeb200_policy_act(self._ctx, ptr, float(act_noise), float(act_limit)), self._ctx)

    def policy_act_rng(self, seed, offset, act_noise, act_limit=1.0):
        L.check(self._lib.pdeb200_policy_act_rng(self._ctx, int(seed), int(offset), float(act_noise), float(act_limit)),
                self._ctx)

    def rollout(self, n_steps, act_limit=1.0, reward_sum=False):
        """n_steps x {actor forward -> env step} enqueued by one call (evaluation loop, src/plotting.jl:55-73)."""
        out = np.zeros(self.n_envs, dtype=np.float64) if reward_sum else None
        L.check(self._lib.pdeb200_rollout(self._ctx, int(n_steps), float(act_limit),
                                          out.ctypes.data if out is not None else None), self._ctx)
        return out
