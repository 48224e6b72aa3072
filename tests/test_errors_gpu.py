"""Error behaviour of the C ABI on a GPU box: status codes + messages instead of crashes, no silent fallbacks."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _raw_ctx(pkg, **over):
    L = pkg.lib
    lib = L.load()
    cfg = L.Config()
    L.check(lib.pdeb200_default_config(L.KS, C.byref(cfg)))
    cfg.nx, cfg.n_envs, cfg.n_sensors, cfg.n_actuators = 192, 2, 8, 8
    for k, v in over.items():
        setattr(cfg, k, v)
    ctx = C.c_void_p()
    rc = lib.pdeb200_create(C.byref(cfg), 0, C.byref(ctx))
    return lib, ctx, rc


def test_unsupported_grid_and_bad_sizes_are_rejected(pkg):
    L = pkg.lib
    lib, ctx, rc = _raw_ctx(pkg, nx=250)
    assert rc == -3 and b"nx must be one of" in lib.pdeb200_last_error(None)            # PDEB200_EUNSUPPORTED
    lib, ctx, rc = _raw_ctx(pkg, window_size=2)
    assert rc == -1                                                                     # even window: PDEB200_EINVAL
    lib, ctx, rc = _raw_ctx(pkg, problem=L.NS2D, nx=100, ny=100, n_sensors=16, n_actuators=16, sensors_per_axis=4)
    assert rc == -3 and b"NS" in lib.pdeb200_last_error(None)


def test_calls_out_of_order_return_estate(pkg):
    lib, ctx, rc = _raw_ctx(pkg)
    assert rc == 0
    a = np.zeros(16)
    assert lib.pdeb200_step(ctx, a.ctypes.data) == -4                                   # set_bases / set_y0 first
    assert b"set_bases" in lib.pdeb200_last_error(ctx)
    assert lib.pdeb200_reset(ctx, None) == -4
    assert lib.pdeb200_policy_act(ctx, None, 0.0, 1.0) == -4                            # actor not set
    assert lib.pdeb200_ddpg_update(ctx, 0.99, 0.995, 5e-4, 1e-3, 1) == -4               # no batch
    y = np.zeros(5)
    assert lib.pdeb200_get(ctx, pkg.lib.ARR_Y, y.ctypes.data, y.nbytes) == -1           # size mismatch
    assert b"size mismatch" in lib.pdeb200_last_error(ctx)
    used = C.c_int32()
    x = np.zeros((4, 1), np.float32); out = np.zeros((4, 1), np.float32)
    assert lib.pdeb200_net_forward(ctx, 0, 4, x.ctypes.data, out.ctypes.data, 0, C.byref(used)) == -4
    assert lib.pdeb200_destroy(ctx) == 0


def test_actor_shape_mismatch_is_einval(pkg):
    A = pkg.agent
    setup = pkg.setups.KSSetup.ks22(window_size=3)
    env = setup.make_env(n_envs=2, dtype="f64", y0=setup.y0_standard())
    rng = np.random.default_rng(0)
    bad = A.create_chain(na=1, ns=1, is_actor=True, rng=rng, nna_scale=0.6, drop_middle_layer=True)   # ns should be 3
    A.CustomNeuralNetworkApproximator(env, pkg.lib.NET_BEHAVIOR_ACTOR, bad)
    with pytest.raises(pkg.PdeB200Error, match="do not match"):
        env.policy_act()
    with pytest.raises(pkg.PdeB200Error, match="do not match"):
        env.rollout(1)
    with pytest.raises(ValueError):
        env(np.zeros((1, 5)))
    env.close()
