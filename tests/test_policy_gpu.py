"""GPU parity: fused actor forward + env step (rollout) vs oracle MLP + oracle env."""
import numpy as np
import pytest

from conftest import relerr
from oracle import ks_oracle as K

pytestmark = pytest.mark.gpu


def mlp_ref(chain, x):
    """Flux Chain of Dense on a (ns, ncols) matrix, float32."""
    h = x.astype(np.float32)
    for l in chain.layers:
        h = l.W @ h + l.b[:, None]
        if l.act == "relu":
            h = np.maximum(h, 0)
        elif l.act == "tanh":
            h = np.tanh(h)
    return h


def _agent_mod(pkg):
    import importlib
    return importlib.import_module(pkg.__name__ + ".agent")


@pytest.mark.parametrize("window", [1, 3])
def test_rollout_matches_stepwise_oracle(pkg, golden, window):
    A = _agent_mod(pkg)
    cfg = K.ks256_config(window)
    setup = pkg.setups.KSSetup.ks256(window_size=window)
    rng = np.random.default_rng(5)
    B, n_a = 5, cfg.n_actuators
    if window == 1:
        g = golden("ks200_hook")                                   # KS200 bestNNA actor 1->6->1
        chain = A.Chain(A.Dense(g["best_W1"], g["best_b1"], "relu"), A.Dense(g["best_W2"], g["best_b2"], "tanh"))
    else:
        chain = A.create_chain(na=1, ns=3, is_actor=True, rng=rng, nna_scale=0.6, drop_middle_layer=True)
    y0 = setup.generate_random_init(rng, B)
    env = setup.make_env(n_envs=B, dtype="f64", y0=y0)
    A.CustomNeuralNetworkApproximator(env, pkg.lib.NET_BEHAVIOR_ACTOR, chain)
    steps = 4
    rsum = env.rollout(steps, act_limit=1.0, reward_sum=True)
    for b in range(B):
        ref = K.KSEnv(cfg, y0=y0[b])
        acc = 0.0
        for _ in range(steps):
            a = np.clip(mlp_ref(chain, ref.state).astype(np.float64), -1, 1)
            ref.step(a)
            acc += ref.reward.mean()
        # fp32 actor inside an fp64 PDE: the action differs at 1e-7 relative, and so does everything downstream
        assert relerr(env.y[:, b], ref.y) < 1e-5
        assert abs(rsum[b] - acc) < 1e-5 * max(1.0, abs(acc))
        assert relerr(env.action[0, b * n_a:(b + 1) * n_a], ref.action[0]) < 1e-5
    assert np.all(env.steps == steps)
    env.close()


def test_policy_act_then_step_equals_rollout(pkg, golden):
    A = _agent_mod(pkg)
    g = golden("ks200_hook")
    chain = A.Chain(A.Dense(g["best_W1"], g["best_b1"], "relu"), A.Dense(g["best_W2"], g["best_b2"], "tanh"))
    setup = pkg.setups.KSSetup.ks256()
    rng = np.random.default_rng(9)
    B = 8
    y0 = setup.generate_random_init(rng, B)
    e1 = setup.make_env(n_envs=B, dtype="f64", y0=y0)
    e2 = setup.make_env(n_envs=B, dtype="f64", y0=y0)
    for e in (e1, e2):
        A.CustomNeuralNetworkApproximator(e, pkg.lib.NET_BEHAVIOR_ACTOR, chain)
    e1.rollout(3)
    for _ in range(3):
        e2.policy_act()
        e2.step_device()
    e2.synchronize()
    assert np.array_equal(e1.y, e2.y)
    assert np.array_equal(e1.reward, e2.reward)
    assert np.array_equal(e1.state, e2.state)
    e1.close(); e2.close()


def test_policy_noise_and_clamp(pkg):
    A = _agent_mod(pkg)
    setup = pkg.setups.KSSetup.ks22()
    rng = np.random.default_rng(2)
    B = 3
    env = setup.make_env(n_envs=B, dtype="f64", y0=setup.generate_random_init(rng, B))
    chain = A.create_chain(na=1, ns=1, is_actor=True, rng=rng, nna_scale=0.6, drop_middle_layer=True)
    A.CustomNeuralNetworkApproximator(env, pkg.lib.NET_BEHAVIOR_ACTOR, chain)
    noise = rng.standard_normal(B * 8)
    s = env.state
    env.policy_act(noise=noise, act_noise=1.2, act_limit=1.0)
    env.step_device(); env.synchronize()
    expect = np.clip(mlp_ref(chain, s).astype(np.float64) + noise[None, :] * 1.2, -1, 1)
    assert relerr(env.action, expect) < 1e-6
    assert np.max(np.abs(env.action)) <= 1.0
    env.close()


def test_act_step_host_equals_policy_act_then_step_host(pkg, golden):
    """`action = policy(env); env(action)` as ONE C call (pdeb200_act_step_host, one packed result copy) must give exactly
    what the two calls give (PDEagent.jl:175-209 then PDEenv.jl:195-241)."""
    import ctypes as C
    A = _agent_mod(pkg)
    L = pkg.lib
    g = golden("ks200_hook")
    chain = A.Chain(A.Dense(g["best_W1"], g["best_b1"], "relu"), A.Dense(g["best_W2"], g["best_b2"], "tanh"))
    setup = pkg.setups.KSSetup.ks200()
    B = 5
    y0 = setup.generate_random_init(np.random.default_rng(2), B)
    envs = [setup.make_env(n_envs=B, dtype="f64", y0=y0) for _ in range(2)]
    for e in envs:
        A.CustomNeuralNetworkApproximator(e, L.NET_BEHAVIOR_ACTOR, chain.copy())
    noise = np.random.default_rng(3).standard_normal(B * 80)
    a, b = envs
    offs = [C.c_size_t() for _ in range(4)]
    L.check(a._lib.pdeb200_result_layout(a._ctx, *[C.byref(o) for o in offs]), a._ctx)
    r_off, s_off, d_off, total = (o.value for o in offs)
    act = np.zeros(B * 80)
    packed = np.zeros(total, dtype=np.uint8)
    for _ in range(3):
        L.check(a._lib.pdeb200_act_step_host(a._ctx, noise.ctypes.data, 0.3, 1.0, act.ctypes.data, None, packed.ctypes.data, None, None, None), a._ctx)
        b.policy_act(noise, 0.3, 1.0)
        act_b = b.get(L.ARR_ACTION_IN)
        b(act_b.reshape(1, -1))
        assert np.array_equal(act, act_b)
        assert np.array_equal(packed[r_off:r_off + B * 80 * 8].view(np.float64), b.reward)
        assert np.array_equal(packed[s_off:s_off + B * 80 * 8].view(np.float64), b.get(L.ARR_STATE))
        assert np.array_equal(packed[d_off:d_off + B], b.get(L.ARR_DONE))
        assert np.array_equal(a.get(L.ARR_Y), b.get(L.ARR_Y))
    # pdeb200_result_select(ctx, 0): the packed copy shrinks to the [reward | done] prefix of the same block
    assert r_off < d_off < s_off
    L.check(a._lib.pdeb200_result_select(a._ctx, 0), a._ctx)
    short = C.c_size_t()
    L.check(a._lib.pdeb200_result_layout(a._ctx, None, None, None, C.byref(short)), a._ctx)
    assert short.value == s_off
    packed[:] = 0xAB
    L.check(a._lib.pdeb200_act_step_host(a._ctx, noise.ctypes.data, 0.3, 1.0, None, None, packed.ctypes.data, None, None, None), a._ctx)
    b.policy_act(noise, 0.3, 1.0)
    b(b.get(L.ARR_ACTION_IN).reshape(1, -1))
    assert np.array_equal(packed[r_off:r_off + B * 80 * 8].view(np.float64), b.reward)
    assert np.array_equal(packed[d_off:d_off + B], b.get(L.ARR_DONE))
    assert np.all(packed[s_off:] == 0xAB)                                   # the observation stayed on the device ...
    assert np.array_equal(a.get(L.ARR_STATE), b.get(L.ARR_STATE))           # ... and is still there on demand
    for e in envs:
        e.close()


def test_prefetched_noise_gives_the_same_steps_as_noise_handed_over_with_the_call(pkg, golden):
    """pdeb200_noise_prefetch: step i+1's host noise is handed over BEFORE the call for step i (two outstanding at most) and
    consumed in order; results must equal the calls that carry their noise themselves (randn inside the policy call,
    PDEagent.jl:201), and a third outstanding prefetch is refused."""
    import ctypes as C
    A = _agent_mod(pkg)
    L = pkg.lib
    g = golden("ks200_hook")
    chain = A.Chain(A.Dense(g["best_W1"], g["best_b1"], "relu"), A.Dense(g["best_W2"], g["best_b2"], "tanh"))
    setup = pkg.setups.KSSetup.ks200()
    B, steps = 4, 5
    y0 = setup.generate_random_init(np.random.default_rng(5), B)
    a, b = [setup.make_env(n_envs=B, dtype="f64", y0=y0) for _ in range(2)]
    for e in (a, b):
        A.CustomNeuralNetworkApproximator(e, L.NET_BEHAVIOR_ACTOR, chain.copy())
    noise = np.random.default_rng(6).standard_normal((steps, B * 80))
    tot = C.c_size_t()
    L.check(a._lib.pdeb200_result_layout(a._ctx, None, None, None, C.byref(tot)), a._ctx)
    pa, pb = np.zeros(tot.value, dtype=np.uint8), np.zeros(tot.value, dtype=np.uint8)
    lib = a._lib
    L.check(lib.pdeb200_noise_prefetch(a._ctx, noise[0].ctypes.data), a._ctx)
    for i in range(steps):
        if i + 1 < steps:
            L.check(lib.pdeb200_noise_prefetch(a._ctx, noise[i + 1].ctypes.data), a._ctx)
            assert lib.pdeb200_noise_prefetch(a._ctx, noise[i + 1].ctypes.data) == -4          # PDEB200_ESTATE: two are outstanding
        L.check(lib.pdeb200_act_step_host(a._ctx, None, 0.3, 1.0, None, None, pa.ctypes.data, None, None, None), a._ctx)
        L.check(lib.pdeb200_act_step_host(b._ctx, noise[i].ctypes.data, 0.3, 1.0, None, None, pb.ctypes.data, None, None, None), b._ctx)
        assert np.array_equal(pa, pb)
        assert np.array_equal(a.get(L.ARR_Y), b.get(L.ARR_Y))
        assert np.array_equal(a.get(L.ARR_ACTION_IN), b.get(L.ARR_ACTION_IN))
    for e in (a, b):
        e.close()
