"""GPU parity: device replay + DDPG update (critic/actor gradients, ADAM, Polyak) vs the agent oracle.
fp32 tolerance 1e-5 relative on gradients of small batches (summation order differs), see DESIGN.md."""
import importlib

import numpy as np
import pytest

from conftest import relerr
from oracle import agent_oracle as AO
from test_agent_oracle import batch, make_nets

pytestmark = pytest.mark.gpu


def to_chain(A, net):
    return A.Chain(*[A.Dense(W, b, act) for W, b, act in net.layers])


def make_policy(pkg, setup, actor, critic, B=2, **kw):
    A = pkg.agent
    env = setup.make_env(n_envs=B, dtype="f64", y0=setup.y0_standard())
    pol = A.CustomDDPGPolicy(env, behavior_actor=to_chain(A, actor), behavior_critic=to_chain(A, critic),
                             trajectory_length=kw.pop("trajectory_length", 4096), **kw)
    return env, pol


@pytest.mark.parametrize("shape", ["ks", "ks_w3_middle", "kseg_like"])
@pytest.mark.parametrize("B", [3, 257])
@pytest.mark.parametrize("literal", [True, False])
def test_ddpg_update_matches_oracle(pkg, shape, B, literal):
    rng = np.random.default_rng(4)
    if shape == "ks":
        ns, ha, hc, middle, setup = 1, 6, 140, False, pkg.setups.KSSetup.ks22()
    elif shape == "ks_w3_middle":
        ns, ha, hc, middle, setup = 3, 6, 20, True, pkg.setups.KSSetup.ks22(window_size=3)
    else:
        ns, ha, hc, middle, setup = 12, 20, 340, False, pkg.setups.KSSetup.ks22(window_size=3, temporal_steps=4)
    actor, critic = make_nets(rng, ns, 1, ha, hc, middle)
    env, pol = make_policy(pkg, setup, actor, critic, literal_q1=literal)
    ref = AO.DDPG(actor.copy(), critic.copy())
    for it in range(3):                                   # three consecutive updates: ADAM state carries over
        s, a, r, t, s2 = batch(rng, ns, 1, B)
        pol.set_batch(s, a, r, t, s2)
        pol.update()
        gc, ga = ref.update(s, a, r, t, s2, literal)
        g = pol.grads()
        gref = np.concatenate([AO.flat_grads(gc), AO.flat_grads(ga)])
        assert relerr(g, gref) < 2e-5, (it, relerr(g, gref))
        for dev, orc in ((pol.behavior_critic, ref.C), (pol.behavior_actor, ref.A), (pol.target_critic, ref.Ct),
                         (pol.target_actor, ref.At)):
            assert relerr(dev.sync_from_device().flat(), orc.flat()) < 2e-5
        ls = pol.losses
        assert abs(ls["critic_loss"] - float(ref.critic_loss)) < 1e-4 * max(1.0, abs(float(ref.critic_loss)))
        assert abs(ls["actor_loss"] - float(ref.actor_loss)) < 1e-4 * max(1.0, abs(float(ref.actor_loss)))
    env.close()


def test_replay_ring_matches_oracle_including_wrap(pkg):
    """Literal CircularArraySARTTrajectory semantics (capacity+1 / capacity rings) through 2 episodes + wrap."""
    rng = np.random.default_rng(8)
    setup = pkg.setups.KSSetup.ks22()
    actor, critic = make_nets(rng, 1, 1, 6, 140)
    B = 2
    ncols = B * 8
    cap = 5 * ncols + 3
    env, pol = make_policy(pkg, setup, actor, critic, B=B, trajectory_length=cap, start_steps=-1)
    env.set_y0(setup.generate_random_init(rng, B).T)
    tr = AO.Trajectory(cap, 1, 1)
    L = pkg.lib
    for ep in range(2):
        env.reset()
        pol.trajectory.pre_episode(); tr.pre_episode(ncols)
        for step in range(4):
            a = rng.uniform(-1, 1, (1, ncols))
            env.put(L.ARR_ACTION_IN, a.T)
            st = env.state.astype(np.float32)
            pol.trajectory.pre_act(); tr.pre_act(st, a.astype(np.float32))
            env.step_device(); env.synchronize()
            pol.trajectory.post_act(); tr.post_act(env.reward.astype(np.float32), False)
        pol.trajectory.post_episode(); tr.post_episode(env.state.astype(np.float32), 1)
        assert len(pol.trajectory) == len(tr)
    inds = rng.integers(0, len(tr) - ncols, 64)
    pol.sample(inds)
    s, a, r, t, s2 = tr.fetch(inds, ncols)
    # read the sampled batch back through a critic-gradient call is indirect; compare via a probe update instead:
    ref = AO.DDPG(actor.copy(), critic.copy())
    pol.update(local_batch=64)
    gc, ga = ref.update(s, a, r, t.astype(bool), s2, True)
    assert relerr(pol.grads(), np.concatenate([AO.flat_grads(gc), AO.flat_grads(ga)])) < 2e-5
    env.close()


def test_training_loop_smoke(pkg):
    """run_episode drives policy -> PreAct(push, update) -> env -> PostAct in the reference's stage order."""
    rng = np.random.default_rng(1)
    setup = pkg.setups.KSSetup.ks22()
    B = 4
    env = setup.make_env(n_envs=B, dtype="f32", y0=setup.generate_random_init(rng, B))
    pol = pkg.agent.create_agent(env, rng=rng, nna_scale=0.6, nna_scale_critic=7.0, drop_middle_layer=True,
                                 batch_size=64, start_steps=2, update_after=3, update_loops=2, trajectory_length=10_000)
    n = pkg.agent.run_episode(pol, env)
    assert n == 51
    assert len(pol.trajectory) == 51 * B * 8
    assert pol.n_updates == 2 * (51 - 4)                 # updates start once len(traj) > update_after*ncols
    w = pol.behavior_actor.sync_from_device().flat()
    assert np.all(np.isfinite(w))
    n2 = pkg.agent.run_episode(pol, env)
    assert n2 == 51 and len(pol.trajectory) == 2 * 51 * B * 8
    env.close()


@pytest.mark.parametrize("shape,path", [
    ("wide_critic_middle", 0),      # does not fit the fused smem kernels -> layer-wise path, tensor cores for 340x340
    ("wide_critic_middle", 1),      # same, CUDA cores only
    ("wide_both_middle", 2),        # wide actor too, tensor cores wherever possible
    ("ks_small", 3),                # shipped-size networks forced through the layer-wise path
])
@pytest.mark.parametrize("literal", [True, False])
def test_layerwise_ddpg_update_matches_oracle(pkg, shape, path, literal):
    """update! (PDEagent.jl:363-418) for `drop_middle_layer = false` networks: forward, input-gradient and
    weight-gradient GEMMs (tcgen05 3xTF32 / CUDA cores) vs the oracle; three updates so ADAM state carries."""
    rng = np.random.default_rng(12)
    if shape == "wide_critic_middle":
        ns, ha, hc, mid_a, mid_c, B = 12, 20, 340, False, True, 1500
        setup = pkg.setups.KSSetup.ks22(window_size=3, temporal_steps=4)
    elif shape == "wide_both_middle":
        ns, ha, hc, mid_a, mid_c, B = 12, 96, 140, True, True, 777
        setup = pkg.setups.KSSetup.ks22(window_size=3, temporal_steps=4)
    else:
        ns, ha, hc, mid_a, mid_c, B = 1, 6, 140, False, False, 300
        setup = pkg.setups.KSSetup.ks22()
    actor, _ = make_nets(rng, ns, 1, ha, hc, mid_a)
    _, critic = make_nets(rng, ns, 1, ha, hc, mid_c)
    env, pol = make_policy(pkg, setup, actor, critic, literal_q1=literal)
    pol.set_update_path(path)
    ref = AO.DDPG(actor.copy(), critic.copy())
    for it in range(3):
        s, a, r, t, s2 = batch(rng, ns, 1, B)
        pol.set_batch(s, a, r, t, s2)
        pol.update()
        gc, ga = ref.update(s, a, r, t, s2, literal)
        g = pol.grads()
        nc = len(AO.flat_grads(gc))
        # first update: identical weights on both sides -> kernel parity at 2e-5.  Later updates start from weights
        # that already differ by the ADAM-amplified round-off (see below), so only the trajectory is compared.
        tol_g = 2e-5 if it == 0 else 1e-3
        assert relerr(g[:nc], AO.flat_grads(gc)) < tol_g, (it, "critic", relerr(g[:nc], AO.flat_grads(gc)))
        assert relerr(g[nc:], AO.flat_grads(ga)) < tol_g, (it, "actor", relerr(g[nc:], AO.flat_grads(ga)))
        # ADAM's m / sqrt(v) is scale free: for the many near-zero gradient entries of a 1e5-parameter network a
        # summation-order difference of the gradient (checked to 2e-5 of its max above) moves the step by up to
        # ~lr * O(1e-2), so the weights are compared at 2e-4 of their max here (2e-5 for the small networks above).
        for dev, orc in ((pol.behavior_critic, ref.C), (pol.behavior_actor, ref.A), (pol.target_critic, ref.Ct),
                         (pol.target_actor, ref.At)):
            assert relerr(dev.sync_from_device().flat(), orc.flat()) < 2e-4
        ls = pol.losses
        assert abs(ls["critic_loss"] - float(ref.critic_loss)) < 1e-4 * max(1.0, abs(float(ref.critic_loss)))
        assert abs(ls["actor_loss"] - float(ref.actor_loss)) < 1e-4 * max(1.0, abs(float(ref.actor_loss)))
    env.close()
