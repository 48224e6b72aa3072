"""Checkpoint compatibility and the batched PDEhook (8f row 2)."""
from pathlib import Path

import numpy as np
import pytest

REF = Path("/root/reference")


@pytest.mark.skipif(not REF.exists(), reason="the reference tree only exists in the build container")
@pytest.mark.parametrize("rel,name", [("scripts/KS/KS200/saves/hook.jld2", "ks200_hook"), ("scripts/KS/KS22/saves/hook.jld2", "ks22_hook"),
                                      ("scripts/Keller-Segel/Keller-Segel10_16/saves/hook.jld2", "kseg10_16_hook"),
                                      ("scripts/Fluid/Fluid_16/saves/hook.jld2", "fluid16_hook")])
def test_jld2_reader_loads_the_shipped_actors(pkg, golden, rel, name):
    best, cur = pkg.checkpoint.load_hook_actors(REF / rel)
    g = golden(name)
    assert np.array_equal(best.layers[0].W, g["best_W1"]) and np.array_equal(best.layers[1].b, g["best_b2"])
    assert np.array_equal(cur.layers[0].W, g["cur_W1"]) and np.array_equal(cur.layers[1].W, g["cur_W2"])
    assert [l.act for l in best.layers] == ["relu", "tanh"]


def test_chain_from_arrays_rejects_mismatched_shapes(pkg):
    with pytest.raises(ValueError):
        pkg.checkpoint.chain_from_arrays([np.zeros((3, 2), np.float32), np.zeros(4, np.float32)])


@pytest.mark.gpu
def test_hook_and_npz_round_trip_over_a_training_episode(pkg, tmp_path):
    A = pkg.agent
    setup = pkg.setups.KSSetup.ks22(te=0.6)
    B = 4
    env = setup.make_env(n_envs=B, dtype="f64", y0=setup.generate_random_init(np.random.default_rng(0), B))
    pol = A.create_agent(env, rng=np.random.default_rng(1), nna_scale=0.6, nna_scale_critic=7.0, drop_middle_layer=True,
                         batch_size=16, start_steps=1, update_after=1, update_loops=2, trajectory_length=4096)
    hook = pkg.PDEhook(track_env=2)
    for _ in range(2):
        n = A.run_episode(pol, env, hook=hook)
    assert n == 6 and hook.ep == 3 and len(hook.rewards) == 2 and len(hook.rewards_compare) == 2
    assert hook.rewards_per_env[0].shape == (B,) and np.isclose(hook.rewards_per_env[-1].mean(), hook.rewards[-1])
    assert len(hook.bestDF) == 6 and hook.bestDF[0]["y"].shape == (192,) and hook.bestDF[-1]["timestep"] == 6
    assert hook.bestepisode in (1, 2) and hook.bestreward == max(hook.rewards_compare)
    # currentNNA == the device actor after the last episode; save -> perturb -> load restores the device networks
    dev = pol.behavior_actor.sync_from_device().flat()
    assert np.array_equal(hook.currentNNA.flat(), dev)
    path = tmp_path / "agent.npz"
    pkg.checkpoint.save_npz(path, pol, hook)
    before = {k: getattr(pol, k).sync_from_device().flat().copy() for k in ("behavior_actor", "behavior_critic", "target_critic")}
    for k in before:
        app = getattr(pol, k)
        app.model.load_flat(np.zeros_like(before[k]))
        app.upload()
    pkg.checkpoint.load_npz(path, pol)
    for k, v in before.items():
        assert np.array_equal(getattr(pol, k).sync_from_device().flat(), v)
    env.close()


def _weights(pol):
    return np.concatenate([getattr(pol, k).sync_from_device().flat() for k in ("behavior_critic", "behavior_actor", "target_critic", "target_actor")])


@pytest.mark.gpu
def test_resume_from_checkpoint_is_bit_identical_to_the_uninterrupted_run(pkg, tmp_path):
    """load(); train() (KSSetup.jl:392-402, 304-319): run k loop steps, save, load into a FRESH context, continue -- weights,
    ADAM moments, beta powers and the replay rings must equal the run that never stopped, bit for bit."""
    A = pkg.agent
    setup = pkg.setups.KSSetup.ks22()
    B = 3
    y0 = setup.generate_random_init(np.random.default_rng(0), B)

    def fresh():
        env = setup.make_env(n_envs=B, dtype="f64", y0=y0)
        pol = A.create_agent(env, rng=np.random.default_rng(1), nna_scale=0.6, nna_scale_critic=7.0, drop_middle_layer=True,
                             batch_size=32, start_steps=1, update_after=1, update_loops=3, trajectory_length=200, seed=5)
        env.reset()
        pol.trajectory.pre_episode()
        return env, pol

    def steps(env, pol, n):
        for _ in range(n):
            pol(env, learning=True)
            pol.trajectory.pre_act()
            pol.maybe_update()
            env.step_device()
            pol.trajectory.post_act()

    env_a, pol_a = fresh()
    steps(env_a, pol_a, 14)                                   # 14 x 24 columns: the 200-column ring has wrapped
    path = tmp_path / "resume.npz"
    pkg.checkpoint.save_npz(path, pol_a)
    y_saved, st_saved, act_saved = env_a.get(pkg.lib.ARR_Y), env_a.get(pkg.lib.ARR_STATE), env_a.get(pkg.lib.ARR_ACTION)
    time_saved, steps_saved = env_a.get(pkg.lib.ARR_TIME), env_a.get(pkg.lib.ARR_STEPS)
    steps(env_a, pol_a, 9)
    env_b, pol_b = fresh()
    pkg.checkpoint.load_npz(path, pol_b)
    # the reference does not save the env either (KSSetup.jl:382 is commented out); restore it by hand for the comparison
    for which, v in ((pkg.lib.ARR_Y, y_saved), (pkg.lib.ARR_STATE, st_saved), (pkg.lib.ARR_ACTION, act_saved),
                     (pkg.lib.ARR_TIME, time_saved), (pkg.lib.ARR_STEPS, steps_saved)):
        env_b.put(which, v)
    steps(env_b, pol_b, 9)
    assert np.array_equal(_weights(pol_a), _weights(pol_b))
    for k in ("behavior_actor", "behavior_critic"):
        for x, y in zip(getattr(pol_a, k).opt_state(), getattr(pol_b, k).opt_state()):
            assert np.array_equal(x, y), k
    assert pol_a.trajectory.positions() == pol_b.trajectory.positions()
    for x, y in zip(pol_a.trajectory.get(), pol_b.trajectory.get()):
        assert np.array_equal(x, y)
    assert pol_a.sampler_offset() == pol_b.sampler_offset() and pol_a.n_updates == pol_b.n_updates
    env_a.close(); env_b.close()


def test_load_npz_rejects_a_checkpoint_of_another_shape(pkg, tmp_path):
    """shape / activation validation happens before anything is uploaded (no GPU needed: it fails on the host side)"""
    z = {"behavior_actor_sizes": np.asarray([3, 6, 1])}
    np.savez(tmp_path / "bad.npz", **z)

    class FakeApp:
        class model:
            sizes = [1, 6, 1]
            layers = []

    class FakePol:
        behavior_actor = FakeApp()
    with pytest.raises(ValueError):
        pkg.checkpoint.load_npz(tmp_path / "bad.npz", FakePol())


@pytest.mark.skipif(not REF.exists(), reason="the reference tree only exists in the build container")
def test_agent_jld2_importer_reads_everything_the_reference_saved(pkg, golden):
    d = pkg.checkpoint.load_agent_jld2(REF / "scripts/KS/KS200/saves/agent.jld2")
    g = golden("ks200_agent")
    assert {k: v.sizes for k, v in d["nets"].items()} == {"behavior_actor": [1, 6, 1], "behavior_critic": [2, 140, 1],
                                                          "target_actor": [1, 6, 1], "target_critic": [2, 140, 1]}
    rp = d["replay"]
    assert (rp["first_sa"], rp["first_rt"], rp["state"].shape, rp["reward"].shape) == (72317, 72240, (1, 150001), (150000,))
    assert np.array_equal(d["opt"]["behavior_critic"][0], g["opt_m_behavior_critic"])
    assert np.array_equal(rp["reward"][:4096], g["w0_reward"])
    # the hook's best actor is one of the behavior actor's past states: same shapes, Float32
    assert d["nets"]["behavior_actor"].flat().dtype == np.float32


@pytest.mark.gpu
def test_reference_agent_state_loads_into_the_device_and_trains_on(pkg, golden):
    """The shipped KS200 agent (networks + ADAM state from tests/golden/ks200_agent.npz) as the starting point of an update:
    the first ADAM step from the saved (m, v, beta^t) must match the oracle's Adam fed with the same state."""
    from oracle import agent_oracle as AO
    from test_agent_oracle import batch
    g = golden("ks200_agent")
    A = pkg.agent
    setup = pkg.setups.KSSetup.ks200()
    env = setup.make_env(n_envs=1, dtype="f64", y0=setup.y0_standard())

    def chain(name, out_act):
        sizes = [int(x) for x in g["sizes_" + name]]
        c = A.Chain(A.Dense(np.zeros((sizes[1], sizes[0]), np.float32), np.zeros(sizes[1], np.float32), "relu"),
                    A.Dense(np.zeros((sizes[2], sizes[1]), np.float32), np.zeros(sizes[2], np.float32), out_act))
        c.load_flat(g["net_" + name])
        return c
    pol = A.CustomDDPGPolicy(env, behavior_actor=chain("behavior_actor", "tanh"), behavior_critic=chain("behavior_critic", None),
                             target_actor=chain("target_actor", "tanh"), target_critic=chain("target_critic", None), trajectory_length=4096)
    for name in ("behavior_actor", "behavior_critic"):
        getattr(pol, name).set_opt_state(g["opt_m_" + name], g["opt_v_" + name], g["opt_betap_" + name])
        m, v, bp = getattr(pol, name).opt_state()
        assert np.array_equal(m, g["opt_m_" + name]) and np.array_equal(bp, g["opt_betap_" + name])

    def net(c):
        return AO.Net([(l.W, l.b, l.act) for l in c.layers])
    ref = AO.DDPG(net(pol.behavior_actor.model), net(pol.behavior_critic.model))
    ref.At, ref.Ct = net(pol.target_actor.model), net(pol.target_critic.model)
    for opt, name, model in ((ref.opt_a, "behavior_actor", ref.A), (ref.opt_c, "behavior_critic", ref.C)):
        o = 0
        for i, p in enumerate(model.params()):
            n = p.size
            opt.m[i] = g["opt_m_" + name][o:o + n].reshape(p.shape, order="F").copy()
            opt.v[i] = g["opt_v_" + name][o:o + n].reshape(p.shape, order="F").copy()
            opt.bp[i] = [float(g["opt_betap_" + name][0]), float(g["opt_betap_" + name][1])]
            o += n
    rng = np.random.default_rng(2)
    for it in range(2):
        s, a, r, t, s2 = batch(rng, 1, 1, 96)
        pol.set_batch(s, a, r, t, s2)
        pol.update()
        ref.update(s, a, r, t, s2, True)
        for dev, orc in ((pol.behavior_critic, ref.C), (pol.behavior_actor, ref.A), (pol.target_critic, ref.Ct), (pol.target_actor, ref.At)):
            w = dev.sync_from_device().flat()
            assert np.max(np.abs(w - orc.flat())) <= 2e-5 * np.max(np.abs(orc.flat())), it
    env.close()
