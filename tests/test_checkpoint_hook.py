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
