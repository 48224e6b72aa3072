"""The approximator call (custom_nna.jl:13) for wide networks: tcgen05 (3xTF32) Dense layers vs a float64 numpy
restatement of Flux Dense, and vs the CUDA-core path.  fp32 tolerance 1e-5 (north star)."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


def _ref_forward(layers, x):
    h = x.astype(np.float64)
    for W, b, act in layers:
        h = W.astype(np.float64) @ h + b.astype(np.float64)[:, None]
        h = np.maximum(h, 0) if act == "relu" else (np.tanh(h) if act == "tanh" else h)
    return h


def _net(rng, sizes, acts):
    g = lambda o, i: ((rng.random((o, i)) - 0.5) * np.sqrt(24.0 / (o + i))).astype(np.float32)
    return [(g(o, i), (0.1 * rng.standard_normal(o)).astype(np.float32), a) for i, o, a in zip(sizes[:-1], sizes[1:], acts)]


@pytest.mark.parametrize("sizes,acts,n_cols", [
    ([13, 340, 340, 1], ["relu", "relu", None], 1000),        # NS / KSeg critic with the middle layer (nna_scale_critic = 17)
    ([3, 140, 140, 1], ["relu", "relu", None], 4096),         # KS critic with the middle layer
    ([36, 128, 128, 1], ["relu", "relu", "tanh"], 130),       # wide actor, tile-edge sizes
    ([12, 200, 72, 40, 4], ["relu", "tanh", "relu", None], 257),
])
def test_wide_network_forward_tensor_cores(pkg, sizes, acts, n_cols):
    rng = np.random.default_rng(7)
    layers = _net(rng, sizes, acts)
    A = pkg.agent
    setup = pkg.setups.KSSetup.ks22()
    env = setup.make_env(n_envs=1, dtype="f64", y0=setup.y0_standard())
    app = A.CustomNeuralNetworkApproximator(env, pkg.lib.NET_BEHAVIOR_CRITIC, A.Chain(*[A.Dense(W, b, a) for W, b, a in layers]))
    x = rng.standard_normal((sizes[0], n_cols)).astype(np.float32)
    want = _ref_forward(layers, x)
    y_tc, used = app(x, path=0, return_info=True)
    assert used >= 1, "dense layers must run on the tensor cores"
    y_cc, used_cc = app(x, path=1, return_info=True)
    assert used_cc == 0
    scale = np.max(np.abs(want))
    assert np.max(np.abs(y_cc - want)) / scale < 1e-5
    assert np.max(np.abs(y_tc - want)) / scale < 1e-5, np.max(np.abs(y_tc - want)) / scale
    env.close()


def test_shipped_thin_networks_stay_on_cuda_cores(pkg, golden):
    g = golden("ks200_hook")
    A = pkg.agent
    setup = pkg.setups.KSSetup.ks22()
    env = setup.make_env(n_envs=1, dtype="f64", y0=setup.y0_standard())
    chain = A.Chain(A.Dense(g["best_W1"], g["best_b1"], "relu"), A.Dense(g["best_W2"], g["best_b2"], "tanh"))
    app = A.CustomNeuralNetworkApproximator(env, pkg.lib.NET_BEHAVIOR_ACTOR, chain)
    x = np.random.default_rng(0).standard_normal((1, 5000)).astype(np.float32)
    y, used = app(x, return_info=True)
    assert used == 0
    want = _ref_forward([(g["best_W1"], g["best_b1"], "relu"), (g["best_W2"], g["best_b2"], "tanh")], x)
    assert relerr(y, want) < 1e-6
    env.close()
