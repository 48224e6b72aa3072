"""Host-side mirror of src/custom_nna.jl and src/PDEagent.jl over the C ABI.

`Chain`/`Dense` hold Flux-shaped float32 parameters (W is (out, in)); the
device copies inside libpdeb200 are the ones the kernels use, and `sync_from_device`
pulls them back so that save()/load()-style host code keeps working
(KSSetup.jl:378-402).
"""
import ctypes as C

import numpy as np

from . import _lib as L

_ACT = {"identity": L.ACT_IDENTITY, None: L.ACT_IDENTITY, "relu": L.ACT_RELU, "tanh": L.ACT_TANH}


class Dense:
    """Flux.Dense(in, out, act): y = act.(W*x .+ b)"""

    def __init__(self, W, b, act=None):
        self.W = np.ascontiguousarray(W, dtype=np.float32)
        self.b = np.ascontiguousarray(b, dtype=np.float32).reshape(-1)
        assert self.W.shape[0] == self.b.shape[0]
        self.act = act


class Chain:
    def __init__(self, *layers):
        self.layers = list(layers)

    @property
    def sizes(self):
        return [self.layers[0].W.shape[1]] + [l.W.shape[0] for l in self.layers]

    def flat(self):
        return np.concatenate([np.concatenate([l.W.flatten(order="F"), l.b]) for l in self.layers]).astype(np.float32)

    def load_flat(self, flat):
        o = 0
        for l in self.layers:
            n = l.W.size
            l.W = np.asarray(flat[o:o + n], dtype=np.float32).reshape(l.W.shape, order="F").copy()
            o += n
            l.b = np.asarray(flat[o:o + l.b.size], dtype=np.float32).copy()
            o += l.b.size

    def copy(self):
        return Chain(*[Dense(l.W.copy(), l.b.copy(), l.act) for l in self.layers])


def glorot_uniform(rng, out, inp):
    """Flux.glorot_uniform(rng)(out, in): (rand(Float32) - 0.5) * sqrt(24 / (in + out))."""
    return ((rng.random((out, inp), dtype=np.float32) - np.float32(0.5)) *
            np.float32(np.sqrt(24.0 / (out + inp)))).astype(np.float32)


def create_chain(*, na, ns, is_actor, rng, nna_scale, drop_middle_layer, fun="relu"):
    """create_NNA's network factory, src/PDEagent.jl:14-44."""
    h = int(np.floor((10 if is_actor else 20) * nna_scale))
    if is_actor:
        dims = [(ns, h, fun)] + ([] if drop_middle_layer else [(h, h, fun)]) + [(h, na, "tanh")]
    else:
        dims = [(ns + na, h, fun)] + ([] if drop_middle_layer else [(h, h, fun)]) + [(h, 1, None)]
    return Chain(*[Dense(glorot_uniform(rng, o, i), np.zeros(o, np.float32), a) for i, o, a in dims])


class CustomNeuralNetworkApproximator:
    """src/custom_nna.jl:7-27 -- (model, optimizer) pair bound to one of the four device networks."""

    def __init__(self, env, net_id, model, learning_rate=0.001):
        self.env, self.net_id, self.model, self.learning_rate = env, net_id, model, learning_rate
        self.upload()

    def upload(self):
        m = self.model
        sizes = np.asarray(m.sizes, dtype=np.int32)
        acts = np.asarray([_ACT[l.act] for l in m.layers], dtype=np.int32)
        flat = m.flat()
        L.check(self.env._lib.pdeb200_net_set(self.env._ctx, self.net_id, len(m.layers), sizes.ctypes.data,
                                              acts.ctypes.data, flat.ctypes.data), self.env._ctx)

    def sync_from_device(self):
        n = self.env._lib.pdeb200_net_num_params(self.env._ctx, self.net_id)
        flat = np.empty(n, dtype=np.float32)
        L.check(self.env._lib.pdeb200_net_get(self.env._ctx, self.net_id, flat.ctypes.data, n), self.env._ctx)
        self.model.load_flat(flat)
        return self.model

    def copyto(self, src):
        """Base.copyto!(dest, src), custom_nna.jl:26-27"""
        self.model.load_flat(src.sync_from_device().flat())
        self.upload()
