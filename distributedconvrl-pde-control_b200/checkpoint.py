"""Checkpoint compatibility with the reference.

The reference saves / loads its agent with JLD2 (`save()` / `load()`, scripts/KS/setup/KSSetup.jl:378-402):
`saves/hook.jld2` holds `hook.bestNNA` / `hook.currentNNA` (the actor as a Flux Chain of Dense), `saves/agent.jld2`
the four networks, ADAM state and the replay buffer.  This module

  * reads the numeric arrays of such files (JLD2 is an HDF5 subset; numpy only, no h5py) so that the actors the
    reference ships can be loaded straight into the device networks: `load_hook_actors`, `chain_from_arrays`;
  * moves parameters between host models and the device (`sync_to_host` before a save, `upload` after a load) --
    the Flux-shaped host `Chain` stays the source of truth exactly as in the reference.

Writing JLD2 is out of scope: `save_npz` / `load_npz` store the same content as .npz.
"""
import re
import struct
from pathlib import Path

import numpy as np

from .agent import Chain, Dense

def _parse_ohdr(b, off):
    """Return (dims, dtype, data bytes) for a numeric dataset header, else None."""
    if b[off:off + 4] != b"OHDR" or b[off + 4] != 2:
        return None
    flags = b[off + 5]
    pos = off + 6
    if flags & 0x20:
        pos += 16                      # access/mod/change/birth times
    if flags & 0x10:
        pos += 4                       # max compact / min dense attributes
    szw = 1 << (flags & 3)
    chunk0 = int.from_bytes(b[pos:pos + szw], "little")
    pos += szw
    end = pos + chunk0
    dims = dtype = data = None
    while pos + 4 <= end:
        mtype = b[pos]
        msize = struct.unpack_from("<H", b, pos + 1)[0]
        pos += 4
        if flags & 0x04:
            pos += 2                   # creation order
        body = b[pos:pos + msize]
        if mtype == 0x01 and len(body) >= 4 and body[0] == 2:
            rank = body[1]
            dims = [struct.unpack_from("<Q", body, 4 + 8 * i)[0] for i in range(rank)]
        elif mtype == 0x03 and len(body) >= 8:
            cls = body[0] & 0x0F
            size = struct.unpack_from("<I", body, 4)[0]
            if cls == 1 and size in (4, 8):
                dtype = np.dtype("<f%d" % size)
            elif cls == 0 and size in (1, 2, 4, 8):
                signed = (body[1] >> 3) & 1
                dtype = np.dtype("<%s%d" % ("i" if signed else "u", size))
        elif mtype == 0x08 and len(body) >= 2 and body[0] == 4:
            lclass = body[1]
            if lclass == 0:
                n = struct.unpack_from("<H", body, 2)[0]
                data = body[4:4 + n]
            elif lclass == 1:
                addr, n = struct.unpack_from("<QQ", body, 2)
                if addr != 0xFFFFFFFFFFFFFFFF:
                    data = b[addr:addr + n]
        pos += msize
    if dims is None or dtype is None or data is None:
        return None
    count = int(np.prod(dims)) if dims else 1
    if count * dtype.itemsize != len(data):
        return None
    arr = np.frombuffer(data, dtype=dtype).reshape(dims if dims else ())
    return arr


def numeric_arrays(path, float_only=True):
    """All numeric dataset arrays of a JLD2 file, in file order.

    HDF5 dims are C-order over the same bytes Julia wrote column-major, so the
    returned array is the *transpose* of the Julia array: a Julia (h, ns) weight
    matrix comes back as (ns, h).  We transpose back to Julia's shape.
    """
    b = Path(path).read_bytes()
    out = []
    for m in re.finditer(b"OHDR", b):
        arr = _parse_ohdr(b, m.start())
        if arr is None:
            continue
        if float_only and arr.dtype.kind != "f":
            continue
        out.append((m.start(), np.ascontiguousarray(arr.T)))
    return out




def load_hook_actors(path, hidden_act="relu", out_act="tanh"):
    """(bestNNA, currentNNA) of a reference `hook.jld2` as Chains (PDEhook fields, src/PDEhook.jl:21-26; network
    shape from create_NNA, src/PDEagent.jl:18-30).  The Float32 arrays appear in file order: bestNNA's (W, b) per
    layer first, currentNNA's last."""
    f32 = [a for _, a in numeric_arrays(path) if a.dtype == np.float32]
    if len(f32) < 4 or len(f32) % 2:
        raise ValueError("%s: expected an even number (>= 4) of Float32 parameter arrays, found %d" % (path, len(f32)))
    n = len(f32) // 2
    return chain_from_arrays(f32[:n], hidden_act, out_act), chain_from_arrays(f32[n:], hidden_act, out_act)


def chain_from_arrays(arrays, hidden_act="relu", out_act="tanh"):
    """[W1, b1, W2, b2, ...] (Flux shapes: W (out, in)) -> Chain with the reference's activations."""
    layers = []
    n_layers = len(arrays) // 2
    for l in range(n_layers):
        W, b = np.asarray(arrays[2 * l], dtype=np.float32), np.asarray(arrays[2 * l + 1], dtype=np.float32).reshape(-1)
        if W.ndim != 2 or W.shape[0] != b.shape[0]:
            raise ValueError("layer %d: W %s does not match b %s" % (l, W.shape, b.shape))
        layers.append(Dense(W, b, out_act if l == n_layers - 1 else hidden_act))
    return Chain(*layers)


def sync_to_host(policy):
    """Pull the four device networks into their host Chains (call before saving), returns them by name."""
    return {name: getattr(policy, name).sync_from_device()
            for name in ("behavior_actor", "behavior_critic", "target_actor", "target_critic")}


def save_npz(path, policy, hook=None):
    """The content of the reference's agent.jld2 / hook.jld2 that this path owns: network parameters (+ hook actors)."""
    out = {}
    for name, chain in sync_to_host(policy).items():
        for l, layer in enumerate(chain.layers):
            out["%s_W%d" % (name, l + 1)] = layer.W
            out["%s_b%d" % (name, l + 1)] = layer.b
            out["%s_act%d" % (name, l + 1)] = np.array(layer.act or "identity")
    if hook is not None:
        for tag, chain in (("best", hook.bestNNA), ("current", hook.currentNNA)):
            if chain is not None:
                for l, layer in enumerate(chain.layers):
                    out["hook_%s_W%d" % (tag, l + 1)] = layer.W
                    out["hook_%s_b%d" % (tag, l + 1)] = layer.b
        out["hook_rewards"] = np.asarray(hook.rewards, dtype=np.float64)
    np.savez_compressed(path, **out)


def load_npz(path, policy):
    """Inverse of save_npz for the four networks: host Chains are overwritten and uploaded to the device."""
    z = np.load(path, allow_pickle=False)
    for name in ("behavior_actor", "behavior_critic", "target_actor", "target_critic"):
        app = getattr(policy, name)
        for l, layer in enumerate(app.model.layers):
            layer.W = np.ascontiguousarray(z["%s_W%d" % (name, l + 1)], dtype=np.float32)
            layer.b = np.ascontiguousarray(z["%s_b%d" % (name, l + 1)], dtype=np.float32)
        app.upload()
