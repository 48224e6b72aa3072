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

def _base_address(b):
    """HDF5 superblock (v2/v3) base address: JLD2 puts a 512-byte text header in front of the superblock and every
    file address (contiguous dataset storage) is relative to it."""
    i = b.find(b"\x89HDF\r\n\x1a\n")
    if i < 0 or b[i + 8] < 2:
        return 0
    return struct.unpack_from("<Q", b, i + 12)[0]


def _parse_ohdr(b, off, base=0):
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
            elif cls == 4 and size == 1:
                dtype = np.dtype("u1")             # bitfield: Julia Bool arrays (the replay `terminal` ring)
        elif mtype == 0x08 and len(body) >= 2 and body[0] == 4:
            lclass = body[1]
            if lclass == 0:
                n = struct.unpack_from("<H", body, 2)[0]
                data = body[4:4 + n]
            elif lclass == 1:
                addr, n = struct.unpack_from("<QQ", body, 2)
                if addr != 0xFFFFFFFFFFFFFFFF:
                    data = b[base + addr:base + addr + n]
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
    base = _base_address(b)
    out = []
    for m in re.finditer(b"OHDR", b):
        arr = _parse_ohdr(b, m.start(), base)
        if arr is None:
            continue
        if float_only and arr.dtype.kind != "f":
            continue
        out.append((m.start(), np.ascontiguousarray(arr.T)))
    return out




def circular_buffers(path):
    """The CircularArrayBuffers of a reference `agent.jld2` (RLCore 0.8.13 `CircularArraySARTTrajectory`): for every
    large dataset the struct that references it stores `(buffer, first, nframes, step_size)` -- an 8-byte file-relative
    object reference followed by three Int64.  Returns [(raw array in Julia shape, first (1-based), nframes)] in file
    order: state, action, reward, terminal."""
    b = Path(path).read_bytes()
    base = _base_address(b)
    out = []
    for m in re.finditer(b"OHDR", b):
        arr = _parse_ohdr(b, m.start(), base)
        if arr is None or arr.size <= 4096:
            continue
        ref = struct.pack("<Q", m.start() - base)
        i = b.find(ref)
        first = nframes = None
        while i >= 0:
            f, n, step = struct.unpack_from("<qqq", b, i + 8)
            if step == 1 and 1 <= f <= arr.shape[0] and 0 <= n <= arr.shape[0]:
                first, nframes = f, n
                break
            i = b.find(ref, i + 1)
        if first is None:
            raise ValueError("%s: no CircularArrayBuffer header found for the dataset at %d" % (path, m.start()))
        out.append((np.ascontiguousarray(arr.T), int(first), int(nframes)))
    return out


def load_agent_jld2(path, hidden_act="relu", actor_out_act="tanh"):
    """Everything a reference `agent.jld2` (scripts/KS/setup/KSSetup.jl:378-402 `save()`) holds that this path owns:

      nets      {behavior_actor, behavior_critic, target_actor, target_critic: Chain}
      opt       {behavior_actor, behavior_critic: (m, v, beta_p)} flat in the parameter layout (Flux ADAM's IdDict is
                stored in hash order: the (m, v, beta_p) triples are matched to the parameter arrays by shape)
      replay    state (ns, n_sa), action (na, n_sa), reward (n_rt,), terminal (n_rt,) in LOGICAL order (oldest first) and
                `first_sa`, `first_rt`: the 0-based raw ring positions of the oldest column, so that a restored ring
                continues writing exactly where the reference's would
    Float arrays appear in file order: behavior actor (W, b per layer), its ADAM triples, behavior critic, its triples,
    target actor, target critic, scalars, then the four rings."""
    arrs = [a for _, a in numeric_arrays(path, float_only=False)]
    small = [a for a in arrs if a.size <= 4096 and a.dtype.kind == "f"]

    def take_net(i):
        """consecutive (W, b) float32 pairs starting at small[i] while shapes chain up"""
        layers = []
        while i + 1 < len(small) and small[i].dtype == np.float32 and small[i].ndim == 2 and small[i + 1].ndim == 1 \
                and small[i].shape[0] == small[i + 1].shape[0] and (not layers or layers[-1][0].shape[0] == small[i].shape[1]):
            layers.append((small[i], small[i + 1]))
            i += 2
        return layers, i

    def take_opt(i, layers):
        """2 * n_layers (m, v, beta_p) triples in arbitrary order -> flat m, v in the parameter layout"""
        want = {}
        for l, (W, b) in enumerate(layers):
            want.setdefault(W.shape, []).append(("W", l))
            want.setdefault(b.shape, []).append(("b", l))
        got, bps = {}, []
        for _ in range(2 * len(layers)):
            m, v, bp = small[i], small[i + 1], small[i + 2]
            if m.shape != v.shape or bp.shape != (2,) or bp.dtype != np.float64 or not want.get(m.shape):
                raise ValueError("%s: unexpected ADAM state layout near float array %d" % (path, i))
            got[want[m.shape].pop(0)] = (m, v)
            bps.append(bp)
            i += 3
        flat_m = np.concatenate([np.concatenate([got[("W", l)][0].flatten(order="F"), got[("b", l)][0]]) for l in range(len(layers))])
        flat_v = np.concatenate([np.concatenate([got[("W", l)][1].flatten(order="F"), got[("b", l)][1]]) for l in range(len(layers))])
        if any(not np.array_equal(bp, bps[0]) for bp in bps):
            raise ValueError("%s: ADAM beta powers differ between the arrays of one network" % path)
        return (flat_m.astype(np.float32), flat_v.astype(np.float32), bps[0].astype(np.float64)), i

    i = 0
    la, i = take_net(i)
    oa, i = take_opt(i, la)
    lc, i = take_net(i)
    oc, i = take_opt(i, lc)
    lta, i = take_net(i)
    ltc, i = take_net(i)
    mk = lambda layers, out_act: chain_from_arrays([x for pair in layers for x in pair], hidden_act, out_act)
    nets = {"behavior_actor": mk(la, actor_out_act), "behavior_critic": mk(lc, None),
            "target_actor": mk(lta, actor_out_act), "target_critic": mk(ltc, None)}
    rings = circular_buffers(path)
    if len(rings) != 4:
        raise ValueError("%s: expected 4 replay rings, found %d" % (path, len(rings)))

    def logical(raw, first, n):
        raw = raw.reshape(raw.shape[0], -1) if raw.ndim > 1 else raw.reshape(1, -1)
        idx = (first - 1 + np.arange(n)) % raw.shape[1]
        return raw[:, idx]

    (s_raw, fs, n_sa), (a_raw, fa, n_a), (r_raw, fr, n_rt), (t_raw, ft, n_t) = rings
    if (fs, n_sa) != (fa, n_a) or (fr, n_rt) != (ft, n_t):
        raise ValueError("%s: state/action or reward/terminal rings are not aligned" % path)
    replay = {"state": logical(s_raw, fs, n_sa), "action": logical(a_raw, fa, n_a), "reward": logical(r_raw, fr, n_rt)[0],
              "terminal": logical(t_raw, ft, n_t)[0].astype(bool), "first_sa": fs - 1, "first_rt": fr - 1,
              "capacity": r_raw.size}
    return {"nets": nets, "opt": {"behavior_actor": oa, "behavior_critic": oc}, "replay": replay}


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


NETS = ("behavior_actor", "behavior_critic", "target_actor", "target_critic")


def save_npz(path, policy, hook=None):
    """The content of the reference's agent.jld2 / hook.jld2 that this path owns (save(), KSSetup.jl:378-402), complete
    enough to resume: the four networks, the ADAM state (m, v, beta powers) of every network, the replay rings in
    logical order with their raw ring positions, the policy's counters (update_step, noise / sampler Philox counters),
    and the hook's actors and reward history."""
    out = {}
    for name, chain in sync_to_host(policy).items():
        out["%s_sizes" % name] = np.asarray(chain.sizes, dtype=np.int64)
        for l, layer in enumerate(chain.layers):
            out["%s_W%d" % (name, l + 1)] = layer.W
            out["%s_b%d" % (name, l + 1)] = layer.b
            out["%s_act%d" % (name, l + 1)] = np.array(layer.act or "identity")
        m, v, bp = getattr(policy, name).opt_state()
        out["%s_adam_m" % name], out["%s_adam_v" % name], out["%s_adam_betap" % name] = m, v, bp
    tr = policy.trajectory
    cap, n_sa, n_rt, first_sa, first_rt = tr.positions()
    st, ac, rw, tm = tr.get()
    out.update(replay_state=st, replay_action=ac, replay_reward=rw, replay_terminal=tm,
               replay_meta=np.asarray([cap, n_sa, n_rt, first_sa, first_rt], dtype=np.int64),
               policy_counters=np.asarray([policy.update_step, policy._rng_offset, policy.sampler_offset(), policy.n_updates],
                                          dtype=np.uint64))
    if hook is not None:
        for tag, chain in (("best", hook.bestNNA), ("current", hook.currentNNA)):
            if chain is not None:
                for l, layer in enumerate(chain.layers):
                    out["hook_%s_W%d" % (tag, l + 1)] = layer.W
                    out["hook_%s_b%d" % (tag, l + 1)] = layer.b
        out["hook_rewards"] = np.asarray(hook.rewards, dtype=np.float64)
    np.savez_compressed(path, **out)


def load_npz(path, policy):
    """Inverse of save_npz (load(), KSSetup.jl:392-402): networks (shapes and activations are validated against the
    policy's), ADAM state, replay rings at their saved raw positions and the counters -- `load(); train()` then continues
    bit-identically to an uninterrupted run (tests/test_checkpoint_hook.py)."""
    z = np.load(path, allow_pickle=False)
    for name in NETS:
        app = getattr(policy, name)
        if list(z["%s_sizes" % name]) != list(app.model.sizes):
            raise ValueError("%s: saved layer sizes %s do not match the policy's %s" % (name, list(z["%s_sizes" % name]), app.model.sizes))
        for l, layer in enumerate(app.model.layers):
            act = str(z["%s_act%d" % (name, l + 1)])
            if act != (layer.act or "identity"):
                raise ValueError("%s layer %d: saved activation %r != %r" % (name, l + 1, act, layer.act or "identity"))
            W, b = z["%s_W%d" % (name, l + 1)], z["%s_b%d" % (name, l + 1)]
            if W.shape != layer.W.shape or b.shape != layer.b.shape:
                raise ValueError("%s layer %d: saved parameter shapes do not match" % (name, l + 1))
            layer.W = np.ascontiguousarray(W, dtype=np.float32)
            layer.b = np.ascontiguousarray(b, dtype=np.float32)
        app.upload()
        if "%s_adam_m" % name in z.files:
            app.set_opt_state(z["%s_adam_m" % name], z["%s_adam_v" % name], z["%s_adam_betap" % name])
    if "replay_meta" in z.files:
        cap, n_sa, n_rt, first_sa, first_rt = (int(x) for x in z["replay_meta"])
        if cap != policy.trajectory.capacity:
            raise ValueError("saved replay capacity %d != the policy's %d" % (cap, policy.trajectory.capacity))
        policy.trajectory.set(z["replay_state"], z["replay_action"], z["replay_reward"], z["replay_terminal"],
                              first_sa=first_sa, first_rt=first_rt)
        us, noise_off, samp_off, n_upd = (int(x) for x in z["policy_counters"])
        policy.update_step, policy._rng_offset, policy.n_updates = us, noise_off, n_upd
        policy.set_sampler_offset(samp_off)


def load_reference_agent(path, policy):
    """Import a reference `agent.jld2` into a policy whose networks have the saved shapes: weights of all four networks,
    the behavior networks' ADAM state, and the replay rings at the reference's ring positions (the policy's trajectory
    must have the file's capacity).  The reference keeps no optimiser for the targets (Polyak only)."""
    d = load_agent_jld2(path)
    for name in NETS:
        app = getattr(policy, name)
        src = d["nets"][name]
        if src.sizes != app.model.sizes:
            raise ValueError("%s: file has %s, policy has %s" % (name, src.sizes, app.model.sizes))
        app.model.load_flat(src.flat())
        app.upload()
        if name in d["opt"]:
            app.set_opt_state(*d["opt"][name])
    rp = d["replay"]
    if rp["capacity"] != policy.trajectory.capacity:
        raise ValueError("file replay capacity %d != the policy's %d" % (rp["capacity"], policy.trajectory.capacity))
    policy.trajectory.set(rp["state"], rp["action"], rp["reward"], rp["terminal"], first_sa=rp["first_sa"], first_rt=rp["first_rt"])
    return d
