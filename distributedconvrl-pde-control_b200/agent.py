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

    def __call__(self, x, path=0, return_info=False):
        """(app)(x) = app.model(x), custom_nna.jl:13: x is (rows, n_cols); returns (out, n_cols) float32.
        path: 0 auto, 1 CUDA cores only, 2 tensor cores where possible."""
        x = np.asarray(x, dtype=np.float32)
        xm = np.ascontiguousarray(x.T)
        y = np.empty((xm.shape[0], self.model.sizes[-1]), dtype=np.float32)
        used = C.c_int32(0)
        L.check(self.env._lib.pdeb200_net_forward(self.env._ctx, self.net_id, xm.shape[0], xm.ctypes.data, y.ctypes.data,
                                                  int(path), C.byref(used)), self.env._ctx)
        return (y.T, used.value) if return_info else y.T

    def copyto(self, src):
        """Base.copyto!(dest, src) = Flux.loadparams!(dest.model, params(src)), custom_nna.jl:26-27: the weights are
        overwritten, the optimiser state of `dest` is kept (pdeb200_net_set_params)."""
        flat = src.sync_from_device().flat()
        self.model.load_flat(flat)
        L.check(self.env._lib.pdeb200_net_set_params(self.env._ctx, self.net_id, flat.ctypes.data, flat.size), self.env._ctx)

    # -- optimiser state (Flux ADAM: per-parameter (m, v, beta powers)); save()/load() of KSSetup.jl:378-402 ------
    def opt_state(self):
        n = self.env._lib.pdeb200_net_num_params(self.env._ctx, self.net_id)
        m, v, bp = np.empty(n, np.float32), np.empty(n, np.float32), np.empty(2, np.float64)
        L.check(self.env._lib.pdeb200_opt_get(self.env._ctx, self.net_id, m.ctypes.data, v.ctypes.data, bp.ctypes.data, n), self.env._ctx)
        return m, v, bp

    def set_opt_state(self, m, v, beta_p):
        m = np.ascontiguousarray(m, dtype=np.float32).reshape(-1)
        v = np.ascontiguousarray(v, dtype=np.float32).reshape(-1)
        bp = np.ascontiguousarray(beta_p, dtype=np.float64).reshape(-1)
        n = self.env._lib.pdeb200_net_num_params(self.env._ctx, self.net_id)
        if m.size != n or v.size != n or bp.size != 2:
            raise ValueError("optimizer state does not match the network (%d parameters)" % n)
        L.check(self.env._lib.pdeb200_opt_set(self.env._ctx, self.net_id, m.ctypes.data, v.ctypes.data, bp.ctypes.data, n), self.env._ctx)


class ZeroPolicy:
    """src/PDEagent.jl:420-424"""

    def __call__(self, env):
        return np.zeros((env.a_rows, env.n_envs * env.n_actuators), dtype=env.np_dtype)


class DeviceTrajectory:
    """CircularArraySARTTrajectory kept on the GPU, with the reference's `update!` overload set
    (src/PDEagent.jl:237-314).  One transition per actuator column."""

    def __init__(self, env, capacity):
        self.env = env
        self.capacity = int(capacity)
        L.check(env._lib.pdeb200_traj_create(env._ctx, self.capacity), env._ctx)

    def __len__(self):
        n = C.c_int64()
        L.check(self.env._lib.pdeb200_traj_length(self.env._ctx, C.byref(n)), self.env._ctx)
        return n.value

    def pre_episode(self):           # PreEpisodeStage: pop the dummy tail
        L.check(self.env._lib.pdeb200_traj_pop_tail(self.env._ctx), self.env._ctx)

    def pre_act(self):               # PreActStage: push (s[:, i], a[:, i]) for every column
        L.check(self.env._lib.pdeb200_traj_push_pre(self.env._ctx), self.env._ctx)

    def post_act(self):              # PostActStage: push r[i], terminal for every column
        L.check(self.env._lib.pdeb200_traj_push_post(self.env._ctx), self.env._ctx)

    def post_episode(self):          # PostEpisodeStage: push final state + zero action
        L.check(self.env._lib.pdeb200_traj_episode_end(self.env._ctx), self.env._ctx)

    # -- whole-buffer access in logical order (oldest column first): checkpoints, parity tests -------------------
    def info(self):
        """(capacity, n_sa, n_rt): columns held by the state/action and the reward/terminal rings."""
        return self.positions()[:3]

    def positions(self):
        """(capacity, n_sa, n_rt, first_sa, first_rt): counts and the 0-based raw ring position of the oldest column --
        CircularArrayBuffer's `nframes` and `first - 1` (what agent.jld2 stores)."""
        v = [C.c_int64() for _ in range(5)]
        L.check(self.env._lib.pdeb200_traj_info(self.env._ctx, *[C.byref(x) for x in v]), self.env._ctx)
        return tuple(x.value for x in v)

    def get(self):
        """(state (ns, n_sa), action (na, n_sa), reward (n_rt,), terminal (n_rt,)) in the reference's shapes."""
        _, n_sa, n_rt = self.info()
        ns, na = self.env.ns, (self.env.n_actuators * self.env.a_rows if self.env.cfg.mono else self.env.a_rows)
        s, a = np.empty((n_sa, ns), np.float32), np.empty((n_sa, na), np.float32)
        r, t = np.empty(n_rt, np.float32), np.empty(n_rt, np.uint8)
        L.check(self.env._lib.pdeb200_traj_get(self.env._ctx, s.ctypes.data, a.ctypes.data, r.ctypes.data, t.ctypes.data), self.env._ctx)
        return s.T, a.T, r, t.astype(bool)

    def set(self, state, action, reward, terminal, first_sa=0, first_rt=0):
        s = np.ascontiguousarray(np.asarray(state, dtype=np.float32).T)
        a = np.ascontiguousarray(np.asarray(action, dtype=np.float32).T)
        r = np.ascontiguousarray(reward, dtype=np.float32).reshape(-1)
        t = np.ascontiguousarray(terminal, dtype=np.uint8).reshape(-1)
        if s.shape[0] != a.shape[0] or r.size != t.size:
            raise ValueError("state/action and reward/terminal must have matching column counts")
        L.check(self.env._lib.pdeb200_traj_set(self.env._ctx, s.shape[0], r.size, int(first_sa), int(first_rt), s.ctypes.data,
                                               a.ctypes.data, r.ctypes.data, t.ctypes.data), self.env._ctx)


class CustomDDPGPolicy:
    """src/PDEagent.jl:121-209 + the update trigger :342-361 + update! :363-418, over the C ABI.

    Field names follow the Julia struct (y = discount, p = Polyak factor).  `comm` is an optional
    `parallel.Comm`: with more than one rank the context joins the process group (`pdeb200_comm_init`) and
    every sample / update call becomes a collective whose gradient exchange runs inside the library's kernels
    over NVLink peer memory; None = single GPU.
    """

    def __init__(self, env, *, behavior_actor, behavior_critic, target_actor=None, target_critic=None, y=0.99, p=0.995,
                 batch_size=3, start_steps=6, start_policy=None, update_after=10, update_freq=1, update_loops=20,
                 act_limit=1.0, act_noise=1.2, memory_size=0, learning_rate=5e-4, learning_rate_critic=1e-3,
                 trajectory_length=150_000, literal_q1=True, seed=0, comm=None):
        self.env = env
        lib = L
        self.behavior_actor = CustomNeuralNetworkApproximator(env, lib.NET_BEHAVIOR_ACTOR, behavior_actor, learning_rate)
        self.behavior_critic = CustomNeuralNetworkApproximator(env, lib.NET_BEHAVIOR_CRITIC, behavior_critic, learning_rate_critic)
        self.target_actor = CustomNeuralNetworkApproximator(
            env, lib.NET_TARGET_ACTOR, target_actor if target_actor is not None else behavior_actor.copy(), learning_rate)
        self.target_critic = CustomNeuralNetworkApproximator(
            env, lib.NET_TARGET_CRITIC, target_critic if target_critic is not None else behavior_critic.copy(), learning_rate_critic)
        self.y, self.p, self.batch_size = y, p, batch_size
        self.start_steps, self.start_policy = start_steps, start_policy or ZeroPolicy()
        self.update_after, self.update_freq, self.update_loops = update_after, update_freq, update_loops
        self.act_limit, self.act_noise, self.memory_size = act_limit, act_noise, memory_size
        self.literal_q1 = literal_q1
        self.update_step = 0
        self.seed, self._rng_offset = int(seed), 0
        self.comm = comm
        self.transport = comm.attach(env) if comm is not None and comm.world_size > 1 else L.COMM_NONE
        self.trajectory = DeviceTrajectory(env, trajectory_length)
        self.number_actuators = env.n_envs * env.n_cols       # columns per env step (PDEagent.jl:348-353)
        self.n_updates = 0

    # -- policy forward (PDEagent.jl:175-209); the action is staged on the device ---------------
    def __call__(self, env=None, learning=True, noise=None):
        env = env or self.env
        if learning:
            self.update_step += 1
        if self.update_step <= self.start_steps:
            a = np.ascontiguousarray(np.asarray(self.start_policy(env), dtype=env.np_dtype).T)
            env.put(L.ARR_ACTION_IN, a)
        elif not learning:
            env.policy_act(None, 0.0, self.act_limit)
        elif noise is not None:
            env.policy_act(noise, self.act_noise, self.act_limit)
        else:
            n = env.n_envs * env.n_actuators * env.a_rows
            env.policy_act_rng(self.seed, self._rng_offset, self.act_noise, self.act_limit)
            self._rng_offset += n

    @property
    def losses(self):
        out = np.zeros(2, dtype=np.float32)
        L.check(self.env._lib.pdeb200_get(self.env._ctx, L.ARR_LOSSES, out.ctypes.data, 8), self.env._ctx)
        return {"critic_loss": float(out[0]), "actor_loss": float(out[1])}

    # -- update trigger (PDEagent.jl:342-361) -------------------------------------------------------
    def maybe_update(self):
        if not (len(self.trajectory) > self.update_after * self.number_actuators):
            return 0
        if self.update_step % self.update_freq != 0:
            return 0
        # update_loops x {pde_sample; update!}: one C call, one CUDA graph on the device
        lib, ctx = self.env._lib, self.env._ctx
        L.check(lib.pdeb200_train_updates(ctx, int(self.update_loops), int(self.batch_size), float(self.y), float(self.p),
                                          float(self.behavior_actor.learning_rate), float(self.behavior_critic.learning_rate),
                                          int(self.literal_q1), self.seed ^ 0x5DEECE66D), ctx)
        self.n_updates += self.update_loops
        return self.update_loops

    def sample(self, inds=None):
        lib, ctx = self.env._lib, self.env._ctx
        if inds is not None:
            inds = np.ascontiguousarray(inds, dtype=np.int64)
            L.check(lib.pdeb200_sample(ctx, len(inds), inds.ctypes.data, 0, 0), ctx)
            self._staged_batch = len(inds)
        else:
            self._staged_batch = int(self.batch_size)
            L.check(lib.pdeb200_sample(ctx, int(self.batch_size), None, self.seed ^ 0x5DEECE66D, self._rng_offset), ctx)
            self._rng_offset += int(self.batch_size)

    def set_batch(self, s, a, r, t, snext):
        """Explicit batch in the reference's shapes: s (ns,B), a (na,B), r (B,), t (B,), snext (ns,B)."""
        f = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float32).T)
        s_, a_, s2 = f(s), f(a), f(snext)
        r_ = np.ascontiguousarray(r, dtype=np.float32).reshape(-1)
        t_ = np.ascontiguousarray(t, dtype=np.uint8).reshape(-1)
        self._staged_batch = len(r_)
        L.check(self.env._lib.pdeb200_set_batch(self.env._ctx, len(r_), s_.ctypes.data, a_.ctypes.data, r_.ctypes.data,
                                                t_.ctypes.data, s2.ctypes.data), self.env._ctx)

    def get_batch(self):
        """The staged batch as pde_fetch! returns it (PDEagent.jl:322-340): s (ns,B), a (na,B), r (B,), t (B,), s' (ns,B), inds."""
        lib, ctx, env = self.env._lib, self.env._ctx, self.env
        ns, na = env.ns, (env.n_actuators * env.a_rows if env.cfg.mono else env.a_rows)
        B = int(self._staged_batch)
        s, a, s2 = np.empty((B, ns), np.float32), np.empty((B, na), np.float32), np.empty((B, ns), np.float32)
        r, t, inds = np.empty(B, np.float32), np.empty(B, np.uint8), np.empty(B, np.int64)
        L.check(lib.pdeb200_get_batch(ctx, s.ctypes.data, a.ctypes.data, r.ctypes.data, t.ctypes.data, s2.ctypes.data,
                                      inds.ctypes.data), ctx)
        return s.T, a.T, r, t.astype(bool), s2.T, inds

    # -- DDPG update (PDEagent.jl:363-418) ------------------------------------------------------------
    def update(self):
        """update!(policy, batch) on the staged batch.  With a multi-rank `comm` this is a collective: gradients and
        the quirk-Q1 r-bar are global-batch means, exchanged inside the gradient kernels (no host-side allreduce)."""
        lib, ctx = self.env._lib, self.env._ctx
        L.check(lib.pdeb200_ddpg_update(ctx, float(self.y), float(self.p), float(self.behavior_actor.learning_rate),
                                        float(self.behavior_critic.learning_rate), int(self.literal_q1)), ctx)
        self.n_updates += 1

    def set_sampler_offset(self, offset):
        """Philox counter of the device sampler used by `maybe_update` (part of a resumable checkpoint)."""
        L.check(self.env._lib.pdeb200_rng_set(self.env._ctx, int(offset)), self.env._ctx)

    def sampler_offset(self):
        v = C.c_uint64()
        L.check(self.env._lib.pdeb200_rng_get(self.env._ctx, C.byref(v)), self.env._ctx)
        return v.value

    def set_update_path(self, path):
        """0 auto, 1 layer-wise CUDA cores, 2 layer-wise tensor cores, 3 layer-wise auto (pdeb200_ddpg_set_path)."""
        L.check(self.env._lib.pdeb200_ddpg_set_path(self.env._ctx, int(path)), self.env._ctx)

    def grads(self):
        n = self.env._lib.pdeb200_net_num_params(self.env._ctx, L.NET_BEHAVIOR_CRITIC) + \
            self.env._lib.pdeb200_net_num_params(self.env._ctx, L.NET_BEHAVIOR_ACTOR)
        out = np.empty(n, dtype=np.float32)
        L.check(self.env._lib.pdeb200_get(self.env._ctx, L.ARR_GRADS, out.ctypes.data, out.nbytes), self.env._ctx)
        return out

    def post_episode(self):
        """update!(policy, traj, env, ::PostEpisodeStage): update_step = 0 (PDEagent.jl:215-224)."""
        self.update_step = 0


def create_agent(env, *, rng, nna_scale=1.0, nna_scale_critic=None, drop_middle_layer=False,
                 drop_middle_layer_critic=None, fun="relu", fun_critic=None, mono=False, **kw):
    """create_agent, src/PDEagent.jl:58-119: four networks (targets copied from behaviors) + trajectory."""
    nna_scale_critic = nna_scale if nna_scale_critic is None else nna_scale_critic
    drop_c = drop_middle_layer if drop_middle_layer_critic is None else drop_middle_layer_critic
    fun_critic = fun_critic or fun
    ns = env.ns
    na = env.n_actuators * env.a_rows if mono else env.a_rows
    actor = create_chain(na=na, ns=ns, is_actor=True, rng=rng, nna_scale=nna_scale, drop_middle_layer=drop_middle_layer, fun=fun)
    critic = create_chain(na=na, ns=ns, is_actor=False, rng=rng, nna_scale=nna_scale_critic, drop_middle_layer=drop_c, fun=fun_critic)
    return CustomDDPGPolicy(env, behavior_actor=actor, behavior_critic=critic, **kw)


def run_episode(policy, env, hook=None, learning=True, max_steps=None, check_done=True):
    """One episode in the stage order of RLCore's `run` (spelled out in the reference at
    scripts/Fluid/setup/FluidSetup.jl:455-519):  reset! -> PreEpisode -> loop { policy -> PreAct
    (push s,a; update) -> env(action) -> PostAct (push r,t) } -> PostEpisode (dummy push).

    Batched termination: the B environments share the clock, so the time limit ends the episode for all of them
    at once; an environment that DIVERGES earlier (PDEenv.jl:226-237) ends only its own episode -- its terminal
    flag is pushed, then it is reset in place (`pdeb200_reset_diverged`) while the others keep stepping, so no
    Inf/NaN column is ever integrated further or pushed as a state.  check_done=False skips the per-step
    read-back of the three counters (fixed-length roll-outs)."""
    traj = policy.trajectory
    env.reset()
    if learning:
        traj.pre_episode()
    if hook is not None:
        hook.pre_episode(env, policy)
    steps = 0
    n_diverged = 0
    while True:
        policy(env, learning=learning)
        if learning:
            traj.pre_act()
            policy.maybe_update()
        env.step_device()
        if learning:
            traj.post_act()
        if hook is not None:
            hook.post_act(env)
        steps += 1
        counts = env.reset_diverged(sync=check_done)
        if counts is not None:
            n_diverged += counts[2]
            if counts[1] > 0:                       # the time limit: every environment that was not reset reaches it together
                break
        if max_steps is not None and steps >= max_steps:
            break
    if learning:
        traj.post_episode()
        policy.post_episode()
    if hook is not None:
        hook.post_episode(env, policy)
    run_episode.last_diverged = n_diverged
    return steps
