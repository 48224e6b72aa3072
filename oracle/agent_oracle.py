"""CPU oracle: weight-shared per-actuator DDPG agent (float32 restatement).

TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Restates /root/reference/src/PDEagent.jl:
  create_NNA network shapes        :14-56
  policy forward                   :175-209
  trajectory update! overloads     :237-314  (CircularArraySARTTrajectory of RLCore 0.8.13)
  pde_sample / pde_fetch!          :317-340
  DDPG update!                     :363-418
and Flux's Dense / ADAM (third-party, version unpinned; Flux 0.13-era semantics).

PARITY UNPINNED: the reference ships weights but no (input, output, gradient)
triple, so this restatement is validated by finite differences of its own loss
(tests/test_agent_oracle.py), not by reference outputs.
"""
import numpy as np

F = np.float32


class Net:
    """Flux Chain of Dense layers. layers: list of [W (out,in) f32, b (out,) f32, act]."""

    def __init__(self, layers):
        self.layers = [[np.array(W, dtype=F), np.array(b, dtype=F), act] for W, b, act in layers]

    def copy(self):
        return Net([(W.copy(), b.copy(), a) for W, b, a in self.layers])

    def forward(self, x, keep=False):
        h = np.asarray(x, dtype=F)
        acts = [h]
        for W, b, act in self.layers:
            z = W @ h + b[:, None]
            if act == "relu":
                h = np.maximum(z, F(0))
            elif act == "tanh":
                h = np.tanh(z)
            else:
                h = z
            acts.append(h.astype(F))
        return acts if keep else acts[-1]

    def backward(self, acts, dout):
        """dout: dLoss/d(output) (n_out, B). Returns (param grads [(dW, db)], dLoss/d(input))."""
        grads = []
        d = np.asarray(dout, dtype=F)
        for li in range(len(self.layers) - 1, -1, -1):
            W, b, act = self.layers[li]
            out, inp = acts[li + 1], acts[li]
            if act == "relu":
                d = d * (out > 0)
            elif act == "tanh":
                d = d * (F(1) - out * out)
            grads.append(((d @ inp.T).astype(F), d.sum(axis=1).astype(F)))
            d = (W.T @ d).astype(F)
        return grads[::-1], d

    def params(self):
        return [p for W, b, _ in self.layers for p in (W, b)]

    def flat(self):
        return np.concatenate([np.concatenate([W.flatten(order="F"), b]) for W, b, _ in self.layers]).astype(F)


def flat_grads(grads):
    return np.concatenate([np.concatenate([dW.flatten(order="F"), db]) for dW, db in grads]).astype(F)


class Adam:
    """Flux.Optimise.ADAM(eta, (0.9, 0.999)), epsilon 1e-8: state per parameter array (m, v, beta_p)
    with Float64 hyper-parameters applied to Float32 arrays (each broadcast rounds to Float32)."""

    def __init__(self, eta, n_arrays):
        self.eta = float(eta)
        self.beta = (0.9, 0.999)
        self.eps = 1e-8
        self.m = [None] * n_arrays
        self.v = [None] * n_arrays
        self.bp = [[0.9, 0.999] for _ in range(n_arrays)]

    def step(self, params, grads):
        for i, (x, g) in enumerate(zip(params, grads)):
            if self.m[i] is None:
                self.m[i] = np.zeros_like(x)
                self.v[i] = np.zeros_like(x)
            b1, b2 = self.beta
            g64 = g.astype(np.float64)
            self.m[i] = (b1 * self.m[i].astype(np.float64) + (1 - b1) * g64).astype(F)
            self.v[i] = (b2 * self.v[i].astype(np.float64) + (1 - b2) * g64 * g64).astype(F)
            delta = (self.m[i].astype(np.float64) / (1 - self.bp[i][0]) /
                     (np.sqrt(self.v[i].astype(np.float64) / (1 - self.bp[i][1])) + self.eps) * self.eta).astype(F)
            self.bp[i][0] *= b1
            self.bp[i][1] *= b2
            x -= delta


class DDPG:
    """CustomDDPGPolicy update (PDEagent.jl:363-418)."""

    def __init__(self, actor, critic, lr_actor=5e-4, lr_critic=1e-3, gamma=0.99, polyak=0.995):
        self.A, self.C = actor, critic
        self.At, self.Ct = actor.copy(), critic.copy()
        self.gamma, self.polyak = F(gamma), F(polyak)
        self.opt_a = Adam(lr_actor, len(actor.params()))
        self.opt_c = Adam(lr_critic, len(critic.params()))
        self.actor_loss = self.critic_loss = F(0)

    def critic_loss_and_grads(self, s, a, r, t, snext, literal_q1=True):
        B = s.shape[1]
        anext = self.At.forward(snext)
        qt = self.Ct.forward(np.vstack([snext, anext]))[0]
        T = self.gamma * (F(1) - t.astype(F)) * qt
        acts = self.C.forward(np.vstack([s, a]), keep=True)
        q = acts[-1][0]
        r = np.asarray(r, dtype=F).reshape(-1)
        if literal_q1:
            # qnext = r .+ y .* (1 .- t) .* q_t with r a (1,B) row and the rest length-B vectors
            # => (B,B) matrix [i,j] = r_j + T_i ; loss = mean((qnext .- q).^2), q broadcast along dim 1 (Q1)
            diff = r[None, :] + T[:, None] - q[:, None]
            loss = np.mean(diff * diff, dtype=F)
            dq = (-2.0 / (B * B)) * diff.sum(axis=1)
        else:
            diff = r + T - q
            loss = np.mean(diff * diff, dtype=F)
            dq = (-2.0 / B) * diff
        grads, _ = self.C.backward(acts, dq[None, :].astype(F))
        return F(loss), grads

    def actor_loss_and_grads(self, s):
        B = s.shape[1]
        a_acts = self.A.forward(s, keep=True)
        c_acts = self.C.forward(np.vstack([s, a_acts[-1]]), keep=True)
        loss = -np.mean(c_acts[-1], dtype=F)
        _, dx = self.C.backward(c_acts, np.full((1, B), -1.0 / B, dtype=F))
        da = dx[s.shape[0]:]
        grads, _ = self.A.backward(a_acts, da)
        return F(loss), grads

    def update(self, s, a, r, t, snext, literal_q1=True):
        s, a, snext = (np.asarray(x, dtype=F) for x in (s, a, snext))
        self.critic_loss, gc = self.critic_loss_and_grads(s, a, r, t, snext, literal_q1)
        self.opt_c.step(self.C.params(), [g for pair in gc for g in pair])
        self.actor_loss, ga = self.actor_loss_and_grads(s)
        self.opt_a.step(self.A.params(), [g for pair in ga for g in pair])
        p = self.polyak
        for dst, src in zip(self.At.params() + self.Ct.params(), self.A.params() + self.C.params()):
            dst[...] = p * dst + (F(1) - p) * src
        return gc, ga


class Ring:
    """CircularArrayBuffer: push overwrites the oldest element when full; logical index 0 = oldest."""

    def __init__(self, rows, capacity):
        self.data = np.zeros((rows, capacity), dtype=F)
        self.cap, self.start, self.len = capacity, 0, 0

    def push(self, col):
        if self.len < self.cap:
            self.data[:, (self.start + self.len) % self.cap] = col
            self.len += 1
        else:
            self.data[:, self.start] = col
            self.start = (self.start + 1) % self.cap

    def pop(self):
        self.len -= 1

    def get(self, idx):
        return self.data[:, (self.start + np.asarray(idx)) % self.cap]


class Trajectory:
    """CircularArraySARTTrajectory(capacity; state => ns, action => na, reward => 1) as the reference
    uses it: state/action rings hold capacity+1 columns, reward/terminal rings hold capacity."""

    def __init__(self, capacity, ns, na):
        self.state, self.action = Ring(ns, capacity + 1), Ring(na, capacity + 1)
        self.reward, self.terminal = Ring(1, capacity), Ring(1, capacity)

    def __len__(self):
        return self.terminal.len

    def pre_episode(self, n_cols):          # PDEagent.jl:237-252
        if len(self) > 0:
            for _ in range(n_cols):
                self.state.pop(); self.action.pop()

    def pre_act(self, state, action):       # :254-274
        for i in range(state.shape[1]):
            self.state.push(state[:, i]); self.action.push(action[:, i])

    def post_act(self, reward, done):       # :276-289
        for i in range(len(reward)):
            self.reward.push(reward[i]); self.terminal.push(float(done))

    def post_episode(self, state, na):      # :291-314
        for i in range(state.shape[1]):
            self.state.push(state[:, i]); self.action.push(np.zeros(na, dtype=F))

    def fetch(self, inds, n_cols):          # pde_fetch!, :323-340 (inds 0-based here)
        inds = np.asarray(inds)
        return (self.state.get(inds), self.action.get(inds), self.reward.get(inds)[0],
                self.terminal.get(inds)[0] != 0, self.state.get(inds + n_cols))
