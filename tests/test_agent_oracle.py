"""CPU: the agent oracle checks itself (parity unpinned by the reference: no saved gradients).
Finite-difference gradients of the restated losses, ADAM/Polyak arithmetic, replay-ring semantics."""
import numpy as np

from oracle import agent_oracle as AO


def make_nets(rng, ns, na, ha, hc, middle=False):
    g = lambda o, i: ((rng.random((o, i)) - 0.5) * np.sqrt(24.0 / (o + i))).astype(np.float32)
    b = lambda o: (0.1 * rng.standard_normal(o)).astype(np.float32)
    al = [(g(ha, ns), b(ha), "relu")] + ([(g(ha, ha), b(ha), "relu")] if middle else []) + [(g(na, ha), b(na), "tanh")]
    cl = [(g(hc, ns + na), b(hc), "relu")] + ([(g(hc, hc), b(hc), "relu")] if middle else []) + [(g(1, hc), b(1), None)]
    return AO.Net(al), AO.Net(cl)


def batch(rng, ns, na, B):
    return (rng.standard_normal((ns, B)).astype(np.float32), rng.uniform(-1, 1, (na, B)).astype(np.float32),
            rng.standard_normal(B).astype(np.float32), rng.random(B) < 0.2, rng.standard_normal((ns, B)).astype(np.float32))


def fd_check(loss_fn, net, grads, rng, n=12, eps=1e-3):
    flat = AO.flat_grads(grads)
    params = net.params()
    # map flat index -> (array, index)
    offs = np.cumsum([0] + [p.size for p in params])
    worst = 0.0
    for _ in range(n):
        q = int(rng.integers(0, offs[-1]))
        ai = int(np.searchsorted(offs, q, side="right") - 1)
        arr = params[ai]
        idx = np.unravel_index(q - offs[ai], arr.shape, order="F")
        old = arr[idx]
        arr[idx] = old + eps
        lp = float(loss_fn())
        arr[idx] = old - eps
        lm = float(loss_fn())
        arr[idx] = old
        fd = (lp - lm) / (2 * eps)
        worst = max(worst, abs(fd - flat[q]) / max(1e-3, abs(fd)))
    return worst


def test_critic_and_actor_gradients_match_finite_differences():
    rng = np.random.default_rng(0)
    for middle in (False, True):
        for literal in (True, False):
            A, Cn = make_nets(rng, 3, 1, 6, 14, middle)
            d = AO.DDPG(A, Cn)
            s, a, r, t, s2 = batch(rng, 3, 1, 8)
            _, gc = d.critic_loss_and_grads(s, a, r, t, s2, literal)
            w = fd_check(lambda: d.critic_loss_and_grads(s, a, r, t, s2, literal)[0], d.C, gc, rng)
            assert w < 2e-2, (middle, literal, w)
            _, ga = d.actor_loss_and_grads(s)
            w = fd_check(lambda: d.actor_loss_and_grads(s)[0], d.A, ga, rng)
            assert w < 2e-2, (middle, w)


def test_literal_q1_equals_mse_with_batch_mean_reward():
    """Quirk Q1: the (1,B) x (B,) broadcast makes dLoss/dq the MSE gradient with r_i -> mean(r)."""
    rng = np.random.default_rng(1)
    A, Cn = make_nets(rng, 1, 1, 6, 140)
    d = AO.DDPG(A, Cn)
    s, a, r, t, s2 = batch(rng, 1, 1, 3)
    _, g_lit = d.critic_loss_and_grads(s, a, r, t, s2, True)
    _, g_mean = d.critic_loss_and_grads(s, a, np.full(3, r.mean(), np.float32), t, s2, False)
    assert np.allclose(AO.flat_grads(g_lit), AO.flat_grads(g_mean), rtol=1e-5, atol=1e-7)


def test_adam_first_step_and_polyak():
    x = np.array([1.0, -2.0], dtype=np.float32)
    g = np.array([0.5, -0.25], dtype=np.float32)
    opt = AO.Adam(1e-3, 1)
    opt.step([x], [g])
    # first ADAM step moves every coordinate by eta * sign(g) (up to eps)
    assert np.allclose(x, [1.0 - 1e-3, -2.0 + 1e-3], atol=1e-7)
    assert opt.bp[0] == [0.9 * 0.9, 0.999 * 0.999]
    rng = np.random.default_rng(2)
    A, Cn = make_nets(rng, 1, 1, 6, 140)
    d = AO.DDPG(A, Cn)
    before = d.At.flat().copy()
    s, a, r, t, s2 = batch(rng, 1, 1, 3)
    d.update(s, a, r, t, s2)
    p = np.float32(0.995)
    assert np.allclose(d.At.flat(), p * before + (np.float32(1) - p) * d.A.flat(), atol=1e-7)


def test_trajectory_ring_semantics():
    """capacity+1 state/action ring vs capacity reward ring: aligned until the first wrap, then the
    reference's buffers are offset by one column (documented in DESIGN.md); fetch is literal."""
    cap, ns, na, ncols = 10, 1, 1, 2
    tr = AO.Trajectory(cap, ns, na)
    tr.pre_episode(ncols)
    for step in range(4):
        tr.pre_act(np.full((ns, ncols), step, np.float32) + np.array([[0.0, 0.5]], np.float32), np.full((na, ncols), -step, np.float32))
        tr.post_act(np.array([10.0 + step, 20.0 + step], np.float32), False)
    assert len(tr) == 8 and tr.state.len == 8
    s, a, r, t, s2 = tr.fetch([0, 3], ncols)
    assert list(s[0]) == [0.0, 1.5] and list(r) == [10.0, 21.0] and list(s2[0]) == [1.0, 2.5]
    tr.post_episode(np.full((ns, ncols), 9, np.float32), na)
    assert tr.state.len == 10 and len(tr) == 8
    tr.pre_episode(ncols)
    assert tr.state.len == 8
    for step in range(4, 8):                              # wraps: 16 pushes into cap 10 / 11
        tr.pre_act(np.full((ns, ncols), step, np.float32) + np.array([[0.0, 0.5]], np.float32), np.full((na, ncols), -step, np.float32))
        tr.post_act(np.array([10.0 + step, 20.0 + step], np.float32), False)
    assert len(tr) == 10 and tr.state.len == 11
    s, a, r, t, s2 = tr.fetch([0], ncols)
    assert r[0] == 13.0 and s[0, 0] == 2.5               # state ring holds one more old column than the reward ring
