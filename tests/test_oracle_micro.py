"""CPU: pins the oracle on the hand-computable micro cases of SURVEY.md A.8 (the reference's own tests hold no
numeric vector for this path -- test/runtests.jl:34-48 only smoke-runs -- so these analytic cases, the independent
torch twin (test_oracle_twin.py) and the fp64 finite differences are the pins; parity stays "unpinned")."""
import numpy as np
import pytest

from oracle import zoo_oracle as O
from oracle.sumtree import CSumTree, PySumTree

F = np.float32


def test_huber_td_micro():
    # dqn.jl:133-137: G=[0.5,3.0], q=[0,0] -> per-sample [0.125, 2.5], mean 1.3125, dL/dq = [-0.25, -0.5]
    G, q = np.array([0.5, 3.0], F), np.zeros(2, F)
    per, dq, dG = O.huber_per_sample(G, q)
    assert np.allclose(per, [0.125, 2.5])
    assert np.isclose(per.mean(), 1.3125)
    assert np.allclose(dq / 2, [-0.25, -0.5])
    assert np.allclose(dG, -dq)
    # strict `<` (q4): |e| == delta takes the linear branch; both branches meet there
    per1, dq1, _ = O.huber_per_sample(np.array([1.0], F), np.zeros(1, F))
    assert per1[0] == F(0.5) and dq1[0] == F(-1.0)


def test_per_micro():
    # prioritized_dqn.jl:125-143 with priority=[1,4], beta=0.5
    w = O.per_weights(np.array([1.0, 4.0], F), 0.5)
    assert np.allclose(w, [1.0, 0.5], atol=1e-6)
    per = np.array([0.125, 2.5], F)
    assert np.isclose(np.dot(w, per) / 2, 0.6875, atol=1e-6)
    assert np.allclose(O.new_priorities(per, 0.5), [0.35355, 1.58114], atol=1e-5)


def test_c51_projection_micro():
    # rainbow.jl:204-221: z=[-1,0,1], dz=1, r=0.5, gamma^n(1-t)=1, p=[0.2,0.5,0.3] -> m=[0.10,0.35,0.55]
    z = np.array([-1.0, 0.0, 1.0], F)
    p = np.array([[0.2, 0.5, 0.3]], F)
    Tz = (F(0.5) + z * F(1.0))[None]
    m = O.project_distribution(Tz, p, z, F(1.0), -1.0, 1.0)
    assert np.allclose(m, [[0.10, 0.35, 0.55]], atol=1e-6)
    # terminal: Tz = [0.5,0.5,0.5] -> [0, 0.5, 0.5]
    Tz = (F(0.5) + z * F(0.0))[None]
    m = O.project_distribution(Tz, p, z, F(1.0), -1.0, 1.0)
    assert np.allclose(m, [[0.0, 0.5, 0.5]], atol=1e-6)


@pytest.mark.parametrize("seed", range(5))
def test_c51_projection_properties(seed):
    # columns of the projected distribution sum to 1 when the source does; mass stays inside [vmin, vmax]
    rng = np.random.default_rng(seed)
    B, n = 16, 51
    z = np.linspace(F(-10), F(10), n).astype(F)
    dz = F(z[1] - z[0])
    p = O.softmax(rng.standard_normal((B, n)).astype(F), axis=1)
    r = rng.integers(-1, 2, B).astype(F)
    disc = (F(0.99) ** 3) * (rng.random(B) > 0.2).astype(F)
    Tz = r[:, None] + z[None] * disc[:, None]
    m = O.project_distribution(Tz, p, z, dz, -10.0, 10.0)
    assert np.all(m >= 0)
    assert np.allclose(m.sum(axis=1), 1.0, atol=2e-5)
    # expectation is preserved by the triangular kernel while nothing is clamped
    inside = (Tz.min(axis=1) >= -10) & (Tz.max(axis=1) <= 10)
    assert np.allclose((m * z).sum(axis=1)[inside], (p * Tz).sum(axis=1)[inside], atol=1e-4)


def test_quantile_huber_micro():
    # iqn.jl:207-219 with N = N' = B = 1, kappa = 1
    def one(target, q, tau):
        per, dq = O.quantile_huber(np.array([[target]], F), np.array([[q]], F), np.array([[tau]], F), 1.0)
        return float(per[0]), float(dq[0, 0])
    loss, d = one(1.0, 0.2, 0.3)
    assert np.isclose(loss, 0.096, atol=1e-6) and np.isclose(d, -0.24, atol=1e-6)
    loss, d = one(0.0, 0.5, 0.3)
    assert np.isclose(loss, 0.0875, atol=1e-6) and np.isclose(d, 0.35, atol=1e-6)
    loss, d = one(2.0, 0.0, 0.5)
    assert np.isclose(loss, 0.75, atol=1e-6) and np.isclose(d, -0.5, atol=1e-6)


def test_quantile_huber_sum_over_target_mean_over_online():
    # q2 (iqn.jl:218-219): sum over N' (targets), mean over N (online quantiles) -- differs from the transpose when N != N'
    rng = np.random.default_rng(0)
    B, N, Np = 3, 4, 6
    t, q, tau = rng.standard_normal((B, Np)).astype(F), rng.standard_normal((B, N)).astype(F), rng.random((B, N), dtype=F)
    per, _ = O.quantile_huber(t, q, tau, 1.0)
    TD = t[:, :, None] - q[:, None, :]
    hub = np.where(np.abs(TD) <= 1, 0.5 * TD * TD, np.abs(TD) - 0.5)
    raw = np.abs(tau[:, None, :] - (TD < 0)) * hub
    assert np.allclose(per, raw.sum(axis=1).mean(axis=1), atol=1e-6)


def test_qr_quantile_huber_micro():
    # qr_dqn.jl:3-14 with N = 2, B = 1, kappa = 1: y = [1, 0], y_hat = [0.2, 0.5]; cum_prob = [0.25, 0.75] along the TARGET axis
    per, d = O.quantile_huber_loss_qr(np.array([[0.2, 0.5]], F), np.array([[1.0, 0.0]], F), 1.0)
    # D[i,j] = y[i]-yhat[j] = [[0.8, 0.5], [-0.2, -0.5]]; huber = [[.32, .125], [.02, .125]]; w = [[.25, .25], [.25, .25]]
    assert np.isclose(per[0], ((0.25 * 0.32 + 0.25 * 0.02) + (0.25 * 0.125 + 0.25 * 0.125)) / 2, atol=1e-6)
    assert np.allclose(d[0], [-(0.25 * 0.8 + 0.25 * -0.2) / 2, -(0.25 * 0.5 + 0.25 * -0.5) / 2], atol=1e-6)
    assert np.allclose(O.qr_cum_prob(4), [0.125, 0.375, 0.625, 0.875])


@pytest.mark.parametrize("Tree", [CSumTree, PySumTree])
def test_sumtree_docstring_example(Tree):
    # RLCore docstring example: SumTree(8), push 1..16 -> leaves 9..16, root 100, node2 42, node3 58
    t = Tree(8)
    for i in range(1, 17):
        t.push(F(i))
    tree = t.export() if Tree is CSumTree else t.tree[1:]
    assert len(tree) == 7 + 8
    assert list(tree[7:]) == [9, 10, 11, 12, 13, 14, 15, 16]
    assert tree[0] == 100 and tree[1] == 42 and tree[2] == 58
    assert t.first == 1
    # get(12.5) -> (2, 10.0); stratified n=4, all u=0.5 -> v = 12.5, 37.5, 62.5, 87.5 -> indices 2,4,6,8
    for dt in (np.float32, np.float64):
        idx, pr = t.sample(np.full(4, 0.5, dt), stratified=True)
        assert list(idx) == [2, 4, 6, 8] and list(pr) == [10, 12, 14, 16]
        idx, pr = t.sample(np.array([0.125], dt), stratified=False)
        assert list(idx) == [2] and pr[0] == 10
    # t[2] = 0.5: change -9.5 -> node4 9.5, node2 32.5, root 90.5
    t.update([2], [F(0.5)])
    tree = t.export() if Tree is CSumTree else t.tree[1:]
    assert tree[3] == 9.5 and tree[1] == 32.5 and tree[0] == 90.5


@pytest.mark.parametrize("Tree", [CSumTree, PySumTree])
def test_sumtree_duplicates_in_batch_are_sequential(Tree):
    t = Tree(8)
    for i in range(1, 17):
        t.push(F(i))
    t.update([2, 2], [F(0.5), F(3.0)])  # the second change is 3.0 - 0.5, not 3.0 - 10
    tree = t.export() if Tree is CSumTree else t.tree[1:]
    assert tree[7 + 1] == 3.0 and tree[0] == 93.0


@pytest.mark.parametrize("cap", [5, 8, 100, 1000])
def test_sumtree_c_equals_python(cap):
    rng = np.random.default_rng(cap)
    c, p = CSumTree(cap), PySumTree(cap)
    for v in rng.random(cap + cap // 2 + 1, dtype=F):
        c.push(v)
        p.push(v)
    assert c.first == p.first and len(c) == p.length
    assert np.array_equal(c.export(), p.tree[1:])
    for dt in (np.float32, np.float64):
        for strat in (True, False):
            u = rng.random(17).astype(dt)
            ci, cp = c.sample(u, strat)
            pi, pp = p.sample(u, strat)
            assert np.array_equal(ci, pi) and np.array_equal(cp, pp)
    idx = rng.integers(1, cap + 1, 40)
    idx[7] = idx[3]
    newp = rng.random(40, dtype=F)
    c.update(idx, newp)
    p.update(idx, newp)
    assert np.array_equal(c.export(), p.tree[1:])


def test_sumtree_stratified_sample_lies_in_its_stratum():
    cap, n = 4096, 64
    rng = np.random.default_rng(3)
    t = CSumTree(cap)
    pr = rng.random(cap, dtype=F)
    for v in pr:
        t.push(v)
    idx, got = t.sample(rng.random(n), stratified=True)
    cum = np.concatenate([[0.0], np.cumsum(pr.astype(np.float64))])
    total = float(t.total)
    for i, k in enumerate(idx):
        lo, hi = cum[k - 1], cum[k]  # mass interval of leaf k
        assert hi >= i / n * total - 1e-2 and lo <= (i + 1) / n * total + 1e-2
        assert got[i] == pr[k - 1]


def test_optimiser_first_steps():
    # RMSProp first step eta*g/(sqrt(0.05)|g| + 1e-8) ~ 4.4721*eta*sign(g); ADAM first step ~ eta*sign(g)
    g = [np.array([0.3, -2.0, 1e-3], F)]
    p = [np.zeros(3, F)]
    st = O.rmsprop_init(p)
    new = O.rmsprop_step(p, g, st, 0.00025, 0.95)
    assert np.allclose(-new[0], 0.00025 * 4.472136 * np.sign(g[0]), rtol=1e-4)
    assert np.allclose(st["acc"][0], 0.05 * g[0] ** 2, rtol=1e-6)
    st = O.adam_init(p)
    new = O.adam_step(p, g, st, 6.25e-5)
    assert np.allclose(-new[0], 6.25e-5 * np.sign(g[0]), rtol=1e-4)
    # second ADAM step uses beta^2 in the bias corrections (beta^t carried per tensor, starts at beta)
    new2 = O.adam_step(new, g, st, 6.25e-5)
    m = 0.9 * 0.1 * g[0] + 0.1 * g[0]
    v = 0.999 * 0.001 * g[0] ** 2 + 0.001 * g[0] ** 2
    step = m / (1 - 0.81) / (np.sqrt(v / (1 - 0.999 ** 2)) + 1e-8) * 6.25e-5
    assert np.allclose(new[0] - new2[0], step, rtol=1e-4)


def test_nature_cnn_geometry_and_flatten_order():
    # q9: 84 -> 21 -> 11 -> 11, fc1 fan-in 7744; flatten index = w + 11 h + 121 c (NCHW flatten)
    net = O.chain_net(O.nature_cnn(6))
    shapes = O.param_shapes(net)
    assert shapes == [(32, 4, 8, 8), (32,), (64, 32, 4, 4), (64,), (64, 64, 3, 3), (64,), (512, 7744), (512,), (6, 512), (6,)]
    assert sum(int(np.prod(s)) for s in shapes) == 4_046_502
    p = O.init_params(net, 0)
    x = np.random.default_rng(0).integers(0, 256, (2, 4, 84, 84), dtype=np.uint8)
    q, cache = O.net_forward(net, p, x)
    assert q.shape == (2, 6)
