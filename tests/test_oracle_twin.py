"""CPU: the numpy oracle (hand-written adjoints) must agree with the independent torch-autograd twin on every learner,
and with fp64 central finite differences.  This cross-check is what stands in for reference golden vectors, which
do not exist for this path (parity unpinned)."""
import numpy as np
import pytest
import torch

from helpers import nrel
from oracle import torch_twin as T
from oracle import zoo_oracle as O

TOL = 1e-5


def _cmp(a, b, what, tol=TOL):
    assert abs(float(a["loss"]) - float(b["loss"])) <= tol * max(1.0, abs(float(b["loss"]))), (what, a["loss"], b["loss"])
    assert len(a["grads"]) == len(b["grads"])
    for i, (g, r) in enumerate(zip(a["grads"], b["grads"])):
        assert g.shape == r.shape
        assert nrel(g, r) <= tol, "%s: tensor %d err %.3e" % (what, i, nrel(g, r))
    for k in ("per_sample", "priorities"):
        if k in a or k in b:
            assert nrel(a[k], b[k]) <= tol, (what, k)


def cartpole(seed, **kw):
    return O.synth_batch(seed, 32, (4,), 2, atari=False, **kw)


def test_basic_dqn():
    net = O.chain_net(O.mlp([4, 128, 128, 2]))
    p, b = O.init_params(net, 1, 0.1), cartpole(3)
    _cmp(O.basic_dqn_grads(net, p, b), T.basic_dqn_grads(net, p, b), "basic")


@pytest.mark.parametrize("double,n,mask", [(True, 1, False), (False, 3, False), (True, 3, True)])
def test_dqn(double, n, mask):
    net = O.chain_net(O.mlp([4, 128, 128, 2]))
    p, tp, b = O.init_params(net, 1, 0.1), O.init_params(net, 2, 0.1), cartpole(4, mask=mask)
    _cmp(O.dqn_grads(net, p, tp, b, 0.99, n, double), T.dqn_grads(net, p, tp, b, 0.99, n, double), "dqn")


def test_prioritized_dqn():
    net = O.chain_net(O.mlp([4, 128, 128, 2]))
    p, tp, b = O.init_params(net, 1, 0.1), O.init_params(net, 2, 0.1), cartpole(5, priority=True, mask=True)
    _cmp(O.prioritized_dqn_grads(net, p, tp, b), T.prioritized_dqn_grads(net, p, tp, b), "per")


def test_dueling_dqn():
    net = O.dueling_net(O.mlp([4, 128, 128], relu_last=True), O.mlp([128, 128, 1]), O.mlp([128, 128, 2]))
    p, tp, b = O.init_params(net, 1, 0.1), O.init_params(net, 2, 0.1), cartpole(6)
    _cmp(O.dqn_grads(net, p, tp, b), T.dqn_grads(net, p, tp, b), "dueling")


@pytest.mark.parametrize("per", [True, False])
def test_rainbow(per):
    net = O.chain_net(O.mlp([4, 128, 128, 2 * 51]))
    p, tp, b = O.init_params(net, 1, 0.1), O.init_params(net, 2, 0.1), cartpole(7, priority=per)
    _cmp(O.rainbow_grads(net, p, tp, b, 2, 51, 200.0, 0.0), T.rainbow_grads(net, p, tp, b, 2, 51, 200.0, 0.0), "rainbow")


def test_iqn():
    net = O.cartpole_iqn_net()
    p, tp, b = O.init_params(net, 1, 0.1), O.init_params(net, 2, 0.1), cartpole(8, priority=True)
    rng = np.random.default_rng(0)
    tau, taup = rng.random((32, 8), dtype=np.float32), rng.random((32, 6), dtype=np.float32)  # N != N' exercises q2
    _cmp(O.iqn_grads(net, p, tp, b, tau, taup), T.iqn_grads(net, p, tp, b, tau, taup), "iqn")


def test_qr_dqn():
    # qr_dqn.jl: N quantiles per action, quantile index fastest
    net = O.chain_net(O.mlp([4, 128, 128, 2 * 10]))
    p, tp, b = O.init_params(net, 1, 0.1), O.init_params(net, 2, 0.1), cartpole(14)
    _cmp(O.qr_dqn_grads(net, p, tp, b, 2, 10), T.qr_dqn_grads(net, p, tp, b, 2, 10), "qr-dqn")


@pytest.mark.parametrize("mask", [False, True])
def test_rem_dqn(mask):
    # JuliaRL_REMDQN_CartPole.jl: 16 heads, random convex combination
    net = O.chain_net(O.mlp([4, 128, 128, 2 * 16]))
    p, tp, b = O.init_params(net, 1, 0.1), O.init_params(net, 2, 0.1), cartpole(15, mask=mask)
    w = np.random.default_rng(3).random(16, dtype=np.float32)
    w = (w / w.sum(dtype=np.float32)).astype(np.float32)
    _cmp(O.rem_dqn_grads(net, p, tp, b, 2, 16, w), T.rem_dqn_grads(net, p, tp, b, 2, 16, w), "rem-dqn")


def test_dqn_nature_cnn_small_batch():
    net = O.chain_net(O.nature_cnn(6))
    p, tp = O.init_params(net, 1, 0.05), O.init_params(net, 2, 0.05)
    b = O.synth_batch(11, 2, (4, 84, 84), 6, atari=True, mask=True)
    _cmp(O.dqn_grads(net, p, tp, b), T.dqn_grads(net, p, tp, b), "dqn atari", tol=2e-5)


def test_iqn_atari_net_forward():
    net = O.atari_iqn_net(6)
    p = O.init_params(net, 3, 0.05)
    obs = np.random.default_rng(0).integers(0, 256, (2, 4, 84, 84), dtype=np.uint8)
    tau = np.random.default_rng(1).random((2, 5), dtype=np.float32)
    ref, _ = O.net_forward(net, p, obs, tau=tau)
    tw = T.forward(net, [torch.from_numpy(x) for x in p], torch.from_numpy(obs).float(), torch.from_numpy(tau)).numpy()
    assert nrel(ref, tw) <= 2e-5


@pytest.mark.parametrize("kind", ["dqn", "rainbow", "iqn"])
def test_fp64_finite_differences(kind):
    """Central differences of the fp64 loss w.r.t. a handful of parameters vs the oracle's analytic gradient."""
    rng = np.random.default_rng(42)
    b = O.synth_batch(9, 8, (4,), 2, atari=False, priority=True)
    if kind == "iqn":
        net = O.cartpole_iqn_net()
        tau, taup = rng.random((8, 4)), rng.random((8, 4))
        f = lambda pp: O.iqn_grads(net, pp, tp, b, tau, taup, dtype=np.float64)
    elif kind == "rainbow":
        net = O.chain_net(O.mlp([4, 16, 2 * 11]))
        f = lambda pp: O.rainbow_grads(net, pp, tp, b, 2, 11, 10.0, -10.0, dtype=np.float64)
    else:
        net = O.chain_net(O.mlp([4, 16, 2]))
        f = lambda pp: O.dqn_grads(net, pp, tp, b, dtype=np.float64)
    p = [x.astype(np.float64) for x in O.init_params(net, 1, 0.1)]
    tp = [x.astype(np.float64) for x in O.init_params(net, 2, 0.1)]
    base = f(p)
    h = 1e-6
    for ti in range(len(p)):
        flat = p[ti].reshape(-1)
        for j in rng.integers(0, flat.size, 3):
            old = flat[j]
            flat[j] = old + h
            lp = float(f(p)["loss"])
            flat[j] = old - h
            lm = float(f(p)["loss"])
            flat[j] = old
            fd = (lp - lm) / (2 * h)
            an = base["grads"][ti].reshape(-1)[j]
            assert abs(fd - an) <= 1e-6 + 1e-5 * abs(an), (kind, ti, j, fd, an)


def test_optimisers_match_twin():
    rng = np.random.default_rng(0)
    p = [rng.standard_normal((7, 5)).astype(np.float32), rng.standard_normal(5).astype(np.float32)]
    for name in ("rmsprop", "adam"):
        ps = [torch.from_numpy(x.copy()) for x in p]
        opt = T.RMSProp(ps, 0.00025, 0.95) if name == "rmsprop" else T.Adam(ps, 6.25e-5)
        cur = [x.copy() for x in p]
        st = O.rmsprop_init(cur) if name == "rmsprop" else O.adam_init(cur)
        for step in range(4):
            g = [rng.standard_normal(x.shape).astype(np.float32) for x in p]
            cur = O.rmsprop_step(cur, g, st, 0.00025, 0.95) if name == "rmsprop" else O.adam_step(cur, g, st, 6.25e-5)
            with torch.no_grad():
                opt.step(ps, [torch.from_numpy(x) for x in g])
            for a, r in zip(ps, cur):
                assert np.abs(a.numpy() - r).max() <= 1e-7


def test_twin_with_its_own_relu_masks_pinned_reproduces_the_free_run():
    """The full-size GPU parity tests re-run the twin with the DEVICE's relu masks (tests/helpers.py:full_size_parity);
    pinning the twin's own masks must change nothing -- checks the mask / pre-activation plumbing on the CPU."""
    net = O.atari_iqn_net(6)
    p, tp = O.init_params(net, 1, 0.05), O.init_params(net, 2, 0.05)
    B, N = 2, 8
    b = O.synth_batch(3, B, (4, 84, 84), 6, atari=True)
    rng = np.random.default_rng(0)
    tau, taup = rng.random((B, N), dtype=np.float32), rng.random((B, N), dtype=np.float32)
    rec = []
    free = T.iqn_grads(net, p, tp, b, tau, taup, rec=rec)
    assert [tuple(r.shape) for r in rec] == [(B, 32, 21, 21), (B, 64, 11, 11), (B, 64, 11, 11), (B * N, 7744), (B * N, 512)]
    pinned = T.iqn_grads(net, p, tp, b, tau, taup, masks=[(r > 0).float() for r in rec])
    assert float(free["loss"]) == float(pinned["loss"])
    for x, y in zip(free["grads"], pinned["grads"]):
        assert np.array_equal(x, y)
    net = O.chain_net(O.nature_cnn(6))
    p, tp = O.init_params(net, 1, 0.05), O.init_params(net, 2, 0.05)
    rec = []
    free = T.dqn_grads(net, p, tp, b, rec=rec)
    pinned = T.dqn_grads(net, p, tp, b, masks=[(r > 0).float() for r in rec])
    for x, y in zip(free["grads"], pinned["grads"]):
        assert np.array_equal(x, y)
