"""Committed golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the oracle -- the reference
cannot run here and has no vectors of its own, see that script's header).
CPU: the oracle still reproduces them.  GPU (-m gpu): libzoo_sm100 reproduces them through the C ABI; the SumTree ones bit-exactly."""
import importlib.util
import os

import numpy as np
import pytest

from helpers import make_approx
from oracle.sumtree import CSumTree

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
MG = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(MG)
CASES = MG.cases()


def load(name):
    return dict(np.load(os.path.join(HERE, "golden", name + ".npz")))


def check_digest(got, gold, tol, what):
    assert abs(float(got["loss"]) - float(gold["loss"])) <= tol * max(1.0, abs(float(gold["loss"]))), (what, got["loss"], gold["loss"])
    for k in ("per_sample", "priorities"):
        if k in gold:
            assert np.abs(got[k] - gold[k]).max() <= tol * max(1e-30, np.abs(gold[k]).max()), (what, k)
    n = int(gold["n_tensors"])
    assert int(got["n_tensors"]) == n
    for i in range(n):
        norm = float(gold["g%d_norm" % i])
        assert abs(float(got["g%d_norm" % i]) - norm) <= tol * max(norm, 1e-30), (what, i, "norm")
        head = gold["g%d_head" % i]
        scale = max(float(np.abs(head).max()), norm / np.sqrt(max(1, head.size)))
        assert np.abs(got["g%d_head" % i] - head).max() <= tol * scale + 1e-12, (what, i, "head")


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    # 1e-5: BLAS summation order may differ between hosts; the fixtures pin the algorithm, not the last ulp
    check_digest(MG.digest(MG.run_oracle(CASES[name])), load(name), 1e-5, name)


@pytest.mark.parametrize("name", sorted(MG.SUMTREE_CASES))
def test_sumtree_oracle_reproduces_golden(name):
    gold = load(name)
    got = MG.sumtree_case(*MG.SUMTREE_CASES[name])
    for k, v in gold.items():
        assert np.array_equal(got[k], v), (name, k)  # integer / bit-exact fp32


def device_learner(zoo, c, prec="fp32"):
    p, tp, b, tau, taup = MG.inputs(c)
    B, k = c["B"], c["kind"]
    iqn = k == "iqn"
    opt = zoo.ADAM()
    app = make_approx(zoo, c["net"], p, opt, max_batch=B, max_q=c.get("N", 0) if iqn else 0, precision=prec)
    tapp = None if k == "basic" else make_approx(zoo, c["net"], tp, max_batch=B, max_q=c.get("Np", 0) if iqn else 0, precision=prec)
    if k == "basic":
        lrn = zoo.BasicDQNLearner(app, batch_size=B)
    elif k == "dqn":
        lrn = zoo.DQNLearner(app, tapp, batch_size=B, update_horizon=c["n"], is_enable_double_DQN=c["double"])
    elif k == "per":
        lrn = zoo.PrioritizedDQNLearner(app, tapp, batch_size=B)
    elif k == "rainbow":
        lrn = zoo.RainbowLearner(app, tapp, n_actions=c["A"], Vmax=c["vmax"], Vmin=c["vmin"], update_horizon=c["n"], batch_size=B)
    elif k == "qr":
        lrn = zoo.QRDQNLearner(app, tapp, n_actions=c["A"], n_quantile=c["N"], batch_size=B)
    elif k == "rem":
        lrn = zoo.REMDQNLearner(app, tapp, n_actions=c["A"], ensemble_num=c["N"], batch_size=B)
    else:
        lrn = zoo.IQNLearner(app, tapp, n_actions=c["A"], N=c["N"], N_prime=c["Np"], N_em=c["net"]["phi"][0][1], batch_size=B)
    app.set_params(p)  # undo the constructor's online := target force-sync
    kw = dict(tau=tau, tau_prime=taup) if iqn else (dict(convex_polygon=tau) if k == "rem" else {})
    return lrn, b, kw


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "fp32_simt"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_device_reproduces_golden(zoo, name, prec):
    c = CASES[name]
    lrn, b, kw = device_learner(zoo, c, prec)
    loss, grads, prio = lrn.compute_grads(b, **kw)
    r = {"loss": loss, "grads": grads}
    gold = load(name)
    if "per_sample" in gold:
        r["per_sample"] = lrn.per_sample_loss()
    if "priorities" in gold:
        r["priorities"] = prio
    check_digest(MG.digest(r), gold, 1e-4, name)  # north_star: fp32 path within 1e-4 relative


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MG.SUMTREE_CASES))
def test_device_sumtree_reproduces_golden(zoo, name):
    gold = load(name)
    cap, n_push, B, seed = MG.SUMTREE_CASES[name]
    rng = np.random.default_rng(seed)
    pr = np.power(rng.random(n_push, dtype=np.float32) + np.float32(1e-10), np.float32(0.5)).astype(np.float32)
    pr[rng.random(n_push) < 0.01] = np.float32(100.0)
    t = zoo.SumTree(cap)
    t.push_many(pr)
    assert t.total == gold["total0"]
    u64, u32 = rng.random(B), rng.random(B, dtype=np.float32)
    for u, strat, tag in [(u64, True, "strat_f64"), (u64, False, "indep_f64"), (u32, True, "strat_f32"), (u32, False, "indep_f32")]:
        idx, p = t.sample(u, stratified=strat)
        assert np.array_equal(idx, gold["idx_" + tag]) and np.array_equal(p, gold["p_" + tag]), (name, tag)
    t.update(gold["upd_idx"], gold["upd_p"])
    tree = t.export()
    assert t.total == gold["total1"]
    assert np.array_equal(tree[:64], gold["tree_head"])
    assert np.uint32(np.bitwise_xor.reduce(tree.view(np.uint32))) == gold["tree_xor"]
