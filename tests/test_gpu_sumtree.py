"""GPU: SumTree sample / priority update through the C ABI, bit-exact against the C oracle (north_star: "SumTree sample
indices and priority updates bit-exact given the same uniform draws"), plus size-independent properties at the full
BASELINE.json size (capacity 1M, batch 32 -> 4096)."""
import numpy as np
import pytest

from oracle.sumtree import CSumTree

pytestmark = pytest.mark.gpu
F = np.float32


def priorities(rng, n):
    p = np.power(rng.random(n, dtype=F) + F(1e-10), F(0.5)).astype(F)
    p[rng.random(n) < 0.01] = F(100.0)  # default_priority (prioritized_dqn.jl:57)
    return p


def pair(zoo, cap, n_push, seed):
    rng = np.random.default_rng(seed)
    pr = priorities(rng, n_push)
    t, o = zoo.SumTree(cap), CSumTree(cap)
    t.push_many(pr)
    o.push_many(pr)
    return t, o, rng


def test_docstring_example(zoo):
    t = zoo.SumTree(8)
    for i in range(1, 17):
        t.push(F(i))
    tree = t.export()
    assert list(tree[7:]) == [9, 10, 11, 12, 13, 14, 15, 16] and tree[0] == 100 and tree[1] == 42 and tree[2] == 58
    idx, pr = t.sample(np.full(4, 0.5), stratified=True)
    assert list(idx) == [2, 4, 6, 8] and list(pr) == [10, 12, 14, 16]
    t.update([2, 2], [F(0.5), F(3.0)])  # duplicates see each other's writes
    assert t[2] == 3.0 and t.total == 93.0


@pytest.mark.parametrize("cap,n_push,B", [(5, 3, 2), (8, 16, 4), (1000, 1300, 32), (1000, 700, 1), (65536, 70000, 1024)])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_sample_and_update_bit_exact(zoo, cap, n_push, B, dt):
    t, o, rng = pair(zoo, cap, n_push, cap + B)
    assert len(t) == len(o) and t.first == o.first
    assert np.array_equal(t.export(), o.export())
    for rnd in range(3):
        for strat in (True, False):
            u = rng.random(B).astype(dt)
            idx, pv = t.sample(u, stratified=strat)
            oidx, opv = o.sample(u, stratified=strat)
            assert np.array_equal(idx, oidx) and np.array_equal(pv, opv), (rnd, strat)
        newp = np.power(rng.random(B, dtype=F) + F(1e-10), F(0.5))
        if B > 3:
            idx[B // 2] = idx[0]
            idx[B - 1] = idx[0]  # several duplicates of one leaf, in batch order
        t.update(idx, newp)
        o.update(idx, newp)
        assert np.array_equal(t.export(), o.export()), rnd
        assert np.array_equal(t[idx], np.array([o[i] for i in idx], F))


def test_empty_and_edge_inputs(zoo):
    t, o, rng = pair(zoo, 16, 16, 0)
    before = t.export()
    t.update(np.zeros(0, np.int64), np.zeros(0, F))  # empty batch: no-op
    assert np.array_equal(t.export(), before)
    idx, pv = t.sample(np.zeros(0), stratified=True)
    assert idx.size == 0 and pv.size == 0
    # u -> 0 picks the first leaf with mass, u just below 1 the last
    idx, _ = t.sample(np.array([0.0, np.nextafter(1.0, 0.0)]), stratified=False)
    oidx, _ = o.sample(np.array([0.0, np.nextafter(1.0, 0.0)]), stratified=False)
    assert np.array_equal(idx, oidx)
    with pytest.raises(zoo.ZooError):
        t.update(np.array([17], np.int64), np.array([1.0], F))  # out of range logical index
    with pytest.raises(zoo.ZooError):
        t.update(np.array([0], np.int64), np.array([1.0], F))


@pytest.mark.parametrize("B", [32, 4096])
def test_full_size_capacity_1m(zoo, B):
    """BASELINE.json config 3: capacity 1M, batch 32 -> 4096.  Bit-exact vs the C oracle, and the properties:
    sample i lies in stratum i; after updates the root equals the fp64 sum of the leaves to fp32 rounding noise."""
    cap = 1_000_000
    t, o, rng = pair(zoo, cap, cap + 12345, 7)
    assert t.first == o.first == 12346
    for rnd in range(4):
        u = rng.random(B)
        idx, pv = t.sample(u, stratified=True)
        oidx, opv = o.sample(u, stratified=True)
        assert np.array_equal(idx, oidx) and np.array_equal(pv, opv)
        ok = pv > 0  # fp32 drift of the root lets the last draws overshoot onto an empty node (reference quirk)
        assert ok.mean() > 0.99 and np.array_equal(t[idx[ok]], pv[ok])
        newp = priorities(rng, B)
        t.update(idx, newp)
        o.update(idx, newp)
    tree, otree = t.export(), o.export()
    assert np.array_equal(tree, otree)
    P = (1 << 20) - 1
    leaves = tree[P:].astype(np.float64)
    assert abs(float(tree[0]) - leaves.sum()) <= 2e-3 * leaves.sum()  # fp32 `+= change` drift only
    # stratum property: the descent walks cumulative mass in PHYSICAL leaf order; sample i lies in stratum i
    idx, pv = t.sample(rng.random(B), stratified=True)
    ok = pv > 0
    phys = idx + t.first - 1
    phys = np.where(phys > cap, phys - cap, phys)
    cum = np.cumsum(leaves)
    lo = np.concatenate([[0.0], cum])[phys - 1]
    hi = cum[phys - 1]
    total = float(tree[0])
    i = np.arange(B)
    slack = 4e-3 * total  # accumulated fp32 drift of interior nodes
    assert np.all((hi >= i / B * total - slack)[ok]) and np.all((lo <= (i + 1) / B * total + slack)[ok])


def test_export_import_roundtrip(zoo):
    t, o, rng = pair(zoo, 1000, 1500, 5)
    tree = t.export()
    t2 = zoo.SumTree(1000)
    t2.load(tree, len(t), t.first)
    u = rng.random(64)
    a, b = t.sample(u), t2.sample(u)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
