"""CPU: independent third witnesses for the arithmetic the oracle restates from UN-VENDORED dependencies (SURVEY.md §8c:
Flux 0.11 optimisers / losses, RLCore SumTree -- all [RECALLED], nothing on disk under /root/reference pins them).
Each test compares the oracle with a separately maintained implementation of the same published definition:
  * torch.optim.RMSprop / torch.optim.Adam (eps outside the square root, like Flux)  vs  oracle.rmsprop_step / adam_step
  * torch.nn.functional.huber_loss / cross_entropy (soft targets) + autograd       vs  oracle.huber_per_sample / logit-CE adjoint
  * the C51 paper's floor/ceil projection (Bellemare et al. 2017, Alg. 1)            vs  oracle.project_distribution (rainbow.jl:204-221)
  * inverse-CDF sampling on an exact cumulative sum (np.searchsorted)               vs  the SumTree descent (call site common.jl:18)
  * the direct n-step return  sum_k gamma^k r_k  in float64                         vs  the replay oracle's fp32 Horner fold
What stays single-sourced is listed in DESIGN.md section 2."""
import numpy as np
import torch
import torch.nn.functional as Fn

from oracle import zoo_oracle as O
from oracle.sumtree import CSumTree, PySumTree


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def test_rmsprop_matches_torch_optim_over_10_steps():
    rng = np.random.default_rng(0)
    shapes = [(32, 4, 8, 8), (512, 64), (512,)]
    p0 = [(rng.standard_normal(s) * 0.1).astype(np.float32) for s in shapes]
    tp = [torch.tensor(p.copy(), requires_grad=True) for p in p0]
    opt = torch.optim.RMSprop(tp, lr=0.00025, alpha=0.95, eps=1e-8, momentum=0, centered=False)  # Dopamine_DQN_Atari.jl:42
    p, st = [x.copy() for x in p0], O.rmsprop_init(p0)
    for step in range(10):
        g = [(rng.standard_normal(s) * (10.0 ** rng.integers(-4, 1))).astype(np.float32) for s in shapes]
        for t, gi in zip(tp, g):
            t.grad = torch.tensor(gi.copy())
        opt.step()
        p = O.rmsprop_step(p, g, st, 0.00025, 0.95)
        for a, b in zip(p, tp):
            assert _rel(a, b.detach().numpy()) <= 2e-7
    # the update itself (not just the parameter, which hides it behind its own magnitude)
    for a, b, q in zip(p, tp, p0):
        assert _rel(a - q, b.detach().numpy() - q) <= 1e-5


def test_adam_matches_torch_optim_over_10_steps():
    rng = np.random.default_rng(1)
    shapes = [(64, 32, 4, 4), (7744, 16), (16,)]
    for eta, beta in ((6.25e-5, (0.9, 0.999)), (1e-3, (0.9, 0.999)), (5e-5, (0.8, 0.99))):
        p0 = [(rng.standard_normal(s) * 0.1).astype(np.float32) for s in shapes]
        tp = [torch.tensor(p.copy(), requires_grad=True) for p in p0]
        opt = torch.optim.Adam(tp, lr=eta, betas=beta, eps=1e-8)  # Dopamine_Rainbow_Atari.jl:41
        p, st = [x.copy() for x in p0], O.adam_init(p0, beta)
        for step in range(10):
            g = [(rng.standard_normal(s) * (10.0 ** rng.integers(-4, 1))).astype(np.float32) for s in shapes]
            for t, gi in zip(tp, g):
                t.grad = torch.tensor(gi.copy())
            opt.step()
            p = O.adam_step(p, g, st, eta, beta)
        for a, b, q in zip(p, tp, p0):
            assert _rel(a, b.detach().numpy()) <= 2e-7
            assert _rel(a - q, b.detach().numpy() - q) <= 1e-4  # difference of two fp32 parameters ten steps apart


def test_huber_matches_torch_functional_and_autograd():
    rng = np.random.default_rng(2)
    G = (rng.standard_normal(4096) * 2).astype(np.float32)
    q = (rng.standard_normal(4096) * 2).astype(np.float32)
    G[:4], q[:4] = [0.5, 3.0, 1.0, -1.0], [0.0, 0.0, 0.0, 0.0]  # incl. |e| == delta exactly (strict <: linear branch)
    per, dq, dG = O.huber_per_sample(G, q)
    tq = torch.tensor(q, requires_grad=True)
    tl = Fn.huber_loss(tq, torch.tensor(G), reduction="none", delta=1.0)
    assert _rel(per, tl.detach().numpy()) <= 1e-6
    tl.sum().backward()
    assert _rel(dq, tq.grad.numpy()) <= 1e-6
    assert np.allclose(per[:2], [0.125, 2.5]) and np.allclose(dq[:2], [-0.5, -1.0])  # SURVEY.md A.8 (per-sample, before the mean)


def test_logit_cross_entropy_matches_torch_soft_target_cross_entropy():
    rng = np.random.default_rng(3)
    B, n = 64, 51
    logits = (rng.standard_normal((B, n)) * 3).astype(np.float32)
    y = rng.random((B, n)).astype(np.float32)
    y /= y.sum(1, keepdims=True)
    per = -(y * O.logsoftmax(logits, axis=1)).sum(1)  # Flux logitcrossentropy(ŷ, y; agg = identity), Dopamine_Rainbow_Atari.jl:54
    tl = torch.tensor(logits, requires_grad=True)
    ce = Fn.cross_entropy(tl, torch.tensor(y), reduction="none")
    assert _rel(per, ce.detach().numpy()) <= 2e-6
    ce.sum().backward()
    d = O.softmax(logits, axis=1) * y.sum(1, keepdims=True) - y  # the general adjoint (sum(y) != 1 exactly)
    assert _rel(d, tl.grad.numpy()) <= 2e-6


def _c51_floor_ceil(Tz, w, z, dz, vmin, vmax):
    """Bellemare, Dabney, Munos 2017, Algorithm 1 (the scatter form Dopamine and every other C51 implementation uses)."""
    B, n = Tz.shape
    out = np.zeros((B, n), np.float64)
    for b in range(B):
        for i in range(n):
            t = min(max(float(Tz[b, i]), vmin), vmax)
            pos = min(max((t - float(z[0])) / dz, 0.0), n - 1.0)
            l, u = int(np.floor(pos)), int(np.ceil(pos))
            if l == u:
                out[b, l] += float(w[b, i])
            else:
                out[b, l] += float(w[b, i]) * (u - pos)
                out[b, u] += float(w[b, i]) * (pos - l)
    return out


def test_c51_dense_projection_matches_the_papers_floor_ceil_scatter():
    rng = np.random.default_rng(4)
    n, vmax = 51, 10.0
    z = np.linspace(-vmax, vmax, n).astype(np.float32)
    dz = float(z[1] - z[0])
    B = 24
    r = rng.integers(-1, 2, B).astype(np.float32)
    disc = (0.99 ** 3) * (rng.random(B) > 0.1)
    Tz = (r[:, None] + disc[:, None] * z[None, :]).astype(np.float32)
    w = rng.random((B, n)).astype(np.float32)
    w /= w.sum(1, keepdims=True)
    got = O.project_distribution(Tz, w, z, np.float32(dz), -vmax, vmax)
    ref = _c51_floor_ceil(Tz, w, z, dz, -vmax, vmax)
    assert np.abs(got - ref).max() <= 1e-5  # fp32: |c - z_j| near 10 carries 1e-6, divided by dz = 0.4
    assert np.abs(got.sum(1) - 1).max() <= 1e-5


def test_sumtree_descent_is_inverse_cdf_sampling_when_sums_are_exact():
    """Small-integer priorities make every fp32 partial sum exact, so the tree descent (v <= left ? left : v -= left, right) must
    return exactly the index np.searchsorted finds on the cumulative sum: first leaf whose prefix sum is >= v."""
    rng = np.random.default_rng(5)
    for cap, pushes in ((8, 16), (1000, 1000), (1000, 1777), (4096, 5000)):
        pr = rng.integers(1, 9, pushes).astype(np.float32)
        for Tree in (CSumTree, PySumTree):
            if Tree is PySumTree and cap > 1000:
                continue
            t = Tree(cap)
            for x in pr:
                t.push(float(x))
            # the descent walks the PHYSICAL slots (ring storage); push i lands in slot i % cap, `first` is the oldest one
            phys = np.zeros(cap, np.float32)
            for i, x in enumerate(pr):
                phys[i % cap] = x
            first = (pushes % cap) + 1 if pushes > cap else 1
            cs = np.cumsum(phys.astype(np.float64))
            assert float(t.total) == cs[-1]
            n = 64
            u = rng.random(n)
            idx, pv = t.sample(u, stratified=True)
            v = (np.arange(n) + u) / n * cs[-1]
            k = np.searchsorted(cs, v, side="left") + 1  # 1-based physical slot: first prefix sum >= v
            want = np.where(k >= first, k - first + 1, k + cap - first + 1)  # logical index, as RLCore returns it
            assert np.array_equal(np.asarray(idx), want)
            assert np.array_equal(np.asarray(pv, np.float32), phys[k - 1])


def test_nstep_fold_matches_the_direct_discounted_sum():
    from oracle.replay import ReplayOracle
    rng = np.random.default_rng(6)
    n, gamma = 3, 0.99
    R = ReplayOracle(64, frame_shape=(4, 4), stack_size=4, n_step=n, gamma=gamma)
    rew, term = rng.standard_normal(60).astype(np.float32), (rng.random(60) < 0.15)
    for i in range(60):
        R.push(rng.integers(0, 256, (4, 4), dtype=np.uint8), int(rng.integers(0, 6)), float(rew[i]), bool(term[i]))
    lo, hi = R.valid_range()
    inds = np.arange(lo, hi + 1)
    b = R.fetch(inds)
    for j, ind in enumerate(inds):
        acc, t = 0.0, False
        for k in range(n):  # sum_{k<m} gamma^k r_{ind+k}, truncated at the first terminal
            acc += gamma ** k * float(rew[ind - 1 + k])
            if term[ind - 1 + k]:
                t = True
                break
        assert abs(float(b["reward"][j]) - acc) <= 1e-5 * max(1.0, abs(acc)), (ind, b["reward"][j], acc)
        assert bool(b["terminal"][j]) == t
