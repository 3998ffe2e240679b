"""Replay batch assembly (RLCore's NStepBatchSampler.fetch! behind common.jl:18).
CPU: the restatement oracle/replay.py on hand-checkable cases.  GPU (-m gpu): the device shard (csrc/replay.cu) is bit-exact
against it -- frames, actions, terminals, n-step rewards, uniform and prioritized indices -- and
``update!(learner, trajectory)`` from the shard equals ``update!(learner, batch)`` on the batch the oracle assembles."""
import numpy as np
import pytest

from helpers import make_approx, nrel
from oracle import zoo_oracle as O
from oracle.replay import ReplayOracle

F = np.float32


def fill(o, n, seed, shape=(8, 8), shard=None, p_term=0.1):
    rng = np.random.default_rng(seed)
    frames = rng.integers(0, 256, (n,) + shape, dtype=np.uint8)
    frames[:, 0, 0] = np.arange(n) % 251  # a tag that identifies each frame
    act = rng.integers(0, 6, n).astype(np.int32)
    rew = rng.integers(-1, 2, n).astype(F)
    term = (rng.random(n) < p_term).astype(np.uint8)
    pri = np.power(rng.random(n, dtype=F) + F(1e-10), F(0.5))
    for i in range(n):
        o.push(frames[i], act[i], rew[i], term[i], pri[i])
    if shard is not None:
        shard.push_many(frames, act, rew, term, pri)
    return frames, act, rew, term


def test_fetch_micro():
    o = ReplayOracle(16, (2, 2), stack_size=2, n_step=3, gamma=0.5)
    for i in range(1, 11):  # frames filled with i, reward i, terminal only at transition 6
        o.push(np.full((2, 2), i, np.uint8), i % 3, float(i), i == 6)
    b = o.fetch([3, 5, 6])
    assert b["state"][:, :, 0, 0].tolist() == [[2, 3], [4, 5], [5, 6]]       # frames[i-1 : i]
    assert b["next_state"][:, :, 0, 0].tolist() == [[5, 6], [7, 8], [8, 9]]  # frames[i+2 : i+3]
    assert b["action"].tolist() == [0, 2, 0]
    # i=3: no terminal in 3,4,5 -> 3 + 0.5*4 + 0.25*5 = 6.25;  i=5: terminal at k=2 (transition 6) -> 5 + 0.5*6 = 8;  i=6: k=1 -> 6
    assert b["reward"].tolist() == [6.25, 8.0, 6.0] and b["terminal"].tolist() == [0, 1, 1]
    assert o.valid_range() == (2, 7)
    assert o.indices(np.array([0.0, 0.5, 0.999999]), "uniform")[0].tolist() == [2, 5, 7]


def test_ring_wraps():
    o = ReplayOracle(8, (2, 2), stack_size=2, n_step=1, gamma=0.9)
    for i in range(1, 14):  # 13 pushes into 8 slots: logical 1..8 now hold pushes 6..13
        o.push(np.full((2, 2), i, np.uint8), 0, float(i), False)
    assert o.first == 6 and o.length == 8
    b = o.fetch([2, 7])
    assert b["state"][:, :, 0, 0].tolist() == [[6, 7], [11, 12]] and b["next_state"][:, :, 0, 0].tolist() == [[7, 8], [12, 13]]
    assert b["reward"].tolist() == [7.0, 12.0]


@pytest.mark.gpu
@pytest.mark.parametrize("cap,n_push,stack,n,dt", [(64, 40, 4, 1, np.float64), (64, 150, 4, 3, np.float32), (1000, 2500, 2, 5, np.float64)])
def test_device_fetch_is_bit_exact(zoo, cap, n_push, stack, n, dt):
    lrn = _tiny_learner(zoo, 16, stack)
    shard = zoo.ReplayShard(cap, (8, 8), stack, n, 0.99, max_batch=16)
    o = ReplayOracle(cap, (8, 8), stack, n, 0.99)
    fill(o, n_push, cap + n, shard=shard)
    assert len(shard) == o.length
    rng = np.random.default_rng(1)
    for _ in range(3):
        u = rng.random(16).astype(dt)
        b, inds = shard.sample_batch(lrn, u, want_inds=True)
        oi, _, _ = o.indices(u, "uniform")
        assert np.array_equal(inds, oi)
        ref = o.fetch(oi)
        got = _read_batch(b, 16, stack, (8, 8))
        for k in ref:
            assert np.array_equal(got[k], ref[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["independent", "stratified"])
def test_device_prioritized_sample_and_write_back(zoo, mode):
    cap, stack, n, B = 512, 4, 3, 32
    lrn = _tiny_learner(zoo, B, stack, per=True)
    shard = zoo.ReplayShard(cap, (8, 8), stack, n, 0.99, max_batch=B, prioritized=True)
    o = ReplayOracle(cap, (8, 8), stack, n, 0.99, prioritized=True)
    fill(o, 700, 3, shard=shard)
    from rlzoo_b200 import _lib as ZL
    rng = np.random.default_rng(2)
    u = rng.random(B)
    b, inds = shard.sample_batch(lrn, u, mode=ZL.SAMPLE_STRATIFIED if mode == "stratified" else ZL.SAMPLE_INDEPENDENT, want_inds=True)
    oi, opr, nbad = o.indices(u, mode)
    assert np.array_equal(inds, oi) and shard.invalid_count() == nbad
    got = _read_batch(b, B, stack, (8, 8), priority=True)
    ref = o.fetch(oi)
    for k in ref:
        assert np.array_equal(got[k], ref[k]), k
    assert np.array_equal(got["priority"], opr)
    # update!(learner, t): the new priorities land in the tree exactly as t[:priority][inds] .= p would put them (common.jl:22)
    shard.update(lrn, u, mode=ZL.SAMPLE_STRATIFIED if mode == "stratified" else ZL.SAMPLE_INDEPENDENT)
    newp = _read_dev(lrn, B)
    o.tree.update(oi, newp)
    assert np.array_equal(shard.priority_tree_export(), o.tree.export())


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_device_prioritized_redraw_matches_the_oracle(zoo, dt):
    """RLCore redraws prioritized indices outside stack .. length - n (call site common.jl:18).  Early in training most leaves
    still hold default_priority 100, the newest n of them are invalid, so redraws are frequent: here 40 % of the mass sits on
    the invalid ends.  The device redraw loop consumes the host's spare uniforms exactly as the oracle does; with too few spares
    the leftover is clamped, counted, and weighted with the priority of the transition actually fetched."""
    cap, stack, n, B = 256, 4, 3, 32
    lrn = _tiny_learner(zoo, B, stack, per=True)
    shard = zoo.ReplayShard(cap, (8, 8), stack, n, 0.99, max_batch=B, prioritized=True)
    o = ReplayOracle(cap, (8, 8), stack, n, 0.99, prioritized=True)
    rng = np.random.default_rng(11)
    frames = rng.integers(0, 256, (200, 8, 8), dtype=np.uint8)
    act, rew, term = rng.integers(0, 6, 200).astype(np.int32), rng.integers(-1, 2, 200).astype(F), (rng.random(200) < 0.1).astype(np.uint8)
    pri = (rng.random(200, dtype=F) + F(0.1)).astype(F)
    pri[:3] = F(20.0)   # indices 1..3 < stack: invalid
    pri[-3:] = F(20.0)  # indices > length - n: invalid
    for i in range(200):
        o.push(frames[i], act[i], rew[i], term[i], pri[i])
    shard.push_many(frames, act, rew, term, pri)
    from rlzoo_b200 import _lib as ZL
    u = rng.random(B).astype(dt)
    for R in (8, 1, 0):
        spare = rng.random((B, max(R, 1))).astype(dt)[:, :R] if R else None
        before = shard.invalid_count()
        if spare is None:
            shard.set_spare_draws(lrn, None)
        b, inds = shard.sample_batch(lrn, u, mode=ZL.SAMPLE_INDEPENDENT, want_inds=True, spare=spare)
        oi, opr, nbad = o.indices(u, "independent", spare=spare)
        assert np.array_equal(inds, oi), R
        assert shard.invalid_count() - before == nbad
        got = _read_batch(b, B, stack, (8, 8), priority=True)
        assert np.array_equal(got["priority"], opr), R
        lo, hi = o.valid_range()
        assert inds.min() >= lo and inds.max() <= hi
        if R == 8:
            assert nbad == 0  # 0.4^9 per sample: the redraws resolve every index
        if R == 0:
            assert nbad > 0   # and without spares the old clamp-and-count behaviour remains


@pytest.mark.gpu
def test_update_from_shard_equals_update_from_batch(zoo):
    """sample -> update on the device == the oracle's batch fed through update!(learner, batch)."""
    cap, stack, n, B = 256, 4, 1, 8
    net = O.chain_net(O.nature_cnn(6))
    p, tp = O.init_params(net, 1, 0.05), O.init_params(net, 2, 0.05)
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (300, 84, 84), dtype=np.uint8)
    act, rew, term = rng.integers(0, 6, 300).astype(np.int32), rng.integers(-1, 2, 300).astype(F), (rng.random(300) < 0.05).astype(np.uint8)
    o = ReplayOracle(cap, (84, 84), stack, n, 0.99)
    for i in range(300):
        o.push(frames[i], act[i], rew[i], term[i])
    u = rng.random(B)
    res = []
    for via_shard in (False, True):
        app = make_approx(zoo, net, p, zoo.RMSProp(0.00025, 0.95), max_batch=B)
        tapp = make_approx(zoo, net, tp, max_batch=B)
        lrn = zoo.DQNLearner(app, tapp, batch_size=B)
        app.set_params(p)
        if via_shard:
            shard = zoo.ReplayShard(cap, (84, 84), stack, n, 0.99, max_batch=B)
            shard.push_many(frames, act, rew, term)
            loss = shard.update(lrn, u)
        else:
            lrn.update(o.fetch(o.indices(u, "uniform")[0]))
            loss = lrn.loss
        res.append((float(loss), app.params()))
    assert abs(res[0][0] - res[1][0]) <= 1e-5 * max(1.0, abs(res[0][0]))
    for a, r in zip(res[1][1], res[0][1]):
        assert nrel(a, r) <= 1e-3  # split-K atomics reorder fp32 sums between runs


# ---- helpers --------------------------------------------------------------------------------------
def _tiny_learner(zoo, B, stack, per=False):
    net = O.chain_net([("scale255",), ("conv", stack, 8, 4, 2, 0, True), ("flatten",), ("dense", 8 * 3 * 3, 6, False)])
    p = O.init_params(net, 1, 0.1)
    app = zoo.NeuralNetworkApproximator(zoo.from_oracle_net(net), zoo.ADAM(), (stack, 8, 8), B, 0, "fp32")
    tapp = zoo.NeuralNetworkApproximator(zoo.from_oracle_net(net), None, (stack, 8, 8), B, 0, "fp32")
    app.set_params(p)
    tapp.set_params(p)
    return zoo.PrioritizedDQNLearner(app, tapp, batch_size=B) if per else zoo.DQNLearner(app, tapp, batch_size=B)


def _read_batch(b, B, stack, shape, priority=False):
    import torch
    torch.cuda.synchronize()

    def dev(ptr, n, dtype):
        out = np.empty(n, dtype)
        import ctypes
        cudart = ctypes.CDLL("libcudart.so.12")
        rc = cudart.cudaMemcpy(ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(ptr), ctypes.c_size_t(out.nbytes), 2)
        assert rc == 0
        return out
    n_obs = B * stack * shape[0] * shape[1]
    got = {"state": dev(b.state, n_obs, np.uint8).reshape((B, stack) + shape), "next_state": dev(b.next_state, n_obs, np.uint8).reshape((B, stack) + shape),
           "action": dev(b.action, B, np.int32), "reward": dev(b.reward, B, np.float32), "terminal": dev(b.terminal, B, np.uint8)}
    if priority:
        got["priority"] = dev(b.priority, B, np.float32)
    return got


def _read_dev(lrn, B):
    import ctypes

    import zoo_loader
    zoo = zoo_loader.load()
    lrn.sync()
    p = ctypes.c_void_p()
    assert zoo.lib().zoo_learner_priorities_dev(lrn.h, ctypes.byref(p)) == 0
    out = np.empty(B, np.float32)
    cudart = ctypes.CDLL("libcudart.so.12")
    assert cudart.cudaMemcpy(ctypes.c_void_p(out.ctypes.data), p, ctypes.c_size_t(out.nbytes), 2) == 0
    return out
