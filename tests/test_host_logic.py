"""CPU: host-side logic above the C ABI -- the trajectory-level update! control flow (common.jl:7-26), batch sharding,
and (world_size 2, gloo) the data-parallel arithmetic: shard gradients scaled by 1/(B*world) and summed == the
single-process gradient on the concatenated batch, with the PER max(w) taken over the global batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import zoo_oracle as O


class FakeApprox:
    def __init__(self):
        self.copies = 0

    def copyto(self, src):
        self.copies += 1


class FakeLearner:
    def __init__(self, **kw):
        self.update_horizon, self.min_replay_history, self.update_freq, self.target_update_freq, self.update_step = 3, 10, 4, 8, 0
        self.approximator, self.target_approximator = FakeApprox(), FakeApprox()
        self.updates = 0
        self.__dict__.update(kw)

    def update(self, batch):
        self.updates += 1
        return np.full(4, 0.5, np.float32) if "priority" in batch else None


class FakeTraj:
    def __init__(self, n_terminal, prioritized):
        self.n_terminal, self.is_prioritized = n_terminal, prioritized
        self.written = []

    def sample(self, learner):
        b = {"state": None}
        if self.is_prioritized:
            b["priority"] = np.ones(4, np.float32)
        return np.arange(1, 5), b

    def set_priority(self, inds, p):
        self.written.append((inds.copy(), p.copy()))


def test_update_trajectory_gating(zoo):
    lrn = FakeLearner()
    # common.jl:8: length(t[:terminal]) - n <= min_replay_history -> nothing happens, update_step untouched
    t = FakeTraj(13, False)
    assert zoo.update_trajectory(lrn, t) is None and lrn.update_step == 0 and lrn.updates == 0
    t = FakeTraj(14, False)
    for _ in range(16):
        zoo.update_trajectory(lrn, t)
    assert lrn.update_step == 16
    assert lrn.updates == 4  # every update_freq = 4 steps (common.jl:16)
    assert lrn.target_approximator.copies == 2  # every target_update_freq = 8 steps (common.jl:12-14)


def test_update_trajectory_priority_write_back(zoo):
    lrn = FakeLearner(update_freq=1)
    t = FakeTraj(100, True)
    p = zoo.update_trajectory(lrn, t)
    assert np.array_equal(p, np.full(4, 0.5, np.float32))
    assert len(t.written) == 1 and np.array_equal(t.written[0][0], np.arange(1, 5))  # common.jl:22


def test_shard_bounds(zoo):
    assert [zoo.shard_bounds(4096, r, 8) for r in (0, 7)] == [(0, 512), (3584, 4096)]
    with pytest.raises(ValueError):
        zoo.shard_bounds(30, 0, 4)
    b = O.synth_batch(0, 8, (4,), 2, atari=False, priority=True)
    parts = [zoo.shard_batch(b, r, 2) for r in range(2)]
    for k in b:
        assert np.array_equal(np.concatenate([p[k] for p in parts]), b[k])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import zoo_loader
    zoo = zoo_loader.load()
    net = O.chain_net(O.mlp([4, 32, 2]))
    p, tp = O.init_params(net, 1, 0.1), O.init_params(net, 2, 0.1)
    full = O.synth_batch(5, 16, (4,), 2, atari=False, priority=True)
    mine = zoo.shard_batch(full, rank, world)
    # what libzoo_sm100 does per rank (learner.cu enqueue_grads): w_raw, all-reduce(max), loss scale 1/(B*world), all-reduce(sum)
    w_raw = (np.float32(1) / np.power(mine["priority"] + np.float32(1e-10), np.float32(0.5))).astype(np.float32)
    wmax = torch.tensor([w_raw.max()])
    dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
    w = w_raw / wmax.item()
    q, c = O.net_forward(net, p, mine["state"])
    G = O.dqn_target(net, p, tp, mine, 0.99, 1, False)
    B = len(mine["action"])
    per, dq, _ = O.huber_per_sample(G, q[np.arange(B), mine["action"]])
    scale = w / np.float32(B * world)
    d = np.zeros_like(q)
    d[np.arange(B), mine["action"]] = dq * scale
    grads = O.net_backward(net, p, c, d)
    flat = torch.from_numpy(np.concatenate([g.ravel() for g in grads] + [np.array([np.dot(scale, per)], np.float32)]))
    dist.all_reduce(flat)  # gradients + loss ride in one buffer, as in the library
    if rank == 0:
        ref = O.prioritized_dqn_grads(net, p, tp, full)
        ref_flat = np.concatenate([g.ravel() for g in ref["grads"]] + [np.array([ref["loss"]], np.float32)])
        out.put(float(np.abs(flat.numpy() - ref_flat).max() / np.abs(ref_flat).max()))
    dist.destroy_process_group()


def test_data_parallel_gradients_equal_single_process_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get() <= 1e-5


def test_staging_division_by_255_is_exact_for_every_uint8():
    """learner.cu:div255_u8 replaces the IEEE divide of `x ./ 255` (Dopamine_DQN_Atari.jl:28) by q0 = RN(v r), rem = v - 255 q0,
    q = RN(q0 + rem r), r = RN(1/255).  Checked here in exact rational arithmetic for all 256 frame values: the sequence returns the
    correctly rounded fp32 quotient (a plain multiply by r does not)."""
    from fractions import Fraction

    def rn32(fr):  # exact round-to-nearest-even of a non-negative rational to fp32 (normal range)
        if fr == 0:
            return Fraction(0)
        e = fr.numerator.bit_length() - fr.denominator.bit_length()
        while Fraction(2) ** e > fr:
            e -= 1
        while Fraction(2) ** (e + 1) <= fr:
            e += 1
        ulp = Fraction(2) ** (e - 23)
        m = fr / ulp
        fl = m.numerator // m.denominator
        rem = m - fl
        if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and fl % 2 == 1):
            fl += 1
        return fl * ulp

    r = rn32(Fraction(1, 255))
    assert float(r) == float(np.float32(1.0) / np.float32(255.0))
    plain_multiply_wrong = 0
    for v in range(256):
        exact = rn32(Fraction(v, 255))
        assert float(exact) == float(np.float32(v) / np.float32(255.0))
        q0 = rn32(v * r)
        rem = Fraction(v) - q0 * 255  # representable: the fma computes it without rounding
        assert rem == 0 or abs(rem) == rn32(abs(rem))
        q = rn32(q0 + rem * r) if q0 + rem * r >= 0 else None
        assert q == exact, v
        plain_multiply_wrong += q0 != exact
    assert plain_multiply_wrong > 0
