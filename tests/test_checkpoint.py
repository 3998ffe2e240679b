"""Checkpoints (reference: BSON.@save "policy.bson" policy, Dopamine_DQN_Atari.jl:123-126).
CPU: the BSON codec against bsonspec.org's two worked examples and round trips of the array lowering.
GPU (-m gpu): save -> load into a differently initialised learner restores parameters, optimiser state and step bit-exactly,
and the next update of the restored learner equals the next update of the original."""
import numpy as np
import pytest

from helpers import make_approx, nrel
from oracle import zoo_oracle as O


def ck():
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("zoo_ckpt", os.path.join(root, "reinforcementlearningzoo.jl_b200", "checkpoint.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)  # the codec itself imports nothing from the package (and needs no GPU)
    return m


def test_bsonspec_worked_examples():
    m = ck()
    assert m.bson_dumps({"hello": "world"}) == b"\x16\x00\x00\x00\x02hello\x00\x06\x00\x00\x00world\x00\x00"
    ref = (b"\x31\x00\x00\x00\x04BSON\x00\x26\x00\x00\x00\x020\x00\x08\x00\x00\x00awesome\x00\x011\x00\x33\x33\x33\x33\x33\x33\x14\x40"
           b"\x102\x00\xc2\x07\x00\x00\x00\x00")
    assert m.bson_dumps({"BSON": ["awesome", 5.05, 1986]}) == ref
    assert m.bson_loads(ref) == {"BSON": ["awesome", 5.05, 1986]}


def test_document_round_trip_with_arrays():
    m = ck()
    rng = np.random.default_rng(0)
    doc = {"policy": {"learner": {"type": "DQNLearner", "gamma": 0.99, "update_step": 2 ** 40 + 3, "flag": True, "none": None,
                                  "params": [rng.standard_normal((5, 3)).astype(np.float32), rng.standard_normal(5).astype(np.float32),
                                             rng.standard_normal((2, 3, 4, 4)).astype(np.float32)],
                                  "state": [[np.arange(6, dtype=np.float32).reshape(2, 3)], [np.zeros(4, np.float32)]]}}}
    back = m.bson_loads(m.bson_dumps(doc))
    lr, br = doc["policy"]["learner"], back["policy"]["learner"]
    assert br["type"] == "DQNLearner" and br["gamma"] == 0.99 and br["update_step"] == 2 ** 40 + 3 and br["flag"] is True and br["none"] is None
    for a, b in zip(lr["params"], br["params"]):
        assert a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b)
    assert np.array_equal(br["state"][0][0], lr["state"][0][0]) and np.array_equal(br["state"][1][0], lr["state"][1][0])


def test_arrays_carry_julia_dims_and_column_major_bytes():
    m = ck()
    W = np.arange(6, dtype=np.float32).reshape(2, 3)      # Dense(3, 2): Julia W is 2 x 3, column-major
    d = m.lower_array(W)
    assert d["tag"] == "array" and d["type"] == {"tag": "datatype", "params": [], "name": ["Core", "Float32"]} and d["size"] == [2, 3]
    assert np.frombuffer(d["data"], np.float32).tolist() == [0, 3, 1, 4, 2, 5]  # W[1,1], W[2,1], W[1,2], ...
    K = np.arange(2 * 3 * 4 * 5, dtype=np.float32).reshape(2, 3, 4, 5)  # conv OIHW == Julia (kw, kh, cin, cout) bytes
    d = m.lower_array(K)
    assert d["size"] == [5, 4, 3, 2] and d["data"] == K.tobytes()
    assert np.array_equal(m.raise_array(d), K) and np.array_equal(m.raise_array(m.lower_array(W)), W)


def test_corrupt_documents_are_rejected():
    m = ck()
    good = m.bson_dumps({"a": 1})
    with pytest.raises(ValueError):
        m.bson_loads(good + b"\x00")
    with pytest.raises(ValueError):
        m.bson_loads(good[:-1] + b"\x01")


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["dqn_rmsprop", "rainbow_adam"])
def test_save_load_restores_the_learner(zoo, tmp_path, kind):
    B = 8
    if kind == "dqn_rmsprop":
        net = O.chain_net(O.nature_cnn(6))
        opt = lambda: zoo.RMSProp(0.00025, 0.95)
        mk = lambda a, t: zoo.DQNLearner(a, t, batch_size=B)
    else:
        net = O.chain_net(O.mlp([4, 64, 64, 2 * 51]))
        opt = lambda: zoo.ADAM(0.001)
        mk = lambda a, t: zoo.RainbowLearner(a, t, n_actions=2, Vmax=200.0, Vmin=0.0, batch_size=B)
    atari = kind == "dqn_rmsprop"
    shape = (4, 84, 84) if atari else (4,)

    def build(seed):
        app = make_approx(zoo, net, O.init_params(net, seed, 0.05), opt(), max_batch=B)
        tapp = make_approx(zoo, net, O.init_params(net, seed + 1, 0.05), max_batch=B)
        return mk(app, tapp)

    batches = [O.synth_batch(20 + i, B, shape, 6 if atari else 2, atari=atari, priority=not atari) for i in range(4)]
    a = build(1)
    for b in batches[:3]:
        a.update(b)
    path = str(tmp_path / "policy.bson")
    zoo.save_policy(path, a)
    r = build(7)  # different weights, fresh optimiser
    zoo.load_policy(path, r)
    for x, y in zip(a.approximator.params() + a.target_approximator.params(), r.approximator.params() + r.target_approximator.params()):
        assert np.array_equal(x, y)
    slots = 1 if kind == "dqn_rmsprop" else 2
    for s in range(slots):
        for x, y in zip(a.approximator.opt_state(s), r.approximator.opt_state(s)):
            assert np.array_equal(x, y)
    a.update(batches[3])
    r.update(batches[3])
    assert abs(float(a.loss) - float(r.loss)) <= 1e-5 * max(1.0, abs(float(a.loss)))
    for x, y in zip(a.approximator.params(), r.approximator.params()):
        assert nrel(y, x) <= 1e-4  # split-K atomics reorder fp32 sums between runs; Adam's bias correction needs the restored step
