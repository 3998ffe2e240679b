"""Generates tests/golden/*.npz from the oracle (run from the repo root: ``python tests/golden/make_golden.py``).

The reference (pure Julia) cannot be executed or imported in this image and its own tests hold no numeric vectors for
this path, so these fixtures are ORACLE outputs, not reference outputs: parity stays "unpinned".  They
freeze the oracle (so a later edit cannot silently move the target the GPU is held to) and travel to the GPU box, where
``/root/reference`` does not exist.  Inputs are regenerated from the seeds stored in each case (``cases()`` below);
only outputs are stored: loss, per-sample losses, priorities, and per-tensor gradient digests (L2 norm, sum, the first
32 entries) -- full Atari gradients would be 16 MB per case."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import zoo_oracle as O  # noqa: E402
from oracle.sumtree import CSumTree  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    """name -> (net, learner kind, kwargs).  Mirrors BASELINE.json configs at oracle-friendly batch sizes."""
    mlp = O.chain_net(O.mlp([4, 128, 128, 2]))
    return {
        "basic_dqn_cartpole_b32": dict(net=mlp, kind="basic", B=32, atari=False, A=2, seeds=(1, 2, 3)),
        "dqn_cartpole_b32": dict(net=mlp, kind="dqn", B=32, atari=False, A=2, seeds=(1, 2, 4), n=1, double=True),
        "per_dqn_cartpole_b32": dict(net=mlp, kind="per", B=32, atari=False, A=2, seeds=(1, 2, 5), priority=True, mask=True),
        "rainbow_cartpole_b32": dict(net=O.chain_net(O.mlp([4, 128, 128, 2 * 51])), kind="rainbow", B=32, atari=False, A=2,
                                     seeds=(1, 2, 7), priority=True, vmax=200.0, vmin=0.0, n=1),
        "iqn_cartpole_b32": dict(net=O.cartpole_iqn_net(), kind="iqn", B=32, atari=False, A=2, seeds=(1, 2, 8), priority=True, N=8, Np=8),
        "dqn_atari_b8": dict(net=O.chain_net(O.nature_cnn(6)), kind="dqn", B=8, atari=True, A=6, seeds=(1, 2, 11), n=1, double=True, mask=True),
        "rainbow_atari_b8": dict(net=O.chain_net(O.nature_cnn(6 * 51)), kind="rainbow", B=8, atari=True, A=6, seeds=(1, 2, 12),
                                 priority=True, vmax=10.0, vmin=-10.0, n=3),
        "iqn_atari_b4": dict(net=O.atari_iqn_net(6), kind="iqn", B=4, atari=True, A=6, seeds=(1, 2, 13), N=64, Np=64),
        "qr_dqn_cartpole_b32": dict(net=O.chain_net(O.mlp([4, 128, 128, 2 * 10])), kind="qr", B=32, atari=False, A=2, seeds=(1, 2, 14), N=10),
        "rem_dqn_cartpole_b32": dict(net=O.chain_net(O.mlp([4, 128, 128, 2 * 16])), kind="rem", B=32, atari=False, A=2, seeds=(1, 2, 15), N=16,
                                     mask=True),
    }


def inputs(c):
    sp, st, sb = c["seeds"]
    bs = 0.05 if c["atari"] else 0.1
    p, tp = O.init_params(c["net"], sp, bs), O.init_params(c["net"], st, bs)
    shape = (4, 84, 84) if c["atari"] else (4,)
    b = O.synth_batch(sb, c["B"], shape, c["A"], atari=c["atari"], priority=c.get("priority", False), mask=c.get("mask", False))
    tau = taup = None
    if c["kind"] == "iqn":
        rng = np.random.default_rng(sb + 100)
        tau, taup = rng.random((c["B"], c["N"]), dtype=np.float32), rng.random((c["B"], c["Np"]), dtype=np.float32)
    if c["kind"] == "rem":  # the convex combination travels in the tau slot
        w = np.random.default_rng(sb + 100).random(c["N"], dtype=np.float32)
        tau = (w / w.sum(dtype=np.float32)).astype(np.float32)
    return p, tp, b, tau, taup


def run_oracle(c):
    p, tp, b, tau, taup = inputs(c)
    k = c["kind"]
    if k == "basic":
        return O.basic_dqn_grads(c["net"], p, b)
    if k == "dqn":
        return O.dqn_grads(c["net"], p, tp, b, 0.99, c["n"], c["double"])
    if k == "per":
        return O.prioritized_dqn_grads(c["net"], p, tp, b)
    if k == "rainbow":
        return O.rainbow_grads(c["net"], p, tp, b, c["A"], 51, c["vmax"], c["vmin"], 0.99, c["n"])
    if k == "qr":
        return O.qr_dqn_grads(c["net"], p, tp, b, c["A"], c["N"])
    if k == "rem":
        return O.rem_dqn_grads(c["net"], p, tp, b, c["A"], c["N"], tau)
    return O.iqn_grads(c["net"], p, tp, b, tau, taup)


def digest(r):
    out = {"loss": np.float32(r["loss"])}
    for k in ("per_sample", "priorities", "target_distribution"):
        if k in r:
            out[k] = np.asarray(r[k], np.float32)
    for i, g in enumerate(r["grads"]):
        g = np.asarray(g, np.float32)
        out["g%d_norm" % i] = np.float64(np.linalg.norm(g.astype(np.float64)))
        out["g%d_sum" % i] = np.float64(g.astype(np.float64).sum())
        out["g%d_head" % i] = g.reshape(-1)[:32].copy()
    out["n_tensors"] = np.int64(len(r["grads"]))
    return out


def sumtree_case(cap, n_push, B, seed):
    rng = np.random.default_rng(seed)
    pr = np.power(rng.random(n_push, dtype=np.float32) + np.float32(1e-10), np.float32(0.5)).astype(np.float32)
    pr[rng.random(n_push) < 0.01] = np.float32(100.0)
    t = CSumTree(cap)
    t.push_many(pr)
    out = {"cap": np.int64(cap), "n_push": np.int64(n_push), "B": np.int64(B), "seed": np.int64(seed), "total0": np.float32(t.total)}
    u64, u32 = rng.random(B), rng.random(B, dtype=np.float32)
    out["idx_strat_f64"], out["p_strat_f64"] = t.sample(u64, True)
    out["idx_indep_f64"], out["p_indep_f64"] = t.sample(u64, False)
    out["idx_strat_f32"], out["p_strat_f32"] = t.sample(u32, True)
    out["idx_indep_f32"], out["p_indep_f32"] = t.sample(u32, False)
    idx = out["idx_strat_f64"].copy()
    idx[B // 2] = idx[0]  # a duplicate: sequential semantics
    newp = rng.random(B, dtype=np.float32)
    t.update(idx, newp)
    tree = t.export()
    out["upd_idx"], out["upd_p"] = idx, newp
    out["total1"] = np.float32(t.total)
    out["tree_head"] = tree[:64].copy()
    out["tree_xor"] = np.uint32(np.bitwise_xor.reduce(tree.view(np.uint32)))
    out["tree_sum64"] = np.float64(tree.astype(np.float64).sum())
    return out


SUMTREE_CASES = {"sumtree_cap8": (8, 16, 4, 1), "sumtree_cap1000": (1000, 1300, 32, 2), "sumtree_cap100k": (100_000, 130_000, 4096, 3)}


def main():
    for name, c in cases().items():
        np.savez(os.path.join(HERE, name + ".npz"), **digest(run_oracle(c)))
        print("wrote", name)
    for name, a in SUMTREE_CASES.items():
        np.savez(os.path.join(HERE, name + ".npz"), **sumtree_case(*a))
        print("wrote", name)


if __name__ == "__main__":
    main()
