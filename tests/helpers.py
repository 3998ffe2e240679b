"""Shared helpers of the parity tests: build a device learner from an oracle-style net + numpy params, compare."""
import numpy as np

from oracle import zoo_oracle as O


def nrel(a, b):
    """normwise relative error max|a-b| / max|b| (used for distributions and gradients)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def obs_shape_of(net):
    first = O.net_chains(net)[0]
    for l in first:
        if l[0] == "conv":
            return (l[1], 84, 84)
        if l[0] == "dense":
            return (l[1],)
    raise ValueError


def make_approx(zoo, net, params, optimizer=None, max_batch=32, max_q=0, precision="fp32", obs_shape=None):
    app = zoo.NeuralNetworkApproximator(zoo.from_oracle_net(net), optimizer, obs_shape or obs_shape_of(net), max_batch, max_q, precision)
    app.set_params(params)
    return app


def assert_grads_close(got, ref, tol, what=""):
    assert len(got) == len(ref)
    worst = 0.0
    for i, (g, r) in enumerate(zip(got, ref)):
        assert g.shape == r.shape, (what, i, g.shape, r.shape)
        e = nrel(g, r)
        worst = max(worst, e)
        assert e <= tol, "%s: grad tensor %d %s normwise rel err %.3e > %.1e" % (what, i, r.shape, e, tol)
    return worst


def _bf16_bits_to_f32(raw_u16):
    return (raw_u16.astype(np.uint32) << 16).view(np.float32)


def device_activations(zoo, app, net, rows, n_tau=0):
    """Post-activation outputs of every relu layer of the ONLINE net, as the last forward left them on the device, in the
    twin's evaluation order and logical layout (conv: [rows, C, H, W]; dense: [rows, n]; a feature axis that is the flatten
    of a conv output comes back in the reference's (c, h, w) order).  Uses the zoo_test_net_activation hook."""
    import ctypes as C
    lib = zoo.lib()
    esz = 2 if app.precision == "bf16" else 4
    out = []
    flat_chw = None  # (c, h, w) of the flatten space once a conv stack has been flattened
    kind = net["kind"]
    for ci, layers in enumerate(O.net_chains(net)):
        li = 0
        crow = rows * (n_tau if (kind == "iqn" and ci > 0) else 1)
        h, w = (app.obs_shape[1], app.obs_shape[2]) if len(app.obs_shape) == 3 else (1, 1)
        for l in layers:
            if l[0] == "conv":
                _, cin, cout, k, s, p, relu = l
                h, w = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
                shape, perm = (crow, h, w, cout), (0, 3, 1, 2)
                flat_chw = (cout, h, w)
            elif l[0] == "dense":
                _, nin, nout, relu = l
                shape, perm = (crow, nout), None
            else:
                continue
            if relu:
                n = int(np.prod(shape))
                isbf = C.c_int32(0)
                raw = np.empty(n * esz, np.uint8)
                st = lib.zoo_test_net_activation(app.h, ci, li, crow, raw.ctypes.data, raw.size, C.byref(isbf))
                assert st == 0, lib.zoo_last_error()
                assert bool(isbf.value) == (esz == 2)
                a = (_bf16_bits_to_f32(raw.view(np.uint16)) if esz == 2 else raw.view(np.float32)).reshape(shape)
                if perm:
                    a = np.ascontiguousarray(a.transpose(perm))
                elif kind == "iqn" and ci == 1 and flat_chw is not None and nout == int(np.prod(flat_chw)):
                    c_, h_, w_ = flat_chw  # phi's output axis lives in the flatten space: internal (h, w, c) -> reference (c, h, w)
                    a = np.ascontiguousarray(a.reshape(crow, h_, w_, c_).transpose(0, 3, 1, 2).reshape(crow, nout))
                out.append(a)
            li += 1
    return out


FULL_SIZE_CASES = {  # name -> (learner kind, per-GPU batch): the configurations bench.py times (BASELINE.json configs[1..4])
    "dqn_atari_b32": ("dqn", 32), "per_dqn_atari_b32": ("per_dqn", 32), "rainbow_atari_b32": ("rainbow", 32),
    "rainbow_atari_b1024": ("rainbow", 1024), "iqn_atari_b32": ("iqn", 32), "iqn_atari_b512": ("iqn", 512),
}


def full_size_parity(zoo, case, prec, seed=33):
    """Device vs the torch twin at a bench configuration, relu masks pinned.

    1. the device computes loss + gradients (compute_grads through the C ABI);
    2. the relu masks the device used are read back (zoo_test_net_activation);
    3. the twin (oracle/torch_twin.py, <= 1e-5 of the numpy oracle) is run with its own relus (`free`) and again with the
       device's masks in place of relu' / relu (`pinned`);
    returns a report: loss / per-sample / priority errors, per-tensor normwise gradient errors for both runs, and per relu layer
    the number of units whose branch differs from the twin's together with the largest |twin pre-activation| among them, relative
    to the layer's rms pre-activation."""
    import torch

    from oracle import torch_twin as T
    kind, B = FULL_SIZE_CASES[case]
    N = 64 if kind == "iqn" else 0
    A = 6
    net = O.atari_iqn_net(A) if kind == "iqn" else O.chain_net(O.nature_cnn(A * 51 if kind == "rainbow" else A))
    p, tp = O.init_params(net, 1, 0.05), O.init_params(net, 2, 0.05)
    per = kind in ("per_dqn", "rainbow")  # Dopamine Rainbow uses prioritized replay (Dopamine_Rainbow_Atari.jl:66-69)
    b = O.synth_batch(seed, B, (4, 84, 84), A, atari=True, priority=per)
    rng = np.random.default_rng(seed + 1)
    kw = {"tau": rng.random((B, N), dtype=np.float32), "tau_prime": rng.random((B, N), dtype=np.float32)} if kind == "iqn" else {}
    opt = zoo.RMSProp(0.00025, 0.95) if kind in ("dqn", "per_dqn") else zoo.ADAM(0.0000625)
    app = make_approx(zoo, net, p, opt, max_batch=B, max_q=N, precision=prec)
    tapp = make_approx(zoo, net, tp, max_batch=B, max_q=N, precision=prec)
    if kind == "dqn":
        lrn = zoo.DQNLearner(app, tapp, batch_size=B)
        twin = lambda **m: T.dqn_grads(net, p, tp, b, **m)
    elif kind == "per_dqn":
        lrn = zoo.PrioritizedDQNLearner(app, tapp, batch_size=B)
        twin = lambda **m: T.prioritized_dqn_grads(net, p, tp, b, **m)
    elif kind == "rainbow":
        lrn = zoo.RainbowLearner(app, tapp, n_actions=A, Vmax=10.0, Vmin=-10.0, update_horizon=3, batch_size=B)
        twin = lambda **m: T.rainbow_grads(net, p, tp, b, A, n=3, **m)
    else:
        lrn = zoo.IQNLearner(app, tapp, n_actions=A, N=N, N_prime=N, N_em=64, batch_size=B, update_horizon=3)
        twin = lambda **m: T.iqn_grads(net, p, tp, b, kw["tau"], kw["tau_prime"], **m)
    app.set_params(p)  # the constructor force-synced online := target
    loss, g, prio = lrn.compute_grads(b, **kw)
    per_dev = lrn.per_sample_loss()
    acts = device_activations(zoo, app, net, B, N)
    lrn.close()
    del lrn, app, tapp
    rec = []
    free = twin(rec=rec)
    if kind == "iqn":
        # the library fuses the IQN merge into phi's GEMM epilogue and never stores phi's own output: the hook returns
        # merged = psi .* relu(phi) for that layer, from which relu'(phi) = [merged > 0] wherever psi > 0.  Where psi == 0 the
        # unit's branch is unobservable AND irrelevant (its gradient is multiplied by psi): those take the twin's own branch.
        n_psi = sum(1 for l in net["psi"] if l[0] in ("conv", "dense") and l[-1])
        psi_flat = np.repeat(acts[n_psi - 1].reshape(B, -1), N, axis=0)
        dontcare = psi_flat == 0
        acts[n_psi] = np.where(dontcare, rec[n_psi].numpy(), acts[n_psi])
        del psi_flat, dontcare
    masks = [torch.from_numpy((a > 0).astype(np.float32)) for a in acts]
    assert len(masks) == len(rec)
    flips = []
    for a, pre in zip(acts, rec):
        pre = pre.numpy()
        assert pre.shape == a.shape, (pre.shape, a.shape)
        diff = (pre > 0) != (a > 0)
        scale = float(np.sqrt(np.mean(pre.astype(np.float64) ** 2))) + 1e-30
        flips.append({"units": int(a.size), "flipped": int(diff.sum()), "worst_rel_preact": float(np.abs(pre[diff]).max() / scale) if diff.any() else 0.0})
    del rec, acts
    pinned = twin(masks=masks)
    ref_loss = float(pinned["loss"])
    rep = {"case": case, "prec": prec, "B": B, "loss": float(loss), "twin_loss": ref_loss,
           "loss_rel": abs(float(loss) - ref_loss) / max(1.0, abs(ref_loss)),
           "grad_err_pinned": [nrel(x, y) for x, y in zip(g, pinned["grads"])],
           "grad_err_free": [nrel(x, y) for x, y in zip(g, free["grads"])],
           "flips": flips}
    if "per_sample" in pinned:
        rep["per_sample_err"] = nrel(per_dev, pinned["per_sample"])
    if "priorities" in pinned:
        rep["priority_err"] = nrel(prio, pinned["priorities"])
    return rep


def format_full_size_report(rep):
    fl = " ".join("%d/%d(%.1e)" % (f["flipped"], f["units"], f["worst_rel_preact"]) for f in rep["flips"])
    return ("%-20s %-5s loss %.6f twin %.6f (rel %.1e)%s\n    grads, masks pinned: %s\n    grads, twin's own relus: %s\n    relu units flipped/total(worst |pre|/rms): %s" % (
        rep["case"], rep["prec"], rep["loss"], rep["twin_loss"], rep["loss_rel"],
        "  per-sample %.1e" % rep["per_sample_err"] if "per_sample_err" in rep else "",
        " ".join("%.1e" % e for e in rep["grad_err_pinned"]), " ".join("%.1e" % e for e in rep["grad_err_free"]), fl))
