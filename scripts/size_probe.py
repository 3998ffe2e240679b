"""A learner at a batch size the golden fixtures do not cover: device (per precision mode) vs the torch twin, per tensor.
usage: size_probe.py iqn|rainbow|dqn B [precision ...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import zoo_loader
from helpers import make_approx, nrel
from oracle import zoo_oracle as O
from oracle import torch_twin as T
zoo = zoo_loader.load()
kind = sys.argv[1]
B = int(sys.argv[2])
precs = sys.argv[3:] or ["fp32", "fp32_simt"]
N = 64 if kind == "iqn" else 0
net = O.atari_iqn_net(6) if kind == "iqn" else O.chain_net(O.nature_cnn(6 * 51 if kind == "rainbow" else 6))
p, tp = O.init_params(net, 1, 0.05), O.init_params(net, 2, 0.05)
b = O.synth_batch(33, B, (4, 84, 84), 6, atari=True)
rng = np.random.default_rng(3)
kw = {"tau": rng.random((B, N), dtype=np.float32), "tau_prime": rng.random((B, N), dtype=np.float32)} if kind == "iqn" else {}
t0 = time.time()
if kind == "iqn":
    ref = T.iqn_grads(net, p, tp, b, kw["tau"], kw["tau_prime"])
elif kind == "rainbow":
    ref = T.rainbow_grads(net, p, tp, b, 6, n=3)
else:
    ref = T.dqn_grads(net, p, tp, b)
print("twin: loss %.6f (%.1f s)" % (float(ref["loss"]), time.time() - t0), flush=True)
for prec in precs:
    app = make_approx(zoo, net, p, zoo.ADAM(0.00005), max_batch=B, max_q=N, precision=prec)
    tapp = make_approx(zoo, net, tp, max_batch=B, max_q=N, precision=prec)
    if kind == "iqn":
        lrn = zoo.IQNLearner(app, tapp, n_actions=6, N=N, N_prime=N, N_em=64, batch_size=B)
    elif kind == "rainbow":
        lrn = zoo.RainbowLearner(app, tapp, n_actions=6, Vmax=10.0, Vmin=-10.0, update_horizon=3, batch_size=B)
    else:
        lrn = zoo.DQNLearner(app, tapp, batch_size=B)
    app.set_params(p)
    l, g, _ = lrn.compute_grads(b, **kw)
    print("%-9s %s B=%d: loss %.6f (rel %.1e)  grad errs %s" % (prec, kind, B, l, abs(l - float(ref["loss"])) / abs(float(ref["loss"])),
          " ".join("%.1e" % nrel(x, y) for x, y in zip(g, ref["grads"]))), flush=True)
    if "per_sample" in ref:
        d = np.abs(lrn.per_sample_loss() - np.asarray(ref["per_sample"]))
        print("          per-sample loss: worst abs err %.2e; samples with err > 1e-4: %d" % (d.max(), int((d > 1e-4).sum())), flush=True)
    del lrn, app, tapp
