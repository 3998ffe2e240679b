"""Which batch sizes break 'loss(B) == mean of the losses of its B/base parts'?  (base sizes are the ones held to the oracle)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import zoo_loader
from helpers import make_approx, nrel
from oracle import zoo_oracle as O
zoo = zoo_loader.load()
which = sys.argv[1] if len(sys.argv) > 1 else "rainbow"
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
def part(b, lo, hi): return {k: (v[lo:hi] if v is not None else None) for k, v in b.items()}
if which == "rainbow":
    net = O.chain_net(O.nature_cnn(6 * 51)); base, sizes, N = 8, [8, 16, 32, 64, 128, 256, 1024], 0
else:
    net = O.atari_iqn_net(6); base, sizes, N = 4, [4, 8, 16, 32, 64], 64
p, tp = O.init_params(net, 1, 0.05), O.init_params(net, 2, 0.05)
def build(bs):
    app = make_approx(zoo, net, p, zoo.ADAM(0.0000625), max_batch=bs, max_q=N, precision=prec)
    tapp = make_approx(zoo, net, tp, max_batch=bs, max_q=N, precision=prec)
    if which == "rainbow":
        lrn = zoo.RainbowLearner(app, tapp, n_actions=6, Vmax=10.0, Vmin=-10.0, update_horizon=3, batch_size=bs)
    else:
        lrn = zoo.IQNLearner(app, tapp, n_actions=6, N=N, N_prime=N, N_em=64, batch_size=bs, update_horizon=3)
    app.set_params(p)
    return lrn
Bmax = sizes[-1]
b = O.synth_batch(32, Bmax, (4, 84, 84), 6, atari=True)
rng = np.random.default_rng(3)
kw_all = {} if which == "rainbow" else {"tau": rng.random((Bmax, N), dtype=np.float32), "tau_prime": rng.random((Bmax, N), dtype=np.float32)}
small = build(base)
base_losses, base_grads = [], []
for i in range(Bmax // base):
    kw = {k: v[i * base:(i + 1) * base] for k, v in kw_all.items()}
    l, g, _ = small.compute_grads(part(b, i * base, (i + 1) * base), **kw)
    base_losses.append(float(l)); base_grads.append(g)
    if i >= 127 and which == "rainbow": pass
for S in sizes:
    lrn = build(S)
    kw = {k: v[:S] for k, v in kw_all.items()}
    l, g, _ = lrn.compute_grads(part(b, 0, S), **kw)
    n = S // base
    lm = float(np.mean(base_losses[:n]))
    gm = [np.mean([base_grads[i][t] for i in range(n)], axis=0) for t in range(len(g))]
    errs = ["%.1e" % nrel(x, y) for x, y in zip(g, gm)]
    psl = lrn.per_sample_loss()
    print("%s %s B=%4d: loss %.6f vs mean of parts %.6f (rel %.2e)  grad errs %s  per-sample mean %.6f" % (which, prec, S, l, lm, abs(l - lm) / abs(lm), " ".join(errs), float(np.mean(psl))), flush=True)
    del lrn
