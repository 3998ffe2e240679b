import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import zoo_loader
from helpers import make_approx, nrel
from oracle import zoo_oracle as O
zoo = zoo_loader.load()
B = 8
net = O.chain_net(O.nature_cnn(6))
p, tp = O.init_params(net, 1, 0.05), O.init_params(net, 2, 0.05)
b = O.synth_batch(30, B, (4, 84, 84), 6, atari=True)
res = {}
for prec in ("fp32_simt", "fp32"):
    for mode in ("split", "fused", "fused2"):
        app = make_approx(zoo, net, p, zoo.RMSProp(0.00025, 0.95), max_batch=B, precision=prec)
        tapp = make_approx(zoo, net, tp, max_batch=B, precision=prec)
        lrn = zoo.DQNLearner(app, tapp, batch_size=B)
        app.set_params(p)
        if mode == "split":
            lrn.compute_grads(b); lrn.apply()
        else:
            lrn.update(b)
            if mode == "fused2":
                lrn.sync()
        res[(prec, mode)] = app.params()
    for mode in ("fused", "fused2"):
        print(prec, mode, ["%.1e" % nrel(a, r) for a, r in zip(res[(prec, mode)], res[(prec, "split")])], flush=True)
    print(prec, "moved vs init (split)", ["%.1e" % nrel(a, r) for a, r in zip(res[(prec, "split")], p)])
    print(prec, "moved vs init (fused)", ["%.1e" % nrel(a, r) for a, r in zip(res[(prec, "fused")], p)])
