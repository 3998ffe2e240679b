"""Is the intermittent mid-K split parity failure (ZOO_TMA_CONV=0 ZOO_MIDK_SPLIT=1, test_rainbow_atari[fp32]) a relu' flip?
Repeats compute_grads in one process and compares every outcome with the oracle, tensor by tensor: a flip of one unit in conv1 /
conv2 leaves the tensors behind it untouched and moves the loss only at rounding level."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import zoo_loader
from helpers import make_approx, nrel
from oracle import zoo_oracle as O
from oracle import torch_twin as T
zoo = zoo_loader.load()
B = 8
net = O.chain_net(O.nature_cnn(6 * 51))
p, tp = O.init_params(net, 1, 0.05), O.init_params(net, 2, 0.05)
b = O.synth_batch(12, B, (4, 84, 84), 6, atari=True, priority=True)
ref = T.rainbow_grads(net, p, tp, b, 6, n=3)
app = make_approx(zoo, net, p, zoo.ADAM(0.0000625), max_batch=B)
tapp = make_approx(zoo, net, tp, max_batch=B)
lrn = zoo.RainbowLearner(app, tapp, n_actions=6, Vmax=10.0, Vmin=-10.0, update_horizon=3, batch_size=B)
app.set_params(p)
seen = {}
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
    l, g, _ = lrn.compute_grads(b)
    errs = tuple("%.1e" % nrel(x, y) for x, y in zip(g, ref["grads"]))
    key = errs[0][:3] + ("bad" if max(nrel(x, y) for x, y in zip(g, ref["grads"])) > 1e-4 else "ok")
    if key not in seen:
        seen[key] = 0
        print("outcome %-8s loss %.7f (ref %.7f)  grad errs %s" % (key, l, float(ref["loss"]), " ".join(errs)), flush=True)
    seen[key] += 1
print("counts:", seen)
