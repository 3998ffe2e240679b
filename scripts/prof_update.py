"""Per-launch device times of one learner update (library's in-situ profiler), for a bench workload.
usage: python scripts/prof_update.py [workload] [precision] [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import zoo_loader  # noqa: E402
from oracle import zoo_oracle as O  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "dqn_atari_b32"
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
w = bench.WORKLOADS[wl]
import torch  # noqa: E402
zoo = zoo_loader.load()
net = bench.oracle_net(w)
model = zoo.from_oracle_net(net)
opt = zoo.RMSProp(*w["opt"][1:]) if w["opt"][0] == "rmsprop" else zoo.ADAM(w["opt"][1])
B, kind, nq = w["B"], w["kind"], w.get("N", 0)
shape = (4, 84, 84) if w["atari"] else (4,)
app = zoo.NeuralNetworkApproximator(model, opt, shape, B, nq, prec)
tapp = None if kind == "basic" else zoo.NeuralNetworkApproximator(model, None, shape, B, nq, prec)
app.set_params(O.init_params(net, 1))
if tapp is not None:
    tapp.set_params(O.init_params(net, 2))
if kind == "basic":
    lrn = zoo.BasicDQNLearner(app, batch_size=B)
elif kind == "dqn":
    lrn = zoo.DQNLearner(app, tapp, batch_size=B, update_horizon=w["n"])
elif kind == "rainbow":
    lrn = zoo.RainbowLearner(app, tapp, n_actions=w["A"], Vmax=10.0, Vmin=-10.0, update_horizon=w["n"], batch_size=B)
else:
    lrn = zoo.IQNLearner(app, tapp, n_actions=w["A"], N=nq, N_prime=nq, N_em=64, batch_size=B, update_horizon=w["n"])
app.set_params(O.init_params(net, 1))
b = bench.synth(w, 0)
db = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in b.items()}
for _ in range(3):
    lrn.update(db, want_loss=False, want_priorities=False)
lrn.sync()
acc = None
for r in range(reps):
    pr = lrn.profile(db, delay_us=3000 if B <= 64 else 30000)
    if r == 0:
        continue
    if acc is None:
        acc = [[lab, ms] for lab, ms in pr]
    else:
        for a, (lab, ms) in zip(acc, pr):
            a[1] += ms
tot = 0.0
print("# %s %s: per-launch device time (us), mean of %d profiled eager updates" % (wl, prec, reps - 1))
for lab, ms in acc:
    us = ms / (reps - 1) * 1e3
    tot += us
    print("%9.2f  %s" % (us, lab))
print("%9.2f  TOTAL (%d launches)" % (tot, len(acc)))
