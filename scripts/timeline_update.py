"""Concurrent-lane timeline of one eager learner update (zoo_prof_timeline): which stream each kernel ran on, when."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, zoo_loader
from oracle import zoo_oracle as O
import torch
wl = sys.argv[1] if len(sys.argv) > 1 else "dqn_atari_b32"
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
w = bench.WORKLOADS[wl]
zoo = zoo_loader.load(); lib = zoo.lib()
from rlzoo_b200 import _lib as ZL
net = bench.oracle_net(w); model = zoo.from_oracle_net(net)
opt = zoo.RMSProp(*w["opt"][1:]) if w["opt"][0] == "rmsprop" else zoo.ADAM(w["opt"][1])
B = w["B"]
app = zoo.NeuralNetworkApproximator(model, opt, (4, 84, 84), B, w.get("N", 0), prec)
tapp = zoo.NeuralNetworkApproximator(model, None, (4, 84, 84), B, w.get("N", 0), prec)
app.set_params(O.init_params(net, 1)); tapp.set_params(O.init_params(net, 2))
lrn = zoo.DQNLearner(app, tapp, batch_size=B, update_horizon=w["n"])
app.set_params(O.init_params(net, 1))
b = bench.synth(w, 0)
db = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in b.items()}
for _ in range(3):
    lrn.update(db, want_loss=False)
lrn.sync()
ZL.check(lib.zoo_prof_timeline(1, None))
lrn.update(db, want_loss=False)
n = C.c_int32()
ZL.check(lib.zoo_prof_timeline(0, C.byref(n)))
name, st, a, e = C.create_string_buffer(256), C.c_uint64(), C.c_float(), C.c_float()
streams = {}
rows = []
for i in range(n.value):
    ZL.check(lib.zoo_prof_timeline_get(i, name, 256, C.byref(st), C.byref(a), C.byref(e)))
    lane = streams.setdefault(st.value, len(streams))
    rows.append((a.value, e.value, lane, name.value.decode()))
print("# lane 0 = first stream seen (main); start/end us since the first launch")
for a_, e_, lane, nm in sorted(rows):
    print("%8.1f %8.1f  %6.1f  lane%d  %s" % (a_, e_, e_ - a_, lane, nm))
