"""Phase stamps of CTA (0,0,0) of EVERY tcgen05-engine launch of one graph-replayed learner update (zoo_test_trace_*): the trace
is armed before the update's CUDA graph is captured, so the replays keep stamping; times are us since the first stamp of the step.
usage: python scripts/trace_update.py [workload] [precision]"""
import ctypes as C
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
w = bench.WORKLOADS[wl]
import torch  # noqa: E402
zoo = zoo_loader.load()
lib = zoo.lib()
net = bench.oracle_net(w)
model = zoo.from_oracle_net(net)
opt = zoo.RMSProp(*w["opt"][1:]) if w["opt"][0] == "rmsprop" else zoo.ADAM(w["opt"][1])
B, kind, nq = w["B"], w["kind"], w.get("N", 0)
shape = (4, 84, 84) if w["atari"] else (4,)
app = zoo.NeuralNetworkApproximator(model, opt, shape, B, nq, prec)
tapp = None if kind == "basic" else zoo.NeuralNetworkApproximator(model, None, shape, B, nq, prec)
if kind == "basic":
    lrn = zoo.BasicDQNLearner(app, batch_size=B)
elif kind == "dqn":
    lrn = zoo.DQNLearner(app, tapp, batch_size=B, update_horizon=w["n"])
elif kind == "rainbow":
    lrn = zoo.RainbowLearner(app, tapp, n_actions=w["A"], Vmax=10.0, Vmin=-10.0, update_horizon=w["n"], batch_size=B)
else:
    lrn = zoo.IQNLearner(app, tapp, n_actions=w["A"], N=nq, N_prime=nq, N_em=64, batch_size=B, update_horizon=w["n"])
app.set_params(O.init_params(net, 1))
if tapp is not None:
    tapp.set_params(O.init_params(net, 2))
b = bench.synth(w, 0)
db = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in b.items()}
graph = os.environ.get("TRACE_GRAPH", "1") != "0"
if graph:
    lrn.use_graph(True)
    lrn.update(db, want_loss=False, want_priorities=False)  # warm (eager); the next update captures the graph
    lrn.sync()
assert lib.zoo_test_trace_begin(256) == 0
for _ in range(6):  # the first update captures the graph (armed), the rest replay it
    lrn.update(db, want_loss=False, want_priorities=False)
lrn.sync()
n = C.c_int32(0)
lib.zoo_test_trace_stop(C.byref(n))
names = ["entry", "setup", "tma0", "stage0", "mma_done", "acc_ready", "epi_done", "joined", "freed", "c0_ldtm", "c0_sts", "c0_rows", "kb1", "kb2", "kb_last", "s15"]
recs = []
for i in range(n.value):
    nm = C.create_string_buffer(128)
    st = np.zeros(16, np.uint64)
    lib.zoo_test_trace_get(i, nm, 128, st.ctypes.data)
    recs.append((nm.value.decode(), st.astype(np.int64)))
# one update issues n/updates launches only if every update re-enqueues; with a graph the slots belong to the captured update
t0 = min(int(r[1][0]) for r in recs if r[1][0] > 0)
print("# %s %s: CTA(0,0,0) phases per engine launch, us since the step's first engine CTA; %d launches" % (wl, prec, len(recs)))
recs = sorted(recs, key=lambda r: r[1][0])
if not graph:
    recs = recs[-(len(recs) // 6):]  # the last update
t0 = min(int(r[1][0]) for r in recs if r[1][0] > 0)
for nm, st in recs:
    if st[0] == 0:
        print("%-44s (no stamps)" % nm)
        continue
    d = (st - t0) / 1e3
    print("%-44s " % nm + " ".join("%s %.2f" % (k, x) for k, x in zip(names, d[:16]) if -1e5 < x < 1e6 and k not in ("joined",)))
