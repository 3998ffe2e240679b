"""A few eager learner updates of a bench workload, the last `n` of them inside a cudaProfilerStart/Stop range (for
`ncu --profile-from-start off`).  usage: one_update.py <workload> <precision> [n_profiled=1] [warm=3]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import zoo_loader  # noqa: E402

wl, prec = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 3
zoo = zoo_loader.load()
w = bench.WORKLOADS[wl]
lrn, app, tapp = bench.build_learner(zoo, w, prec, None, 0)
_, pool = bench.make_pools(w, 0, torch.device("cuda", 0), torch, 2)
for i in range(warm):
    lrn.update(pool[i % 2])
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(n):
    lrn.update(pool[i % 2])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("one_update ok: %s %s loss %.6f" % (wl, prec, float(lrn.loss)))
