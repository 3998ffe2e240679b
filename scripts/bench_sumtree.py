"""SumTree throughput at BASELINE.json config 3 (capacity 1M, batch 32 -> 4096): device-resident stratified sample + priority
update through the C ABI (*_dev entry points, CUDA events on the tree's stream), the host-pointer calls (H2D/D2H inside), and
the C oracle on one host core.  Algorithmic bytes: 21 nodes x 4 B per draw, 21 x 8 B per update."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import zoo_loader  # noqa: E402
from oracle.sumtree import CSumTree  # noqa: E402

zoo = zoo_loader.load()
lib = zoo.lib()
from rlzoo_b200 import _lib as ZL  # noqa: E402

cap = 1_000_000
rng = np.random.default_rng(0)
pr = np.power(rng.random(cap, dtype=np.float32) + np.float32(1e-10), np.float32(0.5)).astype(np.float32)
t = zoo.SumTree(cap)
t.push_many(pr)
o = CSumTree(cap)
o.push_many(pr)
sp = C.c_void_p()
ZL.check(lib.zoo_sumtree_stream(t.h, C.byref(sp)))
stream = torch.cuda.ExternalStream(sp.value)
out = []
for B in (32, 256, 1024, 4096):
    u = torch.from_numpy(rng.random(B)).cuda()
    newp = torch.from_numpy(np.power(rng.random(B, dtype=np.float32) + 1e-10, 0.5).astype(np.float32)).cuda()
    idx = torch.empty(B, dtype=torch.int64, device="cuda")
    p = torch.empty(B, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()

    def step():
        ZL.check(lib.zoo_sumtree_sample_dev(t.h, u.data_ptr(), ZL.DRAW_F64, B, ZL.SAMPLE_STRATIFIED, idx.data_ptr(), p.data_ptr()))
        ZL.check(lib.zoo_sumtree_update_dev(t.h, idx.data_ptr(), newp.data_ptr(), B))
    for _ in range(5):
        step()
    K = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(K):
            step()
        e1.record(stream)
    torch.cuda.synchronize()
    dev_us = e0.elapsed_time(e1) * 1e3 / K
    uh, nh = u.cpu().numpy(), newp.cpu().numpy()
    t0 = time.perf_counter()
    for _ in range(50):
        ih, _ = t.sample(uh, stratified=True)
        t.update(ih, nh)
    host_us = (time.perf_counter() - t0) * 1e6 / 50
    t0 = time.perf_counter()
    for _ in range(50):
        io, _ = o.sample(uh, stratified=True)
        o.update(io, nh)
    cpu_us = (time.perf_counter() - t0) * 1e6 / 50
    byts = B * 21 * (4 + 8)
    out.append({"batch": B, "device_us_per_sample_plus_update": round(dev_us, 2), "device_Msamples_per_s": round(B / dev_us, 2),
                "achieved_GBps_algorithmic": round(byts / dev_us / 1e3, 3), "host_api_us": round(host_us, 1), "cpu_oracle_us": round(cpu_us, 1),
                "speedup_vs_cpu_oracle_device": round(cpu_us / dev_us, 1)})
    print(json.dumps(out[-1]), flush=True)

# per-kernel device times (in-situ launch profiler) at B = 4096
B = 4096
u = torch.from_numpy(rng.random(B)).cuda()
newp = torch.from_numpy(np.power(rng.random(B, dtype=np.float32) + 1e-10, 0.5).astype(np.float32)).cuda()
idx = torch.empty(B, dtype=torch.int64, device="cuda")
p = torch.empty(B, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
ZL.check(lib.zoo_prof_begin(sp.value, 2000))
ZL.check(lib.zoo_sumtree_sample_dev(t.h, u.data_ptr(), ZL.DRAW_F64, B, ZL.SAMPLE_STRATIFIED, idx.data_ptr(), p.data_ptr()))
ZL.check(lib.zoo_sumtree_update_dev(t.h, idx.data_ptr(), newp.data_ptr(), B))
n = C.c_int32()
ZL.check(lib.zoo_prof_end(C.byref(n)))
name, ms = C.create_string_buffer(256), C.c_float()
for i in range(n.value):
    ZL.check(lib.zoo_prof_get(i, name, 256, C.byref(ms)))
    print("  %8.2f us  %s" % (ms.value * 1e3, name.value.decode()))
