"""Device time of SumTree sample / update separately at small batches (CUDA events on the tree's stream, 200 back-to-back calls)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, zoo_loader
zoo = zoo_loader.load(); lib = zoo.lib()
from rlzoo_b200 import _lib as ZL
cap = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(0)
t = zoo.SumTree(cap); t.push_many(np.power(rng.random(cap, dtype=np.float32) + np.float32(1e-10), np.float32(0.5)).astype(np.float32))
sp = C.c_void_p(); ZL.check(lib.zoo_sumtree_stream(t.h, C.byref(sp))); stream = torch.cuda.ExternalStream(sp.value)
for B in (32, 64, 256):
    u = torch.from_numpy(rng.random(B)).cuda()
    newp = torch.from_numpy(rng.random(B, dtype=np.float32)).cuda()
    idx = torch.empty(B, dtype=torch.int64, device="cuda"); p = torch.empty(B, dtype=torch.float32, device="cuda")
    def samp(): ZL.check(lib.zoo_sumtree_sample_dev(t.h, u.data_ptr(), ZL.DRAW_F64, B, ZL.SAMPLE_STRATIFIED, idx.data_ptr(), p.data_ptr()))
    def upd(): ZL.check(lib.zoo_sumtree_update_dev(t.h, idx.data_ptr(), newp.data_ptr(), B))
    samp(); upd(); torch.cuda.synchronize()
    for name, fn in (("sample", samp), ("update", upd)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(200): fn()
            e1.record(stream)
        torch.cuda.synchronize()
        print("B=%d %s: %.2f us" % (B, name, e0.elapsed_time(e1) * 1e3 / 200), flush=True)
