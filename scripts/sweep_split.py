import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zoo_loader
lib = zoo_loader.load().lib()
def bench(eng, M, N, K, ak, bk, split, mode=2):
    us = C.c_float()
    s = lib.zoo_test_gemm_bench(eng, M, N, K, ak, bk, split, mode, 20, C.byref(us))
    return us.value if s == 0 else float("nan")
for eng in (3, 1):
    print("engine", eng)
    for tag, M, N, K, ak, bk, splits in [("fc1 fwd swapped B64", 512, 64, 7744, 1, 1, [1, 2, 4, 8, 12, 16, 24, 37, 0]),
                                          ("fc1 fwd swapped B32", 512, 32, 7744, 1, 1, [4, 8, 16, 37, 0]),
                                          ("fc1 fwd plain B64", 64, 512, 7744, 1, 1, [4, 9, 18, 0]),
                                          ("fc1 dgrad swapped", 7744, 32, 512, 0, 1, [1, 2, 4, 0]),
                                          ("conv2 fwd", 7744, 64, 512, 1, 1, [1, 2, 4]),
                                          ("conv3 wgrad", 64, 576, 3872, 0, 0, [1, 4, 8, 16, 0])]:
        print("  %-22s" % tag, "  ".join("s%d:%.1f" % (sp, bench(eng, M, N, K, ak, bk, sp, 1 if "wgrad" in tag else 2)) for sp in splits), flush=True)
names = ["entry", "setup", "tma0", "stage0", "mma_done", "acc_ready", "epi_done", "joined", "freed"]
st = np.zeros(12, np.uint64)
lib.zoo_test_gemm_trace(3, 512, 64, 7744, 1, 1, 0, st.ctypes.data)
d = (st.astype(np.int64) - int(st[0])) / 1e3
print("trace fc1 fwd swapped x3 auto split:", "  ".join("%s %.2f" % (n, x) for n, x in zip(names, d[:9])))
