"""Timing ablations of the 3xTF32 main loop (results are wrong in these modes): where does a K-block's time go?"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zoo_loader
lib = zoo_loader.load().lib()
def bench(eng, M, N, K, ak, bk, split, mode, knob):
    us = C.c_float()
    lib.zoo_test_set_tc_stages(knob)
    s = lib.zoo_test_gemm_bench(eng, M, N, K, ak, bk, split, mode, 20, C.byref(us))
    lib.zoo_test_set_tc_stages(0)
    return us.value if s == 0 else float("nan")
for tag, M, N, K, ak, bk in [("conv2 fwd", 7744, 64, 512, 1, 1), ("conv3 fwd", 7744, 64, 576, 1, 1), ("conv1 fwd", 28224, 32, 256, 1, 1),
                             ("fc1 fwd swapped", 512, 64, 7744, 1, 1), ("fc1 dgrad swapped", 7744, 32, 512, 0, 1), ("conv3 dgrad", 3872, 576, 64, 1, 0),
                             ("fc1 fwd B1024", 1024, 512, 7744, 1, 1), ("square 4096", 4096, 4096, 4096, 1, 1)]:
    row = ["%-18s" % tag]
    row.append("bf16 %.1f" % bench(1, M, N, K, ak, bk, 0, 2, 0))
    row.append("tf32 %.1f" % bench(2, M, N, K, ak, bk, 0, 2, 0))
    for name, x in [("x3", 0), ("x3-noconv", 1), ("x3-1mma", 2), ("x3-noconv-1mma", 3), ("x3-nofence", 4), ("x3-nostore", 8), ("x3-nostore-nofence", 12)]:
        row.append("%s %.1f" % (name, bench(3, M, N, K, ak, bk, 0, 2, x << 8)))
    print("  ".join(row), flush=True)
