"""Device-time sweep of the GEMM engines over the contraction shapes of the Dopamine nets (zoo_test_gemm_bench)."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zoo_loader
zoo = zoo_loader.load()
lib = zoo.lib()
SHAPES = [  # (tag, M, N, K, a_kmajor, b_kmajor, mode)
    ("conv1 fwd B64", 28224, 32, 256, 1, 1, 2), ("conv2 fwd B64", 7744, 64, 512, 1, 1, 2), ("conv3 fwd B64", 7744, 64, 576, 1, 1, 2),
    ("fc1 fwd B64", 64, 512, 7744, 1, 1, 2), ("fc1 wgrad B32", 512, 7744, 32, 0, 0, 1), ("fc1 dgrad B32", 32, 7744, 512, 1, 0, 0),
    ("conv3 wgrad B32", 64, 576, 3872, 0, 0, 1), ("conv3 dgrad B32", 3872, 576, 64, 1, 0, 0), ("conv1 wgrad B32", 32, 256, 14112, 0, 0, 1),
    ("tiny", 128, 128, 64, 1, 1, 0), ("fc1 fwd B1024", 1024, 512, 7744, 1, 1, 2), ("iqn fc1 fwd 32768", 32768, 512, 7744, 1, 1, 2),
    ("square 4096", 4096, 4096, 4096, 1, 1, 0),
]
engines = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 2, 3, 0]
for tag, M, N, K, ak, bk, mode in SHAPES:
    row = []
    for eng in engines:
        if eng == 0 and M * N * K > 2e11:
            row.append("   skip")
            continue
        us = C.c_float()
        s = lib.zoo_test_gemm_bench(eng, M, N, K, ak, bk, 0, mode, 20, C.byref(us))
        row.append("%8.1f" % us.value if s == 0 else "  ERR%d" % s)
    fl = 2.0 * M * N * K
    print("%-20s M%6d N%5d K%6d | us by engine %s: %s | %.2f GFLOP" % (tag, M, N, K, engines, " ".join(row), fl / 1e9), flush=True)
