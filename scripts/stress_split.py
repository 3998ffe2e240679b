"""Repeats split-K GEMMs of the shapes the mid-K heuristic picks and counts wrong products (a stand-alone race hunt)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import zoo_loader
zoo = zoo_loader.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for eng, M, N, K, split in [(3, 306, 8, 512, 8), (3, 648, 64, 512, 8), (3, 392, 64, 576, 9), (3, 306, 8, 512, 0), (3, 3136, 8, 512, 8),
                            (3, 1296, 64, 512, 8), (3, 784, 64, 576, 9), (3, 1296, 64, 512, 0), (3, 784, 64, 576, 0), (3, 7744, 32, 512, 4),
                            (1, 1296, 64, 512, 8), (3, 512, 16, 7744, 0), (3, 3872, 64, 512, 8)]:
    rng = np.random.default_rng(1)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    bad, worst = 0, 0.0
    for r in range(reps):
        D = np.zeros((M, N), np.float32)
        s = zoo.lib().zoo_test_gemm(eng, M, N, K, A.ctypes.data, 1, B.ctypes.data, 1, D.ctypes.data, split)
        assert s == 0
        e = float(np.abs(D - ref).max() / np.abs(ref).max())
        worst = max(worst, e)
        if e > (5e-6 if eng == 3 else 2e-2):
            bad += 1
            if bad <= 2:
                w = np.argwhere(np.abs(D - ref) > 1e-3 * np.abs(ref).max())
                print("   bad rows %s..%s cols %s..%s count %d" % (w[:, 0].min(), w[:, 0].max(), w[:, 1].min(), w[:, 1].max(), len(w)))
    print("engine %d  %5d x %3d x %5d split %d: %d / %d wrong, worst %.2e" % (eng, M, N, K, split, bad, reps, worst), flush=True)
