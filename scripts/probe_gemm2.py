"""GPU probe: every tcgen05 / SIMT GEMM variant and the SumTree kernels against numpy / the C oracle."""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = ctypes.CDLL(os.path.join(ROOT, "reinforcementlearningzoo.jl_b200", "libzoo_sm100.so"))
lib.zoo_last_error.restype = ctypes.c_char_p
V=ctypes.c_void_p
lib.zoo_test_gemm.argtypes=[ctypes.c_int]*4+[V,ctypes.c_int,V,ctypes.c_int,V,ctypes.c_int]
lib.zoo_sumtree_push_many.argtypes=[V,V,ctypes.c_int64]
lib.zoo_sumtree_tree_len.argtypes=[V,V]
lib.zoo_sumtree_export.argtypes=[V,V,ctypes.c_int64]
lib.zoo_sumtree_sample.argtypes=[V,V,ctypes.c_int,ctypes.c_int64,ctypes.c_int,V,V]
lib.zoo_sumtree_update.argtypes=[V,V,V,ctypes.c_int64]
lib.zoo_sumtree_destroy.argtypes=[V]
import faulthandler; faulthandler.enable()
print("arch", lib.zoo_device_arch())

def bf16_round(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)

def gemm(engine, M, N, K, ak, bk, split):
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    Am = np.ascontiguousarray(A if ak else A.T)
    Bm = np.ascontiguousarray(B if bk else B.T)
    D = np.zeros((M, N), np.float32)
    s = lib.zoo_test_gemm(engine, M, N, K, Am.ctypes.data, ak, Bm.ctypes.data, bk, D.ctypes.data, split)
    if s != 0:
        return "ERR %d %s" % (s, lib.zoo_last_error().decode())
    if engine == 1:
        ref = bf16_round(A).astype(np.float64) @ bf16_round(B).astype(np.float64).T
    else:
        ref = A.astype(np.float64) @ B.astype(np.float64).T
    err = np.abs(D - ref).max() / np.abs(ref).max()
    return "relerr %.2e" % err

for engine in (3, 2):
    for (M, N, K) in [(128, 128, 64), (128, 64, 128), (256, 32, 256), (300, 200, 136), (3872, 64, 512), (32, 512, 7744), (512, 7744, 32), (64, 576, 3872), (32, 256, 1000)]:
        for ak in (1, 0):
            for bk in (1, 0):
                for split in (1, 0):
                    if ((M if not ak else K) % 4) or ((N if not bk else K) % 4):
                        continue
                    t0 = time.time()
                    r = gemm(engine, M, N, K, ak, bk, split)
                    print("engine %d M %5d N %5d K %5d ak %d bk %d split %d : %s  (%.0f ms)" % (engine, M, N, K, ak, bk, split, r, (time.time() - t0) * 1e3), flush=True)

