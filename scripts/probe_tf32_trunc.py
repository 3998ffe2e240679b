"""Does tcgen05 kind::tf32 truncate (ignore) the low 13 mantissa bits of fp32 operands?  Compare a tf32 GEMM on raw fp32
inputs with the same GEMM on inputs whose low 13 bits were cleared."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zoo_loader
lib = zoo_loader.load().lib()
rng = np.random.default_rng(0)
M, N, K = 256, 128, 512
A = rng.standard_normal((M, K)).astype(np.float32)
B = rng.standard_normal((N, K)).astype(np.float32)
def run(a, b):
    D = np.zeros((M, N), np.float32)
    assert lib.zoo_test_gemm(2, M, N, K, a.ctypes.data, 1, b.ctypes.data, 1, D.ctypes.data, 1) == 0
    return D
trunc = lambda x: (x.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)
D0, D1 = run(A, B), run(trunc(A), trunc(B))
print("raw vs truncated inputs: bit-identical =", np.array_equal(D0, D1), " max abs diff", np.abs(D0 - D1).max())
rn = lambda x: ((x.view(np.uint32).astype(np.uint64) + 0x1000) & 0xffffe000).astype(np.uint32).view(np.float32)
D2 = run(rn(A), rn(B))
print("raw vs round-to-nearest inputs: bit-identical =", np.array_equal(D0, D2), " max abs diff", np.abs(D0 - D2).max())
