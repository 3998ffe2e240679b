"""fc1-shaped GEMM (skinny batch x long K): device time vs K (row pitch of the K-major operands), engine and split."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zoo_loader
lib = zoo_loader.load().lib()
def bench(eng, M, N, K, ak, bk, split, mode, knob=0):
    us = C.c_float()
    lib.zoo_test_set_tc_stages(knob)
    s = lib.zoo_test_gemm_bench(eng, M, N, K, ak, bk, split, mode, 20, C.byref(us))
    lib.zoo_test_set_tc_stages(0)
    return us.value if s == 0 else float("nan")
for K in (7744, 7680, 7712, 7808, 8192, 3872, 15488):
    print("K %5d | " % K + "  ".join("%s s%d %.1f/%.1f" % (nm, sp, bench(e, 512, 64, K, 1, 1, sp, 2), bench(e, 512, 64, K, 1, 1, sp, 10))
                                      for nm, e in (("bf16", 1), ("x3", 3)) for sp in (0, 37, 1)), flush=True)
print("no split-K epilogue cost: M=512 N=64 K=7744 split 1 above; store-only epilogue (mode 0):")
for sp in (0, 37):
    print("  x3 mode0 s%d %.1f" % (sp, bench(3, 512, 64, 7744, 1, 1, sp, 0)), " bf16 mode0 s%d %.1f" % (sp, bench(1, 512, 64, 7744, 1, 1, sp, 0)))
