"""Split-K x ring-depth sweep of the tcgen05 GEMM on the fc1 / conv shapes of the Atari update; warm (L2-resident) and
cold (every launch reads its own operand copy, copies > L2) device times in microseconds."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zoo_loader
lib = zoo_loader.load().lib()
def bench(eng, M, N, K, ak, bk, split, mode, stages):
    us = C.c_float()
    lib.zoo_test_set_tc_stages(stages)
    s = lib.zoo_test_gemm_bench(eng, M, N, K, ak, bk, split, mode, 20, C.byref(us))
    lib.zoo_test_set_tc_stages(0)
    return us.value if s == 0 else float("nan")
engines = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "3,1").split(",")]
for eng in engines:
    print("engine", eng)
    for tag, M, N, K, ak, bk, splits in [("fc1 fwd swapped B64", 512, 64, 7744, 1, 1, [0, 18, 24, 37, 49, 74]),
                                          ("fc1 fwd swapped B32", 512, 32, 7744, 1, 1, [0, 18, 37, 74]),
                                          ("fc1 dgrad swapped", 7744, 32, 512, 0, 1, [0, 1, 2, 4]),
                                          ("conv1 fwd B64", 28224, 32, 256, 1, 1, [0, 1, 2]),
                                          ("conv2 fwd B64", 7744, 64, 512, 1, 1, [0, 1, 2, 4]),
                                          ("conv2 fwd B32", 3872, 64, 512, 1, 1, [0, 1, 2, 4, 8]),
                                          ("conv3 fwd B64", 7744, 64, 576, 1, 1, [0, 1, 2, 3]),
                                          ("conv3 fwd B32", 3872, 64, 576, 1, 1, [0, 1, 2, 3, 6]),
                                          ("conv3 dgrad", 3872, 576, 64, 1, 0, [0, 1]),
                                          ("conv3 wgrad", 64, 576, 3872, 0, 0, [0, 8, 16, 31]),
                                          ("conv1 wgrad", 32, 256, 14112, 0, 0, [0, 16, 32, 74])]:
        md = 1 if "wgrad" in tag else 2
        for stages in (0, 2, 4):
            print("  %-20s stages %d | " % (tag, stages) + "  ".join("s%d: %.1f/%.1f" % (sp, bench(eng, M, N, K, ak, bk, sp, md, stages), bench(eng, M, N, K, ak, bk, sp, md | 8, stages)) for sp in splits), flush=True)
