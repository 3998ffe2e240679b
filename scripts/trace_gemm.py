"""Phase timestamps (us since CTA 0's entry) of one tcgen05 GEMM launch, CTA (0,0,0) only."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zoo_loader
lib = zoo_loader.load().lib()
names = ["entry", "setup", "tma0", "stage0", "mma_done", "acc_ready", "epi_done", "joined", "freed", "c0_ldtm", "c0_sts", "c0_rows", "kb1", "kb2", "kb_last", "-"]
cases = [("tiny bf16", 1, 128, 128, 64, 1, 1, 0), ("tiny x3", 3, 128, 128, 64, 1, 1, 0), ("conv2 fwd bf16", 1, 7744, 64, 512, 1, 1, 0),
         ("conv2 fwd x3", 3, 7744, 64, 512, 1, 1, 0), ("conv3 dgrad x3", 3, 3872, 576, 64, 1, 0, 0), ("fc1 fwd x3", 3, 64, 512, 7744, 1, 1, 0),
         ("fc1 fwd bf16", 1, 64, 512, 7744, 1, 1, 0), ("fc1 fwd bf16 s1", 1, 512, 64, 7744, 1, 1, 1), ("fc1 fwd bf16 s37", 1, 512, 64, 7744, 1, 1, 37),
         ("fc1 fwd bf16 K/2 s37", 1, 512, 64, 3872, 1, 1, 37), ("fc1 fwd bf16 2K s37", 1, 512, 64, 15488, 1, 1, 37),
         ("fc1 dgrad x3", 3, 32, 7744, 512, 1, 0, 0), ("fc1 wgrad x3", 3, 512, 7744, 32, 0, 0, 0), ("fc1 dgrad bf16", 1, 32, 7744, 512, 1, 0, 0),
         ("conv1 fwd x3", 3, 28224, 32, 256, 1, 1, 0)]
for tag, eng, M, N, K, ak, bk, sp in cases:
    st = np.zeros(16, np.uint64)
    s = lib.zoo_test_gemm_trace(eng, M, N, K, ak, bk, sp, st.ctypes.data)
    d = (st.astype(np.int64) - int(st[0])) / 1e3
    print("%-22s rc %d: " % (tag, s) + "  ".join("%s %.2f" % (n, x) for n, x in zip(names, d) if abs(x) < 1e6 and n != "-"), flush=True)
