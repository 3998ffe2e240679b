"""Device vs torch twin at the bench's own configurations with the relu masks pinned (tests/helpers.py:full_size_parity).
usage: full_size_parity.py [case ...] [--prec fp32,bf16] ; writes one report block per (case, precision) to stdout."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import zoo_loader
from helpers import FULL_SIZE_CASES, format_full_size_report, full_size_parity
zoo = zoo_loader.load()
args = sys.argv[1:]
precs = ["fp32", "bf16"]
if "--prec" in args:
    i = args.index("--prec")
    precs = args[i + 1].split(",")
    del args[i:i + 2]
cases = args or list(FULL_SIZE_CASES)
for c in cases:
    for prec in precs:
        t0 = time.time()
        rep = full_size_parity(zoo, c, prec)
        print(format_full_size_report(rep) + "\n    (%.1f s)" % (time.time() - t0), flush=True)
