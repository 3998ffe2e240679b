"""Blackwell-native evidence without a GPU: per kernel, the count of the SASS opcodes that prove tcgen05 / TMEM / TMA use
(B200_PROFILING.md "What proves a Blackwell-native kernel"), from `cuobjdump -sass` of the built objects.
usage: python scripts/sass_histogram.py > profiles/r02_sass_opcodes.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTMACCTL", "SYNCS", "HMMA", "LDGSTS", "ATOMG", "RED"]
print("# cuobjdump -sass opcode counts per kernel (sm_100a objects under build/).  UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,")
print("# UTMALDG/UTMASTG = cp.async.bulk.tensor load/store (TMA), SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be 0).")
print("# %-110s %s" % ("kernel", " ".join("%8s" % o for o in OPS)))
for obj in sorted(glob.glob(os.path.join(ROOT, "build", "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            for o in OPS:
                if op.startswith(o):
                    counts[cur][o] += 1
    print("## %s" % os.path.basename(obj))
    for k, c in counts.items():
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)[:110]
        if any(c[o] for o in OPS[:11]) or "gemm" in name or "kernel" in name:
            print("  %-110s %s" % (name, " ".join("%8d" % c[o] for o in OPS)))
