"""Summarise ncu artefacts brought back from the GPU box into small text files for profiles/.
usage: python scripts/ncu_summary.py launches <launches.csv> | full <report.ncu-rep>"""
import collections
import csv
import subprocess
import sys

csv.field_size_limit(10 ** 9)
mode, path = sys.argv[1], sys.argv[2]
if mode == "launches":
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows[1:]:
        if "gpu__time_duration" in r[mi]:
            k = r[ki][:90]
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += float(r[vi].replace(",", ""))
            tot += float(r[vi].replace(",", ""))
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)")
    print("# launches %d, total %.1f us" % (sum(a[0] for a in agg.values()), tot / 1e3))
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%5d launches %9.1f us  %5.1f%%  avg %7.2f us  %s" % (n, t / 1e3, 100 * t / tot, t / n / 1e3, k))
else:
    if path.endswith(".csv"):  # already exported on the GPU box (the .ncu-rep itself can exceed the 64 MiB return limit)
        out = open(path).read()
    else:
        out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
            "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    units = rows[1]
    print("# ncu --set full --clock-control none: selected raw metrics per captured launch")
    for r in rows[2:]:
        print("---")
        for w, i in idx:
            print("  %-66s %s %s" % (w, r[i], units[i]))
