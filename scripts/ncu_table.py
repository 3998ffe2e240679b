"""One line per captured launch of an `ncu --set full` raw-page CSV: duration, DRAM bytes (read + write), DRAM and tensor-pipe
utilisation, warps active, registers, dynamic shared memory.  usage: ncu_table.py <raw.csv> [title]"""
import csv
import re
import sys

csv.field_size_limit(10 ** 9)
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def f(r, name):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:
        return float("nan")


def scale(name, v):  # to bytes / microseconds whatever unit ncu chose
    u = units[col[name]].lower()
    k = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "second": 1e6, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "us": 1.0, "ns": 1e-3}
    return v * k.get(u, 1.0)


print("# %s" % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# ncu --set full --clock-control none, one eager learner update inside a cudaProfilerStart/Stop range (scripts/ncu_r02.sh);")
print("# per launch: device time, DRAM bytes read + written, achieved DRAM GB/s (bytes / time), DRAM %% of peak, tensor pipe %% active,")
print("# warps active %%, registers / thread, dynamic smem / CTA.  Times under ncu are cold-cache and serialised.")
print("%-72s %10s %9s %10s %9s %7s %8s %7s %5s %8s" % ("kernel <grid>", "time us", "rd MB", "wr MB", "GB/s", "dram%", "tensor%", "warps%", "regs", "smem KB"))
tot = 0.0
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
    name = name.replace("void zoo::", "").replace("zoo::", "")
    t = scale("gpu__time_duration.sum", f(r, "gpu__time_duration.sum"))
    rd = scale("dram__bytes_read.sum", f(r, "dram__bytes_read.sum"))
    wr = scale("dram__bytes_write.sum", f(r, "dram__bytes_write.sum"))
    tot += t
    grid = r[col["Grid Size"]].replace(" ", "")
    print("%-72s %10.2f %9.2f %10.2f %9.1f %7.1f %8.1f %7.1f %5d %8.1f" % (
        (name + " <" + grid + ">")[:72], t, rd / 1e6, wr / 1e6, (rd + wr) / t / 1e3 if t > 0 else 0.0,
        f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), int(f(r, "launch__registers_per_thread")),
        scale("launch__shared_mem_per_block_dynamic", f(r, "launch__shared_mem_per_block_dynamic")) / 1e3))
print("# %d launches, %.1f us in total" % (len(rows) - 2, tot))
