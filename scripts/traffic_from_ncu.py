"""profiles/traffic.json from an `ncu --set full` raw-page CSV of ONE eager update (scripts/ncu_r02.sh) and the in-situ launch labels
of the same workload (scripts/prof_update.py): DRAM bytes read + written per launch, keyed by the label bench.py's roofline uses.
Launch i of the capture is label i of the profile (marks that are not kernels are skipped); every tcgen05-engine label must land on a
tc_gemm kernel, else the script refuses.   usage: traffic_from_ncu.py <workload:prec> <raw.csv> <insitu.txt> [traffic.json]"""
import csv
import json
import os
import re
import sys

csv.field_size_limit(10 ** 9)
key, raw, insitu = sys.argv[1], sys.argv[2], sys.argv[3]
out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    u = units[col[name]].lower()
    k = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
    return float(r[col[name]].replace(",", "")) * k


launches = [(re.sub(r"\(.*", "", r[col["Kernel Name"]]), val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")) for r in rows[2:]]
labels = []
for line in open(insitu):
    m = re.match(r"\s*[\d.]+\s+(\S.*)$", line)
    if m and not line.startswith("#") and "TOTAL" not in line:
        lab = m.group(1).strip()
        if lab.startswith("stage|"):  # the staging mark is an event, not a kernel
            continue
        labels.append(lab)
assert len(labels) == len(launches), (len(labels), len(launches))
tab = {}
for lab, (name, b) in zip(labels, launches):
    if "launch_tc" in lab or "launch_conv" in lab:
        assert "tc_gemm" in name, (lab, name)
    tab[lab.split("|")[0]] = int(b)
d = json.load(open(out)) if os.path.exists(out) else {}
d[key] = tab
d["_source"] = ("ncu --set full --clock-control none of one eager update (scripts/ncu_r02.sh): dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                "matched to the in-situ profiler's labels in launch order (scripts/traffic_from_ncu.py); tables under profiles/r02_ncu_full_*.txt")
json.dump(d, open(out, "w"), indent=1)
print("%s: %d launches" % (key, len(tab)))
