"""ncu_hot_sass.py <report.ncu-rep> <launch index> [min share]: SASS lines of one captured launch whose warp-stall samples are at
least `min share` of the launch's samples, plus 50-instruction buckets -- the "where does the time go" view of --import-source."""
import csv, subprocess, sys, io
rep, idx = sys.argv[1], int(sys.argv[2])
share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.006
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
ins, name = [], ""
for r in rows:
    if r and r[0] == "Kernel Name": name = r[1]
    if len(r) > 6:
        try: s = int(r[4])
        except ValueError: continue
        ins.append((s, r[1][:88], r[5]))
tot = sum(i[0] for i in ins)
print(f"# {name[:100]}\n# samples {tot}, instructions {len(ins)}")
for k, i in enumerate(ins):
    if i[0] >= tot * share: print(f"{k:5d} {i[0]:6d} {100*i[0]/tot:5.1f}%  x{i[2]:>8}  {i[1]}")
print("# buckets of 50 instructions")
for b in range(0, len(ins), 50):
    s = sum(i[0] for i in ins[b:b+50])
    if s >= tot * 0.01: print(f"{b:5d} {s:6d} {100*s/tot:5.1f}%  {ins[b][1][:60]}")
