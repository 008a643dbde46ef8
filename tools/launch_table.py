"""Print one step's kernels from an ncu launch list (gpu__time_duration.sum CSV): tools/launch_table.py launches.csv"""
import csv
import sys

lines = open(sys.argv[1]).read().splitlines()
st = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(lines[st:]))
names = [r["Kernel Name"].split("(")[0].replace("void <unnamed>::", "").replace("<unnamed>::", "") for r in rows]
idx = [i for i, n in enumerate(names) if n.startswith("k_rhs_build")]
i0, i1 = idx[-2], idx[-1]
tot = 0.0
agg = {}
for r, n in zip(rows[i0:i1], names[i0:i1]):
    t = float(r["Metric Value"]) / 1e3
    tot += t
    agg[n] = agg.get(n, 0.0) + t
    print(f"{n:28s} grid {r['Grid Size']:18s} block {r['Block Size']:14s} {t:8.1f} us")
print(f"total {tot:.1f} us")
for n, t in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"  {n:28s} {t:8.1f} us  {100 * t / tot:5.1f} %")
