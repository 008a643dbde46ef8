"""Per-kernel time / DRAM traffic of ONE step from an ncu launch list
(ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv ...):
    python tools/kernel_summary.py gpurun_out/launches.csv profiles/r01_kernel_summary.json
A step is the launch sequence between two consecutive k_ctrl_add launches (the first kernel of a fused step)."""
import csv
import json
import sys
from collections import OrderedDict

lines = open(sys.argv[1]).read().splitlines()
st = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
k = OrderedDict()
for r in csv.DictReader(lines[st:]):
    name = r["Kernel Name"].split("(")[0].replace("void <unnamed>::", "").replace("<unnamed>::", "")
    name = name.split("<")[0]
    d = k.setdefault(r["ID"], {"name": name})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
ids = list(k)
marks = [i for i, j in enumerate(ids) if k[j]["name"] == "k_ctrl_add"]
if len(marks) < 2:
    sys.exit("need at least two k_ctrl_add launches in the list")
i0, i1 = marks[-2], marks[-1]
agg = OrderedDict()
for j in ids[i0:i1]:
    d = k[j]
    a = agg.setdefault(d["name"], {"launches": 0, "time_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
    a["launches"] += 1
    a["time_us"] += d["gpu__time_duration.sum"] / 1e3
    a["dram_read_bytes"] += d.get("dram__bytes_read.sum", 0.0)
    a["dram_write_bytes"] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(a["time_us"] for a in agg.values())
for n, a in agg.items():
    a["share_of_step"] = a["time_us"] / tot
    a["dram_bytes"] = a["dram_read_bytes"] + a["dram_write_bytes"]
out = {"source": sys.argv[1], "note": "ncu per-launch times are cold-cache and serialised: compare shares, not absolutes",
       "step_time_us_under_ncu": tot, "kernels": agg}
json.dump(out, open(sys.argv[2], "w"), indent=1)
for n, a in agg.items():
    print(f"{n:18s} x{a['launches']:3d} {a['time_us']:8.1f} us {100 * a['share_of_step']:5.1f} %  dram {a['dram_bytes'] / 1e6:8.1f} MB")
print(f"step under ncu: {tot:.1f} us")
