"""One step's kernels from an ncu CSV launch list (any metrics): tools/launch_table.py launches.csv"""
import csv
import sys
from collections import OrderedDict

lines = open(sys.argv[1]).read().splitlines()
st = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
k = OrderedDict()
for r in csv.DictReader(lines[st:]):
    name = r["Kernel Name"].split("(")[0].replace("void <unnamed>::", "").replace("<unnamed>::", "")
    d = k.setdefault(r["ID"], {"name": name, "grid": r["Grid Size"], "block": r["Block Size"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
ids = list(k)
names = [k[i]["name"] for i in ids]
idx = [i for i, n in enumerate(names) if n.startswith("k_rhs_build")]
i0 = idx[0] if idx else 0
i1 = idx[1] if len(idx) > 1 else len(ids)
tot, agg = 0.0, {}
for i in ids[i0:i1]:
    d = k[i]
    t = d["gpu__time_duration.sum"] / 1e3
    tot += t
    agg[d["name"]] = agg.get(d["name"], 0.0) + t
    extra = ""
    if "dram__bytes_read.sum" in d:
        extra += f"  dram R {d['dram__bytes_read.sum'] / 1e6:7.1f} W {d['dram__bytes_write.sum'] / 1e6:7.1f} MB"
    if "lts__t_sectors_srcunit_tex_op_read.sum" in d:
        extra += f"  l2->sm {d['lts__t_sectors_srcunit_tex_op_read.sum'] * 32 / 1e6:8.1f} MB"
    if "sm__inst_executed.sum" in d:
        extra += f"  inst {d['sm__inst_executed.sum'] / 1e6:6.2f} M"
    print(f"{d['name'][:24]:24s} {d['grid']:14s} {d['block']:12s} {t:7.1f} us{extra}")
print(f"total {tot:.1f} us")
for n, t in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"  {n:28s} {t:8.1f} us  {100 * t / tot:5.1f} %")
