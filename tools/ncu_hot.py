"""Hot SASS lines of one profiled launch: tools/ncu_hot.py report.ncu-rep <launch index> [top]"""
import csv
import subprocess
import sys

rep, k = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [r for r in rows[2:] if len(r) > iex and (r[isamp] or "0").isdigit()]
tot_s = sum(int(r[isamp] or 0) for r in data)
tot_e = sum(int(r[iex] or 0) for r in data)
print("total samples", tot_s, "total warp-instructions", tot_e)
for n, r in enumerate(data):
    r.append(n)
for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:top]:
    print(f"{r[-1]:5d} {int(r[isamp]):7d} {100 * int(r[isamp]) / max(tot_s, 1):5.1f}% {int(r[iex]):9d}  {r[isrc][:120]}")
