"""Text summary of a single-kernel `ncu --set full` report (the numbers DESIGN.md quotes):
    python tools/ncu_summary.py gpurun_out/r01_spmm_mma.ncu-rep profiles/r01_ncu_spmm_mma.txt
Key throughput metrics, DRAM bytes, occupancy limits, warp stall reasons above 2 %, and the 20 hottest SASS lines."""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, v = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}
want = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
]
lines = [f"ncu --set full --clock-control none, one launch; report {rep}", ""]
for k in want:
    if k in col:
        lines.append(f"{k:75s} {v[col[k]]} {units[col[k]]}")
lines += ["", "warp stall reasons (pc sampling, % of samples, > 2 %):"]
samp = {}
for h, i in col.items():
    if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
        try:
            samp[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(v[i])
        except ValueError:
            pass
stot = sum(samp.values()) or 1.0
for k, x in sorted(samp.items(), key=lambda t: -t[1]):
    if 100 * x / stot > 2:
        lines.append(f"  {k:28s} {100 * x / stot:6.1f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
sh = srows[1]
isrc, isamp, iex = sh.index("Source"), sh.index("# Samples"), sh.index("Instructions Executed")
data = [r for r in srows[2:] if len(r) > iex and (r[isamp] or "0").isdigit()]
tot = sum(int(r[isamp]) for r in data) or 1
bars = [i for i, r in enumerate(data) if "BAR.SYNC" in r[isrc]]
lines += ["", f"SASS: {len(data)} instructions, {tot} samples; samples between barriers (instruction index ranges):"]
edges = [0] + bars + [len(data)]
for a, b in zip(edges[:-1], edges[1:]):
    s = sum(int(r[isamp]) for r in data[a:b])
    lines.append(f"  [{a:5d},{b:5d})  {100 * s / tot:5.1f} %")
lines += ["", "hottest SASS lines (index, samples, %, executions, instruction):"]
for i, r in sorted(enumerate(data), key=lambda t: -int(t[1][isamp]))[:20]:
    lines.append(f"  {i:5d} {int(r[isamp]):6d} {100 * int(r[isamp]) / tot:5.1f}% {int(r[iex] or 0):9d}  {r[isrc][:100]}")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:45]))
