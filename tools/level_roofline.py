"""Per-launch floor of the multifrontal sweeps of the benchmark problem (cylinder, 256 trajectories): the FP64 tensor-core
time of the launch's dense blocks at the measured DMMA issue rate, and the bytes its jobs move between L2 and the SMs
(gathered rows once per 32-row tile and source plane, seed rows, stored rows), beside the measured start-to-start time of
the launch taken from the per-CTA timeline (profiles/r02_sweep_timeline.txt).  Host only.

    python tools/level_roofline.py [B=256] > profiles/r02_level_roofline.txt
"""
import os
import re
import sys
from pathlib import Path

os.environ.setdefault("FCB_BALANCED", "0")  # the archived timeline was taken with the free dissection (23 launches)

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import build_problem  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
DMMA_FLOPS = 148 * 4 * 512 / 16 * 1.965e9  # one m8n8k4 (512 flop) per 16 cycles and SM sub-partition (tools/bench_src/dmma_chains.cu)
fs, prob = build_problem()
plan = prob.plans_for_batch(B)[2]  # the BDF2 operator: the one every step after the first solves with
starts = []
for line in (ROOT / "profiles" / "r02_sweep_timeline.txt").read_text().splitlines():
    m = re.match(r"\s*(\d+) (fwd|bwd)\s+\S+\s+\d+\s+\d+\s+([\d.]+)\s+(-?[\d.]+)\s+([\d.]+)", line)
    if m:
        starts.append((float(m.group(3)), float(m.group(5))))
nl = len(plan.launch_ptr) - 1
print(f"cylinder, {B} trajectories, n = {plan.n}, factor entries {plan.vals.size / 1e6:.2f} M, {nl} launches")
print(f"DMMA rate {DMMA_FLOPS / 1e12:.1f} TFLOP/s (measured issue rate); bytes = rows x {B} x 8")
print(" l dir  blocks   K med/max   M med/max  entries(M)  dmma(us)  L2<->SM(MB)  at 12.4TB/s(us)  measured(us)")
tot = np.zeros(4)
for l in range(nl):
    b0, b1 = int(plan.launch_ptr[l]), int(plan.launch_ptr[l + 1])
    K, M, ns = plan.blk_K[b0:b1].astype(np.int64), plan.blk_M[b0:b1].astype(np.int64), plan.blk_nsrc[b0:b1].astype(np.int64)
    tiles = np.maximum((M + 31) // 32, 1)
    seeds = 0
    for b in range(b0, b1):
        e0, e1 = int(plan.blk_eptr[b]), int(plan.blk_eptr[b]) + int(plan.blk_M[b])
        seeds += int((plan.e0[e0:e1] >= 0).sum() + (plan.e1[e0:e1] >= 0).sum())
    rows = int((tiles * K * ns).sum()) + seeds + int(M.sum()) + int(K[plan.blk_ystore[b0:b1] >= 0].sum())
    if l >= plan.n_forward_launches:
        rows += int(M.sum())  # the backward sweep also writes x in canonical numbering
    ent = float((M * K).sum())
    t_dmma = 2 * ent * B / DMMA_FLOPS * 1e6
    mb = rows * B * 8 / 1e6
    t_l2 = mb / 12.4
    meas = (starts[l + 1][0] - starts[l][0]) if l + 1 < len(starts) else (starts[l][1] if l < len(starts) else float("nan"))
    tot += (ent, t_dmma, mb, meas)
    d = "fwd" if l < plan.n_forward_launches else "bwd"
    print(f"{l:2d} {d} {b1 - b0:7d} {int(np.median(K)):6d}/{K.max():4d} {int(np.median(M)):6d}/{M.max():4d} {ent / 1e6:10.2f} {t_dmma:9.1f} {mb:12.0f} {t_l2:16.1f} {meas:13.1f}")
print(f"sum {'':36s} {tot[0] / 1e6:10.2f} {tot[1]:9.1f} {tot[2]:12.0f} {tot[2] / 12.4:16.1f} {tot[3]:13.1f}")
