"""Per-launch timeline of the sweep kernels from an FCB_SWEEP_DEBUG dump (globaltimer ns):
    FCB_SWEEP_DEBUG=gpurun_out/sweep_dbg.bin python tools/gpu_check.py 256 4 ; python tools/sweep_timeline.py gpurun_out/sweep_dbg.bin
columns are microseconds: launch start relative to the first launch, then medians / maxima over CTAs of
  entry   CTA entry after the launch's first CTA entry (ramp)
  issue   producer issued its first stage, after CTA entry
  full    consumer warp 0 saw its first stage, after CTA entry
  job1    duration of the first job (after `full`)
  life    CTA entry -> consumer done"""
import struct
import sys

import numpy as np

raw = open(sys.argv[1], "rb").read()
nlaunch, nfwd = struct.unpack_from("ii", raw, 0)
off = 8
t_first = None
prev_end = None
print(f"{'l':>2} {'dir':3} {'grid':>9} {'nwc':>3} {'st':>2} {'start':>8} {'gap':>6} {'span':>6} | entry med/max | issue med | full med/max | job1 med/max | life med/max | jobs/cta max")
tot = 0.0
for l in range(nlaunch):
    grid, nslab, nwc, nst = struct.unpack_from("iiii", raw, off)
    off += 16
    n = grid * nslab
    a = np.frombuffer(raw, dtype=np.uint64, count=n * 8, offset=off).reshape(n, 8).astype(np.int64)
    off += n * 64
    if n == 0:
        continue
    t0 = a[:, 0].min()
    tend = max(a[:, 4].max(), a[:, 5].max())
    if t_first is None:
        t_first = t0
    gap = (t0 - prev_end) / 1e3 if prev_end is not None else 0.0
    prev_end = tend
    us = lambda x: x / 1e3  # noqa: E731
    entry = us(a[:, 0] - t0)
    issue = us(a[:, 1] - a[:, 0])
    full = us(a[:, 2] - a[:, 0])
    job1 = us(a[:, 3] - a[:, 2])
    life = us(a[:, 4] - a[:, 0])
    span = us(tend - t0)
    tot += span + max(gap, 0)
    print(f"{l:2d} {'fwd' if l < nfwd else 'bwd'} {grid:5d}x{nslab:<3d} {nwc:3d} {nst:2d} {us(t0 - t_first):8.1f} {gap:6.1f} {span:6.1f} | "
          f"{np.median(entry):5.1f} {entry.max():5.1f} | {np.median(issue):5.1f}     | {np.median(full):5.1f} {full.max():5.1f} | "
          f"{np.median(job1):5.1f} {job1.max():5.1f} | {np.median(life):5.1f} {life.max():5.1f} | {int(a[:, 6].max())}")
print(f"sum of spans + gaps: {tot:.1f} us")
