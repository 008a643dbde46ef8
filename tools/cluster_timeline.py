"""Per-CTA timeline of the cluster sweeps from an FCB_CLUSTER_DEBUG dump (globaltimer ns):
    FCB_CLUSTER_DEBUG=gpurun_out/cl_dbg.bin python tools/gpu_check.py 256 4 ; python tools/cluster_timeline.py gpurun_out/cl_dbg.bin
stamps per CTA: 0 entry, 1 program arrived, 2 resident vector loaded, 3 first A chunk arrived, 4 operations done, 5 exit; 6 = number of operations."""
import struct
import sys

import numpy as np

raw = open(sys.argv[1], "rb").read()
(nt,) = struct.unpack_from("i", raw, 0)
off = 4
for t in range(nt):
    for d in ("fwd", "bwd"):
        (nc,) = struct.unpack_from("i", raw, off)
        off += 4
        a = np.frombuffer(raw, dtype=np.uint64, count=nc * 8, offset=off).reshape(nc, 8).astype(np.int64)
        off += nc * 64
        if nc == 0 or a[:, 0].max() == 0:
            continue
        us = lambda x: x / 1e3  # noqa: E731
        t0 = a[:, 0].min()
        span = us(a[:, 5].max() - t0)
        prog, load, first, ops, out = us(a[:, 1] - a[:, 0]), us(a[:, 2] - a[:, 1]), us(a[:, 3] - a[:, 2]), us(a[:, 4] - a[:, 3]), us(a[:, 5] - a[:, 4])
        life = us(a[:, 5] - a[:, 0])
        nops = a[:, 6]
        print(f"tier {t} {d}: {nc} CTAs, span {span:.1f} us | life med {np.median(life):.1f} max {life.max():.1f} | program {np.median(prog):.2f} | "
              f"load {np.median(load):.2f} | first chunk {np.median(first):.2f} | ops {np.median(ops):.2f} (max {ops.max():.1f}, {np.median(nops):.0f} ops -> "
              f"{np.median(ops / np.maximum(nops, 1)):.2f} us/op) | out {np.median(out):.2f}")
