"""Experiment: k independent sub-ensembles (handles) of B/k trajectories on ONE GPU, each on its own stream, stepped
concurrently from k host threads.  Prints aggregate trajectory-steps/s for a total of B trajectories.

    python tools/concurrent_check.py [B_total=256] [nsteps=200] [ks=1,2,4]
"""
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import build_problem, controller_bank  # noqa: E402
from flowcontrol_b200.ensemble import Ensemble  # noqa: E402

Btot = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
ks = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1,2,4").split(",")]
fs, prob = build_problem()
tab = prob.tab
ic = fs._default_initial_perturbation()
import torch  # noqa: E402

for k in ks:
    Bk = Btot // k
    ens = []
    for i in range(k):
        e = Ensemble(prob, Bk)
        e.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
        e.set_controllers(controller_bank(prob, i * Bk, (i + 1) * Bk, Btot))
        e.run_closed_loop(8, log=False)
        ens.append(e)
    torch.cuda.synchronize()
    best = None
    for rep in range(3):
        bar = threading.Barrier(k + 1)
        def work(e):
            bar.wait()
            e.run_closed_loop(nsteps, log=False)
        th = [threading.Thread(target=work, args=(e,)) for e in ens]
        for t in th:
            t.start()
        bar.wait()
        t0 = time.perf_counter()
        for t in th:
            t.join()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print(f"k={k} sub-ensembles x {Bk} trajectories: {best / nsteps * 1e3:.3f} ms per step of all {Btot} -> {Btot * nsteps / best:.0f} trajectory-steps/s", flush=True)
    for e in ens:
        e.close()
