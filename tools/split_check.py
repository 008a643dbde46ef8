"""Does splitting the ensemble into S concurrent sub-ensembles (own handle + stream each) raise throughput?
    python tools/split_check.py [B_total] [nsplit] [nsteps]"""
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tools.gpu_check import build_cylinder_problem, default_ic  # noqa: E402
from flowcontrol_b200.controller import Controller, ControllerBank  # noqa: E402
from flowcontrol_b200.ensemble import Ensemble  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = int(sys.argv[3]) if len(sys.argv) > 3 else 100
prob, UP0 = build_cylinder_problem()
tab = prob.tab
ic = default_ic(tab, UP0)
k = np.load(ROOT / "tests/golden/Kopt_reduced13.npz")
ens = []
for s in range(S):
    b = B // S
    e = Ensemble(prob, b)
    e.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    bank = ControllerBank([Controller(k["A"], k["B"], k["C"], k["D"]) for _ in range(b)], prob.dt,
                          Ky=np.array([[-1.0, 0.0, 0.0]]), Fu=np.array([[1.0], [1.0]]))
    e.set_controllers(bank)
    e.run_closed_loop(5, log=False)
    ens.append(e)


def work(e):
    e.run_closed_loop(n, log=False)


for rep in range(3):
    th = [threading.Thread(target=work, args=(e,)) for e in ens]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    print(f"B={B} split={S}: {dt / n * 1e3:.3f} ms per step of the whole ensemble -> {B * n / dt:.0f} trajectory-steps/s")
