"""First-light check on a B200: cylinder ensemble, CUDA path vs the CPU oracle.

    python tools/gpu_check.py [B] [nsteps]

Builds the cylinder Re=100 problem from the committed mesh + base-flow fixtures,
steps B trajectories closed-loop with fcb_step and compares trajectory 0 against the
oracle stepped on the host; then prints per-phase device times.
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from flowcontrol_b200.actuator import ActuatorBCParabolicV  # noqa: E402
from flowcontrol_b200.controller import Controller, ControllerBank  # noqa: E402
from flowcontrol_b200.ensemble import Ensemble  # noqa: E402
from flowcontrol_b200.fem import ScalarBlocks  # noqa: E402
from flowcontrol_b200.mesh import TaylorHoodTables  # noqa: E402
from flowcontrol_b200.problem import DirichletBC, FlowProblem  # noqa: E402
from flowcontrol_b200.sensor import SENSOR_TYPE, SensorPoint  # noqa: E402


def near(a, b, tol=3e-16):
    return np.abs(a - b) <= tol


def build_cylinder_problem(leaf_cells=16, top_levels=2, time_scheme="bdf"):
    tab = TaylorHoodTables.from_file(ROOT / "data/meshes/cylinder_O1.npz")
    blocks = ScalarBlocks(tab)
    r = 0.5
    L = ActuatorBCParabolicV.angular_size_deg_to_width(10, r)
    acts = [ActuatorBCParabolicV(width=L, position_x=0.0), ActuatorBCParabolicV(width=L, position_x=0.0)]
    close = lambda x, y: (x >= -r) & (x <= r) & (y >= -r) & (y <= r)
    bcs = [
        DirichletBC(lambda x, y: near(x, -10.0), (0, 1), (0.0, 0.0)),
        DirichletBC(lambda x, y: near(y, -10.0) | near(y, 10.0), (1,), (0.0,)),
        DirichletBC(lambda x, y: close(x, y) & (((x >= -r) & (x <= -L)) | ((x >= L) & (x <= r))), (0, 1), (0.0, 0.0)),
        DirichletBC(lambda x, y: close(x, y) & (x >= -L - 0.01) & (x <= L + 0.01) & (y >= 0) & (y <= r), (0, 1), acts[0]),
        DirichletBC(lambda x, y: close(x, y) & (x >= -L - 0.01) & (x <= L + 0.01) & (y >= -r) & (y <= 0), (0, 1), acts[1]),
    ]
    sensors = [SensorPoint(sensor_type=SENSOR_TYPE.V, position=np.array(p)) for p in ((3.0, 0.0), (3.1, 1.0), (3.1, -1.0))]
    UP0 = np.load(ROOT / "tests/golden/cylinder_baseflow.npz")["UP0"]
    t0 = time.time()
    prob = FlowProblem(tab, blocks, 100.0, 0.005, bcs, acts, sensors, UP0, leaf_cells=leaf_cells, top_levels=top_levels,
                       time_scheme=time_scheme)
    print(f"problem setup {time.time() - t0:.1f}s  n_free={prob.sym.n} factor entries={prob.sym.factor_entries() / 1e6:.2f}M "
          f"launches={len(prob.plans[2].launch_ptr) - 1}", flush=True)
    return prob, UP0


def default_ic(tab, UP0, xloc=0.0, yloc=0.0, radius=1.0, amp=1.0):
    x, y = tab.node_xy[:, 0], tab.node_xy[:, 1]
    psi = 0.25 * np.exp(-0.5 * ((x - xloc) ** 2 + (y - yloc) ** 2) / radius**2)
    ic = np.zeros(tab.N)
    ic[: tab.nN] = amp * psi * (-(y - yloc) / radius**2)
    ic[tab.nN : tab.Nv] = -amp * psi * (-(x - xloc) / radius**2)
    ic[tab.Nv :] = amp * UP0[tab.Nv :]
    return ic


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    import os
    scheme = os.environ.get("FCB_SCHEME", "bdf")
    prob, UP0 = build_cylinder_problem(leaf_cells=int(os.environ.get("FCB_LEAF", "16")), top_levels=int(os.environ.get("FCB_TOP", "2")),
                                       time_scheme=scheme)
    tab = prob.tab
    ic = default_ic(tab, UP0)
    ens = Ensemble(prob, B)
    y0 = ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1).copy()
    print("y0", y0[:, 0], "dE0", ens.dE[0])
    gold = np.load(ROOT / "tests/golden/cylinder_traj.npz")
    print("gold y0", gold["y_meas"][0], "dE0", gold["dE"][0])
    k = np.load(ROOT / "tests/golden/Kopt_reduced13.npz")
    K = Controller(k["A"], k["B"], k["C"], k["D"])
    worst = 0.0
    for i in range(nsteps):
        u = K.step(-ens.y_meas[0, 0], prob.dt)
        uc = np.full((2, B), u[0])
        ens.step(uc)
        if i + 1 < len(gold["y_meas"]):
            ey = np.abs(ens.y_meas[:, 0] - gold["y_meas"][i + 1]).max() / np.abs(gold["y_meas"][i + 1]).max()
            eE = abs(ens.dE[0] - gold["dE"][i + 1]) / gold["dE"][i + 1]
            worst = max(worst, ey, eE)
            print(f"step {i + 1}: u={u[0]:.6e} y={ens.y_meas[:, 0]} dE={ens.dE[0]:.15g} rel.err y {ey:.2e} dE {eE:.2e}")
    print("worst rel err vs oracle golden trajectory:", worst, " diverged:", int(ens.diverged.sum()))
    if "up_final" in gold.files and nsteps == len(gold["y_meas"]) - 1:
        up = ens.fields(0)[:, 0]
        print("final field rel L2 err:", np.linalg.norm(up - gold["up_final"]) / np.linalg.norm(gold["up_final"]))
        print("spread across trajectories:", np.abs(ens.fields(0) - up[:, None]).max())
    prof = [ens.profile_step(uc) for _ in range(5)][-1]
    tot = sum(v["ms"] for v in prof.values())
    for name, v in prof.items():
        print(f"  phase {name:9s} {v['ms']:8.3f} ms  launches {v['launches']}")
    print(f"  total {tot:.3f} ms -> {B / tot * 1e3:.0f} trajectory-steps/s (un-graphed, with phase events)")
    if scheme == "cn":
        # k_spmm: algorithmic bytes = CSR (12 nnz + 4(n+1)) read once + 8 B per input row and output row per trajectory
        ldb = 32 if B <= 32 else 64 if B <= 64 else -(-B // 128) * 128
        nnz, n = prob.E_cn.nnz, prob.sym.n
        by = 12 * nnz + 4 * (n + 1) + 8 * (tab.Nv + n) * ldb
        ms = prof["spmm"]["ms"]
        print(f"  k_spmm: nnz {nnz} rows {n}: {by / 1e6:.1f} MB algorithmic in {ms:.4f} ms -> {by / ms / 1e6:.0f} GB/s "
              f"({by / ms / 1e6 / 6536:.1%} of the measured 6536 GB/s HBM peak)")
    # graph-replayed steps with device-resident inputs
    import torch

    ucd = torch.zeros((2, B), dtype=torch.float64, device="cuda")
    for _ in range(5):
        ens.step_device(ucd)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 50
    for _ in range(n):
        ens.step_device(ucd)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"graph step (device inputs, wall): {dt * 1e3:.3f} ms/step -> {B / dt:.0f} trajectory-steps/s; launches so far {ens.launch_count()}")
