"""north_star tolerance over a long run: sensor and lift time series of 2000 closed-loop steps (cylinder Re=100, shipped
controller) on the device against the CPU oracle stepped on the host, plus the final fields.

    python tools/long_run_check.py [nsteps] [out.json]

Takes ~4 minutes on the GPU box (the oracle does ~10 steps/s on one core).  Result recorded in profiles/."""
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from flowcontrol_b200.controller import Controller, ControllerBank  # noqa: E402
from flowcontrol_b200.ensemble import Ensemble  # noqa: E402
from flowcontrol_b200.examples.cylinder import CylinderFlowSolver  # noqa: E402
from flowcontrol_b200.flowfield import Field  # noqa: E402
from flowcontrol_b200.problem import FlowProblem  # noqa: E402
from oracle import cases  # noqa: E402
from oracle.flow_oracle import FlowOracle, ZOHController, force_coefficients  # noqa: E402

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
out = Path(sys.argv[2]) if len(sys.argv) > 2 else ROOT / "gpurun_out" / "r01_long_run.json"
UP0 = np.load(ROOT / "tests/golden/cylinder_baseflow.npz")["UP0"]
fs = CylinderFlowSolver.make_default(path_out=Path(tempfile.mkdtemp()))
tab = fs.tables
fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
sensors = list(fs.params_control.sensor_list) + fs.force_sensors()
prob = FlowProblem(tab, fs.blocks, 100.0, 0.005, fs.bc.bcu, fs.params_control.actuator_list, sensors, UP0)
ic = fs._default_initial_perturbation(2.0, 0.0, 0.5)  # ParamIC of run_cylinder_example.py:55
B = 32
k = np.load(ROOT / "tests/golden/Kopt_reduced13.npz")
ens = Ensemble(prob, B)
ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
ens.set_controllers(ControllerBank([Controller(k["A"], k["B"], k["C"], k["D"]) for _ in range(B)], prob.dt,
                                   np.array([[-1.0, 0.0, 0.0, 0.0, 0.0]]), np.array([[1.0], [1.0]])))
t0 = time.time()
series = ens.run_closed_loop(nsteps)  # columns: dE, u1, u2, y1..y3, cl, cd
t_gpu = time.time() - t0
case = cases.cylinder(100.0)
case.ic = (2.0, 0.0, 0.5, 1.0)
xy, tri = cases.load_mesh(case.mesh_file)
orc = FlowOracle(case, xy, tri)
orc.set_base_flow(UP0)
orc.init_time_stepping()
K = ZOHController(k["A"], k["B"], k["C"], k["D"])
body = lambda x, y: np.hypot(x, y) < 0.6  # noqa: E731
ref = np.zeros((nsteps, 7))
t0 = time.time()
for s in range(nsteps):
    u = K.step(-orc.y_meas[0], prob.dt)
    orc.step([u[0], u[0]])
    ref[s, 0] = orc.dE
    ref[s, 1] = u[0]
    ref[s, 2:5] = orc.y_meas
    ref[s, 5:] = force_coefficients(orc.mesh, body, orc.up, 0.01, 1.0, 1.0)
t_cpu = time.time() - t0
got = series[:, [0, 1, 3, 4, 5, 6, 7], 0]
names = ["dE", "u_ctrl", "y1", "y2", "y3", "cl", "cd"]
err = np.abs(got - ref).max(axis=0) / np.abs(ref).max(axis=0)
up = ens.fields(0)[:, 0]
rec = {
    "nsteps": nsteps, "B": B, "series_max_rel_err": dict(zip(names, map(float, err))),
    "field_rel_l2_err": {"u": float(np.linalg.norm(up[: tab.Nv] - orc.up[: tab.Nv]) / np.linalg.norm(orc.up[: tab.Nv])),
                         "p": float(np.linalg.norm(up[tab.Nv :] - orc.up[tab.Nv :]) / np.linalg.norm(orc.up[tab.Nv :]))},
    "identical_trajectories_bit_identical": bool(np.abs(series - series[:, :, :1]).max() == 0.0),
    "tolerances": {"series": 1e-6, "fields": 1e-9}, "gpu_seconds_32_trajectories": t_gpu, "oracle_seconds_1_trajectory": t_cpu,
}
rec["within_tolerance"] = bool(max(err) < 1e-6 and max(rec["field_rel_l2_err"].values()) < 1e-9)
out.parent.mkdir(exist_ok=True)
out.write_text(json.dumps(rec, indent=1))
print(json.dumps(rec, indent=1))
