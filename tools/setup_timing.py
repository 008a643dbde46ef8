"""Setup-time phases on the GPU box: symbolic analysis, numeric factorisation on the host (numpy / LAPACK per front) vs on
the device (fcb_factorize), plan compilation.   python tools/setup_timing.py [cylinder pinball]"""
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from flowcontrol_b200.devfactor import DeviceBlockFactor  # noqa: E402
from flowcontrol_b200.flowfield import Field  # noqa: E402
from flowcontrol_b200.multifrontal import BlockFactor, SymbolicFactor, build_plan  # noqa: E402
from flowcontrol_b200.problem import DirichletSet  # noqa: E402


def case(name):
    if name == "cylinder":
        from flowcontrol_b200.examples.cylinder import CylinderFlowSolver

        fs = CylinderFlowSolver.make_default(path_out=Path(tempfile.mkdtemp()))
        UP0 = np.load(ROOT / "tests/golden/cylinder_baseflow.npz")["UP0"]
        return fs, UP0, 100.0, 0.005
    from flowcontrol_b200.actuator import CYLINDER_ACTUATION_MODE
    from flowcontrol_b200.examples.pinball import PinballFlowSolver

    fs = PinballFlowSolver.make_default(Re=100.0, mode_actuation=CYLINDER_ACTUATION_MODE.ROTATION, path_out=Path(tempfile.mkdtemp()))
    return fs, np.load(ROOT / "tests/golden/pinball_Re100_baseflow.npz")["UP0"], 100.0, 0.005


for name in sys.argv[1:] or ["cylinder", "pinball"]:
    fs, UP0, Re, dt = case(name)
    tab = fs.tables
    dset = DirichletSet(tab, fs.bc.bcu, fs.params_control.actuator_list)
    t0 = time.time(); sym = SymbolicFactor(tab, dset.free, leaf_cells=16); t_sym = time.time() - t0
    A = fs.blocks.saddle_point(1.5 / dt, Re, UP0[: tab.Nv])
    t0 = time.time(); fh = BlockFactor(sym, A); t_host = time.time() - t0
    t0 = time.time(); fd = DeviceBlockFactor(sym, A); t_dev1 = time.time() - t0
    t0 = time.time(); fd2 = DeviceBlockFactor(sym, A, maps=fd.maps); t_dev2 = time.time() - t0
    t0 = time.time(); build_plan(fd2); t_plan = time.time() - t0
    b = np.random.default_rng(0).standard_normal(sym.n)
    err = np.linalg.norm(fd2.solve(b) - fh.solve(b)) / np.linalg.norm(fh.solve(b))
    print(f"{name}: N={tab.N} fronts={len(sym.supernodes)} factor entries={sym.factor_entries() / 1e6:.1f}M | symbolic {t_sym:.1f}s | numeric host {t_host:.2f}s | "
          f"numeric device {t_dev1:.2f}s (maps built) / {t_dev2:.2f}s (maps reused) | plan {t_plan:.1f}s | solve rel diff {err:.1e}", flush=True)
