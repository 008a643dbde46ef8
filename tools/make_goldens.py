"""Generate golden fixtures under tests/golden/ with the CPU oracle.

For each benchmark configuration of the reference's regression tests
(tests/integration/test_{cylinder,cavity,lidcavity,pinball}.py) this runs the
oracle's base-flow recipe and the test's short trajectory, then stores

* ``<case>_baseflow.npz``  : UP0 in canonical numbering [ux|uy|p]
* ``<case>_traj.npz``      : per-step u_ctrl, y_meas, dE, and final-state summaries
                             (full final (u,p) only for the two small meshes)

The reference's own golden constants are checked by tests/test_oracle_goldens.py
against these files; the script prints the comparison too.

    python tools/make_goldens.py [cylinder cavity lidcavity pinball]
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import cases  # noqa: E402
from oracle.flow_oracle import FlowOracle, ZOHController  # noqa: E402

OUT = ROOT / "tests" / "golden"

# Kopt_reduced13.mat is data shipped by the reference (cylinder/data_input); its four
# matrices are stored as a fixture so the GPU box does not need /root/reference.
KOPT = OUT / "Kopt_reduced13.npz"


def export_controller():
    import scipy.io

    m = scipy.io.loadmat("/root/reference/src/examples/cylinder/data_input/Kopt_reduced13.mat")
    np.savez(KOPT, A=m["A"], B=m["B"], C=m["C"], D=m["D"])


RECIPES = {
    # name: (case factory, picard iters, newton iters, nsteps, closed loop)
    "cylinder": (lambda: cases.cylinder(100.0), 3, 25, 20, True),
    "cavity": (lambda: cases.cavity(7500.0), 10, 10, 10, False),
    "lidcavity": (lambda: cases.lidcavity(1000.0), 40, 0, 10, False),
    "pinball": (lambda: cases.pinball(30.0, "suction"), 15, 10, 10, False),
}


def run(name):
    factory, npic, nnewt, nsteps, closed = RECIPES[name]
    case = factory()
    xy, tri = cases.load_mesh(case.mesh_file)
    t0 = time.time()
    fo = FlowOracle(case, xy, tri)
    na = len(case.actuators)
    zero = [0.0] * na
    UP = fo.initial_guess()
    UP = fo.picard(UP, zero, max_iter=npic, tol=1e-7, log=print)
    if nnewt:
        UP = fo.newton(UP, zero, max_iter=nnewt, log=print)
    fo.set_base_flow(UP)
    U0 = UP[: fo.mesh.Nv]
    print(f"[{name}] base flow u0_max={U0.max():.16g} u0_mean={U0.mean():.16g}  ({time.time() - t0:.0f}s)")
    np.savez_compressed(OUT / f"{name}_baseflow.npz", UP0=UP, Re=case.Re)
    y0 = fo.init_time_stepping()
    K = None
    if closed:
        k = np.load(KOPT)
        K = ZOHController(k["A"], k["B"], k["C"], k["D"])
    us, ys, dEs = [], [y0.copy()], [0.5 * fo.u_n @ (fo.ops.Mv @ fo.u_n)]
    for _ in range(nsteps):
        if closed:
            u = K.step(-fo.y_meas[0], case.dt)
            uc = [u[0]] * na
        else:
            uc = zero
        fo.step(uc)
        us.append(uc)
        ys.append(fo.y_meas.copy())
        dEs.append(fo.dE)
    U = fo.full_velocity()
    print(f"[{name}] u_max={U.max():.16g} u_mean={U.mean():.16g} y={fo.y_meas} dE={fo.dE:.16g} ({time.time() - t0:.0f}s)")
    out = dict(
        u_ctrl=np.array(us), y_meas=np.array(ys), dE=np.array(dEs), dt=case.dt,
        u_max=U.max(), u_mean=U.mean(), u0_max=U0.max(), u0_mean=U0.mean(),
        up_norm=np.linalg.norm(fo.up), up_sum=fo.up.sum(),
    )
    if fo.mesh.N < 60000:
        out["up_final"] = fo.up
    np.savez_compressed(OUT / f"{name}_traj.npz", **out)


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    if Path("/root/reference").exists():
        export_controller()
    for n in sys.argv[1:] or list(RECIPES):
        run(n)
