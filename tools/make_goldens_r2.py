"""Round-2 golden fixtures: the BASELINE.json configurations at their STATED parameters, generated with the CPU oracle.

    python tools/make_goldens_r2.py lid8000 | pinball100 | cavity_gain | cyl_b1 | cyl_long | cavity_force

* lid8000     lid-driven cavity Re=8000 (mesh64): base flow by Re-continuation exactly as the reference's
              examples/lidcavity/compute_steady_state_increasing_Re.py:71-118 (Picard 10 @1e-7 then Newton 25 per Re in
              [1000,2000,...,7000,7500,8000], each started from the previous Re); then 50 open-loop steps of three of the
              1024 random-vortex trajectories of BASELINE configs[4] (seed 0).
* pinball100  fluidic pinball Re=100, rotation mode: base flow as examples/pinball/run_pinball_rotation_example.py:88-97
              (antisymmetric_bot guess, Picard 15 @1e-7, Newton 10); then 160 steps of Gaussian-pulse rotation (:100-112 with
              per-trajectory amplitudes a_bk ~ U(-2,2), seed 0, SURVEY config 3) for three of the 512 trajectories.
* cavity_gain open cavity Re=7500, static-gain loop u = -k_b y_1 on the Gaussian body-force actuator (SURVEY config 4),
              three of the 256 log-spaced gains, 50 steps.
* cavity_force the same flow, open loop with a prescribed non-zero force amplitude (BDF force actuator with u_ctrl != 0).
* cyl_b1      BASELINE configs[0]: cylinder Re=100, single open-loop trajectory, ParamIC(2,0,0.5,1), 100 steps.
* cyl_long    2000 closed-loop cylinder steps (series of dE, u_ctrl, y1..y3, cl, cd + final fields): north_star's long-run bar.

Outputs go to tests/golden/*.npz; the GPU tests in tests/test_gpu_configs.py read them (never /root/reference).
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import cases  # noqa: E402
from oracle.flow_oracle import FlowOracle, ZOHController, force_coefficients  # noqa: E402

OUT = ROOT / "tests" / "golden"


def _oracle(case):
    xy, tri = cases.load_mesh(case.mesh_file)
    return FlowOracle(case, xy, tri)


# --------------------------------------------------------------------------------------------------------------------
def lid_ics(B=1024):
    """(xloc, yloc) of the 1024 random-vortex initial conditions of config 5 (seed 0)."""
    return np.random.default_rng(0).uniform(0.2, 0.8, size=(B, 2))


LID_PROBES = (0, 511, 1023)


def lid8000():
    t0 = time.time()
    UP = None
    for Re in [1000, 2000, 3000, 4000, 5000, 6000, 7000, 7500, 8000]:
        fo = _oracle(cases.lidcavity(float(Re)))
        if UP is None:
            UP = fo.initial_guess()
        UP = fo.picard(UP, [0.0], max_iter=10, tol=1e-7, log=print)
        UP = fo.newton(UP, [0.0], max_iter=25, log=print)
        print(f"[lid] Re={Re} u0_max={UP[: fo.mesh.Nv].max():.15g} u0_mean={UP[: fo.mesh.Nv].mean():.15g} ({time.time() - t0:.0f}s)", flush=True)
    np.savez_compressed(OUT / "lidcavity_Re8000_baseflow.npz", UP0=UP, Re=8000.0)
    fo.set_base_flow(UP)
    fo.prepare()
    loc = lid_ics()
    nsteps = 50
    ys, dEs, ups = [], [], []
    for b in LID_PROBES:
        fo.case.ic = (float(loc[b, 0]), float(loc[b, 1]), 0.1, 0.1)
        fo.init_time_stepping()
        y, e = [fo.y_meas.copy()], []
        for _ in range(nsteps):
            fo.step([0.0])
            y.append(fo.y_meas.copy())
            e.append(fo.dE)
        ys.append(y); dEs.append(e); ups.append(fo.up.copy())
    np.savez_compressed(OUT / "lidcavity_Re8000_traj.npz", probes=np.array(LID_PROBES), y_meas=np.array(ys), dE=np.array(dEs),
                        up_final=np.array(ups), nsteps=nsteps, dt=fo.case.dt)
    print(f"[lid] done ({time.time() - t0:.0f}s)")


# --------------------------------------------------------------------------------------------------------------------
PIN_PROBES = (0, 255, 511)
PIN_TPEAK = np.array([0.25, 0.5, 0.75])


def pinball_amplitudes(B=512):
    return np.random.default_rng(0).uniform(-2.0, 2.0, size=(3, B))


def pinball_u(t, amp):
    """amp[3] (one trajectory) or amp[3, B]."""
    g = np.exp(-0.5 * (t - PIN_TPEAK) ** 2 / 0.10**2)
    return (amp.T * g).T


def pinball100():
    t0 = time.time()
    case = cases.pinball(100.0, "rotation")
    s = 1.0 / np.sqrt(2.0)
    case.initial_guess = lambda x, y: (np.full_like(x, s), np.full_like(x, -s))  # PinballCustomInitialGuess("antisymmetric_bot")
    case.ic = (2.0, 0.0, 0.5, 1.0)
    fo = _oracle(case)
    bf = OUT / "pinball_Re100_baseflow.npz"
    if bf.exists():
        UP = np.load(bf)["UP0"]
    else:
        UP = fo.picard(fo.initial_guess(), [0.0] * 3, max_iter=15, tol=1e-7, log=print)
        UP = fo.newton(UP, [0.0] * 3, max_iter=10, log=print)
        np.savez_compressed(bf, UP0=UP, Re=100.0)
    print(f"[pinball] base flow u0_max={UP[: fo.mesh.Nv].max():.15g} u0_mean={UP[: fo.mesh.Nv].mean():.15g} ({time.time() - t0:.0f}s)", flush=True)
    fo.set_base_flow(UP)
    fo.prepare()
    amp = pinball_amplitudes()
    nsteps = 160
    ys, dEs, ups, us = [], [], [], []
    for b in PIN_PROBES:
        fo.init_time_stepping()
        y, e, u = [fo.y_meas.copy()], [], []
        for k in range(nsteps):
            uc = pinball_u((k + 1) * case.dt, amp[:, b])
            fo.step(uc)
            y.append(fo.y_meas.copy()); e.append(fo.dE); u.append(uc)
        ys.append(y); dEs.append(e); us.append(u)
        ups.append(fo.up.copy())
        print(f"[pinball] probe {b} done ({time.time() - t0:.0f}s)", flush=True)
    np.savez_compressed(OUT / "pinball_Re100_traj.npz", probes=np.array(PIN_PROBES), y_meas=np.array(ys), dE=np.array(dEs),
                        u_ctrl=np.array(us), up_final=ups[0], up_norms=np.array([np.linalg.norm(u) for u in ups]),
                        up_sample=np.array([u[::97] for u in ups]), nsteps=nsteps, dt=case.dt)
    print(f"[pinball] done ({time.time() - t0:.0f}s)")


# --------------------------------------------------------------------------------------------------------------------
CAV_PROBES = (0, 128, 255)


def cavity_gains(B=256):
    return np.logspace(-3, -1, B)


def _cavity_oracle():
    fo = _oracle(cases.cavity(7500.0))
    fo.set_base_flow(np.load(OUT / "cavity_baseflow.npz")["UP0"])
    fo.prepare()
    return fo


def cavity_gain():
    t0 = time.time()
    fo = _cavity_oracle()
    gains = cavity_gains()
    nsteps = 50
    ys, dEs, us = [], [], []
    for b in CAV_PROBES:
        fo.init_time_stepping()
        y, e, u = [fo.y_meas.copy()], [], []
        for _ in range(nsteps):
            uc = -gains[b] * fo.y_meas[0]  # static gain on the pre-update wall-shear measurement
            fo.step([uc])
            y.append(fo.y_meas.copy()); e.append(fo.dE); u.append(uc)
        ys.append(y); dEs.append(e); us.append(u)
        print(f"[cavity_gain] probe {b} done ({time.time() - t0:.0f}s)", flush=True)
    np.savez_compressed(OUT / "cavity_gain_traj.npz", probes=np.array(CAV_PROBES), gains=gains[list(CAV_PROBES)], y_meas=np.array(ys),
                        dE=np.array(dEs), u_ctrl=np.array(us), nsteps=nsteps)


def cavity_force_u(k, dt=0.0004):
    t = (k + 1) * dt
    return 0.8 * np.sin(2 * np.pi * t / 0.008) + 0.3


def cavity_force():
    t0 = time.time()
    fo = _cavity_oracle()
    fo.init_time_stepping()
    nsteps = 30
    y, e = [fo.y_meas.copy()], []
    for k in range(nsteps):
        fo.step([cavity_force_u(k)])
        y.append(fo.y_meas.copy()); e.append(fo.dE)
    np.savez_compressed(OUT / "cavity_force_traj.npz", y_meas=np.array(y), dE=np.array(e), up_norm=np.linalg.norm(fo.up),
                        up_sample=fo.up[::53], nsteps=nsteps)
    print(f"[cavity_force] done ({time.time() - t0:.0f}s)")


# --------------------------------------------------------------------------------------------------------------------
def _cyl_oracle():
    case = cases.cylinder(100.0)
    case.ic = (2.0, 0.0, 0.5, 1.0)  # ParamIC of run_cylinder_example.py:55
    fo = _oracle(case)
    fo.set_base_flow(np.load(OUT / "cylinder_baseflow.npz")["UP0"])
    fo.init_time_stepping()
    return fo


def cyl_b1():
    fo = _cyl_oracle()
    nsteps = 100
    y, e = [fo.y_meas.copy()], []
    for _ in range(nsteps):
        fo.step([0.0, 0.0])
        y.append(fo.y_meas.copy()); e.append(fo.dE)
    np.savez_compressed(OUT / "cylinder_b1_traj.npz", y_meas=np.array(y), dE=np.array(e), up_final=fo.up, nsteps=nsteps)
    print("[cyl_b1] done")


def cyl_long():
    t0 = time.time()
    fo = _cyl_oracle()
    k = np.load(OUT / "Kopt_reduced13.npz")
    K = ZOHController(k["A"], k["B"], k["C"], k["D"])
    body = lambda x, y: np.hypot(x, y) < 0.6  # noqa: E731
    nsteps = 2000
    ref = np.zeros((nsteps, 7))
    for s in range(nsteps):
        u = K.step(-fo.y_meas[0], fo.case.dt)
        fo.step([u[0], u[0]])
        ref[s, 0] = fo.dE
        ref[s, 1] = u[0]
        ref[s, 2:5] = fo.y_meas
        ref[s, 5:] = force_coefficients(fo.mesh, body, fo.up, 0.01, 1.0, 1.0)
        if s % 200 == 0:
            print(f"[cyl_long] step {s} ({time.time() - t0:.0f}s)", flush=True)
    np.savez_compressed(OUT / "cylinder_long_traj.npz", series=ref, columns=np.array(["dE", "u_ctrl", "y1", "y2", "y3", "cl", "cd"]),
                        up_final=fo.up, nsteps=nsteps)
    print(f"[cyl_long] done ({time.time() - t0:.0f}s)")


if __name__ == "__main__":
    for name in sys.argv[1:]:
        globals()[name]()
