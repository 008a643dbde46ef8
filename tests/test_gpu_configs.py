"""Every BASELINE.json configuration at its STATED parameters and full ensemble width, through the C ABI, against oracle
trajectories committed as fixtures (tools/make_goldens_r2.py wrote them with oracle/flow_oracle.py; nothing here reads
/root/reference or runs the oracle).  Tolerances are BASELINE.json's: sensor / energy series within 1e-6 relative,
velocity / pressure fields within 1e-9 relative L2.

    configs[0]  cylinder Re=100, single open-loop trajectory            test_config0_cylinder_single_trajectory_100_steps
    configs[1]  cylinder Re=100 closed loop, 256 controllers            tests/test_gpu_parity.py (golden closed loop, gain sweep)
                                                                        + test_long_run_2000_steps (north_star's 2000-step bar)
    configs[2]  pinball Re=100, 3 rotation actuators, B=512             test_config2_pinball_re100_rotation_512
    configs[3]  open cavity Re=7500, force actuator + wall shear, B=256 test_config3_cavity_static_gains_256,
                                                                        test_cavity_bdf_force_actuator_nonzero_control
    configs[4]  lid-driven cavity Re=8000, B=1024 open loop             test_config4_lidcavity_re8000_1024
"""
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELD_TOL = 1e-9
SERIES_TOL = 1e-6
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


def series_err(got, ref):
    """max |got - ref| over the series, relative to the largest |ref| of each column (got, ref: [steps, ncol])."""
    got, ref = np.asarray(got), np.asarray(ref)
    return float((np.abs(got - ref).max(axis=0) / np.abs(ref).max(axis=0)).max())


def _problem(fs, Re, dt, UP0, **kw):
    from flowcontrol_b200.flowfield import Field
    from flowcontrol_b200.problem import FlowProblem

    tab = fs.tables
    fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
    sensors = kw.pop("sensors", fs.params_control.sensor_list)
    return FlowProblem(tab, fs.blocks, Re, dt, fs.bc.bcu, fs.params_control.actuator_list, sensors, UP0, **kw)


def test_config0_cylinder_single_trajectory_100_steps(root, built_lib):
    """BASELINE configs[0]: cylinder Re=100, ONE open-loop trajectory (B = 1, the latency case), ParamIC(2, 0, 0.5, 1)
    of run_cylinder_example.py:55, 100 steps, stepped through fcb_run_open_loop."""
    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.examples.cylinder import CylinderFlowSolver

    UP0 = np.load(root / "tests/golden/cylinder_baseflow.npz")["UP0"]
    gold = np.load(root / "tests/golden/cylinder_b1_traj.npz")
    fs = CylinderFlowSolver.make_default(path_out=Path(tempfile.mkdtemp()))
    prob = _problem(fs, 100.0, 0.005, UP0)
    tab = prob.tab
    ic = fs._default_initial_perturbation(2.0, 0.0, 0.5)
    ens = Ensemble(prob, 1)
    y0 = ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    assert np.allclose(y0[:, 0], gold["y_meas"][0], rtol=1e-9)
    n = int(gold["nsteps"])
    series = ens.run_open_loop(np.zeros((n, 2, 1)))  # columns dE, u1, u2, y1..y3
    assert series_err(series[:, 3:, 0], gold["y_meas"][1:]) < SERIES_TOL
    assert series_err(series[:, :1, 0], gold["dE"][:, None]) < SERIES_TOL
    up = ens.fields(0)[:, 0]
    assert rel(up[: tab.Nv], gold["up_final"][: tab.Nv]) < FIELD_TOL
    assert rel(up[tab.Nv :], gold["up_final"][tab.Nv :]) < FIELD_TOL
    assert not ens.diverged.any()
    ens.close()


def test_long_run_2000_steps(root, built_lib):
    """north_star: "sensor/lift time series within 1e-6 relative over 2000 steps".  2000 closed-loop steps of the cylinder
    (shipped 13-state controller on the device, lift/drag rows logged every step) against the oracle's series; the run is
    longer than one series chunk, so the chunked streaming of fcb_run_closed_loop is on the path."""
    from flowcontrol_b200.controller import Controller, ControllerBank
    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.examples.cylinder import CylinderFlowSolver

    UP0 = np.load(root / "tests/golden/cylinder_baseflow.npz")["UP0"]
    gold = np.load(root / "tests/golden/cylinder_long_traj.npz")
    fs = CylinderFlowSolver.make_default(path_out=Path(tempfile.mkdtemp()))
    prob = _problem(fs, 100.0, 0.005, UP0, sensors=list(fs.params_control.sensor_list) + fs.force_sensors())
    tab = prob.tab
    ic = fs._default_initial_perturbation(2.0, 0.0, 0.5)
    B = 32
    k = np.load(root / "tests/golden/Kopt_reduced13.npz")
    ens = Ensemble(prob, B)
    ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    ens.set_controllers(ControllerBank([Controller(k["A"], k["B"], k["C"], k["D"]) for _ in range(B)], prob.dt,
                                       np.array([[-1.0, 0.0, 0.0, 0.0, 0.0]]), np.array([[1.0], [1.0]])))
    n = int(gold["nsteps"])
    assert n == 2000
    series = ens.run_closed_loop(n)  # columns: dE, u1, u2, y1..y3, cl, cd
    got = series[:, [0, 1, 3, 4, 5, 6, 7], 0]
    assert series_err(got, gold["series"]) < SERIES_TOL
    up = ens.fields(0)[:, 0]
    assert rel(up[: tab.Nv], gold["up_final"][: tab.Nv]) < FIELD_TOL
    assert rel(up[tab.Nv :], gold["up_final"][tab.Nv :]) < FIELD_TOL
    assert np.abs(series - series[:, :, :1]).max() == 0.0  # identical controllers: bit-identical trajectories
    ens.close()


def test_config2_pinball_re100_rotation_512(root, built_lib):
    """BASELINE configs[2]: fluidic pinball Re=100, three ActuatorBCRotation (actuator.py:225-252; lifting columns of the
    three cylinder surfaces), B = 512 trajectories with Gaussian-pulse rotation u_bk(t) = a_bk exp(-(t-t_k)^2 / 2 0.1^2),
    t_k = 0.25, 0.5, 0.75 (run_pinball_rotation_example.py:100-112), a_bk ~ U(-2,2) seed 0, 160 steps (all three pulses),
    base flow computed at Re=100 by the example's own recipe; three trajectories against the oracle."""
    import make_goldens_r2 as mk
    from flowcontrol_b200.actuator import CYLINDER_ACTUATION_MODE
    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.examples.pinball import PinballFlowSolver

    UP0 = np.load(root / "tests/golden/pinball_Re100_baseflow.npz")["UP0"]
    gold = np.load(root / "tests/golden/pinball_Re100_traj.npz")
    fs = PinballFlowSolver.make_default(Re=100.0, mode_actuation=CYLINDER_ACTUATION_MODE.ROTATION, path_out=Path(tempfile.mkdtemp()))
    prob = _problem(fs, 100.0, 0.005, UP0)
    tab = prob.tab
    ic = fs._default_initial_perturbation(2.0, 0.0, 0.5)
    B, n = 512, int(gold["nsteps"])
    amp = mk.pinball_amplitudes(B)
    useries = np.stack([mk.pinball_u((k + 1) * 0.005, amp) for k in range(n)])  # [n, 3, B]
    probes = gold["probes"]
    assert np.allclose(useries[:, :, probes].transpose(2, 0, 1), gold["u_ctrl"], rtol=0, atol=0)
    ens = Ensemble(prob, B)
    y0 = ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    assert np.allclose(y0[:, 0], gold["y_meas"][0, 0], rtol=1e-9)
    series = ens.run_open_loop(useries)  # columns dE, u1..u3, y1..y3
    up = ens.fields(0)
    for i, b in enumerate(probes):
        assert series_err(series[:, 4:, b], gold["y_meas"][i, 1:]) < SERIES_TOL
        assert series_err(series[:, :1, b], gold["dE"][i][:, None]) < SERIES_TOL
        assert np.isclose(np.linalg.norm(up[:, b]), gold["up_norms"][i], rtol=1e-9)
        assert rel(up[::97, b], gold["up_sample"][i]) < FIELD_TOL
    assert rel(up[: tab.Nv, probes[0]], gold["up_final"][: tab.Nv]) < FIELD_TOL
    assert rel(up[tab.Nv :, probes[0]], gold["up_final"][tab.Nv :]) < FIELD_TOL
    assert np.isfinite(series).all() and not ens.diverged.any()
    ens.close()


@pytest.fixture(scope="module")
def cavity_problem(root, built_lib):
    from flowcontrol_b200.examples.cavity import CavityFlowSolver

    UP0 = np.load(root / "tests/golden/cavity_baseflow.npz")["UP0"]
    fs = CavityFlowSolver.make_default(path_out=Path(tempfile.mkdtemp()))
    return fs, _problem(fs, 7500.0, 0.0004, UP0)


def test_config3_cavity_static_gains_256(root, cavity_problem):
    """BASELINE configs[3]: open cavity Re=7500, Gaussian body-force actuator + wall-shear sensor, B = 256 static-gain loops
    u = -k_b y_1 (k_b log-spaced 1e-3..1e-1) run on the device; gains 0, 128, 255 vs the oracle."""
    import make_goldens_r2 as mk
    from flowcontrol_b200.controller import Controller, ControllerBank
    from flowcontrol_b200.ensemble import Ensemble

    fs, prob = cavity_problem
    tab = prob.tab
    gold = np.load(root / "tests/golden/cavity_gain_traj.npz")
    B, n = 256, int(gold["nsteps"])
    gains = mk.cavity_gains(B)
    assert np.array_equal(gains[gold["probes"]], gold["gains"])
    ctrls = [Controller(np.array([[-1.0]]), np.zeros((1, 1)), np.zeros((1, 1)), np.array([[-g]])) for g in gains]
    Ky = np.zeros((1, prob.ns))
    Ky[0, 0] = 1.0
    ic = fs._default_initial_perturbation()
    ens = Ensemble(prob, B)
    ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    ens.set_controllers(ControllerBank(ctrls, prob.dt, Ky, np.ones((prob.na, 1))))
    series = ens.run_closed_loop(n)  # columns dE, u, y1, y2
    for i, b in enumerate(gold["probes"]):
        assert series_err(series[:, 2:, b], gold["y_meas"][i, 1:]) < SERIES_TOL
        assert series_err(series[:, :1, b], gold["dE"][i][:, None]) < SERIES_TOL
        assert series_err(series[:, 1:2, b], gold["u_ctrl"][i][:, None]) < SERIES_TOL
    assert not ens.diverged.any()
    ens.close()


def test_cavity_bdf_force_actuator_nonzero_control(root, cavity_problem):
    """ActuatorForceGaussianV with u_ctrl != 0 through the BDF path (rhs += F u_ctrl, actuator.py:297-312): 30 steps of a
    prescribed force amplitude, trajectory-dependent scaling, trajectory 5 (scale 1) vs the oracle."""
    import make_goldens_r2 as mk
    from flowcontrol_b200.ensemble import Ensemble

    fs, prob = cavity_problem
    tab = prob.tab
    gold = np.load(root / "tests/golden/cavity_force_traj.npz")
    B, n, track = 32, int(gold["nsteps"]), 5
    scale = np.linspace(-1.5, 2.0, B)
    scale[track] = 1.0
    ic = fs._default_initial_perturbation()
    ens = Ensemble(prob, B)
    ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    ys, es = [], []
    for k in range(n):
        ens.step((mk.cavity_force_u(k) * scale)[None, :])
        ys.append(ens.y_meas[:, track].copy())
        es.append(ens.dE[track])
    assert series_err(np.array(ys), gold["y_meas"][1:]) < SERIES_TOL
    assert series_err(np.array(es)[:, None], gold["dE"][:, None]) < SERIES_TOL
    up = ens.fields(0)
    assert np.isclose(np.linalg.norm(up[:, track]), float(gold["up_norm"]), rtol=1e-9)
    assert rel(up[::53, track], gold["up_sample"]) < FIELD_TOL
    assert rel(up[:, 0], up[:, track]) > 1e-3  # the force amplitude matters: other scalings give other fields
    assert not ens.diverged.any()
    ens.close()


def test_config4_lidcavity_re8000_1024(root, built_lib):
    """BASELINE configs[4]: lid-driven cavity Re=8000 (base flow by the Re-continuation of
    compute_steady_state_increasing_Re.py:71-118), B = 1024 open-loop trajectories from random Gaussian-vortex initial
    conditions ParamIC(xloc, yloc ~ U(0.2, 0.8), radius 0.1, amplitude 0.1), seed 0; trajectories 0, 511, 1023 vs the oracle."""
    import make_goldens_r2 as mk
    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.examples import lidcavity as ex
    from flowcontrol_b200.flowfield import Field

    UP0 = np.load(root / "tests/golden/lidcavity_Re8000_baseflow.npz")["UP0"]
    gold = np.load(root / "tests/golden/lidcavity_Re8000_traj.npz")
    prob = ex.make_problem(Re=8000.0, UP0=UP0)
    tab = prob.tab
    fs = ex.LidCavityFlowSolver.make_default(Re=8000.0, path_out=Path(tempfile.mkdtemp()))
    fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
    B, n = 1024, int(gold["nsteps"])
    loc = mk.lid_ics(B)
    ic = np.stack([0.1 * fs._default_initial_perturbation(xloc=x, yloc=y, radius=0.1) for x, y in loc], axis=1)
    ens = Ensemble(prob, B)
    y0 = ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    probes = gold["probes"]
    for i, b in enumerate(probes):
        assert np.allclose(y0[:, b], gold["y_meas"][i, 0], rtol=1e-9, atol=1e-15)
    series = ens.run_open_loop(np.zeros((n, 1, B)))  # columns dE, u, y1, y2
    up = ens.fields(0)
    for i, b in enumerate(probes):
        assert series_err(series[:, 2:, b], gold["y_meas"][i, 1:]) < SERIES_TOL
        assert series_err(series[:, :1, b], gold["dE"][i][:, None]) < SERIES_TOL
        assert rel(up[: tab.Nv, b], gold["up_final"][i][: tab.Nv]) < FIELD_TOL
        # enclosed flow: the pressure is defined up to the pinned constant, which both sides fix at the same dof
        assert rel(up[tab.Nv :, b], gold["up_final"][i][tab.Nv :]) < 1e-7
    assert not ens.diverged.any()
    ens.close()
