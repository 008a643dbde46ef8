"""Parity of the CUDA path (through the C ABI) against the CPU oracle — runs on the B200 box.

Tolerances are BASELINE.json's: per-step velocity/pressure relative L2 error <= 1e-9 after 100
steps, sensor time series within 1e-6 relative; integer maps bit-exact (test_host_setup.py).
Nothing here reads /root/reference.
"""
import numpy as np
import pytest

from oracle import cases
from oracle.flow_oracle import FlowOracle, ZOHController

pytestmark = pytest.mark.gpu

FIELD_TOL = 1e-9
SERIES_TOL = 1e-6


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


@pytest.fixture(scope="module")
def cyl(root, built_lib):
    """Cylinder Re=100 FlowSolver facade with the cached base flow + a live oracle on the same inputs."""
    import tempfile
    from pathlib import Path

    from flowcontrol_b200.examples.cylinder import CylinderFlowSolver
    from flowcontrol_b200.flowfield import Field
    from flowcontrol_b200.problem import FlowProblem

    UP0 = np.load(root / "tests/golden/cylinder_baseflow.npz")["UP0"]
    fs = CylinderFlowSolver.make_default(path_out=Path(tempfile.mkdtemp()))
    tab = fs.tables
    fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
    prob = FlowProblem(tab, fs.blocks, 100.0, 0.005, fs.bc.bcu, fs.params_control.actuator_list,
                       fs.params_control.sensor_list, UP0)
    case = cases.cylinder(100.0)
    xy, tri = cases.load_mesh(case.mesh_file)
    orc = FlowOracle(case, xy, tri)
    orc.set_base_flow(UP0)
    orc.prepare()
    return fs, prob, orc, UP0


def test_cylinder_closed_loop_golden_trajectory(root, cyl):
    """The reference's regression scenario (test_cylinder.py:78-126): 20 closed-loop steps."""
    from flowcontrol_b200.controller import Controller
    from flowcontrol_b200.ensemble import Ensemble

    fs, prob, orc, UP0 = cyl
    tab = prob.tab
    gold = np.load(root / "tests/golden/cylinder_traj.npz")
    ic = fs._default_initial_perturbation()
    B = 32
    ens = Ensemble(prob, B)
    y0 = ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    assert np.allclose(y0[:, 0], gold["y_meas"][0], rtol=1e-12)
    assert np.isclose(ens.dE[0], gold["dE"][0], rtol=1e-12)
    K = Controller.from_file(root / "tests/golden/Kopt_reduced13.npz")
    for i in range(20):
        u = K.step(-ens.y_meas[0, 0], prob.dt)
        ens.step(np.full((2, B), u[0]))
        assert np.allclose(ens.y_meas[:, 0], gold["y_meas"][i + 1], rtol=SERIES_TOL, atol=0)
        assert np.isclose(ens.dE[0], gold["dE"][i + 1], rtol=SERIES_TOL)
    # reference goldens themselves (test_cylinder.py:71-74)
    assert np.allclose(ens.y_meas[:, 0], [0.011615482723602308, 0.003860524805395703, 0.0038461597025207803], rtol=1e-9)
    assert np.isclose(ens.dE[0], 0.09462807324653322, rtol=1e-9)
    up = ens.fields(0)
    assert rel(up[: tab.Nv, 0], gold["up_final"][: tab.Nv]) < FIELD_TOL
    assert rel(up[tab.Nv :, 0], gold["up_final"][tab.Nv :]) < FIELD_TOL
    assert np.abs(up - up[:, :1]).max() == 0.0  # identical inputs -> bit-identical trajectories
    assert not ens.diverged.any()
    ens.close()


def test_cylinder_100_steps_actuated_vs_oracle(cyl):
    """100 open-loop steps with time-varying, trajectory-dependent slot actuation (BC lifting path)."""
    from flowcontrol_b200.ensemble import Ensemble

    fs, prob, orc, UP0 = cyl
    tab = prob.tab
    orc.case.ic = (2.0, 0.0, 0.5, 1.0)  # ParamIC of run_cylinder_example.py:55
    orc.init_time_stepping()
    B = 32
    amp = np.linspace(-1.0, 1.0, B)
    track = 7  # the oracle follows this trajectory
    ens = Ensemble(prob, B)
    ens.set_state(orc.ic[: tab.Nv], None, orc.ic[tab.Nv :], order=1)
    worst_y = 0.0
    for k in range(100):
        t = (k + 1) * prob.dt
        base = np.array([np.sin(40 * t), 0.5 * np.cos(25 * t)])
        ens.step(base[:, None] * amp[None, :])
        orc.step(base * amp[track])
        worst_y = max(worst_y, np.abs(ens.y_meas[:, track] - orc.y_meas).max() / np.abs(orc.y_meas).max())
    up = ens.fields(0)[:, track]
    assert rel(up[: tab.Nv], orc.up[: tab.Nv]) < FIELD_TOL
    assert rel(up[tab.Nv :], orc.up[tab.Nv :]) < FIELD_TOL
    assert worst_y < SERIES_TOL
    assert abs(ens.dE[track] - orc.dE) / orc.dE < SERIES_TOL
    ens.close()


def test_device_closed_loop_matches_host_loop_and_oracle(root, cyl):
    """fcb_run_closed_loop (controller on the GPU, CUDA-graph replay) vs stepping from the host,
    for a gain-swept family of controllers (config 2 of BASELINE.json)."""
    from flowcontrol_b200.controller import Controller, ControllerBank
    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.sharding import controller_gain_sweep

    fs, prob, orc, UP0 = cyl
    tab = prob.tab
    ic = fs._default_initial_perturbation()
    B, nsteps = 64, 24
    k = np.load(root / "tests/golden/Kopt_reduced13.npz")
    gains = controller_gain_sweep(B)
    ctrls = [Controller(k["A"], g * k["B"], k["C"], g * k["D"]) for g in gains]
    Ky, Fu = np.array([[-1.0, 0.0, 0.0]]), np.array([[1.0], [1.0]])
    ens = Ensemble(prob, B)
    ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    ens.set_controllers(ControllerBank(ctrls, prob.dt, Ky, Fu))
    series = ens.run_closed_loop(nsteps)  # [nsteps, 1+na+ns, B]
    assert series.shape == (nsteps, 6, B)
    # (a) same thing driven from the host through fcb_step: bitwise identical
    ens2 = Ensemble(prob, B)
    ens2.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    hostK = [Controller(c.A, c.B, c.C, c.D) for c in ctrls]
    for s in range(nsteps):
        u = np.array([hostK[b].step(-ens2.y_meas[0, b], prob.dt)[0] for b in range(B)])
        ens2.step(np.stack([u, u]))
        assert np.allclose(series[s, 1, :], u, rtol=1e-12, atol=1e-300)
        assert np.allclose(series[s, 3:, :], ens2.y_meas, rtol=1e-10, atol=0)
        assert np.allclose(series[s, 0, :], ens2.dE, rtol=1e-12)
    # (b) oracle on three of the trajectories
    for b in (0, 37, 63):
        orc.case.ic = (0.0, 0.0, 1.0, 1.0)
        orc.init_time_stepping()
        K = ZOHController(k["A"], gains[b] * k["B"], k["C"], gains[b] * k["D"])
        for s in range(nsteps):
            u = K.step(-orc.y_meas[0], prob.dt)
            orc.step([u[0], u[0]])
            assert np.allclose(series[s, 3:, b], orc.y_meas, rtol=SERIES_TOL, atol=0)
            assert np.isclose(series[s, 0, b], orc.dE, rtol=SERIES_TOL)
    x = ens.controller_state()
    assert np.allclose(x[:, 5], hostK[5].x, rtol=1e-9, atol=1e-14)
    ens.close()
    ens2.close()


def test_restart_with_order2_continues_identically(cyl):
    """set_state(order=2) from (u_n, u_nn) continues a run exactly (flowsolver.py:599-663)."""
    from flowcontrol_b200.ensemble import Ensemble

    fs, prob, orc, UP0 = cyl
    tab = prob.tab
    ic = fs._default_initial_perturbation()
    B = 32
    uc = np.full((2, B), 0.05)
    a = Ensemble(prob, B)
    a.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    for _ in range(3):
        a.step(uc)
    u_n, u_nn = a.fields(0), a.fields(1)
    b = Ensemble(prob, B)
    y = b.set_state(u_n[: tab.Nv], u_nn, u_n[tab.Nv :], order=2)
    assert np.array_equal(y, a.y_meas)
    for _ in range(3):
        a.step(uc)
        b.step(uc)
    assert np.array_equal(a.fields(0), b.fields(0))
    assert np.array_equal(a.y_meas, b.y_meas) and np.array_equal(a.dE, b.dE)
    a.close()
    b.close()


def test_divergence_is_per_trajectory(cyl):
    from flowcontrol_b200.ensemble import Ensemble

    fs, prob, orc, UP0 = cyl
    tab = prob.tab
    ic = fs._default_initial_perturbation()
    B = 32
    u0 = np.repeat(ic[: tab.Nv, None], B, axis=1)
    u0[100, 3] = np.inf
    ens = Ensemble(prob, B)
    ens.set_state(u0, None, None, order=1)
    ens.step(np.zeros((2, B)))
    assert ens.diverged[3] == 1 and ens.diverged.sum() == 1
    assert np.all(np.isfinite(ens.y_meas[:, np.arange(B) != 3]))
    ens.close()


def test_facade_regression_with_restart(root, tmp_path, cyl):
    """Port of the reference's test_cylinder_regression: FlowSolver API, closed loop, JSON-sidecar restart."""
    from flowcontrol_b200.controller import Controller
    from flowcontrol_b200.examples.cylinder import CylinderFlowSolver
    from flowcontrol_b200.exporter import write_checkpoint
    from flowcontrol_b200.flowfield import Field

    _, _, _, UP0 = cyl
    fs = CylinderFlowSolver.make_default(Re=100, path_out=tmp_path, num_steps=10, save_every=5)
    tab = fs.tables
    # base flow from the fixture (computing it takes ~40 s of SuperLU; covered by the slow CPU test)
    write_checkpoint(fs.paths.U0, "U0", UP0[: tab.Nv], 0.0, append=False, tab=tab)
    write_checkpoint(fs.paths.P0, "P0", UP0[tab.Nv :], 0.0, append=False, tab=tab)
    fs.load_steady_state()
    assert np.isclose(fs.fields.U0.vector().get_local().max(), 1.1921615450014942, rtol=1e-9)
    fs.initialize_time_stepping(ic=None)
    Kss = Controller.from_file(root / "tests/golden/Kopt_reduced13.npz")
    for _ in range(fs.params_time.num_steps):
        u_ctrl = Kss.step(y=-fs.y_meas[0], dt=fs.params_time.dt)
        fs.step(u_ctrl=[u_ctrl[0], u_ctrl[0]])
    fs.write_timeseries()
    fs2 = CylinderFlowSolver.make_default(Re=100, path_out=tmp_path, num_steps=10, save_every=5, Tstart=0.05)
    fs2.load_steady_state()
    fs2.initialize_time_stepping(Tstart=fs2.params_time.Tstart)
    for _ in range(fs2.params_time.num_steps):
        u_ctrl = Kss.step(y=-fs2.y_meas[0], dt=fs2.params_time.dt)
        fs2.step(u_ctrl=np.repeat(u_ctrl, repeats=2, axis=0))
    fs2.write_timeseries()
    last = fs2.timeseries.iloc[-1]
    Usave = fs2.fields.Usave.vector().get_local()
    assert np.isclose(Usave.max(), 1.325070045534714, rtol=1e-9)
    assert np.isclose(Usave.mean(), 0.3376859329866094, rtol=1e-9)
    assert np.isclose(last["time"], 0.1, rtol=1e-12)
    assert np.isclose(last["y_meas_1"], 0.011615482723602308, rtol=1e-9)
    assert np.isclose(last["y_meas_2"], 0.003860524805395703, rtol=1e-9)
    assert np.isclose(last["y_meas_3"], 0.0038461597025207803, rtol=1e-9)
    assert np.isclose(last["dE"], 0.09462807324653322, rtol=1e-9)
    assert np.all(np.isfinite(fs2.fields.u_.vector().get_local()))
    with pytest.raises(ValueError):
        fs2.step(u_ctrl=[0.0])


def test_lidcavity_actuated_and_linear_superposition(root, built_lib):
    """Lid cavity (pinned pressure): (a) non-zero lid speed vs the oracle; (b) with
    is_eq_nonlinear=False the step is linear in (state, u_ctrl): superposition holds to round-off
    at the full ensemble width."""
    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.examples import lidcavity as ex
    from flowcontrol_b200.problem import FlowProblem

    UP0 = np.load(root / "tests/golden/lidcavity_baseflow.npz")["UP0"]
    prob = ex.make_problem(Re=1000.0, UP0=UP0)
    tab = prob.tab
    case = cases.lidcavity(1000.0)
    xy, tri = cases.load_mesh(case.mesh_file)
    orc = FlowOracle(case, xy, tri)
    orc.set_base_flow(UP0)
    orc.init_time_stepping()
    B = 32
    ens = Ensemble(prob, B)
    ens.set_state(orc.ic[: tab.Nv], None, orc.ic[tab.Nv :], order=1)
    for k in range(10):
        uc = 0.1 * np.sin(0.7 * k)
        orc.step([uc])
        ens.step(np.full((1, B), uc))
        assert np.allclose(ens.y_meas[:, 0], orc.y_meas, rtol=SERIES_TOL, atol=0)
    up = ens.fields(0)[:, 0]
    assert rel(up[: tab.Nv], orc.up[: tab.Nv]) < FIELD_TOL
    assert rel(up[tab.Nv :], orc.up[tab.Nv :]) < 1e-8  # pressure level is pinned at dof 0 in both
    ens.close()
    # (b) linear mode
    import tempfile
    from pathlib import Path

    fs = ex.LidCavityFlowSolver.make_default(Re=1000.0, path_out=Path(tempfile.mkdtemp()))
    lin = FlowProblem(tab, fs.blocks, 1000.0, fs.params_time.dt, fs.bc.bcu, fs.params_control.actuator_list,
                      fs.params_control.sensor_list, UP0, nonlinear=False, pin_pressure=True, symbolic=prob.sym)
    B = 256
    rng = np.random.default_rng(0)
    u0 = np.zeros((tab.Nv, B))
    u0[:, 0] = rng.standard_normal(tab.Nv)
    u0[:, 1] = rng.standard_normal(tab.Nv)
    u0[:, 2] = u0[:, 0] + u0[:, 1]
    e = Ensemble(lin, B)
    e.set_state(u0, None, None, order=1)
    for k in range(5):
        uc = np.zeros((1, B))
        uc[0, 0], uc[0, 1] = 0.3 * (k + 1), -0.2
        uc[0, 2] = uc[0, 0] + uc[0, 1]
        e.step(uc)
    f = e.fields(0)
    assert rel(f[:, 2], f[:, 0] + f[:, 1]) < 1e-12
    assert np.abs(f[:, 3:]).max() == 0.0  # untouched trajectories stay exactly zero
    e.close()


@pytest.mark.parametrize("B", [1, 40, 100, 300, 512])
def test_ensemble_widths_match_oracle(root, built_lib, B):
    """Every ensemble width (padding to 32/64/128-multiples, 1/2/4-warp and k-split sweep CTAs, several
    128-trajectory slabs) gives the oracle's trajectory, for trajectory-dependent lid actuation."""
    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.examples import lidcavity as ex

    UP0 = np.load(root / "tests/golden/lidcavity_baseflow.npz")["UP0"]
    prob = ex.make_problem(Re=1000.0, UP0=UP0)
    tab = prob.tab
    case = cases.lidcavity(1000.0)
    xy, tri = cases.load_mesh(case.mesh_file)
    probe = sorted({0, B // 2, B - 1})
    amp = 0.02 + 0.1 * np.arange(B) / max(B - 1, 1)  # lid speed of trajectory b
    oracles = []
    for b in probe:
        orc = FlowOracle(case, xy, tri)
        orc.set_base_flow(UP0)
        orc.init_time_stepping()
        oracles.append(orc)
    ens = Ensemble(prob, B)
    ens.set_state(oracles[0].ic[: tab.Nv], None, oracles[0].ic[tab.Nv :], order=1)
    for k in range(4):
        uc = (amp * np.cos(0.9 * k))[None, :]
        ens.step(uc)
        for orc, b in zip(oracles, probe):
            orc.step([uc[0, b]])
            assert np.allclose(ens.y_meas[:, b], orc.y_meas, rtol=SERIES_TOL, atol=0)
            assert np.isclose(ens.dE[b], orc.dE, rtol=SERIES_TOL)
    up = ens.fields(0)
    for orc, b in zip(oracles, probe):
        assert rel(up[: tab.Nv, b], orc.up[: tab.Nv]) < FIELD_TOL
    assert not ens.diverged.any()
    ens.close()


def test_cavity_force_actuator_golden_trajectory(root, built_lib):
    """Open cavity Re=7500 (235 k dofs, body-force actuator, wall-shear sensor): the reference's
    10-step regression scenario (test_cavity.py:58-90) through the CUDA path."""
    import tempfile
    from pathlib import Path

    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.examples.cavity import CavityFlowSolver
    from flowcontrol_b200.flowfield import Field
    from flowcontrol_b200.problem import FlowProblem

    UP0 = np.load(root / "tests/golden/cavity_baseflow.npz")["UP0"]
    gold = np.load(root / "tests/golden/cavity_traj.npz")
    fs = CavityFlowSolver.make_default(path_out=Path(tempfile.mkdtemp()))
    tab = fs.tables
    fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
    prob = FlowProblem(tab, fs.blocks, 7500.0, 0.0004, fs.bc.bcu, fs.params_control.actuator_list,
                       fs.params_control.sensor_list, UP0)
    ic = fs._default_initial_perturbation()
    B = 32
    ens = Ensemble(prob, B)
    y0 = ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    assert np.allclose(y0[:, 0], gold["y_meas"][0], rtol=1e-9)
    for i in range(10):
        ens.step(np.zeros((1, B)))
        assert np.allclose(ens.y_meas[:, 0], gold["y_meas"][i + 1], rtol=SERIES_TOL, atol=0)
        assert np.isclose(ens.dE[0], gold["dE"][i + 1], rtol=SERIES_TOL)
    assert np.allclose(ens.y_meas[:, 0], [6.0488687475121505, 0.024799707355708498], rtol=1e-9)  # test_cavity.py:52-53
    assert np.isclose(ens.dE[0], 0.005000924582291293, rtol=1e-9)
    up = ens.fields(0)[:, 0]
    assert np.isclose(np.linalg.norm(up), float(gold["up_norm"]), rtol=1e-9)
    # the force actuator with u_ctrl != 0 is compared with the oracle in
    # tests/test_gpu_configs.py::test_cavity_bdf_force_actuator_nonzero_control (open loop) and
    # ::test_config3_cavity_static_gains_256 (closed loop, B = 256)
    assert not ens.diverged.any()
    ens.close()


def test_pinball_golden_trajectory_and_rotation_actuators(root, built_lib):
    """Fluidic pinball (302 k dofs): (a) the reference's regression scenario (test_pinball.py:69-110: Re=30, suction
    mode, 10 steps) through the CUDA path against the oracle trajectory; (b) rotation mode (BASELINE configs[2]):
    the three cylinder surfaces carry exactly the rotation profile x u_ctrl of each trajectory (actuator.py:241-251)."""
    import tempfile
    from pathlib import Path

    from flowcontrol_b200.actuator import CYLINDER_ACTUATION_MODE
    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.examples.pinball import PinballFlowSolver
    from flowcontrol_b200.flowfield import Field
    from flowcontrol_b200.problem import FlowProblem

    UP0 = np.load(root / "tests/golden/pinball_baseflow.npz")["UP0"]
    gold = np.load(root / "tests/golden/pinball_traj.npz")
    fs = PinballFlowSolver.make_default(Re=30.0, mode_actuation=CYLINDER_ACTUATION_MODE.SUCTION, path_out=Path(tempfile.mkdtemp()))
    tab = fs.tables
    fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
    prob = FlowProblem(tab, fs.blocks, 30.0, 0.005, fs.bc.bcu, fs.params_control.actuator_list,
                       fs.params_control.sensor_list, UP0)
    ic = fs._default_initial_perturbation()
    B = 32
    ens = Ensemble(prob, B)
    y0 = ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    assert np.allclose(y0[:, 0], gold["y_meas"][0], rtol=1e-9)
    for i in range(10):
        ens.step(np.repeat(gold["u_ctrl"][i][:, None], B, axis=1))
        assert np.allclose(ens.y_meas[:, 0], gold["y_meas"][i + 1], rtol=SERIES_TOL, atol=0)
        assert np.isclose(ens.dE[0], gold["dE"][i + 1], rtol=SERIES_TOL)
    assert np.isclose(np.linalg.norm(ens.fields(0)[:, 0]), float(gold["up_norm"]), rtol=1e-9)
    ens.close()
    # (b) rotation mode on the same mesh (the suction base flow is only a linearisation point here)
    fr = PinballFlowSolver.make_default(Re=30.0, mode_actuation=CYLINDER_ACTUATION_MODE.ROTATION, path_out=Path(tempfile.mkdtemp()))
    fr._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
    acts = fr.params_control.actuator_list
    pr = FlowProblem(fr.tables, fr.blocks, 30.0, 0.005, fr.bc.bcu, acts, fr.params_control.sensor_list, UP0)
    er = Ensemble(pr, B)
    er.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    rng = np.random.default_rng(3)
    uc = rng.uniform(-2, 2, size=(3, B))
    for _ in range(2):
        er.step(uc)
    up = er.fields(0)
    d = pr.dirichlet
    assert np.abs(up[d.dofs] - (d.const[:, None] + d.shape.T @ uc)).max() == 0.0  # Dirichlet rows are imposed exactly
    xy = fr.tables.node_xy
    for k, a in enumerate(acts):  # the surface of cylinder k rotates with u_ctrl[k] * d/2
        sx, sy = a.shape(xy[:, 0], xy[:, 1])
        on = np.flatnonzero(np.abs(np.hypot(xy[:, 0] - a.position_x, xy[:, 1] - a.position_y) - 0.5) < 1e-6)
        on = on[np.isin(on, d.dofs)]
        assert len(on) > 50
        assert np.allclose(up[on], np.outer(sx[on], uc[k]), atol=1e-14)
        assert np.allclose(np.hypot(sx[on], sy[on]), 0.5, atol=1e-12)
    assert np.isfinite(up).all() and not er.diverged.any()
    er.close()


def test_lift_drag_time_series_200_steps(root, cyl):
    """Lift/drag coefficient rows (examples/cylinder/cylinderflowsolver.py:115-126 made a per-step measurement)
    logged on the device over 200 closed-loop steps vs the oracle: sensor and lift series within 1e-6
    (BASELINE.json north_star tolerance for the time series)."""
    from flowcontrol_b200.controller import Controller, ControllerBank
    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.problem import FlowProblem
    from oracle.flow_oracle import force_coefficients

    fs, prob0, orc, UP0 = cyl
    tab = prob0.tab
    sensors = list(fs.params_control.sensor_list) + fs.force_sensors()
    prob = FlowProblem(tab, fs.blocks, 100.0, 0.005, fs.bc.bcu, fs.params_control.actuator_list, sensors, UP0,
                       symbolic=prob0.sym)
    ic = fs._default_initial_perturbation()
    B, nsteps = 32, 200
    k = np.load(root / "tests/golden/Kopt_reduced13.npz")
    ctrls = [Controller(k["A"], k["B"], k["C"], k["D"]) for _ in range(B)]
    Ky, Fu = np.array([[-1.0, 0.0, 0.0, 0.0, 0.0]]), np.array([[1.0], [1.0]])
    ens = Ensemble(prob, B)
    ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    ens.set_controllers(ControllerBank(ctrls, prob.dt, Ky, Fu))
    series = ens.run_closed_loop(nsteps)  # columns: dE, u1, u2, y1..y3, cl, cd
    assert series.shape == (nsteps, 8, B)
    orc.case.ic = (0.0, 0.0, 1.0, 1.0)
    orc.init_time_stepping()
    K = ZOHController(k["A"], k["B"], k["C"], k["D"])
    body = lambda x, y: np.hypot(x, y) < 0.6  # noqa: E731  every boundary facet of the cylinder
    ref = np.zeros((nsteps, 5))
    for s_ in range(nsteps):
        u = K.step(-orc.y_meas[0], prob.dt)
        orc.step([u[0], u[0]])
        ref[s_, :3] = orc.y_meas
        ref[s_, 3:] = force_coefficients(orc.mesh, body, orc.up, 0.01, 1.0, 1.0)
    got = series[:, 3:, 0]
    scale = np.abs(ref).max(axis=0)
    assert (np.abs(got - ref).max(axis=0) / scale).max() < SERIES_TOL
    assert np.abs(series - series[:, :, :1]).max() == 0.0  # identical controllers -> identical trajectories
    # host-side evaluation on the full field agrees with base-flow value + logged perturbation part
    cl0, cd0 = fs.compute_force_coefficients(UP0[: tab.Nv], UP0[tab.Nv :])
    up = ens.fields(0)[:, 0]
    cl, cd = fs.compute_force_coefficients(UP0[: tab.Nv] + up[: tab.Nv], UP0[tab.Nv :] + up[tab.Nv :])
    assert np.isclose(cl, cl0 + series[-1, 6, 0], rtol=1e-9, atol=1e-12) and np.isclose(cd, cd0 + series[-1, 7, 0], rtol=1e-9)
    ens.close()


def test_crank_nicolson_cylinder_vs_oracle(root, cyl):
    """time_scheme='cn' (nsforms.py:191-236) on the device: k_spmm (E u_n) + element pass (-N(u_n)) + current and
    previous-step control terms, for slot (BC lifting) and body-force actuation (force averaged over the step), host-driven
    and closed loop on the device, against the oracle's Crank-Nicolson step."""
    from flowcontrol_b200.actuator import ActuatorForceGaussianV
    from flowcontrol_b200.controller import Controller, ControllerBank
    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.problem import FlowProblem
    from oracle import flow_oracle as fo

    fs, prob_bdf, _, UP0 = cyl
    tab = prob_bdf.tab
    acts = list(fs.params_control.actuator_list) + [ActuatorForceGaussianV(sigma=0.3, position=np.array([1.5, 0.2]))]
    prob = FlowProblem(tab, fs.blocks, 100.0, 0.005, fs.bc.bcu, acts, fs.params_control.sensor_list, UP0,
                       time_scheme="cn", symbolic=prob_bdf.sym)
    case = cases.cylinder(100.0)
    case.actuators = list(case.actuators) + [fo.ActuatorSpec("force", fo.gaussian_v(0.3, (1.5, 0.2)))]
    case.ic = (2.0, 0.0, 0.5, 1.0)
    xy, tri = cases.load_mesh(case.mesh_file)
    orc = FlowOracle(case, xy, tri, time_scheme="cn")
    orc.set_base_flow(UP0)
    orc.init_time_stepping()
    B, track, nsteps = 40, 11, 60
    amp = np.linspace(-1.0, 1.0, B)
    ens = Ensemble(prob, B)
    y0 = ens.set_state(orc.ic[: tab.Nv], None, orc.ic[tab.Nv :], order="cn")
    assert np.allclose(y0[:, track], orc.y_meas, rtol=1e-12)
    worst_y = 0.0
    for k in range(nsteps):
        t = (k + 1) * prob.dt
        base = np.array([np.sin(40 * t), 0.5 * np.cos(25 * t), 2.0 * np.cos(30 * t)])
        ens.step(base[:, None] * amp[None, :])
        orc.step(base * amp[track])
        worst_y = max(worst_y, np.abs(ens.y_meas[:, track] - orc.y_meas).max() / np.abs(orc.y_meas).max())
    up = ens.fields(0)[:, track]
    assert rel(up[: tab.Nv], orc.up[: tab.Nv]) < FIELD_TOL
    assert rel(up[tab.Nv :], orc.up[tab.Nv :]) < FIELD_TOL
    assert worst_y < SERIES_TOL
    assert abs(ens.dE[track] - orc.dE) / orc.dE < SERIES_TOL
    assert not ens.diverged.any()
    # restart from the state reached (u_ctrl^n = 0 again, like a reference restart): continue on the device in closed
    # loop (CUDA-graph replay) and compare with the oracle driven by the same controller
    u_n = ens.fields(0)[:, track].copy()
    kk = np.load(root / "tests/golden/Kopt_reduced13.npz")
    Ky, Fu = np.array([[-1.0, 0.0, 0.0]]), np.array([[1.0], [1.0], [0.5]])
    ens.set_state(u_n[: tab.Nv], None, u_n[tab.Nv :], order="cn")
    ens.set_controllers(ControllerBank([Controller(kk["A"], kk["B"], kk["C"], kk["D"]) for _ in range(B)], prob.dt, Ky, Fu))
    series = ens.run_closed_loop(16)
    orc.init_time_stepping(ic=u_n)
    K = ZOHController(kk["A"], kk["B"], kk["C"], kk["D"])
    for s in range(16):
        u = K.step(-orc.y_meas[0], prob.dt)
        orc.step([u[0], u[0], 0.5 * u[0]])
        assert np.allclose(series[s, 4:, 0], orc.y_meas, rtol=SERIES_TOL, atol=0)
        assert np.isclose(series[s, 0, 0], orc.dE, rtol=SERIES_TOL)
    assert np.abs(series - series[:, :, :1]).max() == 0.0
    ens.close()


def test_crank_nicolson_through_the_facade(root, cyl, tmp_path):
    """ParamSolver(time_scheme='cn') through FlowSolver.step: self-starting, order stays 'cn' (flowsolver.py:513, 742)."""
    from flowcontrol_b200.examples.cylinder import CylinderFlowSolver
    from flowcontrol_b200.flowfield import Field

    _, _, _, UP0 = cyl
    fs = CylinderFlowSolver.make_default(path_out=tmp_path, num_steps=5)
    fs.params_solver.time_scheme = "cn"
    tab = fs.tables
    fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
    fs.params_ic.xloc, fs.params_ic.yloc, fs.params_ic.radius, fs.params_ic.amplitude = 2.0, 0.0, 0.5, 1.0
    fs.initialize_time_stepping(Tstart=0.0)
    case = cases.cylinder(100.0)
    case.ic = (2.0, 0.0, 0.5, 1.0)
    xy, tri = cases.load_mesh(case.mesh_file)
    orc = FlowOracle(case, xy, tri, time_scheme="cn")
    orc.set_base_flow(UP0)
    orc.init_time_stepping()
    for k in range(5):
        y = fs.step(u_ctrl=[0.1 * k, -0.05 * k])
        orc.step([0.1 * k, -0.05 * k])
        assert fs.order == "cn"
        assert np.allclose(y, orc.y_meas, rtol=SERIES_TOL, atol=0)


def test_device_cost_sums_match_reference_cost_functions(root, cyl):
    """fcb_get_costs (running sums in k_log) vs the reference's cost functions applied to the logged series
    (utils/optim.py:231-288), for a gain-swept controller family."""
    from flowcontrol_b200.controller import Controller, ControllerBank
    from flowcontrol_b200.costs import compute_control_cost, compute_signal_cost
    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.sharding import controller_gain_sweep

    fs, prob, _, _ = cyl
    tab = prob.tab
    ic = fs._default_initial_perturbation()
    B, nsteps = 40, 30
    k = np.load(root / "tests/golden/Kopt_reduced13.npz")
    ctrls = [Controller(k["A"], g * k["B"], k["C"], g * k["D"]) for g in controller_gain_sweep(B)]
    ens = Ensemble(prob, B)
    ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    ens.set_controllers(ControllerBank(ctrls, prob.dt, np.array([[-1.0, 0.0, 0.0]]), np.array([[1.0], [1.0]])))
    series = ens.run_closed_loop(nsteps)
    Tnorm = prob.dt / (nsteps * prob.dt)
    c = ens.costs(Tnorm)
    assert np.allclose(c["energy_integral"], compute_signal_cost(series[:, 0, :], Tnorm, "integral"), rtol=1e-13)
    assert np.array_equal(c["energy_terminal"], compute_signal_cost(series[:, 0, :], Tnorm, "terminal"))
    assert np.allclose(c["control"], compute_control_cost(series[:, 1:3, :], Tnorm), rtol=1e-13)
    b = 7  # one trajectory through the reference's scalar signature
    assert np.isclose(c["control"][b], compute_control_cost(series[:, 1:3, b], Tnorm), rtol=1e-13)
    # a run without logging accumulates the same sums
    ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    ens.set_controllers(ControllerBank(ctrls, prob.dt, np.array([[-1.0, 0.0, 0.0]]), np.array([[1.0], [1.0]])))
    ens.run_closed_loop(nsteps, log=False)
    c2 = ens.costs(Tnorm)
    assert all(np.array_equal(c[k_], c2[k_]) for k_ in c)
    with pytest.raises(ValueError):
        compute_signal_cost(series[:, 0, 0], Tnorm, "mean")
    ens.close()


def test_create_rejects_inconsistent_crank_nicolson_input(cyl):
    """fcb_create fails loudly (negative status + message) when scheme = 1 comes without its operator, and the scalar CSR
    SpMM kept for A/B runs (FCB_SPMM_CSR=1) gives the same step as the tensor-core panels to round-off."""
    import ctypes as C
    import os

    from flowcontrol_b200 import libfcb
    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.problem import FlowProblem

    fs, prob_bdf, _, UP0 = cyl
    tab = prob_bdf.tab
    lib = libfcb.load()
    pack = libfcb.ProblemPack(prob_bdf)
    pack.struct.scheme = 1  # claims Crank-Nicolson but carries no CSR operator
    h = C.c_void_p()
    rc = lib.fcb_create(C.byref(pack.struct), 32, 0, C.byref(h))
    assert rc < 0 and not h.value
    assert b"cn_ptr" in lib.fcb_last_error(None)
    pack.struct.scheme = 7
    assert lib.fcb_create(C.byref(pack.struct), 32, 0, C.byref(h)) < 0
    prob = FlowProblem(tab, fs.blocks, 100.0, 0.005, fs.bc.bcu, fs.params_control.actuator_list,
                       fs.params_control.sensor_list, UP0, time_scheme="cn", symbolic=prob_bdf.sym)
    ic = fs._default_initial_perturbation(2.0, 0.0, 0.5)
    ups = []
    for csr in ("0", "1"):
        os.environ["FCB_SPMM_CSR"] = csr
        try:
            ens = Ensemble(prob, 40)
        finally:
            os.environ.pop("FCB_SPMM_CSR")
        ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order="cn")
        for k in range(5):
            ens.step(np.full((2, 40), 0.1 * k))
        ups.append(ens.fields(0)[:, 17].copy())
        ens.close()
    assert rel(ups[0], ups[1]) < 1e-12  # two summation orders of E u_n, five solves apart (field tolerance of the parity tests: 1e-9)


def test_crank_nicolson_any_ensemble_width(cyl):
    """The Crank-Nicolson path (SpMM slices, seeded element pass) at ragged widths: B = 1, 33 and 300 give the same
    trajectory to round-off, and identical inputs stay bit-identical inside an ensemble."""
    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.problem import FlowProblem

    fs, prob_bdf, _, UP0 = cyl
    tab = prob_bdf.tab
    prob = FlowProblem(tab, fs.blocks, 100.0, 0.005, fs.bc.bcu, fs.params_control.actuator_list,
                       fs.params_control.sensor_list, UP0, time_scheme="cn", symbolic=prob_bdf.sym)
    ic = fs._default_initial_perturbation(2.0, 0.0, 0.5)
    ref = None
    for B in (1, 33, 300):
        ens = Ensemble(prob, B)
        ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order="cn")
        for k in range(6):
            ens.step(np.full((2, B), 0.05 * k))
        up = ens.fields(0)
        assert np.isfinite(up).all() and not ens.diverged.any()
        assert np.abs(up - up[:, :1]).max() == 0.0
        if ref is None:
            ref = up[:, 0].copy()
        else:
            assert rel(up[:, -1], ref) < 1e-12
        ens.close()


@pytest.mark.parametrize("rows,height", [(512, 6), (300, 2)])
def test_subtree_cluster_sweeps_match_oracle(root, built_lib, rows, height):
    """The shared-memory subtree clusters (k_cluster_sweep; optional path of the constant-LHS solve, off by default): same
    trajectory as the oracle for two cuts of the elimination tree (one and several tiers, k-split operations, imported
    update vectors), ragged ensemble width, trajectory-dependent lid actuation."""
    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.examples import lidcavity as ex
    from flowcontrol_b200.flowfield import Field
    from flowcontrol_b200.problem import FlowProblem

    UP0 = np.load(root / "tests/golden/lidcavity_baseflow.npz")["UP0"]
    import tempfile
    from pathlib import Path

    fs = ex.LidCavityFlowSolver.make_default(Re=1000.0, path_out=Path(tempfile.mkdtemp()))
    tab = fs.tables
    fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
    prob = FlowProblem(tab, fs.blocks, 1000.0, 0.005, fs.bc.bcu, fs.params_control.actuator_list, fs.params_control.sensor_list,
                       UP0, pin_pressure=True, cluster_rows=rows, cluster_height=height)
    assert len(prob.plans[2].cl_fptr) - 1 > 0  # the plan really has clusters
    case = cases.lidcavity(1000.0)
    xy, tri = cases.load_mesh(case.mesh_file)
    B = 72
    probe = [0, 35, 71]
    amp = 0.02 + 0.1 * np.arange(B) / (B - 1)
    oracles = []
    for b in probe:
        orc = FlowOracle(case, xy, tri)
        orc.set_base_flow(UP0)
        orc.init_time_stepping()
        oracles.append(orc)
    ens = Ensemble(prob, B)
    ens.set_state(oracles[0].ic[: tab.Nv], None, oracles[0].ic[tab.Nv :], order=1)
    for k in range(6):
        uc = (amp * np.cos(0.9 * k))[None, :]
        ens.step(uc)
        for orc, b in zip(oracles, probe):
            orc.step([uc[0, b]])
            assert np.allclose(ens.y_meas[:, b], orc.y_meas, rtol=SERIES_TOL, atol=0)
            assert np.isclose(ens.dE[b], orc.dE, rtol=SERIES_TOL)
    up = ens.fields(0)
    for orc, b in zip(oracles, probe):
        assert rel(up[: tab.Nv, b], orc.up[: tab.Nv]) < FIELD_TOL
    assert not ens.diverged.any()
    ens.close()


@pytest.mark.parametrize("variant", ["free_dissection", "presummed_gathers"])
def test_optional_solve_plans_match_oracle(root, built_lib, variant):
    """Solve plans that are supported but not the default: the free (not depth-bounded) dissection the round started with,
    and forward blocks that gather one pre-summed plane (gather-sums as programmatic dependent launches between the sweep
    launches).  Same oracle comparison as the cluster test: ragged width, trajectory-dependent lid actuation."""
    import tempfile
    from pathlib import Path

    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.examples import lidcavity as ex
    from flowcontrol_b200.flowfield import Field
    from flowcontrol_b200.problem import FlowProblem

    UP0 = np.load(root / "tests/golden/lidcavity_baseflow.npz")["UP0"]
    fs = ex.LidCavityFlowSolver.make_default(Re=1000.0, path_out=Path(tempfile.mkdtemp()))
    tab = fs.tables
    fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
    kw = dict(balanced=False) if variant == "free_dissection" else dict(presum_height=1)
    prob = FlowProblem(tab, fs.blocks, 1000.0, 0.005, fs.bc.bcu, fs.params_control.actuator_list, fs.params_control.sensor_list,
                       UP0, pin_pressure=True, cluster_rows=0, **kw)
    plan = prob.plans[2]
    fwd = np.arange(len(plan.blk_K)) < plan.launch_ptr[plan.n_forward_launches]
    if variant == "presummed_gathers":
        assert (plan.blk_nsrc[fwd] == 1).all() and plan.asm_lptr[plan.n_forward_launches] > 0
    else:
        assert (plan.blk_nsrc[fwd] == 3).any()
    case = cases.lidcavity(1000.0)
    xy, tri = cases.load_mesh(case.mesh_file)
    B = 72
    probe = [0, 35, 71]
    amp = 0.02 + 0.1 * np.arange(B) / (B - 1)
    oracles = []
    for b in probe:
        orc = FlowOracle(case, xy, tri)
        orc.set_base_flow(UP0)
        orc.init_time_stepping()
        oracles.append(orc)
    ens = Ensemble(prob, B)
    ens.set_state(oracles[0].ic[: tab.Nv], None, oracles[0].ic[tab.Nv :], order=1)
    for k in range(6):
        uc = (amp * np.cos(0.9 * k))[None, :]
        ens.step(uc)
        for orc, b in zip(oracles, probe):
            orc.step([uc[0, b]])
            assert np.allclose(ens.y_meas[:, b], orc.y_meas, rtol=SERIES_TOL, atol=0)
            assert np.isclose(ens.dE[b], orc.dE, rtol=SERIES_TOL)
    up = ens.fields(0)
    for orc, b in zip(oracles, probe):
        assert rel(up[: tab.Nv, b], orc.up[: tab.Nv]) < FIELD_TOL
    assert not ens.diverged.any()
    ens.close()


def test_facade_ensemble_checkpoint_and_restart(root, tmp_path, built_lib):
    """Ensemble (batch = 5) through the FlowSolver facade: the XDMF/HDF5 checkpoints hold one function per trajectory, a
    restarted ensemble continues EVERY trajectory exactly where it was (the restart goes through full fields and back, so
    to round-off of one addition), the reference-shaped accessors give trajectory 0 and ``.ensemble`` all of them, and a
    single-trajectory run restarts from the same files (trajectory 0)."""
    from flowcontrol_b200.examples import lidcavity as ex
    from flowcontrol_b200.exporter import read_checkpoint, write_checkpoint
    from flowcontrol_b200.hdf5_lite import HDF5LiteFile

    UP0 = np.load(root / "tests/golden/lidcavity_baseflow.npz")["UP0"]
    B, dt = 5, 0.005

    def make(batch, Tstart=0.0, num_steps=6):
        fs = ex.LidCavityFlowSolver.make_default(Re=1000, path_out=tmp_path, num_steps=num_steps, save_every=3, Tstart=Tstart, batch=batch)
        tab = fs.tables
        write_checkpoint(fs.paths.U0, "U0", UP0[: tab.Nv], 0.0, append=False, tab=tab)
        write_checkpoint(fs.paths.P0, "P0", UP0[tab.Nv :], 0.0, append=False, tab=tab)
        fs.load_steady_state()
        return fs

    fs = make(B)
    tab = fs.tables
    rng = np.random.default_rng(4)
    loc = rng.uniform(0.3, 0.7, size=(B, 2))
    ic = np.stack([0.1 * fs._default_initial_perturbation(xloc=x, yloc=y, radius=0.1) for x, y in loc], axis=1)
    fs.params_ic.amplitude = 0.0
    fs.initialize_time_stepping(ic=ic)
    amp = np.linspace(-0.05, 0.08, B)
    for k in range(6):
        y = fs.step(u_ctrl=(amp * np.cos(0.7 * k))[:, None])
        assert y.shape == (B, 2)
    up_full = fs.fields.up_.ensemble.copy()  # [N, B]
    assert up_full.shape == (tab.N, B) and np.array_equal(fs.fields.up_.vector().get_local(), up_full[:, 0])
    assert np.array_equal(fs.fields.u_nn.ensemble.shape, (tab.Nv, B))
    assert rel(up_full[:, 1], up_full[:, 0]) > 1e-3  # the trajectories really differ
    h5 = HDF5LiteFile(fs.paths.U_restart.with_suffix(".h5"))
    assert h5.keys("/") == ["U"] + [f"U_traj{b:04d}" for b in range(1, B)] and h5.keys("/U_traj0003") == [f"U_traj0003_{k}" for k in range(3)]
    # snapshot 1 (t = 3 dt) is the full field U0 + u of every trajectory
    snap = read_checkpoint(fs.paths.U_restart, 1, tab=tab, batch=B)
    assert snap.shape == (tab.Nv, B)
    # restart every trajectory at t = 3 dt and redo steps 4..6
    fs2 = make(B, Tstart=3 * dt, num_steps=3)
    fs2.initialize_time_stepping(Tstart=3 * dt)
    assert fs2.order == 2
    for k in range(3, 6):
        fs2.step(u_ctrl=(amp * np.cos(0.7 * k))[:, None])
    up2 = fs2.fields.up_.ensemble
    for b in range(B):
        assert rel(up2[: tab.Nv, b], up_full[: tab.Nv, b]) < 1e-12
    assert np.allclose(fs2.y_meas, fs.y_meas, rtol=1e-9, atol=1e-14)
    # a single run restarts from the same files: trajectory 0
    fs1 = make(1, Tstart=3 * dt, num_steps=3)
    fs1.initialize_time_stepping(Tstart=3 * dt)
    for k in range(3, 6):
        fs1.step(u_ctrl=[amp[0] * np.cos(0.7 * k)])
    assert rel(fs1.fields.u_.vector().get_local(), up_full[: tab.Nv, 0]) < 1e-12


def test_fun_array_runs_a_controller_population_as_ensembles(root, cyl):
    """fun_array (utils/optim.py:48-66) with an EnsembleCost: a population of controller parameters evaluated as ensembles
    on the device (two chunks, the second partial) gives the costs of the point-by-point loop of the reference -- each point
    simulated on its own through the host loop with compute_signal_cost / compute_control_cost."""
    from flowcontrol_b200.controller import Controller
    from flowcontrol_b200.costs import EnsembleCost, compute_control_cost, compute_signal_cost, fun_array
    from flowcontrol_b200.ensemble import Ensemble

    fs, prob, _, _ = cyl
    tab = prob.tab
    ic = fs._default_initial_perturbation()
    k = np.load(root / "tests/golden/Kopt_reduced13.npz")

    def make_controller(x):  # two parameters: loop gain and a scaling of the feed-through
        return Controller(k["A"], x[0] * k["B"], k["C"], x[0] * x[1] * k["D"])

    B, nsteps, pen = 32, 12, 0.3
    rng = np.random.default_rng(8)
    X = np.column_stack([rng.uniform(0.5, 1.5, 40), rng.uniform(0.0, 2.0, 40)])
    ens = Ensemble(prob, B)
    cost = EnsembleCost(ens, make_controller, ic[: tab.Nv], nsteps, Ky=[[-1.0, 0.0, 0.0]], Fu=[[1.0], [1.0]], u_penalty=pen, p_n=ic[tab.Nv :])
    J = fun_array(X, cost)
    assert J.shape == (40, 1) and np.all(np.isfinite(J)) and not cost.last["diverged"].any()
    assert np.isclose(cost(X[3]), J[3, 0], rtol=1e-13)  # scalar signature
    # the reference's loop, point by point, for a few points (host-stepped controller, logged series, reference cost functions)
    for i in (0, 17, 39):
        K = make_controller(X[i])
        e1 = Ensemble(prob, 1)
        e1.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
        dE, us = [], []
        for _ in range(nsteps):
            u = K.step(-e1.y_meas[0, 0], prob.dt)
            e1.step(np.array([[u[0]], [u[0]]]))
            dE.append(e1.dE[0])
            us.append([u[0], u[0]])
        Tnorm = prob.dt / (nsteps * prob.dt)
        ref = compute_signal_cost(np.array(dE), Tnorm, "integral") + pen * compute_control_cost(np.array(us), Tnorm)
        assert np.isclose(J[i, 0], ref, rtol=1e-10)
        e1.close()
    # any other callable is looped over as in the reference
    assert np.array_equal(fun_array(X[:3], lambda x: x.sum()), X[:3].sum(axis=1)[:, None])
    ens.close()


def test_device_matrix_assembly_matches_host_and_oracle(root, built_lib):
    """fcb_assemble_advection (k_assemble_advection: per-element matrices of C(U) and D^ij(U), scattered into CSR through the
    position map, one launch per colour) for an ensemble of velocity fields vs the host blocks (fem.py) and vs the oracle's
    independent assembly (16-point rule), and bit-reproducible."""
    from flowcontrol_b200.assembly import DeviceAdvectionAssembler
    from flowcontrol_b200.fem import ScalarBlocks
    from flowcontrol_b200.mesh import TaylorHoodTables
    from oracle.flow_oracle import Operators, TaylorHoodMesh

    tab = TaylorHoodTables.from_file(root / "data/meshes/lidcavity_mesh64.npz")
    blocks = ScalarBlocks(tab)
    asm = DeviceAdvectionAssembler(tab, blocks)
    rng = np.random.default_rng(11)
    B = 5
    U = rng.standard_normal((tab.Nv, B))
    Cv, Dv = asm.values(U)
    Cv2, Dv2 = asm.values(U)
    assert np.array_equal(Cv, Cv2) and np.array_equal(Dv, Dv2)
    xy, tri = cases.load_mesh("lidcavity_mesh64")
    ops = Operators(TaylorHoodMesh(xy, tri))
    for b in (0, 3):
        Ch, Dh = blocks.advection(U[:, b])
        Cd = asm._csr(Cv[:, b])
        assert abs(Cd - Ch).max() < 1e-13 * abs(Ch).max()
        Co, Do = ops.advection_blocks(U[:, b])
        assert abs(Cd - Co).max() < 1e-12 * abs(Co).max()
        for i in range(2):
            for j in range(2):
                Dd = asm._csr(Dv[2 * i + j, :, b])
                assert abs(Dd - Dh[i, j]).max() < 1e-13 * abs(Dh[i, j]).max()
                assert abs(Dd - Do[i, j]).max() < 1e-12 * abs(Do[i, j]).max()
    A_dev = asm.saddle_point(1.5 / 0.005, 100.0, U[:, 1])
    A_host = blocks.saddle_point(1.5 / 0.005, 100.0, U[:, 1])
    assert abs(A_dev - A_host).max() < 1e-12 * abs(A_host).max()


def test_steady_state_newton_with_device_assembly(root, tmp_path, built_lib):
    """the base flow of the cylinder (Re=100; 3 Picard + Newton as tests/integration/test_cylinder.py)
    with every iteration matrix assembled on the GPU reproduces the reference's golden constants and the committed base flow."""
    from flowcontrol_b200.examples.cylinder import CylinderFlowSolver

    fs = CylinderFlowSolver.make_default(Re=100, path_out=tmp_path)
    fs.compute_steady_state(method="picard", max_iter=3, tol=1e-7, u_ctrl=[0.0, 0.0], assembly="device")
    fs.compute_steady_state(method="newton", max_iter=25, u_ctrl=[0.0, 0.0], initial_guess=fs.fields.UP0, assembly="device")
    U0 = fs.fields.U0.vector().get_local()
    assert np.isclose(U0.max(), 1.1921615450014942, rtol=1e-9) and np.isclose(U0.mean(), 0.336746427968607, rtol=1e-9)  # test_cylinder.py:66-67
    UP0 = np.load(root / "tests/golden/cylinder_baseflow.npz")["UP0"]
    assert rel(fs.fields.UP0.vector().get_local(), UP0) < 1e-9
    # lid-driven cavity Re=1000 (enclosed flow, pinned pressure; 40 Picard iterations as tests/integration/test_lidcavity.py)
    from flowcontrol_b200.examples.lidcavity import LidCavityFlowSolver

    fl = LidCavityFlowSolver.make_default(Re=1000, path_out=tmp_path / "lid")
    fl.compute_steady_state(method="picard", max_iter=40, tol=1e-7, u_ctrl=[0.0], assembly="device")
    L0 = np.load(root / "tests/golden/lidcavity_baseflow.npz")["UP0"]
    Nv = fl.tables.Nv
    assert rel(fl.fields.UP0.vector().get_local()[:Nv], L0[:Nv]) < 1e-9
    assert np.isclose(fl.fields.U0.vector().get_local().mean(), 0.0020234251738529907, rtol=1e-6)  # test_lidcavity.py:48


def test_run_loops_chunked_series_open_loop_and_controller_state(root, cyl, monkeypatch):
    """The device-resident loops (include/fcb200.h): (a) the series is streamed in chunks -- a run with 7-step chunks gives
    bit for bit the series of a run with one chunk; (b) fcb_run_open_loop == stepping with fcb_step from the host;
    (c) fcb_set_controller_state restores the controller states (Controller.reset, controller.py:161-163): a repeated
    closed-loop run reproduces the first one exactly, and fcb_set_state alone does not touch them."""
    from flowcontrol_b200.controller import Controller, ControllerBank
    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.sharding import controller_gain_sweep

    fs, prob, _, _ = cyl
    tab = prob.tab
    ic = fs._default_initial_perturbation()
    B, n = 40, 23
    k = np.load(root / "tests/golden/Kopt_reduced13.npz")
    bank = ControllerBank([Controller(k["A"], g * k["B"], k["C"], g * k["D"]) for g in controller_gain_sweep(B)], prob.dt,
                          np.array([[-1.0, 0.0, 0.0]]), np.array([[1.0], [1.0]]))

    def closed(chunk):
        if chunk:
            monkeypatch.setenv("FCB_SERIES_CHUNK", str(chunk))
        else:
            monkeypatch.delenv("FCB_SERIES_CHUNK", raising=False)
        e = Ensemble(prob, B)
        e.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
        e.set_controllers(bank)
        return e, e.run_closed_loop(n)

    e1, s1 = closed(0)
    e7, s7 = closed(7)
    assert np.array_equal(s1, s7) and np.array_equal(e1.fields(0), e7.fields(0))
    e7.close()
    # (c) same handle: new state, controller states still those of the end of the run -> different series; reset -> identical
    x_end = e1.controller_state().copy()
    assert np.abs(x_end).max() > 0
    e1.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    assert np.array_equal(e1.controller_state(), x_end)
    e1.set_controller_state(None)
    assert np.abs(e1.controller_state()).max() == 0.0
    s1b = e1.run_closed_loop(n)
    assert np.array_equal(s1b, s1)
    e1.set_controller_state(x_end)
    assert np.array_equal(e1.controller_state(), x_end)
    e1.close()
    # (b) open loop: device-resident run vs host-driven steps
    rng = np.random.default_rng(5)
    useries = 0.1 * rng.standard_normal((n, 2, B))
    eo = Ensemble(prob, B)
    eo.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    so = eo.run_open_loop(useries)
    eh = Ensemble(prob, B)
    eh.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    for s_ in range(n):
        eh.step(useries[s_])
        assert np.array_equal(so[s_, 3:, :], eh.y_meas) and np.array_equal(so[s_, 0, :], eh.dE)
        assert np.array_equal(so[s_, 1:3, :], useries[s_])
    assert np.array_equal(eo.fields(0), eh.fields(0))
    eo.close()
    eh.close()


def test_device_factorisation_matches_host_blocks_and_drives_the_step(root, built_lib):
    """fcb_factorize (k_ff_assemble / k_ff_invert / k_ff_gemm: multifrontal fronts assembled, inverted with partial pivoting
    and multiplied on the GPU) gives the blocks of the host factorisation (multifrontal.BlockFactor, LAPACK) to round-off,
    flags a singular matrix, and a FlowProblem factorised on the device steps like the oracle."""
    import tempfile
    from pathlib import Path

    import scipy.sparse as sp

    from flowcontrol_b200.devfactor import DeviceBlockFactor, FrontMaps
    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.examples import lidcavity as ex
    from flowcontrol_b200.flowfield import Field
    from flowcontrol_b200.multifrontal import BlockFactor
    from flowcontrol_b200.problem import FlowProblem

    UP0 = np.load(root / "tests/golden/lidcavity_baseflow.npz")["UP0"]
    fs = ex.LidCavityFlowSolver.make_default(Re=1000.0, path_out=Path(tempfile.mkdtemp()))
    tab = fs.tables
    fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
    kw = dict(pin_pressure=True)
    prob_h = FlowProblem(tab, fs.blocks, 1000.0, 0.005, fs.bc.bcu, fs.params_control.actuator_list, fs.params_control.sensor_list, UP0, **kw)
    prob_d = FlowProblem(tab, fs.blocks, 1000.0, 0.005, fs.bc.bcu, fs.params_control.actuator_list, fs.params_control.sensor_list, UP0,
                         symbolic=prob_h.sym, factor_device=0, **kw)
    for order in (1, 2):
        worst = 0.0
        for (E, Fi, G), (E2, Fi2, G2) in zip(prob_h.factors[order].blocks, prob_d.factors[order].blocks):
            for a, b in ((E, E2), (Fi, Fi2), (G, G2)):
                if a.size:
                    worst = max(worst, float(np.abs(a - b).max() / np.abs(a).max()))
        assert worst < 1e-9, worst
    rng = np.random.default_rng(2)
    b = rng.standard_normal(prob_h.sym.n)
    xh, xd = prob_h.factors[2].solve(b), prob_d.factors[2].solve(b)
    assert rel(xd, xh) < 1e-11
    # the maps are reused for a matrix with the same sparsity; a singular matrix is reported
    A2 = prob_h.A_raw[2]
    maps = prob_d.factors[2].maps
    again = DeviceBlockFactor(prob_h.sym, A2, maps=maps)
    assert again.maps is maps and np.array_equal(again.blocks[0][1], prob_d.factors[2].blocks[0][1])
    dead = sp.csr_matrix(A2).copy().tolil()
    r = int(prob_h.sym.perm[0])
    dead[r, :] = 0.0
    dead[:, r] = 0.0
    with pytest.raises(np.linalg.LinAlgError):
        DeviceBlockFactor(prob_h.sym, dead.tocsr() + 0.0 * sp.csr_matrix(A2), maps=None)
    # the device-factorised problem steps like the oracle
    case = cases.lidcavity(1000.0)
    xy, tri = cases.load_mesh(case.mesh_file)
    orc = FlowOracle(case, xy, tri)
    orc.set_base_flow(UP0)
    orc.init_time_stepping()
    B = 48
    ens = Ensemble(prob_d, B)
    ens.set_state(orc.ic[: tab.Nv], None, orc.ic[tab.Nv :], order=1)
    for k in range(5):
        uc = 0.04 * np.cos(0.8 * k)
        orc.step([uc])
        ens.step(np.full((1, B), uc))
        assert np.allclose(ens.y_meas[:, 7], orc.y_meas, rtol=SERIES_TOL, atol=0)
    assert rel(ens.fields(0)[: tab.Nv, 7], orc.up[: tab.Nv]) < FIELD_TOL
    ens.close()


def test_steady_state_newton_fully_on_device_assembly_and_factorisation(root, tmp_path, built_lib):
    """Newton / Picard base flow of the cylinder with every iteration matrix assembled AND factorised on
    the GPU (the two sweeps of the single right-hand side run on the host from the device-computed blocks)."""
    from flowcontrol_b200.examples.cylinder import CylinderFlowSolver

    fs = CylinderFlowSolver.make_default(Re=100, path_out=tmp_path)
    fs.compute_steady_state(method="picard", max_iter=3, tol=1e-7, u_ctrl=[0.0, 0.0], assembly="device", factor="device")
    fs.compute_steady_state(method="newton", max_iter=25, u_ctrl=[0.0, 0.0], initial_guess=fs.fields.UP0, assembly="device", factor="device")
    U0 = fs.fields.U0.vector().get_local()
    assert np.isclose(U0.max(), 1.1921615450014942, rtol=1e-9) and np.isclose(U0.mean(), 0.336746427968607, rtol=1e-9)  # test_cylinder.py:66-67
    UP0 = np.load(root / "tests/golden/cylinder_baseflow.npz")["UP0"]
    assert rel(fs.fields.UP0.vector().get_local(), UP0) < 1e-9
