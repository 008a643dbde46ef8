"""Host-side mirror of the reference API: controller, exporter, parameters, facade validation."""
import json

import numpy as np
import pandas as pd
import pytest

from flowcontrol_b200 import flowsolverparameters as fsp
from flowcontrol_b200.controller import Controller, ControllerBank
from flowcontrol_b200.exporter import FlowExporter, read_checkpoint
from flowcontrol_b200.flowfield import Field, FlowFieldCollection, SimPaths


def _siso():
    return Controller.from_matrices(
        A=np.array([[1, 1, 1], [0.2, -1, 0], [0.0, 1.0, 1.0]]), B=np.array([[0], [1], [0.5]]),
        C=np.array([0.5, 0.2, 0]), D=0, x0=np.array([1.0, 2.0, 3.0]),
    )


def test_controller_step_matches_zoh_recursion():
    """tests/test_controller.py:182-196 of the reference: multi-step recursion vs c2d."""
    from scipy.signal import cont2discrete

    K = _siso()
    Ad, Bd, Cd, Dd, _ = cont2discrete((K.A, K.B, K.C, K.D), 0.1, method="zoh")
    x = K.x.copy()
    for y in (1.2, -0.3, 0.7):
        u = K.step(np.array([y]), 0.1)
        assert np.allclose(u, Cd @ x + Dd @ np.array([y]))
        x = Ad @ x + Bd @ np.array([y])
        assert np.allclose(K.x, x)
    K.step(np.array([0.1]), 0.05)  # dt change re-discretises (test_controller.py:198-213)
    assert K._dt == 0.05
    K.reset()
    assert np.all(K.x == 0)


def test_controller_from_npz_fixture(root):
    K = Controller.from_file(root / "tests/golden/Kopt_reduced13.npz")
    assert K.nstates == 13 and K.ninputs == 1 and K.noutputs == 1
    assert np.isclose(K.D[0, 0], -7.4e-4, rtol=0.05)


def test_controller_bank_packing_and_padding():
    K1, K2 = _siso(), Controller.from_matrices(A=[[-1.0]], B=[[2.0]], C=[[3.0]], D=[[0.5]])
    bank = ControllerBank([K1, K2], 0.1, Ky=np.array([[-1.0, 0, 0]]), Fu=np.array([[1.0], [1.0]]))
    assert (bank.nx, bank.ny, bank.nu, bank.B) == (3, 1, 1, 2)
    Ad2 = bank.Ad[:, 1].reshape(3, 3)
    assert np.isclose(Ad2[0, 0], np.exp(-0.1)) and np.all(Ad2[1:, :] == 0) and np.all(Ad2[:, 1:] == 0)
    assert np.allclose(bank.x0[:, 0], [1, 2, 3]) and np.all(bank.x0[:, 1] == 0)


def _paths(tmp_path):
    names = ["U0", "P0", "U", "P", "Uprev", "U_restart", "Uprev_restart", "P_restart"]
    kw = {n: tmp_path / f"{n}.xdmf" for n in names}
    return SimPaths(timeseries=tmp_path / "ts.csv", metadata=tmp_path / "meta.json", steady_meta=tmp_path / "steady.json",
                    mesh=tmp_path / "m.xdmf", **kw)


def _small_tab():
    from flowcontrol_b200.mesh import TaylorHoodTables
    from util import unit_square_mesh

    xy, tri = unit_square_mesh(3, jitter=0.1, seed=2)
    return TaylorHoodTables.from_arrays(xy, tri)


def test_exporter_columns_csv_and_sidecar(tmp_path):
    """Column names/order and JSON keys of exporter.py:186-262; field snapshots through the XDMF/HDF5 checkpoint files."""
    tab = _small_tab()
    rng = np.random.default_rng(0)
    U0, P0 = rng.standard_normal(tab.Nv), rng.standard_normal(tab.nV)
    fields = FlowFieldCollection(U0=Field(U0), P0=Field(P0))
    ex = FlowExporter(_paths(tmp_path), fields, Tstart=0.0, dt=0.005, save_every=5, tab=tab)
    ex.log_ic(0.0, np.array([0.1, 0.2]), 0.5)
    ex.log(np.array([1.0]), np.array([0.3, 0.4]), 0.4, 0.005, 1e-3)
    df = ex.to_dataframe()
    assert list(df.columns) == ["time", "dE", "runtime", "y_meas_1", "y_meas_2", "u_ctrl_1"]
    assert np.isnan(df.loc[0, "u_ctrl_1"]) and df.loc[1, "u_ctrl_1"] == 1.0
    u1, u2, p1 = rng.standard_normal(tab.Nv), rng.standard_normal(tab.Nv), rng.standard_normal(tab.nV)
    ex.export_xdmf(Field(u1), Field(np.zeros(tab.Nv)), Field(p1), time=0.025, append=False, adjust_baseflow=1.0)
    ex.export_xdmf(Field(u2), Field(u1), Field(p1), time=0.05, adjust_baseflow=1.0)
    assert np.array_equal(read_checkpoint(tmp_path / "U_restart.xdmf", 1, tab=tab), u2 + U0)
    assert np.array_equal(read_checkpoint(tmp_path / "U_restart.xdmf", 0, tab=tab), u1 + U0)
    assert np.array_equal(read_checkpoint(tmp_path / "Uprev_restart.xdmf", -1, tab=tab), u1 + U0)
    assert np.array_equal(read_checkpoint(tmp_path / "P_restart.xdmf", -1, tab=tab), p1 + P0)
    assert np.allclose(fields.Usave.array, u2 + U0)
    ex.write_metadata(restart_order=2)
    ex.write_timeseries()
    meta = json.loads((tmp_path / "meta.json").read_text())
    assert meta == {"Tstart": 0.0, "dt": 0.005, "save_every": 5, "checkpoints_written": 2, "restart_order": 2,
                    "files": {"U": "U_restart.xdmf", "Uprev": "Uprev_restart.xdmf", "P": "P_restart.xdmf"}}
    assert list(pd.read_csv(tmp_path / "ts.csv").columns) == list(df.columns)
    ex.reset()
    assert ex.to_dataframe().empty


def test_xdmf_checkpoint_layout_roundtrip_and_foreign_numbering(tmp_path):
    """The dolfin write_checkpoint layout (utils/io.py:21-50): XDMF attributes and HDF5 dataset tree, appended time steps
    that share the first mesh, one function per trajectory of an ensemble, and a file written with ANOTHER dof numbering,
    cell order and cell-vertex order (what dolfin itself would write) read back into canonical numbering."""
    from flowcontrol_b200 import xdmf_checkpoint as xc
    from flowcontrol_b200.exporter import write_checkpoint
    from flowcontrol_b200.hdf5_lite import HDF5LiteFile, read_all, write_hdf5

    tab = _small_tab()
    rng = np.random.default_rng(1)
    f = tmp_path / "U.xdmf"
    snaps = rng.standard_normal((3, tab.Nv))
    for k, t in enumerate((0.0, 0.05, 0.1)):
        xc.write_checkpoint(f, "U", tab, snaps[k], "V", t, append=k > 0)
    text = f.read_text()
    assert text.count('ItemType="FiniteElementFunction" ElementFamily="CG" ElementDegree="2" ElementCell="triangle" Name="U"') == 3
    assert text.count("<Topology") == 1 and text.count("xi:include") == 2  # later steps share the first step's mesh
    assert xc.checkpoint_times(f, "U") == [0.0, 0.05, 0.1]
    h5 = HDF5LiteFile(tmp_path / "U.h5")
    assert h5.keys("/U") == ["U_0", "U_1", "U_2"]
    assert h5.keys("/U/U_0") == ["cell_dofs", "cells", "mesh", "vector", "x_cell_dofs"] and h5.keys("/U/U_0/mesh") == ["geometry", "topology"]
    assert h5.read("/U/U_1/vector").shape == (tab.Nv, 1) and h5.read("/U/U_0/cell_dofs").shape == (12 * tab.nT, 1)
    assert np.array_equal(h5.read("/U/U_0/x_cell_dofs").ravel(), 12 * np.arange(tab.nT + 1))
    topo = h5.read("/U/U_0/mesh/topology")
    assert np.all(np.diff(topo, axis=1) > 0)  # dolfin orders the vertices of a cell
    for k in range(3):
        assert np.array_equal(xc.read_checkpoint(f, "U", tab, "V", k), snaps[k])
    # pressure function added to the same file
    pv = rng.standard_normal(tab.nV)
    xc.write_checkpoint(f, "P", tab, pv, "P", 0.1, append=True)
    assert np.array_equal(xc.read_checkpoint(f, "P", tab, "P", -1), pv) and np.array_equal(xc.read_checkpoint(f, "U", tab, "V", -1), snaps[2])
    # ensemble: one function per trajectory
    ens = rng.standard_normal((tab.Nv, 4))
    g = tmp_path / "E.xdmf"
    write_checkpoint(g, "U", ens, 0.0, append=False, tab=tab)
    write_checkpoint(g, "U", 2 * ens, 0.05, append=True, tab=tab)
    assert np.array_equal(read_checkpoint(g, 0, tab=tab, batch=4), ens) and np.array_equal(read_checkpoint(g, -1, tab=tab, batch=4), 2 * ens)
    assert np.array_equal(read_checkpoint(g, -1, tab=tab), 2 * ens[:, 0])  # a single run restarts from trajectory 0
    assert HDF5LiteFile(tmp_path / "E.h5").keys("/") == ["U", "U_traj0001", "U_traj0002", "U_traj0003"]
    # foreign numbering: permute dofs, cell order and the vertex order inside the cells, keep the dolfin conventions
    data = read_all(tmp_path / "U.h5")
    perm = rng.permutation(tab.Nv)  # file dof j holds canonical dof perm[j]
    inv = np.argsort(perm)
    corder = rng.permutation(tab.nT)
    cd = xc.cell_dofs(tab, "V")[corder]
    foreign = {"/U/U_0/vector": snaps[1][perm][:, None], "/U/U_0/cell_dofs": inv[cd].ravel()[:, None],
               "/U/U_0/x_cell_dofs": data["/U/U_0/x_cell_dofs"], "/U/U_0/cells": corder[:, None].astype(np.int64),
               "/U/U_0/mesh/topology": data["/U/U_0/mesh/topology"][corder], "/U/U_0/mesh/geometry": data["/U/U_0/mesh/geometry"]}
    write_hdf5(tmp_path / "F.h5", foreign)
    (tmp_path / "F.xdmf").write_text(text.replace("U.h5", "F.h5"))
    assert np.array_equal(xc.read_checkpoint(tmp_path / "F.xdmf", "U", tab, "V", 0), snaps[1])


def test_param_defaults_match_reference():
    assert fsp.ParamIC() == fsp.ParamIC(xloc=0.0, yloc=0.0, radius=1.0, amplitude=1.0)
    assert fsp.ParamSolver().time_scheme == "bdf" and fsp.ParamSolver().throw_error is True
    assert fsp.ParamTime(num_steps=10, dt=0.5, Tstart=0.0).Tfinal == 5.0
    assert fsp.ParamSave(path_out=".", save_every=0).energy_every == 1
    pc = fsp.ParamControl(sensor_list=[1, 2], actuator_list=[3])
    assert (pc.sensor_number, pc.actuator_number) == (2, 1)


def test_facade_validation_and_setup(tmp_path):
    from flowcontrol_b200.examples.cylinder import CylinderFlowSolver

    with pytest.raises(ValueError):
        fs = CylinderFlowSolver.make_default(path_out=tmp_path)
        fs._validate_params(fsp.ParamFlow(Re=-1), fs.params_time, fs.params_save, fs.params_solver, fs.params_mesh,
                            fs.params_control, fs.params_ic)
    with pytest.raises(FileNotFoundError):
        CylinderFlowSolver.make_default(path_out=tmp_path, meshpath=tmp_path / "nope.npz")
    fs = CylinderFlowSolver.make_default(path_out=tmp_path)
    assert list(fs.boundaries.index) == ["inlet", "outlet", "walls", "cylinder", "actuator_up", "actuator_lo"]
    assert fs.paths.timeseries.name == "timeseries1D_restart0,000.csv"
    with pytest.raises(ValueError):
        fs.set_actuators_u_ctrl([0.0])
    fs.set_actuators_u_ctrl([0.1, 0.2])
    assert fs.get_actuators_u_ctrl() == [0.1, 0.2]
    with pytest.raises(RuntimeError):
        fs.initialize_time_stepping()  # no base flow yet


def test_mesh_reader_on_reference_files(root):
    """hdf5_lite against the shipped XDMF/HDF5 meshes (only where /root/reference exists) and the fixtures."""
    from pathlib import Path

    from flowcontrol_b200.hdf5_lite import read_xdmf_mesh

    ref = Path("/root/reference/src/examples")
    if not ref.exists():
        pytest.skip("reference tree not present on this machine")
    for rel, fixture in (("cylinder/data_input/O1.xdmf", "cylinder_O1"), ("lidcavity/data_input/mesh64.xdmf", "lidcavity_mesh64")):
        xy, tri = read_xdmf_mesh(ref / rel)
        d = np.load(root / "data" / "meshes" / f"{fixture}.npz")
        assert np.array_equal(xy, d["vertices"]) and np.array_equal(tri, d["triangles"])


def test_make_solver_hook_solves_the_bc_applied_system(root, tmp_path):
    """FlowSolver._make_solver (flowsolver.py:812-814): object with set_operator / solve(x, b); here a host view of the
    multifrontal factor.  Checked against SuperLU on the oracle's BC-applied BDF2 matrix (lid cavity, pinned pressure)."""
    import scipy.sparse.linalg as spla

    from flowcontrol_b200.examples.lidcavity import LidCavityFlowSolver
    from flowcontrol_b200.flowfield import Field
    from oracle import cases
    from oracle.flow_oracle import FlowOracle

    UP0 = np.load(root / "tests/golden/lidcavity_baseflow.npz")["UP0"]
    fs = LidCavityFlowSolver.make_default(Re=1000.0, path_out=tmp_path)
    tab = fs.tables
    fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
    solver = fs._make_solver(order=2)
    solver.set_operator(None)
    case = cases.lidcavity(1000.0)
    xy, tri = cases.load_mesh(case.mesh_file)
    orc = FlowOracle(case, xy, tri)
    orc.set_base_flow(UP0)
    orc.prepare()
    rng = np.random.default_rng(0)
    b = rng.standard_normal(tab.N)
    b[orc.bc_pert.dofs] = 0.3 * rng.standard_normal(len(orc.bc_pert.dofs))
    b[tab.Nv] = 0.0  # pinned pressure dof
    x = np.zeros(tab.N)
    assert solver.solve(x, b) == 1
    ref = spla.splu(orc.A_bc[2]).solve(b)
    assert np.linalg.norm(x - ref) / np.linalg.norm(ref) < 1e-9
    assert fs.ensemble is None  # the hook alone never needs the GPU
