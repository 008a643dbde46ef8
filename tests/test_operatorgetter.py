"""OperatorGetter (reference: src/flowcontrol/operatorgetter.py, tests/integration/test_operatorgetter.py) on the
cylinder at Re=100: Frobenius-norm regression constant of the reference, finite-difference validation of A against the
steady residual, lifting consistency of B, C rows vs Sensor.eval, and the frequency response against a dense solve."""
import tempfile
from pathlib import Path

import numpy as np
import pytest
import scipy.sparse as sp

from flowcontrol_b200.operatorgetter import OperatorGetter, get_frequency_response_sequential


@pytest.fixture(scope="module")
def fs_cylinder(root):
    from flowcontrol_b200.examples.cylinder import CylinderFlowSolver
    from flowcontrol_b200.flowfield import Field

    UP0 = np.load(root / "tests/golden/cylinder_baseflow.npz")["UP0"]
    fs = CylinderFlowSolver.make_default(path_out=Path(tempfile.mkdtemp()))
    tab = fs.tables
    fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
    return fs, UP0


def test_get_A_frobenius_regression_and_shapes(fs_cylinder):
    fs, UP0 = fs_cylinder
    og = OperatorGetter(fs)
    A, E, B, C = og.get_all()
    N = fs.tables.N
    assert A.shape == (N, N) and E.shape == (N, N) and B.shape == (N, 2) and C.shape == (3, N)
    assert np.isclose(np.sqrt((A.data**2).sum()), 55.37024024761875, rtol=1e-11)  # test_operatorgetter.py:23-26
    assert abs(E[fs.tables.Nv :, :]).sum() == 0 and abs(E[:, fs.tables.Nv :]).sum() == 0  # no pressure mass
    assert np.isclose(E.sum(), 2 * 20 * 30 - 2 * np.pi * 0.25, rtol=1e-3)  # two components x area of the domain minus cylinder


def test_get_A_matches_finite_differences_of_the_steady_residual(fs_cylinder):
    """A x = -(F(UP0 + h x) - F(UP0)) / h on interior dofs (test_operatorgetter.py:111-140)."""
    from flowcontrol_b200.problem import DirichletSet
    from flowcontrol_b200.steadystate import SteadyStateSolver

    fs, UP0 = fs_cylinder
    tab = fs.tables
    A = OperatorGetter(fs).get_A()
    dset = DirichletSet(tab, fs.bc.bcu, fs.params_control.actuator_list)
    sol = SteadyStateSolver(tab, fs.blocks, fs.params_flow.Re, dset)
    interior = np.setdiff1d(np.arange(tab.N), dset.dofs)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(tab.N)
    x[dset.dofs] = 0.0
    h = 1e-6
    fd = -(sol.residual(UP0 + h * x) - sol.residual(UP0)) / h
    Ax = A @ x
    assert np.linalg.norm(Ax[interior] - fd[interior]) / np.linalg.norm(fd[interior]) < 1e-5
    assert np.array_equal(Ax[dset.dofs], x[dset.dofs])  # identity rows


def test_B_lifting_and_C_rows(fs_cylinder):
    fs, UP0 = fs_cylinder
    tab = fs.tables
    og = OperatorGetter(fs)
    B, C = og.get_B(), og.get_C()
    # a BC actuator only loads the rows of the cells touching its slot
    assert 0 < np.count_nonzero(B[:, 0]) < 600 and 0 < np.count_nonzero(B[:, 1]) < 600
    # the two slots are mirror images on a nearly symmetric mesh: similar column norms
    assert np.isclose(np.linalg.norm(B[:, 0]), np.linalg.norm(B[:, 1]), rtol=0.05)
    up = np.random.default_rng(1).standard_normal(tab.N)
    for s, sensor in enumerate(fs.params_control.sensor_list):
        assert np.isclose(C[s] @ up, sensor.eval(up), rtol=1e-12)  # test_sensor.py:172-181 / test_operatorgetter.py:238-254


def test_frequency_response_matches_dense_solve():
    rng = np.random.default_rng(3)
    n = 30
    A = sp.csr_matrix(-2.0 * np.eye(n) + 0.3 * rng.standard_normal((n, n)))
    Q = sp.diags(rng.uniform(0.5, 2.0, n)).tocsr()
    B, C = rng.standard_normal((n, 2)), rng.standard_normal((3, n))
    ww = np.array([0.0, 0.7, 5.0])
    H, w = get_frequency_response_sequential(A, B, C, Q, ww)
    assert H.shape == (3, 2, 3) and np.array_equal(w, ww)
    for i, wi in enumerate(ww):
        ref = C @ np.linalg.solve(1j * wi * Q.toarray() - A.toarray(), B)
        assert np.allclose(H[:, :, i], ref, rtol=1e-10, atol=1e-12)


def test_doubled_symbolic_structure_solves_the_complex_shifted_system():
    """devfactor.DoubledSymbolic: the real 2n x 2n form of (jwE - A) x = B (utils/linalg.py:215) ordered by the mesh's
    nested dissection with real and imaginary parts adjacent; factorised here with the host BlockFactor (the GPU test runs
    the same structure through fcb_factorize) and compared with a dense complex solve."""
    import scipy.sparse as sp

    from flowcontrol_b200.devfactor import DoubledSymbolic
    from flowcontrol_b200.fem import ScalarBlocks
    from flowcontrol_b200.mesh import TaylorHoodTables
    from flowcontrol_b200.multifrontal import BlockFactor, SymbolicFactor
    from util import unit_square_mesh

    xy, tri = unit_square_mesh(7, jitter=0.15, seed=5)
    tab = TaylorHoodTables.from_arrays(xy, tri)
    bl = ScalarBlocks(tab)
    rng = np.random.default_rng(1)
    L = bl.saddle_point(0.0, 50.0, 0.3 * rng.standard_normal(tab.Nv))
    x_, y_ = tab.node_xy[:, 0], tab.node_xy[:, 1]
    bn = np.flatnonzero((np.abs(x_) < 1e-12) | (np.abs(x_ - 1) < 1e-12) | (np.abs(y_) < 1e-12) | (np.abs(y_ - 1) < 1e-12))
    dofs = np.concatenate([bn, bn + tab.nN, [tab.Nv]])
    keep = np.ones(tab.N)
    keep[dofs] = 0
    A = (sp.diags(keep) @ sp.csr_matrix(-L) + sp.csr_matrix((np.ones(len(dofs)), (dofs, dofs)), shape=(tab.N, tab.N))).tocsr()
    E = (sp.diags(keep) @ sp.bmat([[bl.Mv, None], [None, sp.csr_matrix((tab.nV, tab.nV))]], format="csr")).tocsr()
    sym2 = DoubledSymbolic(SymbolicFactor(tab, np.ones(tab.N, dtype=bool), leaf_cells=4))
    n = tab.N
    Bm = rng.standard_normal((n, 2))
    rhs = np.vstack([Bm, np.zeros_like(Bm)])
    for w in (0.0, 1.3, 40.0):
        fac = BlockFactor(sym2, sp.bmat([[-A, -w * E], [w * E, -A]], format="csr"))
        x = np.empty_like(rhs)
        x[sym2.perm] = fac.solve(rhs[sym2.perm])
        ref = np.linalg.solve(1j * w * E.toarray() - A.toarray(), Bm)
        assert np.abs((x[:n] + 1j * x[n:]) - ref).max() < 1e-11 * np.abs(ref).max()


@pytest.mark.gpu
def test_frequency_response_on_the_device_matches_host_lu(root, built_lib):
    """H(w) = C (jwE - A)^-1 B of the linearised lid-driven cavity with one multifrontal factorisation
    per frequency on the GPU (devfactor.frequency_response_device through OperatorGetter.get_frequency_response) against
    the reference's algorithm, one host sparse LU of the real 2n x 2n form per frequency (utils/linalg.py:192-240)."""
    from flowcontrol_b200.examples.lidcavity import LidCavityFlowSolver
    from flowcontrol_b200.flowfield import Field

    UP0 = np.load(root / "tests/golden/lidcavity_baseflow.npz")["UP0"]
    fs = LidCavityFlowSolver.make_default(Re=1000, path_out=Path(tempfile.mkdtemp()))
    tab = fs.tables
    fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
    og = OperatorGetter(fs)
    # the enclosed flow's pressure is defined up to a constant: pin it, as the stepping path does
    A = og.get_A().tolil()
    A[tab.Nv, :] = 0.0
    A[tab.Nv, tab.Nv] = 1.0
    A = A.tocsr()
    E, B, C = og.get_mass_matrix(), og.get_B(), og.get_C()
    from flowcontrol_b200.devfactor import frequency_response_device

    ww = np.array([0.0, 0.8, 12.0])
    Hd, _ = frequency_response_device(A, B, C, E, ww, tab, device=0)
    Hh, _ = get_frequency_response_sequential(A, B, C, E, ww)
    assert Hd.shape == (2, 1, 3)
    assert np.abs(Hd - Hh).max() < 1e-8 * np.abs(Hh).max()
    assert np.abs(Hd.imag[:, :, 0]).max() == 0.0  # w = 0: a real solve
