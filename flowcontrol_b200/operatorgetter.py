"""State-space operators of the linearised flow, E dq/dt = A q + B u, y = C q  (reference: src/flowcontrol/operatorgetter.py)
and the frequency response H(w) = C (jwE - A)^-1 B (reference: src/utils/linalg.py:192-240).

Host-side setup-time tooling (scipy), built from the same P2-P1 blocks, Dirichlet sets, actuator shapes and sensor rows
as the time-stepping path, in the canonical numbering [ux | uy | p].  Same names, arguments and sign conventions as the
reference (A = -dF/dq with the perturbation Dirichlet rows replaced by identity rows, B = -dF/du_ctrl); matrices come
back as scipy CSR instead of dolfin.PETScMatrix.  Pinned by the reference's own regression constant for ||A||_F
(tests/integration/test_operatorgetter.py:23-26) and by its finite-difference check (:111-140)."""
from __future__ import annotations

import logging

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .actuator import ACTUATOR_TYPE
from .problem import sensor_matrix

logger = logging.getLogger(__name__)


class OperatorGetter:
    def __init__(self, flowsolver):
        self.flowsolver = flowsolver

    # ── helpers ──────────────────────────────────────────────────────────────
    def _UP0(self, UP0):
        if UP0 is None:
            UP0 = self.flowsolver.fields.UP0
        return np.asarray(getattr(UP0, "array", UP0), dtype=np.float64)

    def _raw_jacobian(self, UP0) -> sp.csr_matrix:
        """-dF/dq without boundary conditions: -(C + D + K/Re) on the velocity block, +(p, div v) and +(div u, q)."""
        fs = self.flowsolver
        tab = fs.tables
        L = fs.blocks.saddle_point(0.0, fs.params_flow.Re, self._UP0(UP0)[: tab.Nv], shift=0.0, linearised=True)
        return (-L).tocsr()

    # ── operatorgetter.py:25-83 ──────────────────────────────────────────────
    def get_A(self, UP0=None, autodiff: bool = True, u_ctrl=None) -> sp.csr_matrix:
        """Linearised dynamic matrix A = -dF/dUP0 with the Dirichlet rows of the perturbation BCs replaced by identity
        rows (``bc.apply(Jac)``).  ``autodiff`` is accepted for call compatibility (both reference paths assemble the
        same bilinear form); ``u_ctrl`` does not enter the Jacobian (the actuation is affine)."""
        logger.info("Computing jacobian A...")
        fs = self.flowsolver
        A = self._raw_jacobian(UP0)
        dofs = self._dirichlet().dofs
        keep = np.ones(A.shape[0])
        keep[dofs] = 0.0
        ident = sp.csr_matrix((np.ones(len(dofs)), (dofs, dofs)), shape=A.shape)
        return (sp.diags(keep) @ A + ident).tocsr()

    def _dirichlet(self):
        from .problem import DirichletSet

        fs = self.flowsolver
        return DirichletSet(fs.tables, fs.bc.bcu, fs.params_control.actuator_list)

    # ── operatorgetter.py:85-106 ─────────────────────────────────────────────
    def get_mass_matrix(self) -> sp.csr_matrix:
        """Velocity mass matrix E on the mixed space (zero pressure block)."""
        logger.info("Computing mass matrix E...")
        tab = self.flowsolver.tables
        return sp.bmat([[self.flowsolver.blocks.Mv, None], [None, sp.csr_matrix((tab.nV, tab.nV))]], format="csr")

    # ── operatorgetter.py:108-193 ────────────────────────────────────────────
    def get_B(self, UP0=None) -> np.ndarray:
        """Actuation matrix [n_dof, n_actuators]: FORCE columns are the load vectors of the (P2-interpolated) unit
        profile, BC columns come from the lifting ``A_raw . w`` with ``w`` the unit actuator profile on its boundary dofs."""
        logger.info("Computing actuation matrix B...")
        fs = self.flowsolver
        tab = fs.tables
        acts = fs.params_control.actuator_list
        B = np.zeros((tab.N, len(acts)))
        dset = self._dirichlet()
        A_raw = self._raw_jacobian(UP0) if any(a.actuator_type is ACTUATOR_TYPE.BC for a in acts) else None
        for ii, a in enumerate(acts):
            if a.actuator_type is ACTUATOR_TYPE.FORCE:
                sx, sy = a.shape(tab.node_xy[:, 0], tab.node_xy[:, 1])
                B[: tab.Nv, ii] = fs.blocks.Mv @ np.concatenate([sx, sy])
            elif a.actuator_type is ACTUATOR_TYPE.BC:
                w = np.zeros(tab.N)
                w[dset.dofs] = dset.shape[ii]
                B[:, ii] = A_raw @ w
            else:
                raise NotImplementedError(f"Actuator type {a.actuator_type} not supported in get_B")
        return B

    # ── operatorgetter.py:195-240 ────────────────────────────────────────────
    def get_C(self) -> np.ndarray:
        """Measurement matrix [n_sensors, n_dof]: the sparse rows the device evaluates every step, densified."""
        logger.info("Computing measurement matrix C...")
        fs = self.flowsolver
        tab = fs.tables
        sensors = fs.params_control.sensor_list
        ptr, idx, val = sensor_matrix(tab, sensors)
        C = np.zeros((len(sensors), tab.N))
        for s in range(len(sensors)):
            np.add.at(C[s], idx[ptr[s] : ptr[s + 1]], val[ptr[s] : ptr[s + 1]])
        return C

    def get_frequency_response(self, ww, device: int | None = None, UP0=None):
        """H(w) = C (jwE - A)^-1 B of the linearised flow (utils/linalg.py:192-240).  ``device=None``: one host sparse LU
        per frequency (get_frequency_response_sequential); ``device=d``: one multifrontal factorisation per frequency on
        GPU ``d`` (devfactor.frequency_response_device)."""
        A, E, B, C = self.get_A(UP0=UP0), self.get_mass_matrix(), self.get_B(UP0=UP0), self.get_C()
        if device is None:
            return get_frequency_response_sequential(A, B, C, E, ww)
        from .devfactor import frequency_response_device

        pe = getattr(self.flowsolver, "params_ensemble", None)
        return frequency_response_device(A, B, C, E, ww, self.flowsolver.tables, device=device, leaf_cells=getattr(pe, "leaf_cells", 16))

    def get_all(self, autodiff: bool = True, u_ctrl=None) -> tuple:
        """(A, E, B, C) in one call (operatorgetter.py:242-267)."""
        return self.get_A(autodiff=autodiff, u_ctrl=u_ctrl), self.get_mass_matrix(), self.get_B(), self.get_C()


def get_frequency_response_sequential(A, B, C, Q, ww, verbose: bool = False):
    """H(w) = C (jwQ - A)^-1 B for every w in ``ww`` (utils/linalg.py:192-240): one real 2n x 2n sparse LU per
    frequency, [[-A, -wQ], [wQ, -A]] [xr; xi] = [B; 0].  Returns (H [ny, nu, nw] complex, ww)."""
    A = sp.csc_matrix(A)
    Q = sp.csc_matrix(Q)
    B = np.asarray(B, dtype=np.float64).reshape(A.shape[0], -1)
    C = np.asarray(C, dtype=np.float64).reshape(-1, A.shape[0])
    ww = np.asarray(ww, dtype=np.float64)
    n = A.shape[0]
    H = np.zeros((C.shape[0], B.shape[1], len(ww)), dtype=complex)
    rhs = np.vstack([B, np.zeros_like(B)])
    for ii, w in enumerate(ww):
        lu = spla.splu(sp.bmat([[-A, -w * Q], [w * Q, -A]], format="csc"))
        x = lu.solve(rhs)
        H[:, :, ii] = C @ x[:n] + 1j * (C @ x[n:])
        if verbose:
            logger.info("  [%d/%d] w=%.4e | max|H|=%.4e", ii + 1, len(ww), w, np.max(np.abs(H[:, :, ii])))
    return H, ww
