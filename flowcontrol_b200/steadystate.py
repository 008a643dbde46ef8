"""Base-flow (steady-state) solver — one-time host setup shared by an ensemble.

Restates /root/reference/src/flowcontrol/steadystate.py:60-159 (Newton through
``dolfin.solve(F == 0, ...)`` with dolfin's NewtonSolver defaults, hand-rolled
Picard loop) on the scalar blocks of fem.py with SciPy's SuperLU as the direct
solver.  Not on the per-step hot path.  SURVEY.md section 8(f) row f1: the iteration matrices can be assembled on the GPU
(``assembler=DeviceAdvectionAssembler(...)``: per-element matrices scattered into CSR through a position map, coloured)
and factorised on the GPU (``factor="device"``: devfactor.DeviceBlockFactor, multifrontal fronts inverted and multiplied
by hand-written kernels); the two triangular sweeps per iterate (one right-hand side) run on the host from those blocks.
"""

from __future__ import annotations

import logging

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .fem import ScalarBlocks
from .mesh import TaylorHoodTables
from .problem import DirichletSet

logger = logging.getLogger(__name__)


def _constrain_rows(A: sp.csr_matrix, dofs: np.ndarray) -> sp.csc_matrix:
    keep = np.ones(A.shape[0])
    keep[dofs] = 0.0
    return (sp.diags(keep) @ A + sp.diags(1.0 - keep)).tocsc()


class SteadyStateSolver:
    def __init__(self, tab: TaylorHoodTables, blocks: ScalarBlocks, Re: float, dirichlet: DirichletSet,
                 force: np.ndarray | None = None, verbose: bool = False, assembler=None, factor: str = "host", device: int = 0,
                 leaf_cells: int = 16):
        """``assembler``: an object with ``saddle_point(c, Re, U, linearised=...)`` that assembles the iteration matrix;
        None = the host blocks (fem.py), a ``DeviceAdvectionAssembler`` (assembly.py) = element matrices on the GPU.
        ``factor``: "host" = SuperLU on the row-constrained matrix; "device" = multifrontal factorisation of the free-free
        block on the GPU (devfactor.py; the Dirichlet unknowns are eliminated with their prescribed values, which is the
        same linear system)."""
        self.tab, self.blocks, self.Re, self.dirichlet = tab, blocks, Re, dirichlet
        self.asm = assembler if assembler is not None else blocks
        if factor not in ("host", "device"):
            raise ValueError(f"factor must be 'host' or 'device', got {factor!r}")
        self.factor, self.device = factor, int(device)
        self._sym = self._maps = None
        self._leaf_cells = leaf_cells
        self.force = np.zeros(tab.Nv) if force is None else force
        self.verbose = verbose
        self._Kv = sp.block_diag([blocks.K, blocks.K], format="csr")

    def _solve(self, A: sp.csr_matrix, b: np.ndarray, dofs: np.ndarray) -> np.ndarray:
        """Solve the system whose rows ``dofs`` are replaced by identity rows (x[dofs] = b[dofs])."""
        if self.factor == "host":
            return spla.splu(_constrain_rows(A, dofs)).solve(b)
        from .devfactor import DeviceBlockFactor
        from .multifrontal import SymbolicFactor

        if self._sym is None:
            self._sym = SymbolicFactor(self.tab, self.dirichlet.free, leaf_cells=self._leaf_cells)
        perm = self._sym.perm
        A = sp.csr_matrix(A)
        xc = b[dofs]
        rhs = (b - A[:, dofs] @ xc)[perm]
        fac = DeviceBlockFactor(self._sym, A, maps=self._maps, device=self.device)
        self._maps = fac.maps  # same sparsity in the next iteration: only the values are uploaded again
        x = np.empty_like(b)
        x[dofs] = xc
        x[perm] = fac.solve(rhs)
        return x

    def picard(self, UP0: np.ndarray, u_ctrl, max_iter: int = 10, tol: float = 1e-8) -> np.ndarray:
        """Fixed-point iteration with the advection velocity frozen at the previous
        iterate; stops on ||UP1-UP0|| / (||UP0|| + 1e-14) < tol (steadystate.py:139-157)."""
        tab = self.tab
        g = self.dirichlet.values(u_ctrl)
        dofs = self.dirichlet.dofs
        b = np.concatenate([self.force, np.zeros(tab.nV)])
        b[dofs] = g
        UP = np.array(UP0, dtype=np.float64)
        for i in range(max_iter):
            A = self.asm.saddle_point(0.0, self.Re, UP[: tab.Nv], linearised=False)
            UP1 = self._solve(A, b, dofs)
            rel = np.linalg.norm(UP1 - UP) / (np.linalg.norm(UP) + 1e-14)
            UP = UP1
            logger.info("Picard %d/%d  rel_err = %.3e", i + 1, max_iter, rel)
            if rel < tol:
                break
        return UP

    def residual(self, UP: np.ndarray) -> np.ndarray:
        """F(UP) of nsforms.py:137-147 tested against every basis function."""
        tab, bl = self.tab, self.blocks
        U, P = UP[: tab.Nv], UP[tab.Nv :]
        r = bl.convection(U) + self._Kv @ U / self.Re - self.force
        r[: tab.nN] -= bl.Bx.T @ P
        r[tab.nN :] -= bl.By.T @ P
        return np.concatenate([r, -(bl.Bx @ U[: tab.nN] + bl.By @ U[tab.nN :])])

    def newton(self, UP0: np.ndarray, u_ctrl, max_iter: int = 25, rtol: float = 1e-9, atol: float = 1e-10) -> np.ndarray:
        """dolfin NewtonSolver defaults: residual criterion, relaxation 1."""
        tab = self.tab
        g = self.dirichlet.values(u_ctrl)
        dofs = self.dirichlet.dofs
        UP = np.array(UP0, dtype=np.float64)
        r0 = None
        for it in range(max_iter + 1):
            b = self.residual(UP)
            b[dofs] = UP[dofs] - g
            r = float(np.linalg.norm(b))
            r0 = r if r0 is None else r0
            logger.info("Newton %d: r (abs) = %.3e  r (rel) = %.3e", it, r, r / max(r0, 1e-300))
            if r < atol or r < rtol * r0:
                return UP
            if it == max_iter:
                break
            J = self.asm.saddle_point(0.0, self.Re, UP[: tab.Nv], linearised=True)
            UP = UP - self._solve(J, b, dofs)
        raise RuntimeError("Newton solver did not converge")
