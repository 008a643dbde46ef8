"""Device assembly of the U-dependent blocks of the linearised Navier-Stokes operator (setup / steady state / operators).

The reference re-assembles the Jacobian of its steady form in every Newton iteration and the Picard operator in every
fixed-point iteration through dolfin (/root/reference/src/flowcontrol/steadystate.py:95, 139-147 with the forms of
nsforms.py:137-183), and the linearised operator once more in ``OperatorGetter.get_A`` (operatorgetter.py:25-83).  The
only blocks that depend on the velocity field are the advection block ``C(U)`` and the base-gradient blocks ``D^ij(U)``
(SURVEY.md Appendix A); ``fcb_assemble_advection`` (include/fcb200.h, kernel ``k_assemble_advection``) computes their
element matrices on the GPU and scatters them into CSR value arrays through a precomputed position map, one launch per
element colour (atomic-free, bit-reproducible), for a whole ensemble of velocity fields at once.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import libfcb
from .fem import ScalarBlocks
from .mesh import TaylorHoodTables


class DeviceAdvectionAssembler:
    def __init__(self, tab: TaylorHoodTables, blocks: ScalarBlocks, device: int = 0):
        self.tab, self.blocks, self.device = tab, blocks, int(device)
        self.lib = libfcb.load()
        pat = sp.csr_matrix(blocks.M)
        pat.sort_indices()
        self.indptr, self.indices = pat.indptr.astype(np.int32), pat.indices.astype(np.int32)
        nN = tab.nN
        cn = np.asarray(tab.cell_nodes, dtype=np.int64)
        rows = np.repeat(cn, 6, axis=1)  # [nT, 36]: entry a*6+b -> row node a
        cols = np.tile(cn, (1, 6))       #                          column node b
        key = np.repeat(np.arange(nN, dtype=np.int64), np.diff(pat.indptr)) * nN + pat.indices
        want = (rows * nN + cols).ravel()
        loc = np.searchsorted(key, want)
        if np.any(loc >= len(key)) or not np.array_equal(key[np.minimum(loc, len(key) - 1)], want):
            raise ValueError("the scalar P2 pattern does not contain every element-matrix entry")
        self.pos = loc.astype(np.int32)  # bit-exact integer map: checked against the host assembly in the tests
        cptr, ccells = tab.element_colouring()
        self._keep = [np.ascontiguousarray(cn, dtype=np.int32), np.ascontiguousarray(tab.Jinv.reshape(-1, 4), dtype=np.float64),
                      np.ascontiguousarray(tab.detJ, dtype=np.float64), np.ascontiguousarray(cptr, dtype=np.int32),
                      np.ascontiguousarray(ccells, dtype=np.int32), self.pos]
        m = libfcb.fcb_assembly()
        m.nT, m.nN, m.ncolour, m.nnz = tab.nT, nN, len(cptr) - 1, len(self.indices)
        m.cell_nodes = self._keep[0].ctypes.data_as(libfcb.c_i32p)
        m.Jinv = self._keep[1].ctypes.data_as(libfcb.c_f64p)
        m.detJ = self._keep[2].ctypes.data_as(libfcb.c_f64p)
        m.colour_ptr = self._keep[3].ctypes.data_as(libfcb.c_i32p)
        m.colour_cells = self._keep[4].ctypes.data_as(libfcb.c_i32p)
        m.pos = self._keep[5].ctypes.data_as(libfcb.c_i32p)
        self.struct = m

    def values(self, U: np.ndarray, with_D: bool = True):
        """U [Nv] or [Nv, B] -> (C values [nnz, B], D values [4, nnz, B] or None) on the scalar P2 pattern."""
        U = np.asarray(U, dtype=np.float64)
        U2 = np.ascontiguousarray(U[:, None] if U.ndim == 1 else U)
        if U2.shape[0] != self.tab.Nv:
            raise ValueError(f"expected {self.tab.Nv} velocity dofs, got {U2.shape[0]}")
        B, nnz = U2.shape[1], len(self.indices)
        Cv = np.empty((nnz, B))
        Dv = np.empty((4, nnz, B)) if with_D else None
        rc = self.lib.fcb_assemble_advection(C.byref(self.struct), B, self.device, libfcb.as_voidp(U2), libfcb.as_voidp(Cv), libfcb.as_voidp(Dv))
        if rc != 0:
            raise libfcb.FcbError(f"fcb_assemble_advection failed ({rc}): {self.lib.fcb_last_error(None).decode()}")
        return Cv, Dv

    def _csr(self, v: np.ndarray) -> sp.csr_matrix:
        n = self.tab.nN
        return sp.csr_matrix((v, self.indices, self.indptr), shape=(n, n))

    def advection(self, U: np.ndarray):
        """Drop-in for ``ScalarBlocks.advection(U)``: (C, {(i, j): D^ij}) as scipy CSR matrices (one velocity field)."""
        Cv, Dv = self.values(U)
        D = {(i, j): self._csr(Dv[2 * i + j, :, 0]) for i in range(2) for j in range(2)}
        return self._csr(Cv[:, 0]), D

    def saddle_point(self, c_mass: float, Re: float, U: np.ndarray, shift: float = 0.0, linearised: bool = True) -> sp.csr_matrix:
        """``ScalarBlocks.saddle_point`` with the U-dependent blocks assembled on the device."""
        bl = self.blocks
        Cv, Dv = self.values(U, with_D=linearised)
        F = (c_mass - shift) * bl.M + bl.K / Re + self._csr(Cv[:, 0])
        if linearised:
            D = [self._csr(Dv[k, :, 0]) for k in range(4)]
            blk = [[F + D[0], D[1], -bl.Bx.T], [D[2], F + D[3], -bl.By.T], [-bl.Bx, -bl.By, None]]
        else:
            blk = [[F, None, -bl.Bx.T], [None, F, -bl.By.T], [-bl.Bx, -bl.By, None]]
        A = sp.bmat(blk, format="csr")
        A.sum_duplicates()
        A.sort_indices()
        return A
