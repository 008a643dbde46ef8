"""Numeric multifrontal factorisation on the GPU (setup / steady state).

The reference hands every new matrix to MUMPS (``solver.set_operator`` in _prepare_systems, flowsolver.py:694-697; the
Newton / Picard iterates of steadystate.py:95, 139-147 through ``dolfin.solve``), which factorises it numerically on the
host.  ``multifrontal.BlockFactor`` is this build's host factorisation; ``DeviceBlockFactor`` produces the SAME blocks
``E = F21 F11^-1``, ``F11^-1``, ``G = F11^-1 F12`` per front with the dense work on the GPU (``fcb_factorize``, kernels
``k_ff_assemble`` / ``k_ff_invert`` / ``k_ff_gemm`` in csrc/fcb200.cu): fronts are assembled from the CSR values and their
children's Schur complements through index maps computed once per symbolic structure, level by level of the elimination
tree; ``F11`` is inverted in place by Gauss-Jordan elimination with partial pivoting (one CTA per front), the three
products are tiled FP64 GEMMs batched over the fronts of a level.  A new matrix with the same sparsity (the next Newton
iterate, another Reynolds number of a continuation) only re-uploads its values.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import libfcb
from .multifrontal import SymbolicFactor


class FrontMaps:
    """Index maps of the numeric factorisation for one (symbolic structure, sparsity pattern)."""

    def __init__(self, sym: SymbolicFactor, pattern: sp.csr_matrix):
        """``pattern``: the matrix in PERMUTED numbering (CSR, sorted indices); only its structure is used."""
        sns = sym.supernodes
        nS, n = len(sns), sym.n
        self.w = np.array([s.c1 - s.c0 for s in sns], dtype=np.int32)
        self.m = np.array([len(s.struct) for s in sns], dtype=np.int32)
        c0 = np.array([s.c0 for s in sns], dtype=np.int64)
        c1 = np.array([s.c1 for s in sns], dtype=np.int64)
        size = (self.w + self.m).astype(np.int64)
        soff = np.concatenate([[0], np.cumsum(self.m)]).astype(np.int64)
        struct_all = np.concatenate([s.struct for s in sns]).astype(np.int64) if nS else np.zeros(0, np.int64)
        skey = np.repeat(np.arange(nS, dtype=np.int64), self.m) * n + struct_all  # increasing: fronts ascending, rows ascending

        def local(front: np.ndarray, row: np.ndarray) -> np.ndarray:
            """Position of solver row ``row`` in the index set [own | struct] of ``front`` (vectorised)."""
            own = row < c1[front]
            pos = np.searchsorted(skey, front * n + row)
            bad = ~own & ((pos >= len(skey)) | (skey[np.minimum(pos, len(skey) - 1)] != front * n + row))
            if bad.any():
                raise RuntimeError("matrix entry outside the symbolic structure")
            return np.where(own, row - c0[front], self.w[front] + pos - soff[front])

        # matrix entries: (i, j) belongs to the front that eliminates min(i, j)
        coo_i = np.repeat(np.arange(n, dtype=np.int64), np.diff(pattern.indptr))
        coo_j = pattern.indices.astype(np.int64)
        f = sym.sn_of_col[np.minimum(coo_i, coo_j)]
        dst = local(f, coo_i) * size[f] + local(f, coo_j)
        order = np.argsort(f, kind="stable")
        self.a_src = order.astype(np.int32)
        self.a_dst = dst[order].astype(np.int32)
        self.a_ptr = np.concatenate([[0], np.cumsum(np.bincount(f, minlength=nS))]).astype(np.int64)
        # children and the position of their boundary rows in the parent's front
        self.c_ptr = np.concatenate([[0], np.cumsum([len(ch) for ch in sym.children])]).astype(np.int64)
        self.c_front = np.array([c for ch in sym.children for c in ch], dtype=np.int32)
        par = np.repeat(np.arange(nS, dtype=np.int64), np.diff(self.c_ptr))
        self.c_lptr = np.concatenate([[0], np.cumsum(self.m[self.c_front])]).astype(np.int64)
        rows = np.concatenate([sns[c].struct for c in self.c_front]).astype(np.int64) if len(self.c_front) else np.zeros(0, np.int64)
        self.c_loc = local(np.repeat(par, self.m[self.c_front]), rows).astype(np.int32) if len(rows) else np.zeros(0, np.int32)
        # levels: fronts by height, children before parents
        h = np.array([s.height for s in sns], dtype=np.int64)
        self.level_fronts = np.argsort(h, kind="stable").astype(np.int32)
        self.level_ptr = np.concatenate([[0], np.cumsum(np.bincount(h))]).astype(np.int32)
        self.nnz = int(pattern.nnz)
        self.eoff = np.concatenate([[0], np.cumsum(self.m.astype(np.int64) * self.w)])
        self.ioff = np.concatenate([[0], np.cumsum(self.w.astype(np.int64) * self.w)])
        self.indptr, self.indices = pattern.indptr.copy(), pattern.indices.copy()

    def emulate(self, avals: np.ndarray):
        """Numpy restatement of the device algorithm from the flat maps (tests only)."""
        nS = len(self.w)
        F = [None] * nS
        out = []
        for lv in range(len(self.level_ptr) - 1):
            for f in self.level_fronts[self.level_ptr[lv] : self.level_ptr[lv + 1]]:
                w, m = int(self.w[f]), int(self.m[f])
                s = w + m
                Ff = np.zeros(s * s)
                sl = slice(self.a_ptr[f], self.a_ptr[f + 1])
                Ff[self.a_dst[sl]] = avals[self.a_src[sl]]
                Ff = Ff.reshape(s, s)
                for k in range(self.c_ptr[f], self.c_ptr[f + 1]):
                    c = int(self.c_front[k])
                    loc = self.c_loc[self.c_lptr[k] : self.c_lptr[k + 1]]
                    Ff[np.ix_(loc, loc)] += F[c][self.w[c] :, self.w[c] :]
                    F[c] = None
                Finv = np.linalg.inv(Ff[:w, :w]) if w else np.zeros((0, 0))
                E = Ff[w:, :w] @ Finv
                G = Finv @ Ff[:w, w:]
                Ff[w:, w:] -= E @ Ff[:w, w:]
                F[f] = Ff
                out.append((f, E, Finv, G))
        blocks = [None] * nS
        for f, E, Finv, G in out:
            blocks[f] = (E, Finv, G)
        return blocks


class DeviceBlockFactor:
    """Drop-in for ``multifrontal.BlockFactor`` with the numeric work on the GPU."""

    def __init__(self, sym: SymbolicFactor, A: sp.spmatrix, maps: FrontMaps | None = None, device: int = 0, pivot_tol: float = 1e-13):
        self.sym = sym
        P = sym.perm
        Ap = sp.csr_matrix(A)[P][:, P].tocsr()
        Ap.sort_indices()
        if maps is None or maps.nnz != Ap.nnz or not np.array_equal(maps.indptr, Ap.indptr) or not np.array_equal(maps.indices, Ap.indices):
            maps = FrontMaps(sym, Ap)
        self.maps = maps
        lib = libfcb.load()
        keep = [np.ascontiguousarray(a) for a in (maps.w, maps.m, maps.level_ptr, maps.level_fronts, maps.a_ptr, maps.a_src, maps.a_dst,
                                                   maps.c_ptr, maps.c_front, maps.c_lptr, maps.c_loc)]
        s = libfcb.fcb_symbolic()
        s.nfront, s.nlevel = len(maps.w), len(maps.level_ptr) - 1
        for name, a in zip(("w", "m", "level_ptr", "level_fronts", "a_ptr", "a_src", "a_dst", "c_ptr", "c_front", "c_lptr", "c_loc"), keep):
            typ = libfcb.c_i64p if a.dtype == np.int64 else libfcb.c_i32p
            setattr(s, name, a.ctypes.data_as(typ))
        E = np.empty(int(maps.eoff[-1]))
        Finv = np.empty(int(maps.ioff[-1]))
        G = np.empty(int(maps.eoff[-1]))
        growth = np.empty(len(maps.w))
        avals = np.ascontiguousarray(Ap.data, dtype=np.float64)
        rc = lib.fcb_factorize(C.byref(s), int(device), libfcb.as_voidp(avals), int(Ap.nnz), libfcb.as_voidp(E), libfcb.as_voidp(Finv),
                               libfcb.as_voidp(G), libfcb.as_voidp(growth))
        if rc != 0:
            raise libfcb.FcbError(f"fcb_factorize failed ({rc}): {lib.fcb_last_error(None).decode()}")
        self.max_growth = float(growth.max()) if len(growth) else 0.0
        if not np.isfinite(self.max_growth) or self.max_growth > 1.0 / pivot_tol:
            bad = int(np.argmax(~np.isfinite(growth) | (growth > 1.0 / pivot_tol)))
            raise np.linalg.LinAlgError(f"supernode {bad}: fully-summed block is numerically singular (growth {growth[bad]:.2e})")
        self.blocks = []
        for f in range(len(maps.w)):
            w, m = int(maps.w[f]), int(maps.m[f])
            self.blocks.append((E[maps.eoff[f] : maps.eoff[f + 1]].reshape(m, w), Finv[maps.ioff[f] : maps.ioff[f + 1]].reshape(w, w),
                                G[maps.eoff[f] : maps.eoff[f + 1]].reshape(w, m)))

    def solve(self, b_free: np.ndarray) -> np.ndarray:
        """Host sweeps with the device-computed blocks (same as BlockFactor.solve)."""
        from .multifrontal import BlockFactor

        return BlockFactor.solve(self, b_free)


class DoubledSymbolic:
    """Symbolic structure of the real 2n x 2n form [[-A, -wQ], [wQ, -A]] of a complex-shifted system (utils/linalg.py:215)
    on the elimination tree of ``sym``: every unknown k of ``sym`` becomes the pair (real part, imaginary part), adjacent in
    the elimination order, so every front simply doubles.  Original numbering of the doubled system: [real (N) | imaginary (N)]."""

    def __init__(self, sym: SymbolicFactor):
        from .multifrontal import Supernode

        self.N, self.n = 2 * sym.N, 2 * sym.n
        self.perm = np.stack([sym.perm, sym.perm + sym.N], axis=1).ravel()
        self.supernodes = [Supernode(c0=2 * s.c0, c1=2 * s.c1, struct=(2 * s.struct[:, None] + np.arange(2)).ravel(), parent=s.parent,
                                     depth=s.depth, height=s.height) for s in sym.supernodes]
        self.children = sym.children
        self.sn_of_col = np.repeat(sym.sn_of_col, 2)

    def factor_entries(self) -> int:
        return int(sum((s.c1 - s.c0) * ((s.c1 - s.c0) + 2 * len(s.struct)) for s in self.supernodes))


def frequency_response_device(A, B, C, Q, ww, tab, device: int = 0, leaf_cells: int = 16, verbose: bool = False):
    """H(w) = C (jwQ - A)^-1 B for every w in ``ww`` (utils/linalg.py:192-240) with one multifrontal factorisation per
    frequency on the GPU: the real 2n x 2n block form of the reference is ordered by the nested dissection of the mesh with
    the real and imaginary part of every unknown adjacent (DoubledSymbolic), factorised by ``fcb_factorize`` (the index maps
    are built once: every frequency has the same sparsity), and the right-hand sides [B; 0] go through the two sweeps.
    ``A``, ``Q``: n x n over ALL dofs of the mixed space (Dirichlet rows as identity rows, as OperatorGetter returns them).
    Returns (H [ny, nu, nw] complex, ww)."""
    import logging

    log = logging.getLogger(__name__)
    A, Q = sp.csr_matrix(A), sp.csr_matrix(Q)
    n = A.shape[0]
    if n != tab.N:
        raise ValueError(f"operators must be over the {tab.N} dofs of the mixed space, got {n}")
    B = np.asarray(B, dtype=np.float64).reshape(n, -1)
    Cm = np.asarray(C, dtype=np.float64).reshape(-1, n)
    ww = np.asarray(ww, dtype=np.float64)
    # one sparsity for every frequency: A and Q on the union of their patterns (explicit zeros kept)
    pat = (abs(A) + abs(Q)).tocsr()
    pat.sort_indices()
    ones = sp.csr_matrix((np.ones(pat.nnz), pat.indices, pat.indptr), shape=pat.shape)

    def on_pattern(M):
        out = (M + 0.0 * ones).tocsr()
        out.sort_indices()
        if not (np.array_equal(out.indptr, pat.indptr) and np.array_equal(out.indices, pat.indices)):  # scipy pruned a zero
            full = sp.csr_matrix((np.zeros(pat.nnz), pat.indices, pat.indptr), shape=pat.shape)
            coo = M.tocoo()
            key = np.repeat(np.arange(n, dtype=np.int64), np.diff(pat.indptr)) * n + pat.indices
            full.data[np.searchsorted(key, coo.row.astype(np.int64) * n + coo.col)] = coo.data
            out = full
        return out

    Au, Qu = on_pattern(A), on_pattern(Q)
    sym2 = DoubledSymbolic(SymbolicFactor(tab, np.ones(tab.N, dtype=bool), leaf_cells=leaf_cells))
    rhs = np.vstack([B, np.zeros_like(B)])[sym2.perm]
    H = np.zeros((Cm.shape[0], B.shape[1], len(ww)), dtype=complex)
    maps = None
    for ii, w in enumerate(ww):
        wq = sp.csr_matrix((w * Qu.data, Qu.indices, Qu.indptr), shape=Qu.shape)  # explicit zeros stay when w = 0
        na = sp.csr_matrix((-Au.data, Au.indices, Au.indptr), shape=Au.shape)
        blk = sp.bmat([[na, -wq], [wq, na]], format="csr")
        fac = DeviceBlockFactor(sym2, blk, maps=maps, device=device)
        maps = fac.maps
        x = np.empty_like(rhs)
        x[sym2.perm] = fac.solve(rhs)
        H[:, :, ii] = Cm @ x[:n] + 1j * (Cm @ x[n:])
        if verbose:
            log.info("  [%d/%d] w=%.4e | max|H|=%.4e", ii + 1, len(ww), w, np.max(np.abs(H[:, :, ii])))
    return H, ww
