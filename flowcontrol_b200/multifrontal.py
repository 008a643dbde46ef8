"""Multifrontal block factorisation of the constant LHS and the GPU solve plan.

The reference factorises the BDF LHS once with MUMPS and then only performs
forward/backward substitutions per step (``solver.set_operator(A)`` at
/root/reference/src/flowcontrol/flowsolver.py:697, ``solver.solve`` at :729).
This module is the setup-time half of the replacement: it factorises the same
matrix on the host along the nested-dissection tree of ordering.py and emits a
*solve plan* whose per-step application is two sweeps of dense block-row
products over the whole ensemble (csrc/fcb200.cu, kernel ``fcb_block_rows``).

Block form used (no triangular factors inside a supernode): for the front of
supernode t with fully-summed block F11 (w x w) and subdomain-boundary rows
``struct(t)`` (m):

    E_t = F21 F11^-1      G_t = F11^-1 F12      CB_t = F22 - E_t F12

    forward  (leaves -> root):  y_t = b_t - sum_{d below t} E_d[rows in t, :] y_d
    backward (root -> leaves):  x_t = F11^-1 y_t - G_t x_struct(t)

Both sweeps are written in *pull* form: every output row is produced by exactly
one tile, so there are no atomics and the result is bit-reproducible.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

from .mesh import TaylorHoodTables
from .ordering import TreeNode, dissect, postorder


@dataclass
class Supernode:
    c0: int  # first column (permuted numbering)
    c1: int  # one past last column
    struct: np.ndarray  # permuted indices of the boundary rows (sorted, all >= c1)
    parent: int  # supernode id of the parent (-1 for the root)
    depth: int
    height: int = 0
    E: np.ndarray | None = None  # [m, w]
    Finv: np.ndarray | None = None  # [w, w]
    G: np.ndarray | None = None  # [w, m]


class SymbolicFactor:
    """Ordering + supernode structure shared by every matrix on one mesh/BC set."""

    def __init__(self, tab: TaylorHoodTables, free_mask: np.ndarray, leaf_cells: int = 8):
        """``free_mask[N]`` is True for unknowns kept in the solve (non-Dirichlet)."""
        self.N = tab.N
        nN, nV = tab.nN, tab.nV
        tree = dissect(tab, leaf_cells=leaf_cells)
        po = postorder(tree)

        def dofs_of(nodes: np.ndarray) -> np.ndarray:
            if len(nodes) == 0:
                return np.zeros(0, dtype=np.int64)
            # node-major: ux, uy, (p) of one node are adjacent
            out = []
            ux = nodes
            uy = nodes + nN
            pp = np.where(nodes < nV, nodes + 2 * nN, -1)
            trip = np.stack([ux, uy, pp], axis=1).ravel()
            trip = trip[trip >= 0]
            return trip[free_mask[trip]]

        perm_parts = []
        sn_of_tree = {}
        self.supernodes: list[Supernode] = []
        pos = 0
        for t in po:
            own = dofs_of(tree[t].own_nodes)
            sn_of_tree[t] = len(self.supernodes)
            self.supernodes.append(Supernode(c0=pos, c1=pos + len(own), struct=None, parent=-1, depth=tree[t].depth))
            perm_parts.append(own)
            pos += len(own)
        self.perm = np.concatenate(perm_parts).astype(np.int64)  # permuted -> original dof
        self.n = len(self.perm)
        assert self.n == int(free_mask.sum())
        self.iperm = np.full(self.N, -1, dtype=np.int64)
        self.iperm[self.perm] = np.arange(self.n)
        for t in po:
            s = self.supernodes[sn_of_tree[t]]
            b = dofs_of(np.sort(tree[t].bnd_nodes))
            s.struct = np.sort(self.iperm[b])
            assert s.struct.size == 0 or s.struct[0] >= s.c1
            s.parent = sn_of_tree[tree[t].parent] if tree[t].parent >= 0 else -1
        # heights (leaf = 0)
        for i, s in enumerate(self.supernodes):  # post-order: children come first
            if s.parent >= 0:
                p = self.supernodes[s.parent]
                p.height = max(p.height, s.height + 1)
        self.children: list[list[int]] = [[] for _ in self.supernodes]
        for i, s in enumerate(self.supernodes):
            if s.parent >= 0:
                self.children[s.parent].append(i)
        self.sn_of_col = np.empty(self.n, dtype=np.int64)
        for i, s in enumerate(self.supernodes):
            self.sn_of_col[s.c0 : s.c1] = i

    def factor_entries(self) -> int:
        return int(sum((s.c1 - s.c0) * ((s.c1 - s.c0) + 2 * len(s.struct)) for s in self.supernodes))


class BlockFactor:
    """Numeric factorisation of one matrix on a SymbolicFactor."""

    def __init__(self, sym: SymbolicFactor, A: sp.spmatrix, pivot_tol: float = 1e-13):
        self.sym = sym
        P = sym.perm
        Ap = sp.csr_matrix(A)[P][:, P].tocsr()
        Ap.sort_indices()
        ApT = Ap.T.tocsr()
        ApT.sort_indices()
        self.blocks: list[tuple[np.ndarray, np.ndarray, np.ndarray]] = []
        cb_store: dict[int, np.ndarray] = {}
        sns = sym.supernodes
        self.max_growth = 0.0
        for i, s in enumerate(sns):
            w = s.c1 - s.c0
            m = len(s.struct)
            F = np.zeros((w + m, w + m))
            # original entries: rows c0:c1 (all columns >= c0) and columns c0:c1 (rows >= c1)
            self._load_rows(Ap, s, F, transpose=False)
            self._load_rows(ApT, s, F, transpose=True)
            idx_front = np.concatenate([np.arange(s.c0, s.c1), s.struct])
            for c in sym.children[i]:
                cb = cb_store.pop(c)
                loc = np.searchsorted(idx_front[w:], sns[c].struct) + w
                own = sns[c].struct < s.c1
                loc[own] = sns[c].struct[own] - s.c0
                if not np.array_equal(idx_front[loc], sns[c].struct):
                    raise RuntimeError("child boundary not contained in parent front")
                F[np.ix_(loc, loc)] += cb
            F11 = F[:w, :w]
            if w:
                Finv = np.linalg.inv(F11)
                scale = np.abs(F11).max()
                growth = np.abs(Finv).max() * scale
                self.max_growth = max(self.max_growth, growth)
                if not np.isfinite(growth) or growth > 1.0 / pivot_tol:
                    raise np.linalg.LinAlgError(
                        f"supernode {i}: fully-summed block is numerically singular (growth {growth:.2e})"
                    )
            else:
                Finv = np.zeros((0, 0))
            E = F[w:, :w] @ Finv
            G = Finv @ F[:w, w:]
            if m:
                cb_store[i] = F[w:, w:] - E @ F[:w, w:]
            s_blocks = (np.ascontiguousarray(E), np.ascontiguousarray(Finv), np.ascontiguousarray(G))
            self.blocks.append(s_blocks)
        assert not cb_store or all(v.size == 0 for v in cb_store.values())

    @staticmethod
    def _load_rows(M: sp.csr_matrix, s: Supernode, F: np.ndarray, transpose: bool) -> None:
        w = s.c1 - s.c0
        lo, hi = M.indptr[s.c0], M.indptr[s.c1]
        cols = M.indices[lo:hi]
        vals = M.data[lo:hi]
        rows = np.repeat(np.arange(w), np.diff(M.indptr[s.c0 : s.c1 + 1]))
        inside = (cols >= s.c0) & (cols < s.c1)
        above = cols >= s.c1
        if not transpose:
            F[rows[inside], cols[inside] - s.c0] = vals[inside]
        if above.any():
            loc = np.searchsorted(s.struct, cols[above])
            if np.any(loc >= len(s.struct)) or not np.array_equal(s.struct[loc], cols[above]):
                raise RuntimeError("matrix entry outside the symbolic structure")
            if transpose:
                F[w + loc, rows[above]] = vals[above]
            else:
                F[rows[above], w + loc] = vals[above]

    # host reference of the two sweeps (used by tests and by the setup-time Newton)
    def solve(self, b_free: np.ndarray) -> np.ndarray:
        """Solve A x = b for the free unknowns (original numbering of the free set:
        ``b_free`` and the result are indexed like ``sym.perm``-inverse, i.e. full-N vectors
        restricted by the caller).  ``b_free`` has shape [n] or [n, k] in PERMUTED order."""
        y = np.array(b_free, dtype=np.float64, copy=True)
        sns = self.sym.supernodes
        for s, (E, Finv, G) in zip(sns, self.blocks):
            if len(s.struct):
                y[s.struct] -= E @ y[s.c0 : s.c1]
        x = np.empty_like(y)
        for s, (E, Finv, G) in zip(reversed(sns), reversed(self.blocks)):
            x[s.c0 : s.c1] = Finv @ y[s.c0 : s.c1] - (G @ x[s.struct] if len(s.struct) else 0.0)
        return x


@dataclass
class SolvePlan:
    """Flat arrays consumed by the CUDA kernel ``fcb_block_rows``.

    The kernel works on one buffer Z of 2n rows x ldb columns: rows [0,n) hold y
    (forward sweep, initially the permuted RHS), rows [n,2n) hold x.  A *tile*
    produces ``nrows`` consecutive output rows from K gathered input rows:

        Z[out_row + r, :] = (self ? Z[self_row + r, :] : 0) + sum_k vals[vptr + k*RT + r] * Z[cols[kptr + k], :]

    Tiles are grouped into launches; tiles of one launch are independent."""

    n: int
    RT: int
    tile_out: np.ndarray  # int32 [ntiles] first output row in Z
    tile_self: np.ndarray  # int32 [ntiles] row to add (or -1)
    tile_nrows: np.ndarray  # int32 [ntiles]
    tile_kptr: np.ndarray  # int64 [ntiles+1] offsets into cols
    tile_vptr: np.ndarray  # int64 [ntiles] offsets into vals
    cols: np.ndarray  # int32 [sum K]
    vals: np.ndarray  # float64 [sum K * RT]
    launch_ptr: np.ndarray  # int32 [nlaunch+1] tile ranges, forward launches first then backward
    n_forward_launches: int

    @property
    def nnz_padded(self) -> int:
        return int(self.vals.size)


def build_plan(fac: BlockFactor, RT: int = 8) -> SolvePlan:
    sym = fac.sym
    sns = sym.supernodes
    n = sym.n
    nS = len(sns)
    # ---- forward pull structure: for each target tile, the source supernodes touching it
    # tile id of a permuted row: (supernode, (row - c0) // RT) -> flattened
    tile_base = np.zeros(nS + 1, dtype=np.int64)
    for i, s in enumerate(sns):
        tile_base[i + 1] = tile_base[i] + -(-(s.c1 - s.c0) // RT)
    ntile_rows = int(tile_base[-1])
    row_tile = np.empty(n, dtype=np.int64)
    for i, s in enumerate(sns):
        row_tile[s.c0 : s.c1] = tile_base[i] + (np.arange(s.c1 - s.c0) // RT)
    fwd_sources: list[list[tuple[int, np.ndarray, np.ndarray]]] = [[] for _ in range(ntile_rows)]
    for d, s in enumerate(sns):
        if len(s.struct) == 0 or s.c1 == s.c0:
            continue
        tl = row_tile[s.struct]
        # struct is sorted and tiles are monotone in row -> contiguous groups
        cut = np.flatnonzero(np.diff(tl)) + 1
        starts = np.concatenate([[0], cut])
        stops = np.concatenate([cut, [len(tl)]])
        for a, b in zip(starts, stops):
            fwd_sources[tl[a]].append((d, np.arange(a, b), s.struct[a:b]))
    tile_out, tile_self, tile_nrows, kptr, vptr = [], [], [], [0], []
    cols_parts, vals_parts = [], []
    launch_ptr = [0]
    vpos = 0

    def emit(out_row, self_row, nrows, cols, vals_krt):
        nonlocal vpos
        tile_out.append(out_row)
        tile_self.append(self_row)
        tile_nrows.append(nrows)
        cols_parts.append(cols.astype(np.int32))
        kptr.append(kptr[-1] + len(cols))
        vptr.append(vpos)
        vals_parts.append(vals_krt.ravel())
        vpos += vals_krt.size

    max_h = max(s.height for s in sns)
    # forward launches: height 1..max_h (height-0 supernodes have nothing below them)
    for h in range(1, max_h + 1):
        for i, s in enumerate(sns):
            if s.height != h:
                continue
            w = s.c1 - s.c0
            for tix in range(-(-w // RT)):
                src = fwd_sources[tile_base[i] + tix]
                if not src:
                    continue
                r0 = s.c0 + tix * RT
                nrows = min(RT, s.c1 - r0)
                K = sum(sns[d].c1 - sns[d].c0 for d, _, _ in src)
                vals = np.zeros((K, RT))
                cols = np.empty(K, dtype=np.int64)
                k0 = 0
                for d, loc_in_struct, rows in src:
                    wd = sns[d].c1 - sns[d].c0
                    cols[k0 : k0 + wd] = np.arange(sns[d].c0, sns[d].c1)
                    E = fac.blocks[d][0]
                    vals[k0 : k0 + wd, rows - r0] = -E[loc_in_struct, :].T
                    k0 += wd
                emit(r0, r0, nrows, cols, vals)
        if len(tile_out) > launch_ptr[-1]:
            launch_ptr.append(len(tile_out))
    n_fwd = len(launch_ptr) - 1
    # backward launches: depth 0..max
    max_d = max(s.depth for s in sns)
    for dpt in range(max_d + 1):
        for i, s in enumerate(sns):
            if s.depth != dpt:
                continue
            w = s.c1 - s.c0
            if w == 0:
                continue
            E, Finv, G = fac.blocks[i]
            cols = np.concatenate([np.arange(s.c0, s.c1), n + s.struct])
            full = np.concatenate([Finv, -G], axis=1)  # [w, w+m]
            for tix in range(-(-w // RT)):
                r0 = tix * RT
                nrows = min(RT, w - r0)
                vals = np.zeros((w + len(s.struct), RT))
                vals[:, :nrows] = full[r0 : r0 + nrows, :].T
                emit(n + s.c0 + r0, -1, nrows, cols, vals)
        if len(tile_out) > launch_ptr[-1]:
            launch_ptr.append(len(tile_out))
    return SolvePlan(
        n=n,
        RT=RT,
        tile_out=np.array(tile_out, dtype=np.int32),
        tile_self=np.array(tile_self, dtype=np.int32),
        tile_nrows=np.array(tile_nrows, dtype=np.int32),
        tile_kptr=np.array(kptr, dtype=np.int64),
        tile_vptr=np.array(vptr, dtype=np.int64),
        cols=np.concatenate(cols_parts) if cols_parts else np.zeros(0, np.int32),
        vals=np.concatenate(vals_parts) if vals_parts else np.zeros(0),
        launch_ptr=np.array(launch_ptr, dtype=np.int32),
        n_forward_launches=n_fwd,
    )


def apply_plan_host(plan: SolvePlan, b_perm: np.ndarray) -> np.ndarray:
    """Numpy emulation of the CUDA sweeps (tests only; O(ntiles) python loop)."""
    b = np.asarray(b_perm, dtype=np.float64)
    squeeze = b.ndim == 1
    if squeeze:
        b = b[:, None]
    n, RT = plan.n, plan.RT
    Z = np.zeros((2 * n, b.shape[1]))
    Z[:n] = b
    for t in range(len(plan.tile_out)):
        k0, k1 = plan.tile_kptr[t], plan.tile_kptr[t + 1]
        K = k1 - k0
        V = plan.vals[plan.tile_vptr[t] : plan.tile_vptr[t] + K * RT].reshape(K, RT)
        nr = plan.tile_nrows[t]
        acc = V[:, :nr].T @ Z[plan.cols[k0:k1]]
        if plan.tile_self[t] >= 0:
            acc += Z[plan.tile_self[t] : plan.tile_self[t] + nr]
        Z[plan.tile_out[t] : plan.tile_out[t] + nr] = acc
    x = Z[n:]
    return x[:, 0] if squeeze else x
