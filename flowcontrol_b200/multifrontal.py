"""Multifrontal block factorisation of the constant LHS and the GPU solve plan.

The reference factorises the BDF LHS once with MUMPS and then only performs
forward/backward substitutions per step (``solver.set_operator(A)`` at
/root/reference/src/flowcontrol/flowsolver.py:697, ``solver.solve`` at :729).
This module is the setup-time half of the replacement: it factorises the same
matrix on the host along the nested-dissection tree of ordering.py and emits a
*solve plan* whose per-step application is two sweeps of dense block-row
FP64 tensor-core tile products over the whole ensemble (csrc/fcb200.cu, kernel ``k_front_sweep``).

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
    """Flat block arrays handed to the CUDA library (include/fcb200.h: fcb_plan); the library cuts every
    block into tensor-core tiles for the actual ensemble width and SM count (csrc/fcb200.cu: upload_plan).

    The device works on one buffer Z with rows
        [0, n)          b on entry, x on exit            (solver row order)
        [n, 2n)         y (forward-eliminated RHS)
        [2n, 2n + nU)   update vectors u_t of every supernode (its subdomain-boundary rows)
        2n + nU         a row of zeros (``zrow``): the source of absent gathers
    A *block* is a dense product  acc[M x traj] = V[M x K] . x[K x traj]  whose input rows are gathered as
    x_k = Z[i0[k]] (+ Z[i1[k]] + Z[i2[k]] when nsrc == 3) and whose M output rows go to consecutive rows
    Z[out0 + r] = acc[r] (+ Z[e0[r]] + Z[e1[r]], -1 = absent).  ``ystore >= 0`` additionally stores the
    gathered x_k to Z[ystore + k] (the block that owns y_t; M may be 0 for a store-only block).

        forward  (one block per supernode t, launches by tree height):
            x_k = b_t[k] + sum_children u_c[..]            (= y_t, stored)
            u_t[r] = sum_children u_c[..] - (E_t y_t)[r]
        backward (launches by tree depth):
            x_t[r] = (F11^-1 y_t)[r] - (G_t x_struct(t))[r]

    The top ``top_levels`` levels of the tree (a handful of large separators, whose level-by-level sweeps have too
    little parallelism for a GPU) are merged: their right-hand side r_T = b_T + (updates of the supernodes just below)
    is assembled by a gather-sum, and x_T = S^-1 r_T is ONE launch of dense blocks with the explicit inverse of the
    top Schur complement (the "partitioned inverse" of the top separators).

    Every output row is produced by exactly one block: no atomics, bit-reproducible."""

    n: int
    nU: int
    blk_K: np.ndarray  # int32 [nblk] gathered rows
    blk_M: np.ndarray  # int32 output rows (0 = store-only)
    blk_nsrc: np.ndarray  # int32 1 or 3
    blk_out0: np.ndarray  # int32
    blk_ystore: np.ndarray  # int32 (-1 = none)
    blk_iptr: np.ndarray  # int64 offsets into i0/i1/i2 (K entries per block)
    blk_vptr: np.ndarray  # int64 offsets into vals (row-major [M, K] per block)
    blk_eptr: np.ndarray  # int64 offsets into e0/e1 (M entries per block; -1 = no seed gather)
    i0: np.ndarray
    i1: np.ndarray
    i2: np.ndarray
    e0: np.ndarray
    e1: np.ndarray
    vals: np.ndarray
    launch_ptr: np.ndarray  # int32 [nlaunch+1] block ranges, forward launches first
    n_forward_launches: int
    # gather-sum executed between the forward and the remaining launches: Z[asm_dst[i]] = sum_j Z[asm_src[j]],
    # j in [asm_ptr[i], asm_ptr[i+1]) -- assembles the right-hand side of the merged top of the tree
    asm_ptr: np.ndarray = None
    asm_src: np.ndarray = None
    asm_dst: np.ndarray = None

    @property
    def nnz(self) -> int:
        return int(self.vals.size)

    @property
    def zrow(self) -> int:
        return 2 * self.n + self.nU

    @property
    def z_rows(self) -> int:
        return 2 * self.n + self.nU + 1


def top_inverse(fac: BlockFactor, top: list[int]) -> tuple[np.ndarray, np.ndarray]:
    """Explicit inverse of the Schur complement of the supernodes ``top`` (closed under taking ancestors):
    returns (solver rows of the top unknowns, S^-1) with x_T = S^-1 r_T."""
    sns = fac.sym.supernodes
    trows = np.concatenate([np.arange(sns[t].c0, sns[t].c1) for t in top]) if top else np.zeros(0, dtype=np.int64)
    pos = {int(r): j for j, r in enumerate(trows)}
    nT = len(trows)
    Y = np.eye(nT)
    own = {t: np.array([pos[r] for r in range(sns[t].c0, sns[t].c1)], dtype=np.int64) for t in top}
    st = {t: np.array([pos[int(r)] for r in sns[t].struct], dtype=np.int64) for t in top}
    for t in top:  # post-order: children first
        E = fac.blocks[t][0]
        if len(st[t]):
            Y[st[t]] -= E @ Y[own[t]]
    X = np.zeros((nT, nT))
    for t in reversed(top):
        _, Finv, G = fac.blocks[t]
        X[own[t]] = Finv @ Y[own[t]] - (G @ X[st[t]] if len(st[t]) else 0.0)
    return trows, X


def build_plan(fac: BlockFactor, top_levels: int = 2) -> SolvePlan:
    sym = fac.sym
    sns = sym.supernodes
    n = sym.n
    nS = len(sns)
    uoff = np.zeros(nS + 1, dtype=np.int64)
    for i, s in enumerate(sns):
        uoff[i + 1] = uoff[i] + len(s.struct)
    nU = int(uoff[-1])
    UB = 2 * n  # first row of the U region
    ZROW = 2 * n + nU
    top = [i for i, s in enumerate(sns) if s.depth < top_levels]  # post-order, closed under ancestors
    in_top = np.zeros(nS, dtype=bool)
    in_top[top] = True

    def child_sources(i: int, rows: np.ndarray, absent: int) -> tuple[np.ndarray, np.ndarray]:
        """Z rows of the (at most two) children's update vectors that hit the given solver rows."""
        srcs = [np.full(len(rows), absent, dtype=np.int64), np.full(len(rows), absent, dtype=np.int64)]
        ch = sym.children[i]
        if len(ch) > 2:
            raise NotImplementedError("solve plan assumes a binary dissection tree")
        for slot, c in enumerate(ch):
            st = sns[c].struct
            if len(st) == 0 or len(rows) == 0:
                continue
            pos = np.searchsorted(st, rows)
            pos_c = np.minimum(pos, len(st) - 1)
            hit = st[pos_c] == rows
            srcs[slot][hit] = UB + uoff[c] + pos_c[hit]
        return srcs[0], srcs[1]

    blocks: list[dict] = []
    launch_ptr = [0]
    max_h = max(s.height for s in sns)
    for h in range(max_h + 1):
        for i in (i for i, s in enumerate(sns) if s.height == h and not in_top[i]):
            s = sns[i]
            w, m = s.c1 - s.c0, len(s.struct)
            if w == 0 and m == 0:
                continue
            own = np.arange(s.c0, s.c1, dtype=np.int64)
            a1, a2 = child_sources(i, own, ZROW)
            has_children = bool(sym.children[i])
            e0, e1 = child_sources(i, s.struct, -1)
            blocks.append(dict(K=w, M=m, nsrc=3 if has_children else 1, out0=UB + int(uoff[i]), ystore=n + s.c0,
                               i0=own, i1=a1, i2=a2, vals=-fac.blocks[i][0], e0=e0 if has_children else None,
                               e1=e1 if has_children else None))
        if len(blocks) > launch_ptr[-1]:
            launch_ptr.append(len(blocks))
    n_fwd = len(launch_ptr) - 1
    # merged top: assemble r_T into the y rows of the top unknowns, then x_T = S^-1 r_T
    asm_ptr, asm_src, asm_dst = [0], [], []
    if top:
        trows, Sinv = top_inverse(fac, top)
        frontier = [c for t in top for c in sym.children[t] if not in_top[c]]
        for t in top:
            s = sns[t]
            rows = np.arange(s.c0, s.c1, dtype=np.int64)
            srcs = [rows.copy()]  # b
            for c in frontier:
                st = sns[c].struct
                if len(st) == 0:
                    continue
                pos = np.searchsorted(st, rows)
                pos_c = np.minimum(pos, len(st) - 1)
                srcs.append(np.where(st[pos_c] == rows, UB + uoff[c] + pos_c, -1))
            S = np.stack(srcs, axis=1)
            for k, r in enumerate(rows):
                valid = S[k][S[k] >= 0]
                asm_src.extend(valid.tolist())
                asm_ptr.append(len(asm_src))
                asm_dst.append(n + int(r))
        pos = 0
        for t in top:
            s = sns[t]
            w = s.c1 - s.c0
            if w:
                blocks.append(dict(K=len(trows), M=w, nsrc=1, out0=s.c0, ystore=-1, i0=n + trows, i1=None, i2=None,
                                   vals=Sinv[pos : pos + w], e0=None, e1=None))
            pos += w
        if len(blocks) > launch_ptr[-1]:
            launch_ptr.append(len(blocks))
    max_d = max(s.depth for s in sns)
    for dpt in range(max_d + 1):
        for i in (i for i, s in enumerate(sns) if s.depth == dpt and not in_top[i]):
            s = sns[i]
            w = s.c1 - s.c0
            if w == 0:
                continue
            _, Finv, G = fac.blocks[i]
            full = np.concatenate([Finv, -G], axis=1)  # [w, w+m]
            idx = np.concatenate([n + np.arange(s.c0, s.c1, dtype=np.int64), s.struct.astype(np.int64)])
            blocks.append(dict(K=full.shape[1], M=w, nsrc=1, out0=s.c0, ystore=-1, i0=idx, i1=None, i2=None, vals=full,
                               e0=None, e1=None))
        if len(blocks) > launch_ptr[-1]:
            launch_ptr.append(len(blocks))
    nb = len(blocks)
    K = np.array([b["K"] for b in blocks], dtype=np.int32)
    M = np.array([b["M"] for b in blocks], dtype=np.int32)
    iptr = np.zeros(nb + 1, dtype=np.int64)
    iptr[1:] = np.cumsum(K)
    i0 = np.full(int(iptr[-1]), ZROW, dtype=np.int32)
    i1 = np.full(int(iptr[-1]), ZROW, dtype=np.int32)
    i2 = np.full(int(iptr[-1]), ZROW, dtype=np.int32)
    vptr = np.zeros(nb, dtype=np.int64)
    eptr = np.full(nb, -1, dtype=np.int64)
    vparts, e0p, e1p = [], [], []
    vpos = epos = 0
    for q, b in enumerate(blocks):
        sl = slice(iptr[q], iptr[q + 1])
        i0[sl] = b["i0"]
        if b["nsrc"] == 3:
            i1[sl], i2[sl] = b["i1"], b["i2"]
        vptr[q] = vpos
        v = np.ascontiguousarray(b["vals"], dtype=np.float64).reshape(int(M[q]), int(K[q]))
        vparts.append(v.ravel())
        vpos += v.size
        if b["e0"] is not None:
            eptr[q] = epos
            e0p.append(b["e0"])
            e1p.append(b["e1"])
            epos += int(M[q])
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)  # noqa: E731
    return SolvePlan(
        n=n, nU=nU, blk_K=K, blk_M=M,
        blk_nsrc=np.array([b["nsrc"] for b in blocks], dtype=np.int32),
        blk_out0=np.array([b["out0"] for b in blocks], dtype=np.int32),
        blk_ystore=np.array([b["ystore"] for b in blocks], dtype=np.int32),
        blk_iptr=iptr[:-1].copy(), blk_vptr=vptr, blk_eptr=eptr,
        i0=i0, i1=i1, i2=i2, e0=cat(e0p, np.int32), e1=cat(e1p, np.int32), vals=cat(vparts, np.float64),
        launch_ptr=np.array(launch_ptr, dtype=np.int32), n_forward_launches=n_fwd,
        asm_ptr=np.array(asm_ptr, dtype=np.int32), asm_src=np.array(asm_src, dtype=np.int32),
        asm_dst=np.array(asm_dst, dtype=np.int32),
    )


def apply_plan_host(plan: SolvePlan, b_perm: np.ndarray) -> np.ndarray:
    """Numpy emulation of the device sweeps, block by block, from the flat arrays (tests only)."""
    b = np.asarray(b_perm, dtype=np.float64)
    squeeze = b.ndim == 1
    if squeeze:
        b = b[:, None]
    n = plan.n
    Z = np.zeros((plan.z_rows, b.shape[1]))
    Z[:n] = b
    q_asm = int(plan.launch_ptr[plan.n_forward_launches])  # first block after the forward launches
    for q in range(len(plan.blk_K) + 1):
        if q == q_asm:
            for i, dst in enumerate(plan.asm_dst):
                Z[dst] = Z[plan.asm_src[plan.asm_ptr[i] : plan.asm_ptr[i + 1]]].sum(axis=0)
        if q == len(plan.blk_K):
            break
        K, M = int(plan.blk_K[q]), int(plan.blk_M[q])
        sl = slice(plan.blk_iptr[q], plan.blk_iptr[q] + K)
        x = Z[plan.i0[sl]]
        if plan.blk_nsrc[q] == 3:
            x = x + Z[plan.i1[sl]] + Z[plan.i2[sl]]
        if plan.blk_ystore[q] >= 0:
            Z[plan.blk_ystore[q] : plan.blk_ystore[q] + K] = x
        if M == 0:
            continue
        V = plan.vals[plan.blk_vptr[q] : plan.blk_vptr[q] + M * K].reshape(M, K)
        acc = V @ x
        if plan.blk_eptr[q] >= 0:
            es = slice(plan.blk_eptr[q], plan.blk_eptr[q] + M)
            for extra in (plan.e0[es], plan.e1[es]):
                has = extra >= 0
                acc[has] += Z[extra[has]]
        Z[plan.blk_out0[q] : plan.blk_out0[q] + M] = acc
    assert not Z[plan.zrow].any()
    x = Z[:n]
    return x[:, 0] if squeeze else x
