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
    """Flat job arrays consumed by the CUDA kernel ``k_front_sweep`` (csrc/fcb200.cu).

    The kernel works on one buffer Z with rows
        [0, n)          b on entry, x on exit            (solver row order)
        [n, 2n)         y (forward-eliminated RHS)
        [2n, 2n + nU)   update vectors u_t of every supernode (its subdomain-boundary rows)
        2n + nU         a row of zeros (``zrow``): the source of padded / absent gathers
    A *job* is a small dense GEMM  acc[8*nrb x traj] = V[8*nrb x K4] . x[K4 x traj]  (K4 = K rounded up
    to a multiple of 4, the k-depth of one FP64 tensor-core MMA) whose input rows are gathered as
    x_k = Z[i0[k]] (+ Z[i1[k]] + Z[i2[k]] when nsrc == 3) and whose ``nr`` output rows are written to
    consecutive rows  Z[out0 + r] = acc[r] (+ Z[e0[r]] + Z[e1[r]], -1 = absent).
    ``ystore >= 0`` additionally stores the gathered x_k (k < K) to Z[ystore + k] (the job that owns y_t).

        forward  (per supernode t, tiles over its m boundary rows, launches by tree height):
            x_k = b_t[k] + sum_children u_c[..]            (= y_t)
            u_t[r] = sum_children u_c[..] - (E_t y_t)[r]
        backward (tiles over its w own rows, launches by tree depth):
            x_t[r] = (F11^-1 y_t)[r] - (G_t x_struct(t))[r]

    ``vals`` holds every job's V in MMA A-fragment order: [K4/4][nrb][8 rows][4 k] doubles, so the 32
    lanes of a warp read one 8x4 fragment as 256 contiguous bytes.
    Every output row is produced by exactly one job: no atomics, bit-reproducible."""

    n: int
    nU: int
    job_K: np.ndarray  # int32 [njobs] gathered rows (index arrays are padded to K4 with zrow)
    job_nrb: np.ndarray  # int32 8-row blocks of the value tile (0 = store-only job, else 1..4)
    job_nr: np.ndarray  # int32 valid output rows (<= 8*nrb)
    job_nsrc: np.ndarray  # int32 1 or 3
    job_out0: np.ndarray  # int32
    job_ystore: np.ndarray  # int32 (-1 = none)
    job_iptr: np.ndarray  # int64 offsets into i0/i1/i2
    job_vptr: np.ndarray  # int64 offsets into vals
    job_eptr: np.ndarray  # int64 offsets into e0/e1 (-1 = no epilogue gather)
    i0: np.ndarray
    i1: np.ndarray
    i2: np.ndarray
    e0: np.ndarray
    e1: np.ndarray
    vals: np.ndarray
    launch_ptr: np.ndarray  # int32 [nlaunch+1] job ranges, forward launches first
    n_forward_launches: int

    @property
    def nnz_padded(self) -> int:
        return int(self.vals.size)

    @property
    def zrow(self) -> int:
        return 2 * self.n + self.nU

    @property
    def z_rows(self) -> int:
        return 2 * self.n + self.nU + 1


def _row_tiles(nrows: int, max_rb: int = 4) -> list[tuple[int, int, int]]:
    """Split ``nrows`` into (r0, nr, nrb) tiles of at most ``max_rb`` 8-row blocks, evenly sized."""
    nblk = (nrows + 7) // 8
    ntile = (nblk + max_rb - 1) // max_rb
    out, r0 = [], 0
    for t in range(ntile):
        nb = nblk // ntile + (1 if t < nblk % ntile else 0)
        nr = min(8 * nb, nrows - r0)
        out.append((r0, nr, nb))
        r0 += nr
    assert r0 == nrows
    return out


def _pack_fragments(V: np.ndarray, nrb: int) -> np.ndarray:
    """V [nr, K] -> A-fragment order [K4/4][nrb][8][4] (zero padded), flattened."""
    nr, K = V.shape
    K4 = (K + 3) // 4 * 4
    P = np.zeros((8 * nrb, K4))
    P[:nr, :K] = V
    return P.reshape(nrb, 8, K4 // 4, 4).transpose(2, 0, 1, 3).ravel()


def build_plan(fac: BlockFactor, max_rb: int = 4, target_jobs: int = 296) -> SolvePlan:
    """``max_rb``: largest tile height in 8-row blocks; levels with few rows use shorter tiles so that
    a launch has at least ~``target_jobs`` independent CTAs' worth of work where possible."""
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

    jobs: list[dict] = []
    launches: list[list[int]] = []

    def tile_blocks_for(level_rows: list[int]) -> int:
        for rb in range(max_rb, 1, -1):
            if sum((r + 8 * rb - 1) // (8 * rb) for r in level_rows) >= target_jobs:
                return rb
        return 1

    max_h = max(s.height for s in sns)
    for h in range(max_h + 1):
        ids = [i for i, s in enumerate(sns) if s.height == h]
        rb_cap = tile_blocks_for([len(sns[i].struct) for i in ids])
        cur: list[int] = []
        for i in ids:
            s = sns[i]
            w, m = s.c1 - s.c0, len(s.struct)
            own = np.arange(s.c0, s.c1, dtype=np.int64)
            a1, a2 = child_sources(i, own, ZROW)
            nsrc = 3 if sym.children[i] else 1
            E = fac.blocks[i][0]
            if m == 0:
                # root of the tree: nothing to update, only y_t = b_t + children; rows are independent,
                # so the copy is cut into short store-only jobs that spread over the SMs
                for r0 in range(0, w, 16):
                    sl = slice(r0, min(w, r0 + 16))
                    jobs.append(dict(K=sl.stop - r0, nrb=0, nr=0, nsrc=nsrc, out0=0, ystore=n + s.c0 + r0, i0=own[sl],
                                     i1=a1[sl], i2=a2[sl], vals=np.zeros(0), e0=None, e1=None))
                    cur.append(len(jobs) - 1)
                continue
            e0, e1 = child_sources(i, s.struct, -1)
            for t, (r0, nr, nrb) in enumerate(_row_tiles(m, rb_cap)):
                jobs.append(dict(K=w, nrb=nrb, nr=nr, nsrc=nsrc, out0=UB + int(uoff[i]) + r0,
                                 ystore=(n + s.c0) if t == 0 else -1, i0=own, i1=a1, i2=a2,
                                 vals=_pack_fragments(-E[r0 : r0 + nr, :], nrb),
                                 e0=e0[r0 : r0 + nr] if nsrc == 3 else None, e1=e1[r0 : r0 + nr] if nsrc == 3 else None))
                cur.append(len(jobs) - 1)
        if cur:
            launches.append(cur)
    n_fwd = len(launches)
    max_d = max(s.depth for s in sns)
    for dpt in range(max_d + 1):
        ids = [i for i, s in enumerate(sns) if s.depth == dpt]
        rb_cap = tile_blocks_for([sns[i].c1 - sns[i].c0 for i in ids])
        cur = []
        for i in ids:
            s = sns[i]
            w = s.c1 - s.c0
            if w == 0:
                continue
            _, Finv, G = fac.blocks[i]
            full = np.concatenate([Finv, -G], axis=1)  # [w, w+m]
            idx = np.concatenate([n + np.arange(s.c0, s.c1, dtype=np.int64), s.struct.astype(np.int64)])
            for r0, nr, nrb in _row_tiles(w, rb_cap):
                jobs.append(dict(K=full.shape[1], nrb=nrb, nr=nr, nsrc=1, out0=s.c0 + r0, ystore=-1, i0=idx, i1=None,
                                 i2=None, vals=_pack_fragments(full[r0 : r0 + nr, :], nrb), e0=None, e1=None))
                cur.append(len(jobs) - 1)
        if cur:
            launches.append(cur)
    # longest jobs first inside each launch (the kernel deals them round-robin to persistent CTAs)
    order: list[int] = []
    launch_ptr = [0]
    for cur in launches:
        cur = sorted(cur, key=lambda j: -(jobs[j]["K"] * jobs[j]["nsrc"] + jobs[j]["K"] * jobs[j]["nrb"] + 8 * jobs[j]["nrb"]))
        order += cur
        launch_ptr.append(len(order))
    nj = len(order)
    K = np.array([jobs[j]["K"] for j in order], dtype=np.int32)
    K4 = (K.astype(np.int64) + 3) // 4 * 4
    iptr = np.zeros(nj + 1, dtype=np.int64)
    iptr[1:] = np.cumsum(K4)
    i0 = np.full(int(iptr[-1]), ZROW, dtype=np.int32)
    i1 = np.full(int(iptr[-1]), ZROW, dtype=np.int32)
    i2 = np.full(int(iptr[-1]), ZROW, dtype=np.int32)
    vptr = np.zeros(nj, dtype=np.int64)
    eptr = np.full(nj, -1, dtype=np.int64)
    vparts, e0p, e1p = [], [], []
    vpos = epos = 0
    for q, j in enumerate(order):
        jb = jobs[j]
        sl = slice(iptr[q], iptr[q] + jb["K"])
        i0[sl] = jb["i0"]
        if jb["nsrc"] == 3:
            i1[sl], i2[sl] = jb["i1"], jb["i2"]
        vptr[q] = vpos
        vparts.append(jb["vals"])
        vpos += jb["vals"].size
        if jb["e0"] is not None:
            eptr[q] = epos
            e0p.append(jb["e0"])
            e1p.append(jb["e1"])
            epos += jb["nr"]
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)  # noqa: E731
    return SolvePlan(
        n=n, nU=nU, job_K=K,
        job_nrb=np.array([jobs[j]["nrb"] for j in order], dtype=np.int32),
        job_nr=np.array([jobs[j]["nr"] for j in order], dtype=np.int32),
        job_nsrc=np.array([jobs[j]["nsrc"] for j in order], dtype=np.int32),
        job_out0=np.array([jobs[j]["out0"] for j in order], dtype=np.int32),
        job_ystore=np.array([jobs[j]["ystore"] for j in order], dtype=np.int32),
        job_iptr=iptr[:-1].copy(), job_vptr=vptr, job_eptr=eptr,
        i0=i0, i1=i1, i2=i2, e0=cat(e0p, np.int32), e1=cat(e1p, np.int32), vals=cat(vparts, np.float64),
        launch_ptr=np.array(launch_ptr, dtype=np.int32), n_forward_launches=n_fwd,
    )


def apply_plan_host(plan: SolvePlan, b_perm: np.ndarray) -> np.ndarray:
    """Numpy emulation of the CUDA sweeps, job by job, from the packed arrays (tests only)."""
    b = np.asarray(b_perm, dtype=np.float64)
    squeeze = b.ndim == 1
    if squeeze:
        b = b[:, None]
    n = plan.n
    Z = np.zeros((plan.z_rows, b.shape[1]))
    Z[:n] = b
    for q in range(len(plan.job_K)):
        K, nrb, nr = int(plan.job_K[q]), int(plan.job_nrb[q]), int(plan.job_nr[q])
        K4 = (K + 3) // 4 * 4
        sl = slice(plan.job_iptr[q], plan.job_iptr[q] + K4)
        x = Z[plan.i0[sl]]
        if plan.job_nsrc[q] == 3:
            x = x + Z[plan.i1[sl]] + Z[plan.i2[sl]]
        if plan.job_ystore[q] >= 0:
            Z[plan.job_ystore[q] : plan.job_ystore[q] + K] = x[:K]
        if nrb == 0:
            continue
        V = plan.vals[plan.job_vptr[q] : plan.job_vptr[q] + K4 * 8 * nrb]
        V = V.reshape(K4 // 4, nrb, 8, 4).transpose(1, 2, 0, 3).reshape(8 * nrb, K4)
        acc = V[:nr] @ x
        if plan.job_eptr[q] >= 0:
            es = slice(plan.job_eptr[q], plan.job_eptr[q] + nr)
            for extra in (plan.e0[es], plan.e1[es]):
                has = extra >= 0
                acc[has] += Z[extra[has]]
        Z[plan.job_out0[q] : plan.job_out0[q] + nr] = acc
    assert not Z[plan.zrow].any()
    x = Z[:n]
    return x[:, 0] if squeeze else x
