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
from .ordering import TreeNode, amalgamate, dissect, postorder


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

    def __init__(self, tab: TaylorHoodTables, free_mask: np.ndarray, leaf_cells: int = 8, amalgamate_above: int = 0,
                 balanced: bool = False):
        """``free_mask[N]`` is True for unknowns kept in the solve (non-Dirichlet).  ``amalgamate_above`` = h > 0 merges every
        tree node whose children have height >= h with those children (ordering.amalgamate): half the levels above height h.
        ``balanced`` bounds the tree depth by that of an even dissection (ordering.dissect)."""
        self.N = tab.N
        nN, nV = tab.nN, tab.nV
        tree = amalgamate(dissect(tab, leaf_cells=leaf_cells, balanced=balanced), amalgamate_above)
        self.amalgamate_above = int(amalgamate_above)
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
    gathered x_k to Z[ystore + k] (the block that owns y_t; M may be 0 for a store-only block); ``ystore == -2`` marks a
    backward block whose solution rows nobody gathers: the device writes them in canonical numbering only.

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
    asm_lptr: np.ndarray = None  # int32 [nlaunch+1]: gather-sum rows [asm_lptr[l], asm_lptr[l+1]) run right before launch l
    # Subtree clusters (csrc/fcb200.cu: k_cluster_sweep): connected pieces of the lower elimination tree that one CTA sweeps
    # with ALL their unknowns resident in shared memory (right-looking inside the piece, so the update vectors of the
    # fronts inside a cluster never exist in global memory).  Clusters of tier t only need the update vectors of cluster
    # roots of tiers < t; all tiers run before the pull-form launches above in the forward sweep and after them in the
    # backward sweep.  Fronts are listed cluster by cluster in elimination order.
    tier_ptr: np.ndarray = None  # int32 [ntier+1] cluster ranges
    cl_fptr: np.ndarray = None  # int32 [ncluster+1] front ranges
    cl_ustore: np.ndarray = None  # int32 [ncluster] first Z row of the root's update vector (-1: none)
    cl_iptr: np.ndarray = None  # int64 [ncluster+1] ranges of imp_src / imp_dst
    imp_src: np.ndarray = None  # int32 Z row (update vector of a lower cluster's root)
    imp_dst: np.ndarray = None  # int32 solver row it is added to
    fr_c0: np.ndarray = None  # int32 [nfront] first solver row of the front's own unknowns
    fr_w: np.ndarray = None  # int32 own unknowns
    fr_m: np.ndarray = None  # int32 boundary (struct) rows
    fr_sptr: np.ndarray = None  # int64 [nfront+1] ranges of fr_struct
    fr_struct: np.ndarray = None  # int32 solver rows of the boundary
    fr_eptr: np.ndarray = None  # int64 [nfront] offsets into cl_vals: E (m x w, row-major)
    fr_bptr: np.ndarray = None  # int64 [nfront] offsets into cl_vals: [F11^-1 | -G] (w x (w+m), row-major)
    cl_vals: np.ndarray = None

    @property
    def nnz(self) -> int:
        return int(self.vals.size) + (int(self.cl_vals.size) if self.cl_vals is not None else 0)

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


def choose_clusters(sym: SymbolicFactor, in_top: np.ndarray, max_rows: int, max_height: int, max_width: int = 128,
                    min_tier_clusters: int = 24):
    """Greedy bottom-up grouping of supernodes into clusters (connected subtrees) whose resident vector --- the own
    unknowns of every front of the cluster plus the boundary rows of its root --- has at most ``max_rows`` rows.

    Returns (cluster id per supernode or -1, list of clusters as supernode lists in elimination order, tier per cluster).
    A supernode above ``max_height``, in the merged top, wider than ``max_width`` or with a non-clustered child stays in the
    pull-form plan (so everything above a pull supernode is pull as well)."""
    sns = sym.supernodes
    nS = len(sns)
    cl_of = np.full(nS, -1, dtype=np.int64)
    open_cl: dict[int, tuple[int, list[int]]] = {}  # root -> (own rows, members) of clusters that may still grow
    clusters: list[list[int]] = []
    is_pull = np.zeros(nS, dtype=bool)

    def close(c: int) -> None:
        _, members = open_cl.pop(c)
        for j in members:
            cl_of[j] = len(clusters)
        clusters.append(members)

    for i, s in enumerate(sns):  # post-order: children first
        w, m = s.c1 - s.c0, len(s.struct)
        ch = sym.children[i]
        eligible = (not in_top[i]) and s.height <= max_height and w <= max_width and not any(is_pull[c] for c in ch) and max_rows > 0
        if not eligible:
            for c in ch:
                if c in open_cl:
                    close(c)
            is_pull[i] = True
            continue
        own = w + sum(open_cl[c][0] for c in ch if c in open_cl)
        if own + m <= max_rows:
            members: list[int] = []
            for c in ch:
                if c in open_cl:
                    members += open_cl.pop(c)[1]
            open_cl[i] = (own, members + [i])
        else:
            for c in ch:
                if c in open_cl:
                    close(c)
            if w + m <= max_rows:
                open_cl[i] = (w, [i])
            else:
                is_pull[i] = True
    for c in sorted(open_cl):
        close(c)
    # tiers: a cluster comes after the clusters its fronts import from
    tier = np.zeros(len(clusters), dtype=np.int64)
    order = sorted(range(len(clusters)), key=lambda q: clusters[q][-1])  # by root: children's clusters have smaller roots
    for q in order:
        t = 0
        for j in clusters[q]:
            for c in sym.children[j]:
                if cl_of[c] != q and cl_of[c] >= 0:
                    t = max(t, tier[cl_of[c]] + 1)
        tier[q] = t
    # a tier with only a handful of clusters is a launch that leaves most of the GPU idle: its fronts (and everything
    # above them) go back to the pull-form launches, which split one front over many CTAs
    ntier = int(tier.max()) + 1 if len(clusters) else 0
    for t in range(1, ntier):
        if int((tier == t).sum()) < min_tier_clusters:
            keep = [q for q in range(len(clusters)) if tier[q] < t]
            cl_of[:] = -1
            for new_q, q in enumerate(keep):
                cl_of[clusters[q]] = new_q
            clusters = [clusters[q] for q in keep]
            tier = tier[keep]
            break
    return cl_of, clusters, tier


def build_plan(fac: BlockFactor, top_levels: int = 2, cluster_rows: int = 0, cluster_height: int = 6,
               min_tier_clusters: int = 24, presum_height: int = 0, leaf_inplace: bool = False) -> SolvePlan:
    """``cluster_rows`` = 0 disables the shared-memory subtree clusters (every front goes through the pull-form launches).
    ``presum_height`` = h > 0: the forward blocks of fronts of height >= h do not gather three source planes (b and two
    update vectors) once per 32-row tile; a gather-sum right before their launch writes y_t = b_t + sum_children u_c once and
    the tiles gather that single plane (the gathered rows are two thirds of what the sweeps move through L2).
    ``leaf_inplace``: a leaf has no children, so y_t = b_t, and no descendants, so nobody gathers x_t in solver order: its
    forward block stores no y rows, its backward block gathers the (still intact) b rows instead and is marked
    ``ystore = -2`` -- x_t goes out in canonical numbering only.  Leaves own ~60 % of the unknowns: two of the seven
    vector-sized streams of a solve (y written, x written in solver order) shrink to the non-leaf rows."""
    sym = fac.sym
    sns = sym.supernodes
    n = sym.n
    nS = len(sns)
    top = [i for i, s in enumerate(sns) if s.depth < top_levels]  # post-order, closed under ancestors
    in_top = np.zeros(nS, dtype=bool)
    in_top[top] = True
    cl_of, clusters, cl_tier = choose_clusters(sym, in_top, cluster_rows, cluster_height, min_tier_clusters=min_tier_clusters)
    in_cluster = cl_of >= 0
    # update vectors live in global memory only for cluster roots and for the fronts of the pull-form plan
    needs_u = np.array([(not in_top[i]) and len(s.struct) > 0 and ((not in_cluster[i]) or clusters[cl_of[i]][-1] == i)
                        for i, s in enumerate(sns)], dtype=bool)
    uoff = np.zeros(nS + 1, dtype=np.int64)
    for i, s in enumerate(sns):
        uoff[i + 1] = uoff[i] + (len(s.struct) if needs_u[i] else 0)
    UB = 2 * n  # first row of the U region
    # A front gathers b and at most TWO update vectors (the sweep kernel's three source planes and two seed lists).  A front
    # with more children -- the amalgamated levels of the tree -- first sums groups of its children's update vectors into
    # "virtual" update vectors: rows of a gather-sum that runs right before the launch of the front's level.
    nvirt = int(uoff[-1])  # virtual vectors live behind the real ones in the U region
    sources: dict[int, list[tuple[np.ndarray, int]]] = {}  # front -> [(solver rows of the vector, sorted; first Z row)]
    presum: dict[int, list[tuple[int, list[int]]]] = {}  # front -> [(destination Z row, source Z rows in fixed order)]
    for i, s in enumerate(sns):
        if in_top[i] or in_cluster[i]:
            continue
        plain = [(sns[c].struct.astype(np.int64), UB + int(uoff[c])) for c in sym.children[i] if len(sns[c].struct)]
        if len(plain) <= 2:
            sources[i] = plain
            continue
        half = (len(plain) + 1) // 2
        srcs, pre = [], []
        for grp in (plain[:half], plain[half:]):
            if len(grp) == 1:
                srcs.append(grp[0])
                continue
            rows = np.unique(np.concatenate([g[0] for g in grp]))
            base = UB + nvirt
            nvirt += len(rows)
            contrib: list[list[int]] = [[] for _ in rows]
            for st, zb in grp:  # children in order: a fixed summation order per row
                for k, q in enumerate(np.searchsorted(rows, st)):
                    contrib[q].append(zb + k)
            pre += [(base + q, c) for q, c in enumerate(contrib)]
            srcs.append((rows, base))
        sources[i], presum[i] = srcs, pre
    nU = nvirt
    ZROW = 2 * n + nU

    def child_sources(i: int, rows: np.ndarray, absent: int) -> tuple[np.ndarray, np.ndarray]:
        """Z rows of the (at most two) update vectors feeding front ``i`` that hit the given solver rows."""
        srcs = [np.full(len(rows), absent, dtype=np.int64), np.full(len(rows), absent, dtype=np.int64)]
        for slot, (st, zb) in enumerate(sources[i]):
            if len(st) == 0 or len(rows) == 0:
                continue
            pos = np.searchsorted(st, rows)
            pos_c = np.minimum(pos, len(st) - 1)
            hit = st[pos_c] == rows
            srcs[slot][hit] = zb + pos_c[hit]
        return srcs[0], srcs[1]

    blocks: list[dict] = []
    launch_ptr = [0]
    asm_ptr, asm_src, asm_dst = [0], [], []
    asm_lptr = [0]  # gather-sum rows [asm_lptr[l], asm_lptr[l+1]) run right before launch l
    max_h = max(s.height for s in sns)
    inplace_leaf = np.zeros(nS, dtype=bool)
    for h in range(max_h + 1):
        for i in (i for i, s in enumerate(sns) if s.height == h and not in_top[i] and not in_cluster[i]):
            s = sns[i]
            w, m = s.c1 - s.c0, len(s.struct)
            if w == 0 and m == 0:
                continue
            for dst, srcs in presum.get(i, []):
                asm_src.extend(srcs)
                asm_ptr.append(len(asm_src))
                asm_dst.append(dst)
            own = np.arange(s.c0, s.c1, dtype=np.int64)
            a1, a2 = child_sources(i, own, ZROW)
            has_children = bool(sources[i])
            e0, e1 = child_sources(i, s.struct, -1)
            if has_children and m > 0 and w > 0 and presum_height > 0 and h >= presum_height:
                for k in range(w):  # y_t[k] = b + u_c1 + u_c2 in this fixed order
                    asm_src.extend([int(r) for r in (own[k], a1[k], a2[k]) if r != ZROW])
                    asm_ptr.append(len(asm_src))
                    asm_dst.append(n + s.c0 + k)
                blocks.append(dict(K=w, M=m, nsrc=1, out0=UB + int(uoff[i]), ystore=-1, i0=n + own, i1=None, i2=None,
                                   vals=-fac.blocks[i][0], e0=e0, e1=e1))
                continue
            inplace = leaf_inplace and not sym.children[i] and w > 0 and m > 0
            inplace_leaf[i] = inplace
            blocks.append(dict(K=w, M=m, nsrc=3 if has_children else 1, out0=UB + int(uoff[i]), ystore=-1 if inplace else n + s.c0,
                               i0=own, i1=a1, i2=a2, vals=-fac.blocks[i][0], e0=e0 if has_children else None,
                               e1=e1 if has_children else None))
        if len(blocks) > launch_ptr[-1]:
            launch_ptr.append(len(blocks))
            asm_lptr.append(len(asm_dst))
    n_fwd = len(launch_ptr) - 1
    # merged top: assemble r_T into the y rows of the top unknowns, then x_T = S^-1 r_T
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
            asm_lptr.append(len(asm_dst))
    max_d = max(s.depth for s in sns)
    for dpt in range(max_d + 1):
        for i in (i for i, s in enumerate(sns) if s.depth == dpt and not in_top[i] and not in_cluster[i]):
            s = sns[i]
            w = s.c1 - s.c0
            if w == 0:
                continue
            _, Finv, G = fac.blocks[i]
            full = np.concatenate([Finv, -G], axis=1)  # [w, w+m]
            own0 = 0 if inplace_leaf[i] else n  # an in-place leaf reads y_t = b_t where the right-hand side left it
            idx = np.concatenate([own0 + np.arange(s.c0, s.c1, dtype=np.int64), s.struct.astype(np.int64)])
            blocks.append(dict(K=full.shape[1], M=w, nsrc=1, out0=s.c0, ystore=-2 if inplace_leaf[i] else -1, i0=idx, i1=None,
                               i2=None, vals=full, e0=None, e1=None))
        if len(blocks) > launch_ptr[-1]:
            launch_ptr.append(len(blocks))
            asm_lptr.append(len(asm_dst))
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
    # ---- clusters, tier by tier; inside a tier the heaviest first (they start first on the device)
    def cl_work(q: int) -> int:
        return sum((sns[j].c1 - sns[j].c0) * ((sns[j].c1 - sns[j].c0) + 2 * len(sns[j].struct)) for j in clusters[q])

    ntier = int(cl_tier.max()) + 1 if len(clusters) else 0
    cl_order = sorted(range(len(clusters)), key=lambda q: (cl_tier[q], -cl_work(q)))
    tier_ptr = np.zeros(ntier + 1, dtype=np.int32)
    for q in cl_order:
        tier_ptr[cl_tier[q] + 1] += 1
    tier_ptr = np.cumsum(tier_ptr).astype(np.int32)
    cl_fptr, cl_ustore, cl_iptr = [0], [], [0]
    imp_src, imp_dst = [], []
    fr_c0, fr_w, fr_m, fr_sptr, fr_struct, fr_eptr, fr_bptr, cl_vparts = [], [], [], [0], [], [], [], []
    cvpos = 0
    for q in cl_order:
        members = clusters[q]
        root = members[-1]
        cl_ustore.append(UB + int(uoff[root]) if needs_u[root] else -1)
        for j in members:
            sj = sns[j]
            w, m = sj.c1 - sj.c0, len(sj.struct)
            E, Finv, G = fac.blocks[j]
            fr_c0.append(sj.c0); fr_w.append(w); fr_m.append(m)
            fr_struct.append(sj.struct.astype(np.int32))
            fr_sptr.append(fr_sptr[-1] + m)
            fr_eptr.append(cvpos)
            cl_vparts.append(np.ascontiguousarray(E, dtype=np.float64).ravel())
            cvpos += m * w
            fr_bptr.append(cvpos)
            cl_vparts.append(np.ascontiguousarray(np.concatenate([Finv, -G], axis=1), dtype=np.float64).ravel())
            cvpos += w * (w + m)
            for c in sym.children[j]:
                if cl_of[c] != cl_of[j]:  # the root of a lower cluster: its update vector is imported
                    mc = len(sns[c].struct)
                    imp_src.append(UB + uoff[c] + np.arange(mc, dtype=np.int64))
                    imp_dst.append(sns[c].struct.astype(np.int64))
        cl_fptr.append(len(fr_c0))
        cl_iptr.append(int(sum(len(a) for a in imp_src)))
    return SolvePlan(
        n=n, nU=nU, blk_K=K, blk_M=M,
        blk_nsrc=np.array([b["nsrc"] for b in blocks], dtype=np.int32),
        blk_out0=np.array([b["out0"] for b in blocks], dtype=np.int32),
        blk_ystore=np.array([b["ystore"] for b in blocks], dtype=np.int32),
        blk_iptr=iptr[:-1].copy(), blk_vptr=vptr, blk_eptr=eptr,
        i0=i0, i1=i1, i2=i2, e0=cat(e0p, np.int32), e1=cat(e1p, np.int32), vals=cat(vparts, np.float64),
        launch_ptr=np.array(launch_ptr, dtype=np.int32), n_forward_launches=n_fwd,
        asm_ptr=np.array(asm_ptr, dtype=np.int32), asm_src=np.array(asm_src, dtype=np.int32),
        asm_dst=np.array(asm_dst, dtype=np.int32), asm_lptr=np.array(asm_lptr, dtype=np.int32),
        tier_ptr=tier_ptr, cl_fptr=np.array(cl_fptr, dtype=np.int32), cl_ustore=np.array(cl_ustore, dtype=np.int32),
        cl_iptr=np.array(cl_iptr, dtype=np.int64), imp_src=cat(imp_src, np.int32), imp_dst=cat(imp_dst, np.int32),
        fr_c0=np.array(fr_c0, dtype=np.int32), fr_w=np.array(fr_w, dtype=np.int32), fr_m=np.array(fr_m, dtype=np.int32),
        fr_sptr=np.array(fr_sptr, dtype=np.int64), fr_struct=cat(fr_struct, np.int32),
        fr_eptr=np.array(fr_eptr, dtype=np.int64), fr_bptr=np.array(fr_bptr, dtype=np.int64),
        cl_vals=cat(cl_vparts, np.float64),
    )


def _cluster_sweep_host(plan: SolvePlan, Z: np.ndarray, q: int, forward: bool) -> None:
    """Numpy emulation of k_cluster_sweep for cluster ``q`` (same local vector, same order of operations)."""
    n = plan.n
    f0, f1 = int(plan.cl_fptr[q]), int(plan.cl_fptr[q + 1])
    c0, w, m = plan.fr_c0[f0:f1], plan.fr_w[f0:f1], plan.fr_m[f0:f1]
    off = np.concatenate([[0], np.cumsum(w)])
    nown = int(off[-1])
    root_struct = plan.fr_struct[plan.fr_sptr[f1 - 1] : plan.fr_sptr[f1]]
    # solver row -> local row of the resident vector
    loc = {}
    for k in range(f1 - f0):
        for r in range(int(w[k])):
            loc[int(c0[k]) + r] = int(off[k]) + r
    for j, g in enumerate(root_struct):
        loc[int(g)] = nown + j
    S = np.zeros((nown + len(root_struct), Z.shape[1]))
    own_rows = np.concatenate([np.arange(c0[k], c0[k] + w[k]) for k in range(f1 - f0)]) if f1 > f0 else np.zeros(0, dtype=np.int64)
    if forward:
        S[:nown] = Z[own_rows]
        for t in range(int(plan.cl_iptr[q]), int(plan.cl_iptr[q + 1])):
            S[loc[int(plan.imp_dst[t])]] += Z[plan.imp_src[t]]
        for k in range(f1 - f0):
            f = f0 + k
            st = np.array([loc[int(g)] for g in plan.fr_struct[plan.fr_sptr[f] : plan.fr_sptr[f + 1]]], dtype=np.int64)
            E = plan.cl_vals[plan.fr_eptr[f] : plan.fr_eptr[f] + int(m[k]) * int(w[k])].reshape(int(m[k]), int(w[k]))
            if len(st):
                S[st] -= E @ S[off[k] : off[k + 1]]
        Z[n + own_rows] = S[:nown]
        if plan.cl_ustore[q] >= 0:
            Z[plan.cl_ustore[q] : plan.cl_ustore[q] + len(root_struct)] = S[nown:]
    else:
        S[:nown] = Z[n + own_rows]
        S[nown:] = Z[root_struct]
        for k in reversed(range(f1 - f0)):
            f = f0 + k
            st = np.array([loc[int(g)] for g in plan.fr_struct[plan.fr_sptr[f] : plan.fr_sptr[f + 1]]], dtype=np.int64)
            Bm = plan.cl_vals[plan.fr_bptr[f] : plan.fr_bptr[f] + int(w[k]) * int(w[k] + m[k])].reshape(int(w[k]), int(w[k] + m[k]))
            S[off[k] : off[k + 1]] = Bm @ np.concatenate([S[off[k] : off[k + 1]], S[st]])
        Z[own_rows] = S[:nown]


def apply_plan_host(plan: SolvePlan, b_perm: np.ndarray) -> np.ndarray:
    """Numpy emulation of the device sweeps, block by block, from the flat arrays (tests only)."""
    b = np.asarray(b_perm, dtype=np.float64)
    squeeze = b.ndim == 1
    if squeeze:
        b = b[:, None]
    n = plan.n
    Z = np.zeros((plan.z_rows, b.shape[1]))
    Z[:n] = b
    ntier = len(plan.tier_ptr) - 1 if plan.tier_ptr is not None else 0
    for t in range(ntier):
        for q in range(int(plan.tier_ptr[t]), int(plan.tier_ptr[t + 1])):
            _cluster_sweep_host(plan, Z, q, True)
    first_block = {int(plan.launch_ptr[l]): l for l in range(len(plan.launch_ptr) - 1)}
    for q in range(len(plan.blk_K)):
        if q in first_block:  # gather-sums scheduled right before this launch (virtual update vectors, top right-hand side)
            l = first_block[q]
            for i in range(int(plan.asm_lptr[l]), int(plan.asm_lptr[l + 1])):
                acc = np.zeros(Z.shape[1])
                for src in plan.asm_src[plan.asm_ptr[i] : plan.asm_ptr[i + 1]]:  # fixed order, as on the device
                    acc = acc + Z[src]
                Z[plan.asm_dst[i]] = acc
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
    for t in reversed(range(ntier)):
        for q in range(int(plan.tier_ptr[t]), int(plan.tier_ptr[t + 1])):
            _cluster_sweep_host(plan, Z, q, False)
    assert not Z[plan.zrow].any()
    x = Z[:n]
    return x[:, 0] if squeeze else x
