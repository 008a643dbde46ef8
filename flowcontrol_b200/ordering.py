"""Element-based nested dissection of a Taylor-Hood mesh (host, setup time).

The reference hands its constant LHS to MUMPS (flowsolver.py:694-697, 812-814),
whose analysis phase orders the unknowns with METIS/SCOTCH/AMD.  This build
needs an ordering whose elimination tree is shallow and whose supernodes are
dense blocks the GPU solve can stream, so it bisects the *cells* geometrically
and takes the P2 nodes shared by the two halves as the separator:

    tree node t  <->  element set E_t
    sep(t)       =   nodes shared by E_left and E_right, not owned by an ancestor
    leaf         ->  owns the remaining (interior) nodes of its cells
    struct(t)    =   nodes of E_t owned by strict ancestors  (= boundary of the subdomain)

Each tree node becomes one supernode of the multifrontal factorisation
(multifrontal.py); post-order over the tree is the elimination order.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .mesh import TaylorHoodTables


@dataclass
class TreeNode:
    cells: np.ndarray
    parent: int
    children: list = field(default_factory=list)
    own_nodes: np.ndarray | None = None  # P2 nodes eliminated at this tree node
    bnd_nodes: np.ndarray | None = None  # P2 nodes of cells owned by strict ancestors
    depth: int = 0


def _order_along_principal_axis(xy: np.ndarray) -> np.ndarray:
    if xy.shape[0] <= 2:
        return np.arange(xy.shape[0])
    c = xy - xy.mean(axis=0)
    cov = c.T @ c
    w, v = np.linalg.eigh(cov)
    axis = v[:, -1]
    return np.argsort(c @ axis, kind="stable")


# eight cut directions (every 22.5 degrees): 3 % fewer factor entries than four on the cylinder mesh
_CUT_DIRECTIONS = [np.array((np.cos(a), np.sin(a))) for a in np.linspace(0.0, np.pi, 8, endpoint=False)]


def dissect(
    tab: TaylorHoodTables, leaf_cells: int = 8, cut_fractions=(0.4, 0.45, 0.5, 0.55, 0.6), balanced: bool = False
) -> list[TreeNode]:
    """Return the dissection tree as a list (index 0 = root).

    Every bisection tries eight cut directions and a few cut positions around the
    median and keeps the one with the fewest shared P2 nodes.

    ``balanced`` bounds the depth of the tree by D = ceil(log2(cells / leaf_cells)), the depth of a perfectly even
    dissection: a cut is only admissible if both halves still fit ``leaf_cells * 2**(levels left)`` cells (the median
    cut always is).  The device sweeps cost one launch per tree level (csrc/fcb200.cu: k_front_sweep), so the two or
    three extra levels that free cuts leave below their larger halves are launches with a handful of fronts each."""
    cn = tab.cell_nodes
    cent = tab.node_xy[tab.cell_nodes[:, :3]].mean(axis=1)
    nodes: list[TreeNode] = [TreeNode(cells=np.arange(tab.nT), parent=-1, depth=0)]
    max_depth = int(np.ceil(np.log2(max(tab.nT / leaf_cells, 1.0)))) if balanced else None
    stack = [0]
    while stack:
        t = stack.pop()
        cells = nodes[t].cells
        if len(cells) <= leaf_cells:
            continue
        c = cent[cells]
        best = None
        cap = leaf_cells * 2 ** (max_depth - nodes[t].depth - 1) if balanced else len(cells)  # cells a child may hold
        fractions = [f for f in cut_fractions if max(f, 1.0 - f) * len(cells) <= cap] or [0.5]
        for direction in _CUT_DIRECTIONS:
            proj = c @ direction
            order = np.argsort(proj, kind="stable")
            for frac in fractions:
                half = int(round(frac * len(cells)))
                half = min(max(half, 1, len(cells) - cap), len(cells) - 1, cap)
                a = np.unique(cn[cells[order[:half]]].ravel())
                b = np.unique(cn[cells[order[half:]]].ravel())
                nshared = np.intersect1d(a, b, assume_unique=True).size
                # mild penalty for imbalance so that ties prefer the median cut
                score = nshared * (1.0 + 0.5 * abs(frac - 0.5))
                if best is None or score < best[0]:
                    best = (score, order, half)
        _, order, half = best
        for part in (cells[order[:half]], cells[order[half:]]):
            nodes.append(TreeNode(cells=part, parent=t, depth=nodes[t].depth + 1))
            nodes[t].children.append(len(nodes) - 1)
            stack.append(len(nodes) - 1)
    # ownership: top-down, a node shared by both children belongs to the highest such tree node
    owner = np.full(tab.nN, -1, dtype=np.int64)
    cn = tab.cell_nodes
    order_bfs = sorted(range(len(nodes)), key=lambda i: nodes[i].depth)
    for t in order_bfs:
        nd = nodes[t]
        mine = np.unique(cn[nd.cells].ravel())
        if nd.children:
            a = np.unique(cn[nodes[nd.children[0]].cells].ravel())
            b = np.unique(cn[nodes[nd.children[1]].cells].ravel())
            shared = np.intersect1d(a, b, assume_unique=True)
            own = shared[owner[shared] < 0]
        else:
            own = mine[owner[mine] < 0]
        bnd = mine[(owner[mine] >= 0)]
        owner[own] = t
        nd.own_nodes = own[_order_along_principal_axis(tab.node_xy[own])] if len(own) else own
        nd.bnd_nodes = bnd
    assert np.all(owner >= 0)
    return nodes


def postorder(nodes: list[TreeNode]) -> list[int]:
    out: list[int] = []
    stack = [(0, False)]
    while stack:
        t, done = stack.pop()
        if done or not nodes[t].children:
            out.append(t)
        else:
            stack.append((t, True))
            for c in reversed(nodes[t].children):
                stack.append((c, False))
    return out


def amalgamate(nodes: list[TreeNode], min_child_height: int) -> list[TreeNode]:
    """Merge tree nodes pairwise along the depth: a node whose two children both have height >= ``min_child_height``
    absorbs them (its separator becomes the union of the three separators, its children are the four grandchildren), and
    the same rule is then applied to the grandchildren.  The elimination tree above the small fronts gets half as many
    levels -- half as many dependent launches in the GPU sweeps -- for slightly larger dense fronts (the two absorbed
    separators are not coupled to each other, their zero block is stored).  ``min_child_height`` <= 0 returns the tree
    unchanged.  The result is a new list (index 0 = root) with parent / children / depth rebuilt."""
    if min_child_height <= 0:
        return nodes
    height = [0] * len(nodes)
    for t in reversed(postorder(nodes)[::-1]):  # children before parents
        pass
    for t in postorder(nodes):
        for c in nodes[t].children:
            height[t] = max(height[t], height[c] + 1)
    out: list[TreeNode] = []

    def build(t: int, parent: int, depth: int) -> int:
        nd = nodes[t]
        kids = list(nd.children)
        own = nd.own_nodes
        if len(kids) == 2 and all(height[c] >= min_child_height for c in kids) and all(len(nodes[c].children) == 2 for c in kids):
            own = np.concatenate([nodes[kids[0]].own_nodes, nodes[kids[1]].own_nodes, nd.own_nodes])
            kids = [g for c in kids for g in nodes[c].children]
        me = len(out)
        out.append(TreeNode(cells=nd.cells, parent=parent, own_nodes=own, bnd_nodes=nd.bnd_nodes, depth=depth))
        for c in kids:
            out[me].children.append(build(c, me, depth + 1))
        return me

    import sys

    sys.setrecursionlimit(max(sys.getrecursionlimit(), 10000))
    build(0, -1, 0)
    return out
