"""Taylor-Hood (P2-P1) mesh tables built on the host at setup time.

Replaces what ``dolfin.Mesh`` + ``dolfin.FunctionSpace(mesh, MixedElement([P2^2, P1]))``
provide to the reference (/root/reference/src/flowcontrol/flowsolver.py:233-250):
node numbering, cell->dof maps, boundary facets, affine geometry, and (new) the
element colouring the CUDA scatter kernel consumes.

Canonical numbering:
    P2 node ids  = [vertices | nV + edge id], edges = sorted unique vertex pairs
    mixed vector = [ux(nN) | uy(nN) | p(nV)]
Local node order on a triangle (v0,v1,v2): v0, v1, v2, m12, m02, m01
(edge node k+3 is opposite vertex k).
"""

from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path

import numpy as np


@dataclass
class TaylorHoodTables:
    xy: np.ndarray  # [nV,2] vertex coordinates
    tri: np.ndarray  # [nT,3] int32
    edges: np.ndarray  # [nE,2] int32, sorted pairs
    cell_nodes: np.ndarray  # [nT,6] int32 P2 node ids
    node_xy: np.ndarray  # [nN,2] P2 node coordinates
    bnd_edges: np.ndarray  # [nB] int32 edge ids with exactly one incident cell
    bnd_cells: np.ndarray  # [nB] int32 the incident cell
    Jinv: np.ndarray  # [nT,2,2] inverse Jacobian (d ref / d phys)
    detJ: np.ndarray  # [nT] |det J|

    @property
    def nV(self) -> int:
        return self.xy.shape[0]

    @property
    def nT(self) -> int:
        return self.tri.shape[0]

    @property
    def nE(self) -> int:
        return self.edges.shape[0]

    @property
    def nN(self) -> int:
        return self.nV + self.nE

    @property
    def Nv(self) -> int:
        return 2 * self.nN

    @property
    def N(self) -> int:
        return 2 * self.nN + self.nV

    # ------------------------------------------------------------------ build
    @classmethod
    def from_arrays(cls, vertices, triangles) -> "TaylorHoodTables":
        xy = np.ascontiguousarray(vertices, dtype=np.float64)[:, :2]
        tri = np.ascontiguousarray(triangles, dtype=np.int64)
        nV = xy.shape[0]
        nT = tri.shape[0]
        # local edge k joins the two vertices other than k
        lo = np.minimum(tri[:, [1, 0, 0]], tri[:, [2, 2, 1]])
        hi = np.maximum(tri[:, [1, 0, 0]], tri[:, [2, 2, 1]])
        code = (lo * nV + hi).ravel()
        order = np.argsort(code, kind="stable")
        sorted_code = code[order]
        first = np.ones(len(code), dtype=bool)
        first[1:] = sorted_code[1:] != sorted_code[:-1]
        eid_sorted = np.cumsum(first) - 1
        eid = np.empty(len(code), dtype=np.int64)
        eid[order] = eid_sorted
        ucode = sorted_code[first]
        edges = np.stack([ucode // nV, ucode % nV], axis=1)
        nE = edges.shape[0]
        cell_edges = eid.reshape(nT, 3)
        mult = np.bincount(eid, minlength=nE)
        bnd_edges = np.flatnonzero(mult == 1)
        owner = np.empty(nE, dtype=np.int64)
        owner[cell_edges.ravel()] = np.repeat(np.arange(nT), 3)
        cell_nodes = np.concatenate([tri, nV + cell_edges], axis=1)
        node_xy = np.concatenate([xy, 0.5 * (xy[edges[:, 0]] + xy[edges[:, 1]])], axis=0)
        a = xy[tri[:, 1]] - xy[tri[:, 0]]
        b = xy[tri[:, 2]] - xy[tri[:, 0]]
        det = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
        Jinv = np.empty((nT, 2, 2))
        # J = [[a0, b0],[a1, b1]]  (x = x0 + J [xi, eta])
        Jinv[:, 0, 0] = b[:, 1] / det
        Jinv[:, 0, 1] = -b[:, 0] / det
        Jinv[:, 1, 0] = -a[:, 1] / det
        Jinv[:, 1, 1] = a[:, 0] / det
        return cls(
            xy=xy,
            tri=tri.astype(np.int32),
            edges=edges.astype(np.int32),
            cell_nodes=cell_nodes.astype(np.int32),
            node_xy=node_xy,
            bnd_edges=bnd_edges.astype(np.int32),
            bnd_cells=owner[bnd_edges].astype(np.int32),
            Jinv=Jinv,
            detJ=np.abs(det),
        )

    @classmethod
    def from_file(cls, path: str | Path) -> "TaylorHoodTables":
        """Load from ``.xdmf`` (+ sibling ``.h5``) or from an ``.npz`` fixture."""
        path = Path(path)
        if path.suffix == ".npz":
            d = np.load(path)
            return cls.from_arrays(d["vertices"], d["triangles"])
        from .hdf5_lite import read_xdmf_mesh

        xy, tri = read_xdmf_mesh(path)
        return cls.from_arrays(xy, tri)

    # --------------------------------------------------------------- queries
    def mark_boundary_facets(self, inside) -> np.ndarray:
        """Ids (into ``edges``) of exterior facets inside a subdomain.

        dolfin marks a facet iff both end vertices and the midpoint satisfy the
        predicate with ``on_boundary=True``."""
        e = self.bnd_edges
        pa = self.xy[self.edges[e, 0]]
        pb = self.xy[self.edges[e, 1]]
        pm = 0.5 * (pa + pb)
        ok = np.asarray(inside(pa[:, 0], pa[:, 1]), dtype=bool)
        ok = ok & np.asarray(inside(pb[:, 0], pb[:, 1]), dtype=bool)
        ok = ok & np.asarray(inside(pm[:, 0], pm[:, 1]), dtype=bool)
        return e[ok]

    def facet_p2_nodes(self, facet_ids) -> np.ndarray:
        f = np.asarray(facet_ids, dtype=np.int64)
        if f.size == 0:
            return np.zeros(0, dtype=np.int64)
        return np.unique(np.concatenate([self.edges[f, 0], self.edges[f, 1], self.nV + f]))

    def locate_point(self, x: float, y: float, tol: float = 1e-12) -> tuple[int, float, float]:
        """Return (cell, xi, eta) of a cell containing the point."""
        d0 = x - self.xy[self.tri[:, 0], 0]
        d1 = y - self.xy[self.tri[:, 0], 1]
        xi = self.Jinv[:, 0, 0] * d0 + self.Jinv[:, 0, 1] * d1
        eta = self.Jinv[:, 1, 0] * d0 + self.Jinv[:, 1, 1] * d1
        hit = np.flatnonzero((xi >= -tol) & (eta >= -tol) & (xi + eta <= 1.0 + tol))
        if hit.size == 0:
            raise ValueError(f"point ({x}, {y}) is outside the mesh")
        c = int(hit[0])
        return c, float(xi[c]), float(eta[c])

    def element_colouring(self) -> tuple[np.ndarray, np.ndarray]:
        """Greedy colouring of cells such that cells of one colour share no P2 node.

        Returns (colour_ptr[ncol+1], colour_cells[nT]) — the atomic-free scatter
        schedule of the RHS-assembly kernel."""
        nT, nN = self.nT, self.nN
        cn = self.cell_nodes
        # node -> cells adjacency (CSR)
        flat = cn.ravel()
        order = np.argsort(flat, kind="stable")
        cells_sorted = (order // 6).astype(np.int64)
        ptr = np.zeros(nN + 1, dtype=np.int64)
        np.add.at(ptr, flat + 1, 1)
        ptr = np.cumsum(ptr)
        colour = np.full(nT, -1, dtype=np.int64)
        for c in range(nT):
            used = 0
            for n in cn[c]:
                nb = cells_sorted[ptr[n] : ptr[n + 1]]
                cols = colour[nb]
                for k in cols[cols >= 0]:
                    used |= 1 << int(k)
            k = 0
            while used >> k & 1:
                k += 1
            colour[c] = k
        ncol = int(colour.max()) + 1
        order = np.argsort(colour, kind="stable")
        cptr = np.zeros(ncol + 1, dtype=np.int32)
        cptr[1:] = np.cumsum(np.bincount(colour, minlength=ncol))
        return cptr, order.astype(np.int32)
