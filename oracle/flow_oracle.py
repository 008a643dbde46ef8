"""CPU ORACLE (test infrastructure, NOT product code).

A self-contained numpy/scipy restatement of what the reference computes on its
time-stepping hot path, ``FlowSolver.step()`` and everything that feeds it
(/root/reference/src/flowcontrol/flowsolver.py:665-799), including the one-time
setup the reference delegates to FEniCS/dolfin 2019.1.0 + MUMPS (third-party,
un-vendored, pinned in /root/reference/environment.yml:5-11).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  Nothing in
``flowcontrol_b200/`` imports it, and it imports nothing from
``flowcontrol_b200/`` — the two implementations are derived independently
(different quadrature rule, different assembly strategy, different linear
solver) so that agreement between them is evidence and not tautology.

Parity status: PINNED.  ``tests/test_oracle_goldens.py`` checks this oracle
against the golden constants of the reference's own regression tests
(tests/integration/test_cylinder.py:66-74, test_cavity.py:47-54,
test_lidcavity.py:47-54, test_pinball.py:59-65).

Conventions
-----------
* P2 node ids: ``[vertices | nV + edge id]`` with edges = sorted unique vertex
  pairs.  Mixed vector ``W = [ux(nN) | uy(nN) | p(nV)]``.
* Local P2 ordering on a cell ``(v0,v1,v2)``: ``v0,v1,v2,e(v1v2),e(v0v2),e(v0v1)``.
* UFL conventions (nsforms.py:20): ``dot(U0, nabla_grad(u)) = (U0.grad)u``,
  ``dot(u, nabla_grad(U0)) = (u.grad)U0``.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Sequence

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.signal import cont2discrete

DOLFIN_EPS = 3.0e-16


# --------------------------------------------------------------------------- #
# dolfin C++ helper semantics (utils/fem.py:53-70)
# --------------------------------------------------------------------------- #
def near(a, b, tol=DOLFIN_EPS):
    """dolfin ``near(x, x0, eps)``: ``x0-eps <= x <= x0+eps`` (absolute)."""
    return (a >= b - tol) & (a <= b + tol)


def between(a, lo, hi, tol=0.0):
    """``between_cpp`` (utils/fem.py:57-58): inclusive with additive tolerance."""
    return (a >= lo - tol) & (a <= hi + tol)


# --------------------------------------------------------------------------- #
# Quadrature: collapsed Gauss-Legendre (Duffy) rule, n x n points, exact to
# total degree 2n-2 on the reference triangle {xi,eta>=0, xi+eta<=1}.
# --------------------------------------------------------------------------- #
def duffy_rule(n: int = 4):
    g, w = np.polynomial.legendre.leggauss(n)
    g = 0.5 * (g + 1.0)
    w = 0.5 * w
    X, Y = np.meshgrid(g, g, indexing="ij")
    WX, WY = np.meshgrid(w, w, indexing="ij")
    xi = X.ravel()
    eta = (Y * (1.0 - X)).ravel()
    wt = (WX * WY * (1.0 - X)).ravel()
    return xi, eta, wt


def p2_basis(xi, eta):
    """P2 Lagrange basis and reference gradients at points (xi, eta).

    Returns phi[q,6], dphi[q,6,2]."""
    xi = np.asarray(xi, dtype=float)
    eta = np.asarray(eta, dtype=float)
    l0, l1, l2 = 1.0 - xi - eta, xi, eta
    phi = np.stack(
        [l0 * (2 * l0 - 1), l1 * (2 * l1 - 1), l2 * (2 * l2 - 1), 4 * l1 * l2, 4 * l0 * l2, 4 * l0 * l1], axis=-1
    )
    z = np.zeros_like(xi)
    d = np.empty(xi.shape + (6, 2))
    d[..., 0, 0] = -(4 * l0 - 1)
    d[..., 0, 1] = -(4 * l0 - 1)
    d[..., 1, 0] = 4 * l1 - 1
    d[..., 1, 1] = z
    d[..., 2, 0] = z
    d[..., 2, 1] = 4 * l2 - 1
    d[..., 3, 0] = 4 * l2
    d[..., 3, 1] = 4 * l1
    d[..., 4, 0] = -4 * l2
    d[..., 4, 1] = 4 * (l0 - l2)
    d[..., 5, 0] = 4 * (l0 - l1)
    d[..., 5, 1] = -4 * l1
    return phi, d


def p1_basis(xi, eta):
    xi = np.asarray(xi, dtype=float)
    eta = np.asarray(eta, dtype=float)
    return np.stack([1.0 - xi - eta, xi, eta], axis=-1)


# --------------------------------------------------------------------------- #
# Mesh + Taylor-Hood tables
# --------------------------------------------------------------------------- #
class TaylorHoodMesh:
    """Mesh tables for P2-P1 on triangles (replaces dolfin FunctionSpace W,
    flowsolver.py:242-250)."""

    def __init__(self, vertices: np.ndarray, triangles: np.ndarray):
        self.xy = np.asarray(vertices, dtype=np.float64)
        self.tri = np.asarray(triangles, dtype=np.int64)
        nV, nT = len(self.xy), len(self.tri)
        t = self.tri
        # edge opposite local vertex i: (v1,v2), (v0,v2), (v0,v1)
        pairs = np.stack([t[:, [1, 2]], t[:, [0, 2]], t[:, [0, 1]]], axis=1)  # nT,3,2
        pairs = np.sort(pairs, axis=2)
        key = pairs[..., 0] * nV + pairs[..., 1]
        ukey, inv, counts = np.unique(key.ravel(), return_inverse=True, return_counts=True)
        self.edges = np.stack([ukey // nV, ukey % nV], axis=1)
        self.cell_edges = inv.reshape(nT, 3)
        self.nV, self.nT, self.nE = nV, nT, len(ukey)
        self.nN = nV + self.nE
        self.Nv = 2 * self.nN
        self.N = self.Nv + nV
        self.cell_nodes = np.concatenate([t, nV + self.cell_edges], axis=1)  # nT,6
        self.node_xy = np.concatenate([self.xy, 0.5 * (self.xy[self.edges[:, 0]] + self.xy[self.edges[:, 1]])])
        self.bnd_edges = np.flatnonzero(counts == 1)
        # adjacent cell of each boundary edge
        ecell = np.full(self.nE, -1, dtype=np.int64)
        ecell[self.cell_edges.ravel()] = np.repeat(np.arange(nT), 3)
        self.bnd_edge_cell = ecell[self.bnd_edges]
        # geometry
        p0, p1, p2 = self.xy[t[:, 0]], self.xy[t[:, 1]], self.xy[t[:, 2]]
        J = np.stack([p1 - p0, p2 - p0], axis=2)  # nT,2,2  x = p0 + J @ [xi,eta]
        det = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
        Jinv = np.empty_like(J)
        Jinv[:, 0, 0] = J[:, 1, 1] / det
        Jinv[:, 0, 1] = -J[:, 0, 1] / det
        Jinv[:, 1, 0] = -J[:, 1, 0] / det
        Jinv[:, 1, 1] = J[:, 0, 0] / det
        self.J, self.Jinv, self.detJ = J, Jinv, np.abs(det)

    # dof helpers
    def dof_ux(self, nodes):
        return np.asarray(nodes)

    def dof_uy(self, nodes):
        return np.asarray(nodes) + self.nN

    def dof_p(self, verts):
        return np.asarray(verts) + self.Nv

    def mark_facets(self, inside: Callable[[np.ndarray, np.ndarray], np.ndarray]) -> np.ndarray:
        """Boundary-edge ids whose two vertices AND midpoint satisfy ``inside``
        (dolfin SubDomain marking semantics)."""
        e = self.bnd_edges
        a, b = self.edges[e, 0], self.edges[e, 1]
        pa, pb = self.xy[a], self.xy[b]
        pm = 0.5 * (pa + pb)
        ok = inside(pa[:, 0], pa[:, 1]) & inside(pb[:, 0], pb[:, 1]) & inside(pm[:, 0], pm[:, 1])
        return e[ok]

    def facet_nodes(self, edge_ids: np.ndarray) -> np.ndarray:
        e = np.asarray(edge_ids)
        return np.unique(np.concatenate([self.edges[e, 0], self.edges[e, 1], self.nV + e]))

    def locate(self, x: float, y: float):
        """Find a cell containing (x,y); return (cell, xi, eta)."""
        p0 = self.xy[self.tri[:, 0]]
        d = np.stack([x - p0[:, 0], y - p0[:, 1]], axis=1)
        ref = np.einsum("eij,ej->ei", self.Jinv, d)
        xi, eta = ref[:, 0], ref[:, 1]
        tol = 1e-12
        ok = (xi >= -tol) & (eta >= -tol) & (xi + eta <= 1 + tol)
        idx = np.flatnonzero(ok)
        if len(idx) == 0:
            raise ValueError(f"point ({x},{y}) outside mesh")
        c = idx[0]
        return int(c), float(xi[c]), float(eta[c])


# --------------------------------------------------------------------------- #
# Operators
# --------------------------------------------------------------------------- #
class Operators:
    """Assembles the scalar blocks of Appendix A with a 16-point rule."""

    def __init__(self, mesh: TaylorHoodMesh, nquad: int = 4):
        self.m = mesh
        xi, eta, w = duffy_rule(nquad)
        self.w = w
        self.phi, dref = p2_basis(xi, eta)  # q,6 ; q,6,2
        self.psi = p1_basis(xi, eta)  # q,3
        # physical gradients dphi[e,q,a,j] = sum_k dref[q,a,k] Jinv[e,k,j]
        self.dphi = np.einsum("qak,ekj->eqaj", dref, mesh.Jinv)
        cn = mesh.cell_nodes
        self.rows22 = np.repeat(cn, 6, axis=1).ravel()
        self.cols22 = np.tile(cn, (1, 6)).ravel()
        self.rows12 = np.repeat(mesh.tri, 6, axis=1).ravel()
        self.cols12 = np.tile(cn, (1, 3)).ravel()
        nN, nV = mesh.nN, mesh.nV
        det = mesh.detJ
        Me = np.einsum("q,qa,qb->ab", w, self.phi, self.phi)[None] * det[:, None, None]
        Ke = np.einsum("e,q,eqaj,eqbj->eab", det, w, self.dphi, self.dphi)
        Bxe = np.einsum("e,q,qc,eqb->ecb", det, w, self.psi, self.dphi[..., 0])
        Bye = np.einsum("e,q,qc,eqb->ecb", det, w, self.psi, self.dphi[..., 1])
        self.M = self._asm22(Me)
        self.K = self._asm22(Ke)
        self.Bx = sp.coo_matrix((Bxe.ravel(), (self.rows12, self.cols12)), shape=(nV, nN)).tocsr()
        self.By = sp.coo_matrix((Bye.ravel(), (self.rows12, self.cols12)), shape=(nV, nN)).tocsr()
        self.Mv = sp.block_diag([self.M, self.M], format="csr")

    def _asm22(self, Ae):
        nN = self.m.nN
        return sp.coo_matrix((Ae.ravel(), (self.rows22, self.cols22)), shape=(nN, nN)).tocsr()

    def _at_quad(self, U):
        """U: velocity dof vector [2nN] -> values[e,q,2], grads[e,q,i,j]=d_j U_i."""
        m = self.m
        Ue = np.stack([U[: m.nN][m.cell_nodes], U[m.nN :][m.cell_nodes]], axis=2)  # e,a,i
        val = np.einsum("qa,eai->eqi", self.phi, Ue)
        grad = np.einsum("eqaj,eai->eqij", self.dphi, Ue)
        return val, grad

    def advection_blocks(self, U):
        """C_ab = int (U.grad phi_b) phi_a ; D^{ij}_ab = int phi_b (d_j U_i) phi_a."""
        val, grad = self._at_quad(U)
        det, w = self.m.detJ, self.w
        Ce = np.einsum("e,q,qa,eqj,eqbj->eab", det, w, self.phi, val, self.dphi, optimize=True)
        C = self._asm22(Ce)
        D = {}
        for i in range(2):
            for j in range(2):
                De = np.einsum("e,q,qa,qb,eq->eab", det, w, self.phi, self.phi, grad[:, :, i, j], optimize=True)
                D[i, j] = self._asm22(De)
        return C, D

    def convection(self, W):
        """N_i(w)_a = int (w.grad w_i) phi_a  -> vector [2nN] (nsforms.py:263,297-298)."""
        m = self.m
        val, grad = self._at_quad(W)
        conv = np.einsum("eqj,eqij->eqi", val, grad)
        Ne = np.einsum("e,q,qa,eqi->eai", m.detJ, self.w, self.phi, conv, optimize=True)
        out = np.zeros(m.Nv)
        np.add.at(out, m.cell_nodes.ravel(), Ne[..., 0].ravel())
        np.add.at(out, (m.cell_nodes + m.nN).ravel(), Ne[..., 1].ravel())
        return out

    def lhs(self, c_mass: float, Re: float, U0: np.ndarray | None, shift: float = 0.0, newton_terms: bool = True):
        """A(c) of Appendix A.  ``U0=None`` gives the Stokes-like operator."""
        F = c_mass * self.M + self.K / Re - shift * self.M
        Z = None
        if U0 is not None:
            C, D = self.advection_blocks(U0)
            F = F + C
        if U0 is not None and newton_terms:
            A = sp.bmat(
                [[F + D[0, 0], D[0, 1], -self.Bx.T], [D[1, 0], F + D[1, 1], -self.By.T], [-self.Bx, -self.By, Z]],
                format="csr",
            )
        else:
            A = sp.bmat([[F, None, -self.Bx.T], [None, F, -self.By.T], [-self.Bx, -self.By, Z]], format="csr")
        return A


# --------------------------------------------------------------------------- #
# Boundary conditions, actuators, sensors
# --------------------------------------------------------------------------- #
@dataclass
class DirichletSpec:
    """One ``dolfin.DirichletBC``: subdomain predicate, constrained components and value.

    ``value`` is either a constant tuple (one entry per component in ``comps``)
    or ``("actuator", k)`` meaning ``u_ctrl[k] * shape_k(x, y)``."""

    inside: Callable
    comps: tuple
    value: object


@dataclass
class ActuatorSpec:
    kind: str  # "bc" | "force"
    shape: Callable  # (x, y) -> (vx, vy) for u_ctrl = 1 (force: un-normalised)


@dataclass
class SensorSpec:
    kind: str  # "point" | "wall_shear"
    comp: int = 0  # SENSOR_TYPE: 0=U 1=V 2=P
    position: tuple = (0.0, 0.0)
    x_left: float = 0.0
    x_right: float = 0.0
    y: float = 0.0


@dataclass
class CaseSpec:
    name: str
    mesh_file: str
    Re: float
    dt: float
    uinf: float
    bcs_pert: list  # list[DirichletSpec], order matters
    bcs_full: list  # list[DirichletSpec] for the base flow
    actuators: list
    sensors: list
    initial_guess: Callable  # (x, y) -> (ux, uy)
    ic: tuple = (0.0, 0.0, 1.0, 1.0)  # xloc, yloc, radius, amplitude  (ParamIC defaults)
    pin_pressure: bool = False
    extra: dict = field(default_factory=dict)


def parabolic_slot(L, x0):
    """ActuatorBCParabolicV (actuator.py:190-199)."""

    def shape(x, y):
        d = x - x0
        v = np.where((d >= L) | (d <= -L), 0.0, -1.0 * (d + L) * (d - L) / (L * L))
        return np.zeros_like(x), v

    return shape


def rotation_profile(x0, y0, d):
    """ActuatorBCRotation (actuator.py:241-251)."""

    def shape(x, y):
        th = np.arctan2(y - y0, x - x0)
        return -np.sin(th) * d / 2, np.cos(th) * d / 2

    return shape


def uniform_u():
    """ActuatorBCUniformU (actuator.py:269-276)."""

    def shape(x, y):
        return np.ones_like(x), np.zeros_like(x)

    return shape


def gaussian_v(sigma, pos):
    """ActuatorForceGaussianV before normalisation (actuator.py:297-308)."""

    def shape(x, y):
        r2 = (x - pos[0]) ** 2 + (y - pos[1]) ** 2
        return np.zeros_like(x), np.exp(-0.5 * r2 / (sigma * sigma))

    return shape


class Constraints:
    """Dirichlet dof set with constant part and per-actuator shape columns.

    Later entries override earlier ones on shared dofs."""

    def __init__(self, mesh: TaylorHoodMesh, specs: Sequence[DirichletSpec], actuators: Sequence[ActuatorSpec]):
        na = len(actuators)
        const: dict[int, float] = {}
        act: dict[int, tuple] = {}
        for s in specs:
            nodes = mesh.facet_nodes(mesh.mark_facets(s.inside))
            x, y = mesh.node_xy[nodes, 0], mesh.node_xy[nodes, 1]
            if isinstance(s.value, tuple) and len(s.value) == 2 and s.value[0] == "actuator":
                k = s.value[1]
                vx, vy = actuators[k].shape(x, y)
                vals = {0: vx, 1: vy}
                for ci, c in enumerate(s.comps):
                    for n, v in zip(nodes, vals[c]):
                        dof = int(n) + c * mesh.nN
                        const[dof] = 0.0
                        act[dof] = (k, float(v))
            else:
                for ci, c in enumerate(s.comps):
                    for n in nodes:
                        dof = int(n) + c * mesh.nN
                        const[dof] = float(s.value[ci])
                        act.pop(dof, None)
        self.dofs = np.array(sorted(const), dtype=np.int64)
        self.g0 = np.array([const[d] for d in self.dofs])
        self.G = np.zeros((len(self.dofs), na))
        pos = {d: i for i, d in enumerate(self.dofs)}
        for d, (k, v) in act.items():
            self.G[pos[d], k] = v

    def values(self, u_ctrl):
        return self.g0 + self.G @ np.asarray(u_ctrl, dtype=float)


def sensor_row(mesh: TaylorHoodMesh, s: SensorSpec):
    """Sparse row (idx, val) with y = val . up[idx] (sensor.py:96-98, 166-168, 191-223)."""
    if s.kind == "point":
        c, xi, eta = mesh.locate(*s.position)
        if s.comp in (0, 1):
            phi, _ = p2_basis(np.array([xi]), np.array([eta]))
            idx = mesh.cell_nodes[c] + s.comp * mesh.nN
            return idx.astype(np.int64), phi[0]
        psi = p1_basis(np.array([xi]), np.array([eta]))
        return (mesh.tri[c] + mesh.Nv).astype(np.int64), psi[0]
    if s.kind == "wall_shear":

        def inside(x, y):
            return near(y, s.y) & (x >= s.x_left) & (x <= s.x_right)

        edges = mesh.mark_facets(inside)
        acc: dict[int, float] = {}
        ecell = dict(zip(mesh.bnd_edges.tolist(), mesh.bnd_edge_cell.tolist()))
        for e in edges:
            c = ecell[int(e)]
            a, b = mesh.edges[e]
            length = float(np.hypot(*(mesh.xy[a] - mesh.xy[b])))
            mid = 0.5 * (mesh.xy[a] + mesh.xy[b])
            ref = mesh.Jinv[c] @ (mid - mesh.xy[mesh.tri[c, 0]])
            _, dref = p2_basis(np.array([ref[0]]), np.array([ref[1]]))
            dphys = dref[0] @ mesh.Jinv[c]  # 6,2
            for n, g in zip(mesh.cell_nodes[c], dphys[:, 1]):  # d(u_x)/dy
                acc[int(n)] = acc.get(int(n), 0.0) + length * float(g)
        idx = np.array(sorted(acc), dtype=np.int64)
        return idx, np.array([acc[i] for i in idx])
    raise ValueError(s.kind)


def force_coefficients(mesh: TaylorHoodMesh, inside, up: np.ndarray, nu: float, uinf: float = 1.0, D: float = 1.0):
    """(cl, cd) of the body whose boundary facets satisfy ``inside``: int -(2 nu sym(grad u) - p I).n ds / (U^2 D / 2)
    with n = FacetNormal (out of the fluid), evaluated on the mixed vector ``up`` by 2-point Gauss quadrature on every
    facet (examples/cylinder/cylinderflowsolver.py:115-126, utils/physics.py:17-19)."""
    edges = mesh.mark_facets(inside)
    ecell = dict(zip(mesh.bnd_edges.tolist(), mesh.bnd_edge_cell.tolist()))
    gp = 0.5 + np.array([-0.5, 0.5]) / np.sqrt(3.0)
    F = np.zeros(2)
    ux, uy, pr = up[: mesh.nN], up[mesh.nN : mesh.Nv], up[mesh.Nv :]
    for e in edges:
        c = ecell[int(e)]
        a, b = mesh.edges[e]
        xa, xb = mesh.xy[a], mesh.xy[b]
        t = xb - xa
        length = float(np.hypot(*t))
        n = np.array([t[1], -t[0]]) / length
        centroid = mesh.xy[mesh.tri[c]].mean(axis=0)
        if n @ (centroid - xa) > 0:
            n = -n
        nodes = mesh.cell_nodes[c]
        for s_ in gp:
            x = xa + s_ * t
            ref = mesh.Jinv[c] @ (x - mesh.xy[mesh.tri[c, 0]])
            _, dref = p2_basis(np.array([ref[0]]), np.array([ref[1]]))
            psi = p1_basis(np.array([ref[0]]), np.array([ref[1]]))[0]
            dphys = dref[0] @ mesh.Jinv[c]  # [6, 2]
            gu = np.array([ux[nodes] @ dphys, uy[nodes] @ dphys])  # gu[i, j] = d_j u_i
            pq = psi @ pr[mesh.tri[c]]
            sigma = nu * (gu + gu.T) - pq * np.eye(2)
            F += 0.5 * length * (-(sigma @ n))
    drag, lift = F
    q = 0.5 * uinf**2 * D
    return lift / q, drag / q


# --------------------------------------------------------------------------- #
# LTI controller (controller.py:121-159)
# --------------------------------------------------------------------------- #
class ZOHController:
    def __init__(self, A, B, C, D, x0=None):
        self.A = np.atleast_2d(np.asarray(A, dtype=float))
        n = self.A.shape[0]
        self.B = np.asarray(B, dtype=float).reshape(n, -1)
        self.C = np.asarray(C, dtype=float).reshape(-1, n)
        self.D = np.asarray(D, dtype=float).reshape(self.C.shape[0], self.B.shape[1])
        self.x = np.zeros(n) if x0 is None else np.asarray(x0, dtype=float)
        self._dt = None

    def discretize(self, dt):
        Ad, Bd, Cd, Dd, _ = cont2discrete((self.A, self.B, self.C, self.D), dt, method="zoh")
        self.Ad, self.Bd, self.Cd, self.Dd, self._dt = Ad, Bd, Cd, Dd, dt

    def step(self, y, dt):
        if self._dt != dt:
            self.discretize(dt)
        y = np.atleast_1d(y)
        u = self.Cd @ self.x + self.Dd @ y
        self.x = self.Ad @ self.x + self.Bd @ y
        return u


# --------------------------------------------------------------------------- #
# The oracle solver
# --------------------------------------------------------------------------- #
class FlowOracle:
    """Restates FlowSolver: base flow, IC, BDF1->BDF2 stepping, sensors, energy."""

    def __init__(self, case: CaseSpec, vertices, triangles, time_scheme: str = "bdf", nonlinear: bool = True):
        """``time_scheme``: "bdf" (BDF1 start-up then BDF2, nsforms.py:238-305) or "cn" (Crank-Nicolson,
        nsforms.py:191-236, flowsolver.py:678-690, 755-758)."""
        if time_scheme not in ("bdf", "cn"):
            raise ValueError(time_scheme)
        self.time_scheme = time_scheme
        self.nonlinear = bool(nonlinear)  # ParamSolver.is_eq_nonlinear: b0 of nsforms.py:219
        self.case = case
        self.mesh = TaylorHoodMesh(vertices, triangles)
        self.ops = Operators(self.mesh)
        m = self.mesh
        # force actuators: nodal interpolant, unit L2 norm (actuator.py:310-311)
        self.force_vecs = []
        for a in case.actuators:
            if a.kind == "force":
                vx, vy = a.shape(m.node_xy[:, 0], m.node_xy[:, 1])
                s = np.concatenate([vx, vy])
                eta = 1.0 / np.sqrt(s @ (self.ops.Mv @ s))
                self.force_vecs.append(self.ops.Mv @ (eta * s))
            else:
                self.force_vecs.append(None)
        self.bc_pert = Constraints(m, case.bcs_pert, case.actuators)
        self.bc_full = Constraints(m, case.bcs_full, case.actuators)
        self.sensor_rows = [sensor_row(m, s) for s in case.sensors]
        self.UP0 = None
        self.lu = {}

    # -- base flow (steadystate.py:60-159) ------------------------------------
    def _apply_rows(self, A, b, dofs, vals):
        A = A.tolil(copy=True) if False else A.tocsr(copy=True)
        mask = np.ones(A.shape[0])
        mask[dofs] = 0.0
        Dm = sp.diags(mask)
        Id = sp.diags(1.0 - mask)
        A = Dm @ A + Id
        b = b.copy()
        b[dofs] = vals
        return A.tocsc(), b

    def _pin(self, A, b):
        if self.case.pin_pressure:
            d = self.mesh.Nv  # first pressure dof
            return self._apply_rows(A, b, np.array([d]), np.array([0.0]))
        return A, b

    def force_rhs(self, u_ctrl):
        f = np.zeros(self.mesh.Nv)
        for k, fv in enumerate(self.force_vecs):
            if fv is not None and u_ctrl[k] != 0.0:
                f += u_ctrl[k] * fv
        return f

    def initial_guess(self):
        m = self.mesh
        ux, uy = self.case.initial_guess(m.node_xy[:, 0], m.node_xy[:, 1])
        return np.concatenate([ux, uy, np.zeros(m.nV)])

    def picard(self, UP, u_ctrl, max_iter=10, tol=1e-8, log=None):
        m = self.mesh
        g = self.bc_full.values(u_ctrl)
        b0 = np.concatenate([self.force_rhs(u_ctrl), np.zeros(m.nV)])
        UP = UP.copy()
        for i in range(max_iter):
            A = self.ops.lhs(0.0, self.case.Re, UP[: m.Nv], newton_terms=False)
            A, b = self._apply_rows(A, b0, self.bc_full.dofs, g)
            A, b = self._pin(A, b)
            UP1 = spla.splu(A).solve(b)
            rel = np.linalg.norm(UP1 - UP) / (np.linalg.norm(UP) + 1e-14)
            UP = UP1
            if log:
                log(f"picard {i + 1}/{max_iter} rel_err={rel:.3e}")
            if rel < tol:
                break
        return UP

    def steady_residual(self, UP, u_ctrl):
        m, o = self.mesh, self.ops
        U, P = UP[: m.Nv], UP[m.Nv :]
        Kv = sp.block_diag([o.K, o.K], format="csr")
        r_u = o.convection(U) + Kv @ U / self.case.Re
        r_u[: m.nN] -= o.Bx.T @ P
        r_u[m.nN :] -= o.By.T @ P
        r_u -= self.force_rhs(u_ctrl)
        r_p = -(o.Bx @ U[: m.nN] + o.By @ U[m.nN :])
        return np.concatenate([r_u, r_p])

    def newton(self, UP, u_ctrl, max_iter=25, rtol=1e-9, atol=1e-10, log=None):
        """dolfin NewtonSolver defaults."""
        m = self.mesh
        g = self.bc_full.values(u_ctrl)
        dofs = self.bc_full.dofs
        UP = UP.copy()
        r0 = None
        for it in range(max_iter + 1):
            b = self.steady_residual(UP, u_ctrl)
            b[dofs] = UP[dofs] - g
            if self.case.pin_pressure:
                b[m.Nv] = UP[m.Nv]
            r = np.linalg.norm(b)
            if r0 is None:
                r0 = r
            if log:
                log(f"newton {it}: r(abs)={r:.3e} r(rel)={r / max(r0, 1e-300):.3e}")
            if r < atol or r / max(r0, 1e-300) < rtol:
                break
            if it == max_iter:
                raise RuntimeError("Newton did not converge")
            Jm = self.ops.lhs(0.0, self.case.Re, UP[: m.Nv], newton_terms=True)
            Jm, b = self._apply_rows(Jm, b, dofs, b[dofs])
            Jm, b = self._pin(Jm, b)
            UP = UP - spla.splu(Jm).solve(b)
        return UP

    def set_base_flow(self, UP0):
        self.UP0 = UP0.copy()
        self.E0 = 0.5 * UP0[: self.mesh.Nv] @ (self.ops.Mv @ UP0[: self.mesh.Nv])

    # -- time stepping ---------------------------------------------------------
    def default_ic(self):
        """Nodal interpolant of the div-free Gaussian (utils/physics.py:32-56,
        flowsolver.py:522-536); pressure part = amplitude * P0 (flowsolver.py:908-912)."""
        m = self.mesh
        xloc, yloc, radius, amp = self.case.ic
        out = np.zeros(m.N)
        if amp and radius > 0:
            x, y = m.node_xy[:, 0], m.node_xy[:, 1]
            psi = 0.25 * np.exp(-0.5 * ((x - xloc) ** 2 + (y - yloc) ** 2) / radius**2)
            dpsi_dx = psi * (-(x - xloc) / radius**2)
            dpsi_dy = psi * (-(y - yloc) / radius**2)
            out[: m.nN] = amp * dpsi_dy
            out[m.nN : m.Nv] = -amp * dpsi_dx
            out[m.Nv :] = amp * self.UP0[m.Nv :]
        return out

    def prepare(self):
        """_prepare_systems (flowsolver.py:665-701): LHS of BDF1 and BDF2 with
        symmetric Dirichlet elimination, factorised once."""
        m = self.mesh
        dt, Re = self.case.dt, self.case.Re
        self.A_raw, self.A_bc, self.lu = {}, {}, {}
        dofs = self.bc_pert.dofs
        mask = np.ones(m.N)
        mask[dofs] = 0.0
        Dm, Id = sp.diags(mask), sp.diags(1.0 - mask)
        orders = ((1, 1.0 / dt), (2, 1.5 / dt)) if self.time_scheme == "bdf" else (("cn", 1.0 / dt),)
        for order, c in orders:
            A = self.ops.lhs(c, Re, self.UP0[: m.Nv], newton_terms=True)
            if order == "cn":
                # theta = 1/2 (nsforms.py:213-236): the linear velocity terms C + D + K/Re are split half implicit / half
                # explicit; mass, pressure and continuity stay fully implicit
                L = self.ops.lhs(0.0, Re, self.UP0[: m.Nv], newton_terms=True).tocsr()[: m.Nv, : m.Nv]
                Lpad = sp.bmat([[L, None], [None, sp.csr_matrix((m.nV, m.nV))]], format="csr")
                A = (A - 0.5 * Lpad).tocsr()
                self.E_cn = (self.ops.Mv / dt - 0.5 * L).tocsr()  # explicit operator on u_n
            self.A_raw[order] = A
            Abc = (Dm @ A @ Dm + Id).tocsc()
            if self.case.pin_pressure:
                Abc, _ = self._pin(Abc, np.zeros(m.N))
            self.A_bc[order] = Abc
            self.lu[order] = spla.splu(Abc)

    def init_time_stepping(self, ic=None):
        m = self.mesh
        self.ic = self.default_ic() if ic is None else ic.copy()
        self.u_n = self.ic[: m.Nv].copy()
        self.u_nn = self.u_n.copy()
        self.up = self.ic.copy()
        self.order = 1 if self.time_scheme == "bdf" else "cn"
        self.iter = 0
        self.t = 0.0
        self.u_ctrl_prev = np.zeros(len(self.case.actuators))  # f_n_field starts at zero (flowsolver.py:681-690)
        if not self.lu:
            self.prepare()
        self.y_meas = self.measure(self.ic)
        self.N_prev = None
        return self.y_meas

    def rhs(self, u_ctrl):
        """SystemAssembler.assemble(rhs) (flowsolver.py:728) — Appendix A."""
        m, o = self.mesh, self.ops
        dt = self.case.dt
        u_ctrl = np.asarray(u_ctrl, dtype=float)
        # N(u_nn) was N(u_n) of the previous step: reuse it (same value, half the assembly cost)
        Nn = o.convection(self.u_n) if self.nonlinear else np.zeros(m.Nv)
        Nnn = self.N_prev if (self.N_prev is not None and self.order == 2) else None
        fterm = self.force_rhs(u_ctrl)
        if self.order == "cn":
            rv = self.E_cn @ self.u_n - Nn  # b0 = 1 (nsforms.py:219, 229)
            fterm = 0.5 * (fterm + self.force_rhs(self.u_ctrl_prev))  # body force averaged over the step (nsforms.py:233-234)
        elif self.order == 1:
            rv = o.Mv @ self.u_n / dt - Nn
        else:
            if Nnn is None:
                Nnn = o.convection(self.u_nn) if self.nonlinear else np.zeros(m.Nv)
            rv = o.Mv @ (4.0 * self.u_n - self.u_nn) / (2.0 * dt) - 2.0 * Nn + Nnn
        self._N_cur = Nn
        rv = rv + fterm
        b = np.concatenate([rv, np.zeros(m.nV)])
        g = self.bc_pert.values(u_ctrl)
        dofs = self.bc_pert.dofs
        if np.any(g != 0.0):
            gfull = np.zeros(m.N)
            gfull[dofs] = g
            b = b - self.A_raw[self.order] @ gfull
        b[dofs] = g
        if self.case.pin_pressure:
            b[m.Nv] = 0.0
        return b

    def step(self, u_ctrl):
        m = self.mesh
        b = self.rhs(u_ctrl)
        x = self.lu[self.order].solve(b)
        if not np.all(np.isfinite(x[: m.Nv])):
            return None
        self.iter += 1
        self.t = self.iter * self.case.dt
        self.order = 2 if self.time_scheme == "bdf" else "cn"
        self.u_ctrl_prev = np.asarray(u_ctrl, dtype=float).copy()
        self.u_nn = self.u_n
        self.u_n = x[: m.Nv].copy()
        self.N_prev = self._N_cur
        self.up = x
        self.y_meas = self.measure(x)
        self.dE = self.energy()
        return self.y_meas

    def measure(self, up):
        return np.array([val @ up[idx] for idx, val in self.sensor_rows])

    def energy(self):
        return 0.5 * self.u_n @ (self.ops.Mv @ self.u_n)

    def full_velocity(self):
        return self.u_n + self.UP0[: self.mesh.Nv]
