"""Sensors: point probes and boundary-integral measurements.

Mirrors /root/reference/src/flowcontrol/sensor.py.  The reference evaluates a
sensor with ``dolfin.Function.eval`` (bounding-box tree search per call,
sensor.py:96-98 via utils/mpi.py:22-37) or a full ``dolfin.assemble``
(sensor.py:166-168); both are linear in the state, so each sensor is reduced at
setup to one sparse row ``(idx, val)`` over the canonical mixed vector, and the
per-step evaluation is the fused sparse dot product in kernel ``k_measure``.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from enum import IntEnum

import numpy as np

from .fem import p1_shape, p2_shape
from .mesh import TaylorHoodTables

SENSOR_INDEX_DEFAULT = 10000
DOLFIN_EPS = 3.0e-16


class SENSOR_TYPE(IntEnum):
    U = 0
    V = 1
    P = 2
    OTHER = 3


@dataclass(kw_only=True)
class Sensor(ABC):
    sensor_type: SENSOR_TYPE
    require_loading: bool

    @abstractmethod
    def row(self, tab: TaylorHoodTables) -> tuple[np.ndarray, np.ndarray]:
        """Sparse row: y = val . up[idx] over W = [ux | uy | p]."""

    def eval(self, up) -> float:
        """Host evaluation on one mixed vector (numpy array or object with ``.vector()``)."""
        vec = up.vector().get_local() if hasattr(up, "vector") else np.asarray(up)
        idx, val = self._row_cache
        return float(val @ vec[idx])

    def load(self, flowsolver) -> None:
        self._row_cache = self.row(flowsolver.tables)


@dataclass(kw_only=True)
class SensorPoint(Sensor):
    position: np.ndarray
    require_loading: bool = False

    def row(self, tab):
        c, xi, eta = tab.locate_point(float(self.position[0]), float(self.position[1]))
        if self.sensor_type in (SENSOR_TYPE.U, SENSOR_TYPE.V):
            phi, _ = p2_shape(xi, eta)
            return tab.cell_nodes[c].astype(np.int64) + int(self.sensor_type) * tab.nN, phi[0].copy()
        if self.sensor_type == SENSOR_TYPE.P:
            return tab.tri[c].astype(np.int64) + tab.Nv, p1_shape(xi, eta)[0].copy()
        raise ValueError("SensorPoint needs sensor_type U, V or P")


@dataclass(kw_only=True)
class SensorIntegral(Sensor):
    ds: object | None = None
    subdomain: object | None = None
    sensor_index: int = SENSOR_INDEX_DEFAULT
    require_loading: bool = True


@dataclass(kw_only=True)
class SensorHorizontalWallShear(SensorIntegral):
    """Integral of d(u_x)/dy along a horizontal wall segment (sensor.py:191-223)."""

    x_sensor_left: float = 1.0
    x_sensor_right: float = 1.1
    y_sensor: float = 0.0

    def inside(self, x, y):
        return (np.abs(y - self.y_sensor) <= DOLFIN_EPS) & (x >= self.x_sensor_left) & (x <= self.x_sensor_right)

    def row(self, tab):
        facets = tab.mark_boundary_facets(self.inside)
        owner = dict(zip(tab.bnd_edges.tolist(), tab.bnd_cells.tolist()))
        acc: dict[int, float] = {}
        for f in facets.tolist():
            c = owner[f]
            pa, pb = tab.xy[tab.edges[f, 0]], tab.xy[tab.edges[f, 1]]
            length = float(np.hypot(*(pb - pa)))
            mid = 0.5 * (pa + pb) - tab.xy[tab.tri[c, 0]]
            xi, eta = tab.Jinv[c] @ mid
            _, dref = p2_shape(xi, eta)
            dy = dref[0] @ tab.Jinv[c][:, 1]  # d phi_a / dy, constant along the facet normal direction
            for node, g in zip(tab.cell_nodes[c].tolist(), dy.tolist()):
                acc[node] = acc.get(node, 0.0) + length * g  # midpoint rule: exact for the linear integrand
        idx = np.array(sorted(acc), dtype=np.int64)
        return idx, np.array([acc[i] for i in idx.tolist()])


@dataclass(kw_only=True)
class SensorForceCoefficient(SensorIntegral):
    """Lift or drag coefficient of a body as a measurement row.

    The reference computes ``(cl, cd) = int_Gamma -(2 nu sym(grad u) - p I).n ds / (U^2 D / 2)`` with
    ``dolfin.assemble`` after the steady state or at the end of a run
    (examples/cylinder/cylinderflowsolver.py:115-126, examples/pinball/pinballflowsolver.py:202-232,
    utils/physics.py:17-19).  The functional is linear in (u, p), so it is one sparse row over the mixed vector and
    can be logged every step like any other sensor (BASELINE.json north_star asks for a lift time series).
    ``inside(x, y)`` selects the boundary facets of the body (all three surfaces for the cylinder: cylinder,
    actuator_up, actuator_lo); ``component`` 0 = drag, 1 = lift; ``n`` is dolfin's FacetNormal (out of the fluid).
    Like every sensor the row acts on the field it is given: the perturbation on the device, the full field in
    ``FlowSolver.compute_force_coefficients``."""

    inside: object = None
    component: int = 1
    nu: float = 0.01
    uinf: float = 1.0
    D: float = 1.0

    def row(self, tab):
        facets = tab.mark_boundary_facets(self.inside)
        owner = dict(zip(tab.bnd_edges.tolist(), tab.bnd_cells.tolist()))
        acc: dict[int, float] = {}
        i = int(self.component)
        scale = 1.0 / (0.5 * self.uinf**2 * self.D)
        for f in facets.tolist():
            c = owner[f]
            va, vb = int(tab.edges[f, 0]), int(tab.edges[f, 1])
            pa, pb = tab.xy[va], tab.xy[vb]
            t = pb - pa
            length = float(np.hypot(*t))
            nrm = np.array([t[1], -t[0]]) / length
            vo = [v for v in tab.tri[c].tolist() if v not in (va, vb)][0]
            if nrm @ (tab.xy[vo] - pa) > 0:  # the normal points away from the cell's third vertex
                nrm = -nrm
            mid = 0.5 * (pa + pb) - tab.xy[tab.tri[c, 0]]
            xi, eta = tab.Jinv[c] @ mid
            _, dref = p2_shape(xi, eta)
            grad = dref[0] @ tab.Jinv[c]  # [6, 2] physical gradients of the P2 basis at the facet midpoint
            # Fo_i = -nu (d_j u_i + d_i u_j) n_j + p n_i ; the integrand is linear along the facet: midpoint rule is exact
            for a, node in enumerate(tab.cell_nodes[c].tolist()):
                for comp in (0, 1):  # contribution of u_comp at this node
                    v = 0.0
                    if comp == i:
                        v += grad[a] @ nrm  # d_j u_i n_j
                    v += grad[a, i] * nrm[comp]  # d_i u_j n_j with j = comp
                    dof = node + comp * tab.nN
                    acc[dof] = acc.get(dof, 0.0) - self.nu * v * length * scale
            for v in (va, vb):  # P1 pressure: value 1/2 at the midpoint for the two facet vertices
                dof = tab.Nv + v
                acc[dof] = acc.get(dof, 0.0) + 0.5 * nrm[i] * length * scale
        idx = np.array(sorted(acc), dtype=np.int64)
        return idx, np.array([acc[k] for k in idx.tolist()])
