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
