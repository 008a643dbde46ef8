"""Actuators: boundary-condition and body-force control inputs.

Mirrors /root/reference/src/flowcontrol/actuator.py (class names, constructor
fields, ``expression.u_ctrl`` get/set).  The reference wraps a JIT-compiled
``dolfin.Expression``; here every profile is ``u_ctrl * shape(x, y)`` with
``shape`` evaluated once at setup on the P2 node coordinates, which is what the
CUDA step consumes (Dirichlet shape columns / force vectors, fcb_problem.bc_shape
and .ctrl_rhs in include/fcb200.h).
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from enum import IntEnum

import numpy as np


class ACTUATOR_TYPE(IntEnum):
    BC = 1
    FORCE = 2


class CYLINDER_ACTUATION_MODE(IntEnum):
    SUCTION = 1
    ROTATION = 2


class _Expression:
    """Stand-in for ``dolfin.Expression``: carries the mutable amplitude ``u_ctrl``
    (flowsolver.py:296 sets it, :304 reads it) and evaluates the profile."""

    def __init__(self, shape_fn, u_ctrl: float = 0.0):
        self._shape_fn = shape_fn
        self.u_ctrl = u_ctrl

    def shape(self, x, y):
        return self._shape_fn(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64))

    def __call__(self, x, y):
        sx, sy = self.shape(x, y)
        return self.u_ctrl * sx, self.u_ctrl * sy


@dataclass(kw_only=True)
class Actuator(ABC):
    actuator_type: ACTUATOR_TYPE
    expression: _Expression | None = None

    @abstractmethod
    def shape(self, x: np.ndarray, y: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        """Profile (vx, vy) at points for unit amplitude."""

    def load_expression(self, flowsolver=None) -> _Expression:
        self.expression = _Expression(self.shape, 0.0)
        return self.expression


@dataclass(kw_only=True)
class ActuatorBC(Actuator):
    boundary_name: str | None = None
    boundary: object | None = None

    def load_expression(self, flowsolver=None) -> _Expression:
        super().load_expression(flowsolver)
        if self.boundary_name is not None and flowsolver is not None:
            try:
                self.boundary = flowsolver.get_subdomain(self.boundary_name)
            except KeyError:
                available = list(flowsolver.boundaries.index)
                raise KeyError(
                    f"Actuator boundary_name={self.boundary_name!r} not found in "
                    f"FlowSolver.boundaries. Available: {available}"
                ) from None
        return self.expression


@dataclass(kw_only=True)
class ActuatorBCParabolicV(ActuatorBC):
    """Parabolic normal-velocity slot, zero outside [x0-L, x0+L] (actuator.py:190-199)."""

    width: float = 0.0
    position_x: float = 0.0
    actuator_type: ACTUATOR_TYPE = ACTUATOR_TYPE.BC

    def shape(self, x, y):
        L, d = self.width, x - self.position_x
        with np.errstate(divide="ignore", invalid="ignore"):
            v = np.where((d >= L) | (d <= -L), 0.0, -1.0 * (d + L) * (d - L) / (L * L))
        return np.zeros_like(d), v

    @staticmethod
    def angular_size_deg_to_width(angular_size_deg: float, cylinder_radius: float) -> float:
        return cylinder_radius * np.sin(1 / 2 * angular_size_deg * np.pi / 180)


@dataclass(kw_only=True)
class ActuatorBCRotation(ActuatorBC):
    """Tangential velocity of a cylinder spinning at rate u_ctrl (actuator.py:241-251)."""

    position_x: float = 0.0
    position_y: float = 0.0
    diameter: float = 1.0
    actuator_type: ACTUATOR_TYPE = ACTUATOR_TYPE.BC

    def shape(self, x, y):
        th = np.arctan2(y - self.position_y, x - self.position_x)
        return -np.sin(th) * self.diameter / 2, np.cos(th) * self.diameter / 2


@dataclass(kw_only=True)
class ActuatorBCUniformU(ActuatorBC):
    """Uniform streamwise lid velocity (actuator.py:269-276)."""

    actuator_type: ACTUATOR_TYPE = ACTUATOR_TYPE.BC

    def shape(self, x, y):
        return np.ones_like(x, dtype=np.float64), np.zeros_like(x, dtype=np.float64)


@dataclass(kw_only=True)
class ActuatorForceGaussianV(Actuator):
    """Unit-L2-norm Gaussian body force on the v component (actuator.py:297-312).

    ``eta`` is set by ``normalise`` once the mass matrix of the mesh is known."""

    sigma: float
    position: np.ndarray
    actuator_type: ACTUATOR_TYPE = ACTUATOR_TYPE.FORCE
    eta: float = field(default=1.0, compare=False)

    def shape(self, x, y):
        r2 = (x - self.position[0]) ** 2 + (y - self.position[1]) ** 2
        return np.zeros_like(r2), self.eta * np.exp(-0.5 * r2 / (self.sigma * self.sigma))

    def normalise(self, node_xy: np.ndarray, Mv) -> None:
        """eta = 1 / ||P2-interpolant of the profile||_L2  (actuator.py:310-311)."""
        self.eta = 1.0
        sx, sy = self.shape(node_xy[:, 0], node_xy[:, 1])
        s = np.concatenate([sx, sy])
        self.eta = 1.0 / float(np.sqrt(s @ (Mv @ s)))
