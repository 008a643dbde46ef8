"""Containers for fields and file paths (host side).

Same dataclass names and attributes as
/root/reference/src/flowcontrol/flowfield.py:22-105.  ``dolfin.Function`` is
replaced by ``Field``: a numpy-backed object that offers the few accessors the
reference's callers use (``.vector().get_local()``, ``.vector()[:]``, ``copy``,
``assign``).  Fields of a running ensemble are fetched lazily from the GPU.
"""

from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Callable

import numpy as np


class Vector:
    """numpy-backed stand-in for ``dolfin.GenericVector``."""

    def __init__(self, owner: "Field"):
        self._owner = owner

    def get_local(self) -> np.ndarray:
        return np.array(self._owner.array, copy=True)

    def __getitem__(self, key):
        return self._owner.array[key]

    def __setitem__(self, key, value):
        arr = self._owner.array
        arr[key] = value
        self._owner.array = arr

    def apply(self, mode: str = "insert") -> None:  # dolfin API compatibility
        return None

    def __len__(self) -> int:
        return len(self._owner.array)

    def max(self):
        return self._owner.array.max()


class Field:
    """A dof vector on V (2nN), P (nV) or W (2nN+nV) in canonical numbering."""

    def __init__(self, data=None, size: int | None = None, fetch: Callable[[], np.ndarray] | None = None, Nv: int | None = None,
                 fetch_all: Callable[[], np.ndarray] | None = None):
        self._fetch = fetch
        self._fetch_all = fetch_all  # ensemble runs: every trajectory, [n, B]
        self._data = None if data is None else np.array(data, dtype=np.float64, copy=True)
        if self._data is None and fetch is None:
            self._data = np.zeros(int(size))
        self.Nv = Nv  # split point for mixed fields

    @property
    def array(self) -> np.ndarray:
        if self._data is None:
            self._data = np.asarray(self._fetch(), dtype=np.float64)
        return self._data

    @array.setter
    def array(self, value) -> None:
        self._data = np.asarray(value, dtype=np.float64)

    @property
    def ensemble(self) -> np.ndarray:
        """[n, B]: this field for every trajectory of the ensemble (a single run or a host-side field gives [n, 1])."""
        if self._fetch_all is not None:
            return np.asarray(self._fetch_all(), dtype=np.float64)
        return self.array[:, None]

    def vector(self) -> Vector:
        return Vector(self)

    def copy(self, deepcopy: bool = True) -> "Field":
        return Field(self.array, Nv=self.Nv)

    def assign(self, other: "Field") -> None:
        self._data = np.array(other.array, copy=True)

    def split(self, deepcopy: bool = True):
        if self.Nv is None:
            raise ValueError("not a mixed field")
        return Field(self.array[: self.Nv]), Field(self.array[self.Nv :])


@dataclass(frozen=True)
class SimPaths:
    U0: Path
    P0: Path
    U: Path
    P: Path
    Uprev: Path
    U_restart: Path
    Uprev_restart: Path
    P_restart: Path
    timeseries: Path
    metadata: Path
    steady_meta: Path
    mesh: Path


@dataclass
class FlowField:
    up: Field
    u: Field = None
    p: Field = None

    def __post_init__(self) -> None:
        self.u, self.p = self.up.split(deepcopy=True)


@dataclass
class FlowFieldCollection:
    U0: Field | None = None
    P0: Field | None = None
    UP0: Field | None = None
    ic: FlowField | None = None
    u_: Field | None = None
    p_: Field | None = None
    up_: Field | None = None
    u_n: Field | None = None
    u_nn: Field | None = None
    p_n: Field | None = None
    Usave: Field | None = None
    Psave: Field | None = None
    Usave_n: Field | None = None


@dataclass
class BoundaryConditions:
    bcu: list
    bcp: list
