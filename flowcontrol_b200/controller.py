"""LTI controllers: host object (drop-in) and the per-trajectory bank for the GPU.

Mirrors /root/reference/src/flowcontrol/controller.py:22-163.  The reference
subclasses ``control.StateSpace`` (python-control, not available here); only
``A,B,C,D,nstates,ninputs,noutputs`` and the ZOH discretisation are used on the
hot path, so a light state-space base is enough.  ``control.c2d(sys, dt, 'zoh')``
is ``scipy.signal.cont2discrete(..., method='zoh')``.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np
from scipy.signal import cont2discrete


def read_matfile(path) -> dict:
    """A, B, C, D of a MATLAB v5 file (utils/lticontrol.py:20-24) or of an .npz fixture."""
    path = Path(path)
    if path.suffix == ".npz":
        d = np.load(path)
        return {k: np.asarray(d[k], dtype=np.float64) for k in ("A", "B", "C", "D")}
    import warnings

    import scipy.io

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = scipy.io.loadmat(str(path))
    return {k: np.asarray(m[k], dtype=np.float64) for k in ("A", "B", "C", "D")}


class Controller:
    """Continuous-time state-space controller with cached ZOH discretisation."""

    def __init__(self, A, B, C, D, file: Path | None = None, x0=None):
        self.A = np.atleast_2d(np.asarray(A, dtype=np.float64))
        n = self.A.shape[0]
        self.B = np.asarray(B, dtype=np.float64).reshape(n, -1)
        self.C = np.asarray(C, dtype=np.float64).reshape(-1, n)
        self.D = np.asarray(D, dtype=np.float64).reshape(self.C.shape[0], self.B.shape[1])
        self.nstates, self.ninputs, self.noutputs = n, self.B.shape[1], self.C.shape[0]
        self.file = file
        self.x = np.asarray(x0, dtype=np.float64) if x0 is not None else np.zeros((n,))

    @classmethod
    def from_file(cls, file: Path, x0=None) -> "Controller":
        m = read_matfile(file)
        return cls(m["A"], m["B"], m["C"], m["D"], x0=x0, file=file)

    @classmethod
    def from_matrices(cls, A, B, C, D, file: Path | None = None, x0=None) -> "Controller":
        return cls(A, B, C, D, x0=x0, file=file)

    def _discretize(self, dt: float) -> None:
        Ad, Bd, Cd, Dd, _ = cont2discrete((self.A, self.B, self.C, self.D), dt, method="zoh")
        self._Ad, self._Bd, self._Cd, self._Dd, self._dt = Ad, Bd, Cd, Dd, dt

    def discrete(self, dt: float):
        if getattr(self, "_dt", None) != dt:
            self._discretize(dt)
        return self._Ad, self._Bd, self._Cd, self._Dd

    def step(self, y, dt: float) -> np.ndarray:
        """u = Cd x + Dd y with the PRE-update state, then x <- Ad x + Bd y (controller.py:157-158)."""
        Ad, Bd, Cd, Dd = self.discrete(dt)
        y = np.atleast_1d(y)
        u = Cd @ self.x + Dd @ y
        self.x = Ad @ self.x + Bd @ y
        return u

    def reset(self) -> None:
        self.x = np.zeros((self.nstates,))


class ControllerBank:
    """One discrete controller per trajectory, packed trajectory-innermost for
    ``fcb_set_controllers`` (include/fcb200.h).  Controllers with fewer states
    are zero-padded to the common ``nx``.

    ``Ky`` maps the measurement to the controller input (e.g. ``[[-1,0,0]]`` for
    ``y=-y_meas[0]``, run_cylinder_example.py:85) and ``Fu`` fans the controller
    output out to the actuators (``[[1],[1]]`` for ``np.repeat(u, 2)``, :86)."""

    def __init__(self, controllers: list[Controller], dt: float, Ky: np.ndarray, Fu: np.ndarray):
        B = len(controllers)
        nx = max(k.nstates for k in controllers)
        ny, nu = controllers[0].ninputs, controllers[0].noutputs
        self.nx, self.ny, self.nu, self.B = nx, ny, nu, B
        self.Ad = np.zeros((nx * nx, B))
        self.Bd = np.zeros((nx * ny, B))
        self.Cd = np.zeros((nu * nx, B))
        self.Dd = np.zeros((nu * ny, B))
        self.x0 = np.zeros((nx, B))
        for b, k in enumerate(controllers):
            if (k.ninputs, k.noutputs) != (ny, nu):
                raise ValueError("all controllers of a bank need the same input/output sizes")
            Ad, Bd, Cd, Dd = k.discrete(dt)
            n = k.nstates
            A = np.zeros((nx, nx)); A[:n, :n] = Ad
            Bm = np.zeros((nx, ny)); Bm[:n] = Bd
            Cm = np.zeros((nu, nx)); Cm[:, :n] = Cd
            self.Ad[:, b] = A.ravel()
            self.Bd[:, b] = Bm.ravel()
            self.Cd[:, b] = Cm.ravel()
            self.Dd[:, b] = np.asarray(Dd).ravel()
            self.x0[:n, b] = k.x
        self.Ky = np.ascontiguousarray(Ky, dtype=np.float64).reshape(ny, -1)
        self.Fu = np.ascontiguousarray(Fu, dtype=np.float64).reshape(-1, nu)
