"""Device-resident ensemble of trajectories that share one FlowProblem.

Thin Python owner of an ``fcb_handle`` (include/fcb200.h).  Arrays are numpy on
the host side; torch CUDA tensors (or raw device pointers) are accepted wherever
a data pointer is expected, in which case no host<->device copy happens.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import libfcb
from .controller import ControllerBank
from .problem import FlowProblem


class Ensemble:
    def __init__(self, problem: FlowProblem, batch: int, device: int = 0):
        self.problem = problem
        self.B = int(batch)
        self.device = int(device)
        self.lib = libfcb.load()
        self._pack = libfcb.ProblemPack(problem, batch=self.B)
        h = C.c_void_p()
        rc = self.lib.fcb_create(C.byref(self._pack.struct), self.B, self.device, C.byref(h))
        if rc != 0:
            raise libfcb.FcbError(f"fcb_create failed ({rc}): {self.lib.fcb_last_error(None).decode()}")
        self.h = h
        tab = problem.tab
        self.N, self.Nv, self.nV = tab.N, tab.Nv, tab.nV
        self.na, self.ns = problem.na, problem.ns
        self.y_meas = np.zeros((self.ns, self.B))
        self.dE = np.zeros(self.B)
        self.diverged = np.zeros(self.B, dtype=np.int32)
        self._bank = None

    # -- plumbing -----------------------------------------------------------
    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise libfcb.FcbError(f"{what} failed ({rc}): {self.lib.fcb_last_error(self.h).decode()}")

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.fcb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _host(self, a, rows: int) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.ndim == 1:
            a = np.repeat(a[:, None], self.B, axis=1)
        if a.shape != (rows, self.B):
            raise ValueError(f"expected shape ({rows}, {self.B}), got {a.shape}")
        return np.ascontiguousarray(a)

    # -- state --------------------------------------------------------------
    def set_state(self, u_n, u_nn=None, p_n=None, order: int = 1) -> np.ndarray:
        """Load perturbation history; 1-D arrays are broadcast to every trajectory.
        Returns the measurement of the loaded state (flowsolver.py:495)."""
        un = self._host(u_n, self.Nv)
        unn = self._host(u_nn, self.Nv) if u_nn is not None else None
        pn = self._host(p_n, self.nV) if p_n is not None else None
        order = 2 if order == "cn" else int(order)  # Crank-Nicolson has one system; the library ignores the order then
        rc = self.lib.fcb_set_state(self.h, libfcb.as_voidp(un), libfcb.as_voidp(unn), libfcb.as_voidp(pn), order)
        self._check(rc, "fcb_set_state")
        return self.measurement()

    def measurement(self):
        rc = self.lib.fcb_get_measurement(
            self.h, libfcb.as_voidp(self.y_meas), libfcb.as_voidp(self.dE), libfcb.as_voidp(self.diverged)
        )
        self._check(rc, "fcb_get_measurement")
        return self.y_meas

    def set_controllers(self, bank: ControllerBank) -> None:
        if bank.B != self.B:
            raise ValueError("controller bank width != ensemble width")
        c = libfcb.fcb_controllers()
        c.nx, c.ny, c.nu = bank.nx, bank.ny, bank.nu
        keep = [np.ascontiguousarray(getattr(bank, k), dtype=np.float64) for k in ("Ad", "Bd", "Cd", "Dd", "x0", "Ky", "Fu")]
        for name, a in zip(("Ad", "Bd", "Cd", "Dd", "x0", "Ky", "Fu"), keep):
            setattr(c, name, a.ctypes.data_as(libfcb.c_f64p))
        if keep[5].shape != (bank.ny, self.ns) or keep[6].shape != (self.na, bank.nu):
            raise ValueError("Ky must be [ny, ns] and Fu [na, nu]")
        self._check(self.lib.fcb_set_controllers(self.h, C.byref(c)), "fcb_set_controllers")
        self._bank = bank

    # -- stepping -----------------------------------------------------------
    def step(self, u_ctrl) -> np.ndarray:
        """One step for all trajectories with HOST buffers: u_ctrl [na, B] -> y_meas [ns, B]."""
        uc = self._host(u_ctrl, self.na) if self.na else None
        rc = self.lib.fcb_step(
            self.h, libfcb.as_voidp(uc), libfcb.as_voidp(self.y_meas), libfcb.as_voidp(self.dE),
            libfcb.as_voidp(self.diverged),
        )
        self._check(rc, "fcb_step")
        return self.y_meas

    def step_device(self, u_ctrl_dev, y_dev=None, dE_dev=None, div_dev=None) -> None:
        """Same entry point with device-resident buffers (torch tensors / raw pointers)."""
        rc = self.lib.fcb_step(
            self.h, libfcb.as_voidp(u_ctrl_dev), libfcb.as_voidp(y_dev), libfcb.as_voidp(dE_dev), libfcb.as_voidp(div_dev)
        )
        self._check(rc, "fcb_step")

    def run_closed_loop(self, nsteps: int, log: bool = True, out=None):
        """nsteps of controller -> step -> log on the device. Returns series [nsteps, ncol, B]
        with columns (dE, u_ctrl_*, y_meas_*) or None when log=False."""
        series = None
        if log:
            series = out if out is not None else np.empty((nsteps, 1 + self.na + self.ns, self.B))
        rc = self.lib.fcb_run_closed_loop(self.h, int(nsteps), libfcb.as_voidp(series))
        self._check(rc, "fcb_run_closed_loop")
        return series

    def run_open_loop(self, u_series, log: bool = True, out=None):
        """Open-loop run on the device: ``u_series`` [nsteps, na, B] (numpy, or a torch CUDA tensor / device pointer with
        ``nsteps`` given by its shape) holds every step's control inputs.  Returns the series like run_closed_loop."""
        nsteps = int(u_series.shape[0])
        if isinstance(u_series, np.ndarray):
            u_series = np.ascontiguousarray(u_series, dtype=np.float64)
            if u_series.shape != (nsteps, self.na, self.B):
                raise ValueError(f"expected u_series of shape ({nsteps}, {self.na}, {self.B}), got {u_series.shape}")
        series = None
        if log:
            series = out if out is not None else np.empty((nsteps, 1 + self.na + self.ns, self.B))
        rc = self.lib.fcb_run_open_loop(self.h, nsteps, libfcb.as_voidp(u_series), libfcb.as_voidp(series))
        self._check(rc, "fcb_run_open_loop")
        return series

    def set_controller_state(self, x=None) -> None:
        """Controller states [nx, B] (None = zeros): Controller.reset / ensemble restart."""
        xs = None if x is None else self._host(x, self._bank.nx)
        self._check(self.lib.fcb_set_controller_state(self.h, libfcb.as_voidp(xs)), "fcb_set_controller_state")

    def costs(self, Tnorm: float = 1.0) -> dict:
        """Per-trajectory costs of the last run_closed_loop, summed on the device (utils/optim.py:231-288):
        ``energy_integral`` = compute_signal_cost(dE, Tnorm, 'integral'), ``energy_terminal`` = (…, 'terminal'),
        ``control`` = compute_control_cost(u_ctrl, Tnorm).  Diverged trajectories carry inf/nan."""
        out = np.empty((3, self.B))
        self._check(self.lib.fcb_get_costs(self.h, libfcb.as_voidp(out)), "fcb_get_costs")
        return {"energy_integral": out[0] * Tnorm, "control": out[1] * Tnorm, "energy_terminal": out[2].copy()}

    def fields(self, which: int = 0) -> np.ndarray:
        rows = self.N if which == 0 else self.Nv
        out = np.empty((rows, self.B))
        self._check(self.lib.fcb_get_fields(self.h, int(which), libfcb.as_voidp(out)), "fcb_get_fields")
        return out

    def controller_state(self) -> np.ndarray:
        out = np.empty((self._bank.nx, self.B))
        self._check(self.lib.fcb_get_controller_state(self.h, libfcb.as_voidp(out)), "fcb_get_controller_state")
        return out

    def profile_step(self, u_ctrl) -> dict:
        uc = self._host(u_ctrl, self.na) if self.na else None
        ms = (C.c_float * libfcb.FCB_NPHASES)()
        nl = (C.c_int32 * libfcb.FCB_NPHASES)()
        self._check(self.lib.fcb_profile_step(self.h, libfcb.as_voidp(uc), ms, nl), "fcb_profile_step")
        return {n: {"ms": float(ms[i]), "launches": int(nl[i])} for i, n in enumerate(libfcb.PHASE_NAMES)}

    def launch_count(self) -> int:
        return int(self.lib.fcb_launch_count(self.h))

    def synchronize(self) -> None:
        self._check(self.lib.fcb_synchronize(self.h), "fcb_synchronize")

    @property
    def stream(self) -> int:
        return int(self.lib.fcb_stream(self.h) or 0)
