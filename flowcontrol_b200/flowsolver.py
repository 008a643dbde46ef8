"""FlowSolver facade: the reference's public API over the CUDA ensemble step.

Mirrors /root/reference/src/flowcontrol/flowsolver.py (constructor signature,
``compute_steady_state``, ``load_steady_state``, ``initialize_time_stepping``,
``step``, ``write_timeseries``, ``timeseries``, ``fields``, abstract hooks
``_make_boundaries`` / ``_make_bcs`` / ``make_default``).  What changes:

* setup is numpy/scipy (mesh.py, fem.py, steadystate.py) instead of dolfin;
* ``step()`` advances a whole ensemble of ``params_ensemble.batch`` trajectories on
  the GPU through ``fcb_step`` (include/fcb200.h).  With ``batch == 1`` (default) the
  signature and return value are exactly the reference's; with ``batch > 1``
  ``u_ctrl`` may be ``[B, na]`` and ``y_meas`` is ``[B, ns]``.

There is no CPU fallback for the step: without libfcb200.so / a GPU,
``initialize_time_stepping`` raises.
"""

from __future__ import annotations

import json
import logging
import time
from abc import ABC, abstractmethod
from pathlib import Path
from typing import Callable, Iterable, Sequence

import numpy as np
import pandas as pd

from . import flowsolverparameters as fsp
from .actuator import ACTUATOR_TYPE
from .ensemble import Ensemble
from .exporter import FlowExporter, read_checkpoint, write_checkpoint
from .fem import ScalarBlocks
from .flowfield import BoundaryConditions, Field, FlowField, FlowFieldCollection, SimPaths
from .mesh import TaylorHoodTables
from .problem import DirichletBC, DirichletSet, FlowProblem
from .steadystate import SteadyStateSolver

logger = logging.getLogger(__name__)

DOLFIN_EPS = 3.0e-16


class SubDomain:
    """Boundary region given by a vectorised predicate ``inside(x, y) -> bool``
    (the reference compiles C++ strings into ``dolfin.CompiledSubDomain``;
    ``on_boundary`` is implied — only exterior facets are ever tested)."""

    def __init__(self, inside: Callable):
        self.inside = inside


def near(a, b, tol: float = DOLFIN_EPS):
    return (a >= b - tol) & (a <= b + tol)


def between(a, lo, hi, tol: float = 0.0):
    return (a >= lo - tol) & (a <= hi + tol)


class HostLUSolver:
    """``dolfin.LUSolver``-shaped view of the factorised LHS of one time-stepping order: ``solve(x, b)`` with ``b`` the
    assembled right-hand side in canonical numbering (Dirichlet rows holding the boundary values, lifting already
    applied, as ``SystemAssembler`` leaves it) fills ``x``: ``x[free] = A_ff^-1 b[free]``, ``x[bc] = b[bc]``."""

    def __init__(self, problem: FlowProblem, order=2):
        self.problem = problem
        self.order = 2 if order == "cn" else int(order)

    def set_operator(self, A=None) -> None:
        """The operator is fixed at setup (flowsolver.py:697); kept for call compatibility."""

    def solve(self, x, b) -> int:
        prob = self.problem
        b = np.asarray(b, dtype=np.float64)
        x = np.asarray(x)
        x[prob.sym.perm] = prob.factors[self.order].solve(b[prob.sym.perm])
        x[prob.dirichlet.dofs] = b[prob.dirichlet.dofs]
        return 1


class FlowSolver(ABC):
    def __init__(
        self,
        params_flow: fsp.ParamFlow,
        params_time: fsp.ParamTime,
        params_save: fsp.ParamSave,
        params_solver: fsp.ParamSolver,
        params_mesh: fsp.ParamMesh,
        params_control: fsp.ParamControl,
        params_ic: fsp.ParamIC,
        params_restart: fsp.ParamRestart | None = None,
        verbose: int = 1,
        params_ensemble: fsp.ParamEnsemble | None = None,
    ) -> None:
        self._validate_params(params_flow, params_time, params_save, params_solver, params_mesh, params_control,
                              params_ic, params_restart)
        self.params_flow, self.params_time, self.params_save = params_flow, params_time, params_save
        self.params_solver, self.params_mesh, self.params_restart = params_solver, params_mesh, params_restart
        self.params_control, self.params_ic = params_control, params_ic
        self.params_ensemble = params_ensemble or fsp.ParamEnsemble()
        self.verbose = verbose
        self._setup()

    # validation rules of flowsolver.py:108-165
    @staticmethod
    def _validate_params(params_flow, params_time, params_save, params_solver, params_mesh, params_control,
                         params_ic, params_restart=None) -> None:
        if params_time.dt <= 0:
            raise ValueError(f"dt must be positive, got {params_time.dt}")
        if params_time.num_steps < 0:
            raise ValueError(f"num_steps must be non-negative, got {params_time.num_steps}")
        if params_flow.Re <= 0:
            raise ValueError(f"Re must be positive, got {params_flow.Re}")
        if params_save.save_every < 0:
            raise ValueError(f"save_every must be non-negative, got {params_save.save_every}")
        if params_save.energy_every < 0:
            raise ValueError(f"energy_every must be non-negative, got {params_save.energy_every}")
        if len(params_control.actuator_list) != params_control.actuator_number:
            raise ValueError("actuator_list length does not match actuator_number")
        if len(params_control.sensor_list) != params_control.sensor_number:
            raise ValueError("sensor_list length does not match sensor_number")
        if not Path(params_mesh.meshpath).exists():
            raise FileNotFoundError(f"Mesh file not found at {params_mesh.meshpath}")
        if params_restart is not None and params_restart.Trestartfrom < 0:
            raise ValueError(f"Trestartfrom must be non-negative, got {params_restart.Trestartfrom}")
        if params_solver.time_scheme not in ("bdf", "cn"):
            raise ValueError(f"time_scheme must be 'bdf' or 'cn', got {params_solver.time_scheme!r}")

    # ── setup (flowsolver.py:169-201) ─────────────────────────────────────────
    def _setup(self) -> None:
        self.fields = FlowFieldCollection()
        self.E0 = 0.0
        self.paths = self._define_paths()
        self.tables = self.mesh = TaylorHoodTables.from_file(self.params_mesh.meshpath)
        self.blocks = ScalarBlocks(self.tables)
        self.boundaries = self._make_boundaries()
        self.boundaries["idx"] = list(range(len(self.boundaries)))
        for actuator in self.params_control.actuator_list:
            actuator.load_expression(self)
            if actuator.actuator_type == ACTUATOR_TYPE.FORCE and hasattr(actuator, "normalise"):
                actuator.normalise(self.tables.node_xy, self.blocks.Mv)
        for sensor in self.params_control.sensor_list:
            sensor.load(self)
        self.bc = self._make_bcs()
        self.exporter = FlowExporter(self.paths, self.fields, Tstart=self.params_time.Tstart,
                                     dt=self.params_time.dt, save_every=self.params_save.save_every, tab=self.tables)
        self.problem: FlowProblem | None = None
        self.ensemble: Ensemble | None = None
        self.first_step = True

    def _define_paths(self) -> SimPaths:
        def ext(T: float) -> str:
            return f"_restart{T:.3f}".replace(".", ",")

        Tstart = self.params_time.Tstart
        Trestartfrom = self.params_restart.Trestartfrom if self.params_restart else 0.0
        out = Path(self.params_save.path_out)
        return SimPaths(
            U0=out / "steady" / "U0.xdmf", P0=out / "steady" / "P0.xdmf", steady_meta=out / "steady" / "meta.json",
            U=out / ("U" + ext(Trestartfrom) + ".xdmf"), P=out / ("P" + ext(Trestartfrom) + ".xdmf"),
            Uprev=out / ("Uprev" + ext(Trestartfrom) + ".xdmf"),
            U_restart=out / ("U" + ext(Tstart) + ".xdmf"), Uprev_restart=out / ("Uprev" + ext(Tstart) + ".xdmf"),
            P_restart=out / ("P" + ext(Tstart) + ".xdmf"),
            timeseries=out / ("timeseries1D" + ext(Tstart) + ".csv"), metadata=out / ("meta" + ext(Tstart) + ".json"),
            mesh=Path(self.params_mesh.meshpath),
        )

    def get_subdomain(self, name: str) -> SubDomain:
        return self.boundaries.loc[name].subdomain

    # ── actuators / sensors (flowsolver.py:278-325) ───────────────────────────
    def set_actuators_u_ctrl(self, u_ctrl: Iterable) -> None:
        u_ctrl = list(u_ctrl)
        if len(u_ctrl) != self.params_control.actuator_number:
            raise ValueError(f"Expected {self.params_control.actuator_number} control inputs, got {len(u_ctrl)}")
        for actuator, val in zip(self.params_control.actuator_list, u_ctrl):
            actuator.expression.u_ctrl = val

    def flush_actuators_u_ctrl(self) -> None:
        self.set_actuators_u_ctrl([0] * self.params_control.actuator_number)

    def get_actuators_u_ctrl(self) -> list:
        return [a.expression.u_ctrl for a in self.params_control.actuator_list]

    def make_measurement(self, up) -> np.ndarray:
        return np.array([sensor.eval(up=up) for sensor in self.params_control.sensor_list])

    # ── boundary conditions ───────────────────────────────────────────────────
    def _make_BCs(self) -> BoundaryConditions:
        """Full-field BCs: uniform inlet profile + the perturbation BCs (flowsolver.py:329-337)."""
        inlet = DirichletBC(self.get_subdomain("inlet").inside, (0, 1), (self.params_flow.uinf, 0.0))
        bcs = self._make_bcs()
        return BoundaryConditions(bcu=[inlet] + bcs.bcu[1:], bcp=[])

    def _pin_pressure(self) -> bool:
        """True for enclosed flows (no natural outlet): the pressure level is then fixed at one dof."""
        return False

    # ── steady state (flowsolver.py:341-460) ──────────────────────────────────
    def _force_vector(self, u_ctrl) -> np.ndarray:
        tab = self.tables
        f = np.zeros(tab.Nv)
        for a, amp in zip(self.params_control.actuator_list, u_ctrl):
            if a.actuator_type == ACTUATOR_TYPE.FORCE and amp != 0.0:
                sx, sy = a.shape(tab.node_xy[:, 0], tab.node_xy[:, 1])
                f += amp * (self.blocks.Mv @ np.concatenate([sx, sy]))
        return f

    def compute_steady_state(self, u_ctrl: list, method: str = "newton", initial_guess: Field | None = None,
                             max_iter: int = 10, assembly: str = "host", factor: str = "host", **kwargs) -> None:
        """``assembly="device"`` assembles the Newton / Picard matrices on the GPU (assembly.py, fcb_assemble_advection);
        ``factor="device"`` factorises every iterate on the GPU (devfactor.py, fcb_factorize)."""
        self.set_actuators_u_ctrl(u_ctrl)
        tab = self.tables
        extra = [tab.Nv] if self._pin_pressure() else []
        dset = DirichletSet(tab, self._make_BCs().bcu, self.params_control.actuator_list, extra_zero_dofs=extra)
        if assembly not in ("host", "device"):
            raise ValueError(f"assembly must be 'host' or 'device', got {assembly!r}")
        assembler = None
        if assembly == "device":
            from .assembly import DeviceAdvectionAssembler

            assembler = DeviceAdvectionAssembler(tab, self.blocks, device=self.params_ensemble.device)
        ss = SteadyStateSolver(tab, self.blocks, self.params_flow.Re, dset, force=self._force_vector(u_ctrl),
                               verbose=bool(self.verbose), assembler=assembler, factor=factor, device=self.params_ensemble.device,
                               leaf_cells=self.params_ensemble.leaf_cells)
        UP = self._define_initial_guess(initial_guess)
        if method == "newton":
            UP = ss.newton(UP, u_ctrl, max_iter=max_iter, **kwargs)
        elif method == "picard":
            UP = ss.picard(UP, u_ctrl, max_iter=max_iter, **kwargs)
        else:
            raise ValueError(f"method must be 'newton' or 'picard', got {method!r}")
        U0, P0 = Field(UP[: tab.Nv]), Field(UP[tab.Nv :])
        if self.params_save.save_every:
            write_checkpoint(self.paths.U0, "U0", U0.array, 0.0, append=False, tab=tab)
            write_checkpoint(self.paths.P0, "P0", P0.array, 0.0, append=False, tab=tab)
            self.paths.steady_meta.parent.mkdir(parents=True, exist_ok=True)
            self.paths.steady_meta.write_text(json.dumps({"mesh_cells": int(tab.nT)}, indent=2))
        self._assign_steady_state(U0, P0)

    def load_steady_state(self, path_u_p: Sequence[Path] | None = None) -> None:
        paths = path_u_p or (self.paths.U0, self.paths.P0)
        self._check_steady_state_compatible(Path(paths[0]))
        self._assign_steady_state(Field(read_checkpoint(Path(paths[0]), tab=self.tables)),
                                  Field(read_checkpoint(Path(paths[1]), tab=self.tables)))

    def _check_steady_state_compatible(self, u0_path: Path) -> None:
        try:
            meta = json.loads((u0_path.parent / "meta.json").read_text())
        except FileNotFoundError:
            meta = {}
        stored = meta.get("mesh_cells")
        if stored is not None and stored != self.tables.nT:
            raise ValueError(
                f"Steady-state checkpoint at {u0_path.parent} was written with {stored} mesh cells, but the "
                f"current mesh has {self.tables.nT}. Load a checkpoint from the same mesh, or recompute the steady state."
            )

    def _assign_steady_state(self, U0: Field, P0: Field) -> None:
        self.fields.U0, self.fields.P0 = U0, P0
        self.fields.UP0 = self.merge(U0, P0)
        self.E0 = 0.5 * float(U0.array @ (self.blocks.Mv @ U0.array))
        self.problem = None  # LHS depends on the base flow
        if self.ensemble is not None:
            self.ensemble.close()
            self.ensemble = None

    def _define_initial_guess(self, initial_guess=None) -> np.ndarray:
        if initial_guess is not None:
            return np.array(initial_guess.array if isinstance(initial_guess, Field) else initial_guess, dtype=np.float64)
        tab = self.tables
        ux, uy = self._default_steady_state_initial_guess(tab.node_xy[:, 0], tab.node_xy[:, 1])
        return np.concatenate([ux, uy, np.zeros(tab.nV)])

    def _default_steady_state_initial_guess(self, x, y):
        """Uniform flow at uinf (flowsolver.py:887-900)."""
        return np.full_like(x, self.params_flow.uinf), np.zeros_like(x)

    # ── time stepping ─────────────────────────────────────────────────────────
    def _make_problem(self, factor_device: int | None = None) -> FlowProblem:
        """Setup of the constant operators (the reference's _prepare_systems, flowsolver.py:665-701).  ``factor_device``: GPU
        that does the numeric factorisation (the time-stepping path, which needs the GPU anyway); None = host (the
        ``_make_solver`` hook and other host-side tooling)."""
        pe = self.params_ensemble
        return FlowProblem(
            self.tables, self.blocks, self.params_flow.Re, self.params_time.dt, self.bc.bcu,
            self.params_control.actuator_list, self.params_control.sensor_list, self.fields.UP0.array,
            nonlinear=self.params_solver.is_eq_nonlinear, shift=self.params_solver.shift,
            pin_pressure=self._pin_pressure(), leaf_cells=pe.leaf_cells, top_levels=pe.top_levels, factor_device=factor_device,
            time_scheme=self.params_solver.time_scheme,
        )

    def _build_problem(self) -> None:
        pe = self.params_ensemble
        if self.problem is None:
            self.problem = self._make_problem(factor_device=pe.device)
        if self.ensemble is not None:
            self.ensemble.close()
        self.ensemble = Ensemble(self.problem, pe.batch, pe.device)

    def _make_solver(self, order=2) -> "HostLUSolver":
        """The reference's linear-solver hook (flowsolver.py:812-814; docs/numerical-details.md:44-48): an object with
        ``set_operator(A)`` and ``solve(x, b)``.  On this path the factor of the constant LHS lives on the device and
        ``step()`` never calls the hook; the object returned here solves the same BC-applied system on the host with the
        same multifrontal factor (setup checks, operator tooling).  Overriding it does not redirect the device solve."""
        if self.problem is None:
            self.problem = self._make_problem()
        return HostLUSolver(self.problem, "cn" if self.params_solver.time_scheme == "cn" else order)

    def initialize_time_stepping(self, Tstart: float = 0.0, ic=None) -> None:
        """ic: None, a Field / array [N] shared by all trajectories, or an array [N, B]."""
        if self.fields.UP0 is None:
            raise RuntimeError("compute_steady_state or load_steady_state must be called first")
        if self.problem is None or self.ensemble is None:
            self._build_problem()
        if Tstart == 0.0:
            up_ic, u_n, u_nn, order = self._initialize_with_ic(ic)
        else:
            up_ic, u_n, u_nn, order = self._initialize_at_time(Tstart)
        tab = self.tables
        cn = self.params_solver.time_scheme == "cn"
        if cn:
            order = "cn"  # self-starting (flowsolver.py:513)
        self.order = order
        self.iter = 0
        self.t = Tstart if Tstart else self.params_time.Tstart
        self.ensemble.set_state(u_n, u_nn, up_ic[tab.Nv :], order=order)
        self.fields.ic = FlowField(up=Field(up_ic if up_ic.ndim == 1 else up_ic[:, 0], Nv=tab.Nv))
        self._refresh_fields()
        self.first_step = True
        self.exporter.reset()
        self.y_meas = self._squeeze(self.ensemble.y_meas)
        self.exporter.log_ic(t=self.params_time.Tstart, y_meas=self._traj0(self.ensemble.y_meas),
                             dE=float(self.ensemble.dE[0]))

    def _initialize_with_ic(self, ic):
        """Zero or user IC plus amplitude x divergence-free Gaussian (flowsolver.py:502-549)."""
        tab = self.tables
        if ic is None:
            up = np.zeros(tab.N)
        else:
            up = np.array(ic.array if isinstance(ic, Field) else ic, dtype=np.float64)
        if self.params_ic.amplitude:
            pert = self._default_initial_perturbation(self.params_ic.xloc, self.params_ic.yloc, self.params_ic.radius)
            up = up + self.params_ic.amplitude * (pert if up.ndim == 1 else pert[:, None])
        u_n = up[: tab.Nv]
        if self.params_save.save_every:
            B = self.params_ensemble.batch
            if B == 1:
                self.exporter.export_xdmf(Field(up[: tab.Nv]), Field(up[: tab.Nv]), Field(up[tab.Nv :]), time=0.0,
                                          append=False, write_mesh=True, adjust_baseflow=1.0)
            else:  # every trajectory gets its own function from the first snapshot on (the counters stay aligned)
                up_all = up if up.ndim == 2 else np.repeat(up[:, None], B, axis=1)
                self.exporter.export_xdmf(up_all[: tab.Nv], up_all[: tab.Nv], up_all[tab.Nv :], time=0.0,
                                          append=False, write_mesh=True, adjust_baseflow=1.0)
        return up, u_n, u_n, 1

    def _find_restart_from_json(self, Tstart: float):
        path_out = Path(self.params_save.path_out)
        for json_path in sorted(path_out.glob("meta_restart*.json")):
            meta = json.loads(json_path.read_text())
            T0 = meta["Tstart"]
            step = meta["dt"] * meta["save_every"]
            n = meta["checkpoints_written"]
            if n == 0:
                continue
            if T0 - 1e-10 <= Tstart <= T0 + step * n + 1e-10:
                return meta, round((Tstart - T0) / step), path_out
        return None

    def _find_restart_from_params(self, Tstart: float):
        if self.params_restart is None:
            raise FileNotFoundError(
                f"No JSON metadata sidecar found covering Tstart={Tstart} in {self.params_save.path_out}, "
                "and no ParamRestart was provided."
            )
        pr = self.params_restart
        counter = round((Tstart - pr.Trestartfrom) / (pr.dt_old * pr.save_every_old))
        meta = {"restart_order": pr.restart_order,
                "files": {"U": self.paths.U.name, "Uprev": self.paths.Uprev.name, "P": self.paths.P.name}}
        return meta, counter, Path(self.params_save.path_out)

    def _initialize_at_time(self, Tstart: float):
        """Restart from a checkpoint: read full fields, subtract the base flow (flowsolver.py:599-663)."""
        found = self._find_restart_from_json(Tstart) or self._find_restart_from_params(Tstart)
        meta, counter, base = found
        B = self.params_ensemble.batch  # an ensemble restarts trajectory by trajectory (files of a single run are broadcast)
        U = read_checkpoint(base / meta["files"]["U"], counter, tab=self.tables, batch=B)
        Uprev = read_checkpoint(base / meta["files"]["Uprev"], counter, tab=self.tables, batch=B)
        P = read_checkpoint(base / meta["files"]["P"], counter, tab=self.tables, batch=B)
        if self.params_save.save_every:
            self.exporter.export_xdmf(Field(U) if B == 1 else U, Field(Uprev) if B == 1 else Uprev, Field(P) if B == 1 else P,
                                      time=Tstart, append=False, write_mesh=True, adjust_baseflow=0.0)
        U0v, P0v = self.fields.U0.array, self.fields.P0.array
        if B > 1:
            U0v, P0v = U0v[:, None], P0v[:, None]
        u_n, u_nn, p_n = U - U0v, Uprev - U0v, P - P0v
        return np.concatenate([u_n, p_n]), u_n, u_nn, meta["restart_order"]

    def _traj0(self, a: np.ndarray) -> np.ndarray:
        return np.array(a[:, 0], copy=True)

    def _squeeze(self, a: np.ndarray) -> np.ndarray:
        return self._traj0(a) if self.params_ensemble.batch == 1 else np.array(a.T, copy=True)

    def _refresh_fields(self) -> None:
        """Lazy views of the device state.  ``field.vector().get_local()`` / ``field.array`` is trajectory 0 (the
        reference's single-run accessors, flowfield.py:62-97); ``field.ensemble`` is the [n, B] array of EVERY trajectory."""
        tab, ens = self.tables, self.ensemble
        cache: dict = {}

        def cur_all():
            if "cur" not in cache:
                cache["cur"] = ens.fields(0)
            return cache["cur"]

        def prev_all():
            if "prev" not in cache:
                cache["prev"] = ens.fields(1)
            return cache["prev"]

        def view(get_all, lo=None, hi=None, Nv=None):
            return Field(fetch=lambda: get_all()[lo:hi, 0], fetch_all=lambda: get_all()[lo:hi], Nv=Nv)

        f = self.fields
        f.up_ = view(cur_all, Nv=tab.Nv)
        f.u_ = view(cur_all, 0, tab.Nv)
        f.p_ = view(cur_all, tab.Nv, None)
        f.u_n = view(cur_all, 0, tab.Nv)
        f.p_n = view(cur_all, tab.Nv, None)
        f.u_nn = view(prev_all)

    def step(self, u_ctrl) -> np.ndarray | None:
        """Advance every trajectory by one step (flowsolver.py:703-799)."""
        B, na = self.params_ensemble.batch, self.params_control.actuator_number
        uc = np.asarray(u_ctrl, dtype=np.float64)
        if uc.ndim <= 1:
            lst = list(np.atleast_1d(uc))
            if len(lst) != na:
                raise ValueError(f"Expected {na} control inputs, got {len(lst)}")
            self.set_actuators_u_ctrl(lst)
            uc_dev = np.repeat(np.asarray(lst, dtype=np.float64)[:, None], B, axis=1)
        else:
            if uc.shape != (B, na):
                raise ValueError(f"Expected u_ctrl of shape ({B}, {na}), got {uc.shape}")
            self.set_actuators_u_ctrl(list(uc[0]))
            uc_dev = np.ascontiguousarray(uc.T)
        self.first_step = False
        t0 = time.time()
        ens = self.ensemble
        ens.step(uc_dev)
        if ens.diverged.any():
            logger.critical("Solver diverged (Inf detected)")
            if B == 1 or ens.diverged.all():
                if not self.params_solver.throw_error:
                    return None
                raise RuntimeError("Failed solving: Inf found in solution")
        self.iter += 1
        self.t = self.params_time.Tstart + self.iter * self.params_time.dt
        if self.params_solver.time_scheme != "cn":
            self.order = 2
        self._refresh_fields()
        y = np.array(ens.y_meas, copy=True)
        if B > 1:
            y[:, ens.diverged != 0] = np.nan
        self.y_meas = self._squeeze(y)
        runtime = time.time() - t0
        if self._niter_multiple_of(self.iter, self.verbose):
            self.exporter.log_progress(self.iter, self.params_time.num_steps, self.t,
                                       self.params_time.Tfinal + self.params_time.Tstart, runtime)
        at_checkpoint = self._niter_multiple_of(self.iter, self.params_save.save_every)
        dE = float(ens.dE[0]) if self._niter_multiple_of(self.iter, self.params_save.energy_every) else np.nan
        self.exporter.log(u_ctrl=uc_dev[:, 0], y_meas=self._traj0(y), dE=dE, t=self.t, runtime=runtime)
        if at_checkpoint:
            if B == 1:
                self.exporter.export_xdmf(self.fields.u_n, self.fields.u_nn, self.fields.p_n, time=self.t, adjust_baseflow=1.0)
            else:  # every trajectory, so that the ensemble can be restarted trajectory by trajectory
                self.exporter.export_xdmf(self.fields.u_n.ensemble, self.fields.u_nn.ensemble, self.fields.p_n.ensemble,
                                          time=self.t, adjust_baseflow=1.0)
            self.exporter.write_metadata(restart_order="cn" if self.params_solver.time_scheme == "cn" else 2)
            self.exporter.write_timeseries()
        return self.y_meas

    def write_timeseries(self) -> None:
        self.exporter.write_timeseries()

    @property
    def timeseries(self) -> pd.DataFrame:
        return self.exporter.to_dataframe()

    def _niter_multiple_of(self, iter: int, divider: int) -> bool:
        return bool(divider and not iter % divider)

    def compute_perturbation_energy(self) -> float:
        return float(self.ensemble.dE[0])

    def merge(self, u: Field, p: Field) -> Field:
        return Field(np.concatenate([u.array, p.array]), Nv=self.tables.Nv)

    def _default_initial_perturbation(self, xloc: float = 0.0, yloc: float = 0.0, radius: float = 1.0) -> np.ndarray:
        """Nodal interpolant of (d psi/dy, -d psi/dx), psi = exp(-r^2/2 radius^2)/4, merged with the
        base-flow pressure (utils/physics.py:32-56, flowsolver.py:902-912)."""
        tab = self.tables
        out = np.zeros(tab.N)
        if radius <= 0:
            return out
        x, y = tab.node_xy[:, 0] - xloc, tab.node_xy[:, 1] - yloc
        psi = 0.25 * np.exp(-0.5 * (x * x + y * y) / radius**2)
        out[: tab.nN] = psi * (-y / radius**2)
        out[tab.nN : tab.Nv] = -psi * (-x / radius**2)
        out[tab.Nv :] = self.fields.P0.array
        return out

    # ── abstract hooks ────────────────────────────────────────────────────────
    @abstractmethod
    def _make_boundaries(self) -> pd.DataFrame:
        """DataFrame with a 'subdomain' column (SubDomain objects), boundary names as index."""

    @abstractmethod
    def _make_bcs(self) -> BoundaryConditions:
        """Perturbation-field BCs; the first entry of bcu must be the inlet BC."""

    @classmethod
    @abstractmethod
    def make_default(cls, **kwargs) -> "FlowSolver":
        """Instance with the standard parameters of the configuration."""
