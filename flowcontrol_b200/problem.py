"""Host-side setup: everything constant over a run, ready for upload to the GPU.

Replaces the one-time work the reference does through dolfin in
``FlowSolver._setup`` / ``_prepare_systems``
(/root/reference/src/flowcontrol/flowsolver.py:169-201, 665-701): Dirichlet dof
sets and actuator profiles, the BDF1/BDF2 left-hand sides with symmetric
Dirichlet elimination, their factorisation, the lifting/force vectors that make
the per-step RHS linear in ``u_ctrl``, and the sensor rows.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Sequence

import numpy as np
import scipy.sparse as sp

from .actuator import ACTUATOR_TYPE, Actuator
from .fem import ScalarBlocks
from .mesh import TaylorHoodTables
from .multifrontal import BlockFactor, SolvePlan, SymbolicFactor, build_plan
from .sensor import Sensor


@dataclass
class DirichletBC:
    """One velocity Dirichlet condition (``dolfin.DirichletBC`` on ``W.sub(0)`` or
    ``W.sub(0).sub(c)``): a boundary predicate ``inside(x, y) -> bool``, the constrained
    components, and either constant values or the actuator whose profile is imposed."""

    inside: Callable
    components: tuple
    value: object  # tuple of floats (one per component) or an Actuator


class DirichletSet:
    """Union of an ordered BC list; later entries override earlier ones on shared
    dofs (dolfin behaviour)."""

    def __init__(self, tab: TaylorHoodTables, bcs: Sequence[DirichletBC], actuators: Sequence[Actuator],
                 extra_zero_dofs: Sequence[int] = ()):
        N = tab.N
        marked = np.zeros(N, dtype=bool)
        const = np.zeros(N)
        act_id = np.full(N, -1, dtype=np.int64)
        act_val = np.zeros(N)
        for bc in bcs:
            nodes = tab.facet_p2_nodes(tab.mark_boundary_facets(bc.inside))
            if nodes.size == 0:
                continue
            x, y = tab.node_xy[nodes, 0], tab.node_xy[nodes, 1]
            if isinstance(bc.value, Actuator):
                k = next(i for i, a in enumerate(actuators) if a is bc.value)
                sx, sy = bc.value.shape(x, y)
                prof = (sx, sy)
                for c in bc.components:
                    d = nodes + c * tab.nN
                    marked[d] = True
                    const[d] = 0.0
                    act_id[d] = k
                    act_val[d] = prof[c]
            else:
                vals = tuple(bc.value)
                for ci, c in enumerate(bc.components):
                    d = nodes + c * tab.nN
                    marked[d] = True
                    const[d] = float(vals[ci])
                    act_id[d] = -1
                    act_val[d] = 0.0
        for d in extra_zero_dofs:
            marked[d] = True
            const[d] = 0.0
            act_id[d] = -1
        self.dofs = np.flatnonzero(marked).astype(np.int64)
        self.const = const[self.dofs]
        na = len(actuators)
        self.shape = np.zeros((na, len(self.dofs)))
        has = act_id[self.dofs] >= 0
        self.shape[act_id[self.dofs][has], np.flatnonzero(has)] = act_val[self.dofs][has]
        self.free = ~marked

    def values(self, u_ctrl) -> np.ndarray:
        return self.const + np.asarray(u_ctrl, dtype=np.float64) @ self.shape


def sensor_matrix(tab: TaylorHoodTables, sensors: Sequence[Sensor]):
    ptr = [0]
    idx, val = [], []
    for s in sensors:
        i, v = s.row(tab)
        s._row_cache = (i, v)
        idx.append(np.asarray(i, dtype=np.int64))
        val.append(np.asarray(v, dtype=np.float64))
        ptr.append(ptr[-1] + len(i))
    cat = (lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt))
    return np.array(ptr, dtype=np.int32), cat(idx, np.int32), cat(val, np.float64)


class FlowProblem:
    """Constant data of one (mesh, Re, dt, BCs, actuators, sensors, base flow) setup."""

    def __init__(
        self,
        tab: TaylorHoodTables,
        blocks: ScalarBlocks,
        Re: float,
        dt: float,
        bcs: Sequence[DirichletBC],
        actuators: Sequence[Actuator],
        sensors: Sequence[Sensor],
        U0: np.ndarray,
        nonlinear: bool = True,
        shift: float = 0.0,
        pin_pressure: bool = False,
        leaf_cells: int = 16,
        top_levels: int = 2,
        cluster_rows: int | None = None,
        cluster_height: int | None = None,
        amalgamate_above: int | None = None,
        balanced: bool | None = None,
        presum_height: int | None = None,
        factor_device: int | None = None,
        time_scheme: str = "bdf",
        symbolic: SymbolicFactor | None = None,
    ):
        self.tab, self.blocks = tab, blocks
        self.Re, self.dt, self.nonlinear = float(Re), float(dt), bool(nonlinear)
        if time_scheme not in ("bdf", "cn"):
            raise ValueError(f"time_scheme must be 'bdf' or 'cn', got {time_scheme!r}")
        self.time_scheme = time_scheme
        self.actuators, self.sensors = list(actuators), list(sensors)
        na = len(self.actuators)
        # an enclosed flow (all-Dirichlet velocity) has a constant-pressure null space: pin one dof
        extra = [tab.Nv] if pin_pressure else []
        self.dirichlet = DirichletSet(tab, bcs, self.actuators, extra_zero_dofs=extra)
        dset = self.dirichlet
        # amalgamated upper tree (ordering.amalgamate): fronts whose children have height >= amalgamate_above absorb them, which
        # halves the dependent levels above; the merged root then IS the old two-level top, so only one level is inverted densely
        import os

        amal = int(os.environ.get("FCB_AMALGAMATE", 0 if amalgamate_above is None else amalgamate_above))
        # depth-bounded dissection (ordering.dissect): no levels below the depth of an even dissection.  On by default: the
        # cylinder's tree loses 2 of 12 levels = 4 of 23 sweep launches for 10 % more factor entries (measured on B200, 256
        # trajectories: 0.938 -> 0.923 ms per step); FCB_BALANCED=0 restores the free dissection
        bal = bool(int(os.environ.get("FCB_BALANCED", 1 if balanced is None else int(balanced))))
        self.sym = symbolic or SymbolicFactor(tab, dset.free, leaf_cells=leaf_cells, amalgamate_above=amal, balanced=bal)
        if getattr(self.sym, "amalgamate_above", 0) > 0:
            top_levels = min(top_levels, 1)
        if not np.array_equal(np.sort(self.sym.perm), np.flatnonzero(dset.free)):
            raise ValueError("symbolic factorisation was built for a different Dirichlet set")
        # force vectors F_k = blkdiag(M,M) shape_k  (FORCE actuators), zero for BC actuators
        force = np.zeros((na, tab.N))
        for k, a in enumerate(self.actuators):
            if a.actuator_type == ACTUATOR_TYPE.FORCE:
                if hasattr(a, "normalise"):
                    a.normalise(tab.node_xy, blocks.Mv)
                sx, sy = a.shape(tab.node_xy[:, 0], tab.node_xy[:, 1])
                force[k, : tab.Nv] = blocks.Mv @ np.concatenate([sx, sy])
        U0v = np.asarray(U0[: tab.Nv], dtype=np.float64)
        self.A_raw, self.factors, self.plans, self.ctrl_rhs = {}, {}, {}, {}
        G = sp.csr_matrix(
            (dset.shape.ravel(), (np.repeat(np.arange(na), len(dset.dofs)), np.tile(dset.dofs, na))),
            shape=(na, tab.N),
        ) if na else sp.csr_matrix((0, tab.N))
        self.E_cn = None
        self.ctrl_rhs_prev = np.zeros((na, self.sym.n))
        if time_scheme == "bdf":
            orders = ((1, 1.0 / dt), (2, 1.5 / dt))
        else:
            # Crank-Nicolson (nsforms.py:191-236): one self-starting system, used for both plan slots
            orders = ((2, 1.0 / dt),)
        for order, c in orders:
            A = blocks.saddle_point(c, Re, U0v, shift=shift, linearised=True)
            half_force = 1.0
            if time_scheme == "cn":
                # theta = 1/2: the linear velocity terms L = C + D + K/Re are half implicit, half explicit (operator E_cn on
                # u_n); mass, pressure and continuity are fully implicit; the body force is averaged over the step
                L = blocks.saddle_point(0.0, Re, U0v, shift=0.0, linearised=True).tocsr()[: tab.Nv, : tab.Nv]
                A = (A - 0.5 * sp.bmat([[L, None], [None, sp.csr_matrix((tab.nV, tab.nV))]], format="csr")).tocsr()
                self.E_cn = (blocks.Mv / dt - 0.5 * L).tocsr()
                self.E_cn.sort_indices()
                half_force = 0.5
                self.ctrl_rhs_prev = np.ascontiguousarray((0.5 * force)[:, self.sym.perm])
            self.A_raw[order] = A
            if factor_device is None:
                fac = BlockFactor(self.sym, A)
            else:  # numeric factorisation on the GPU (devfactor.py); the index maps are shared by the BDF1 / BDF2 matrices
                from .devfactor import DeviceBlockFactor

                fac = DeviceBlockFactor(self.sym, A, maps=getattr(self, "_front_maps", None), device=factor_device)
                self._front_maps = fac.maps
            self.factors[order] = fac
            # shared-memory subtree clusters of the GPU solve (k_cluster_sweep): OFF by default -- measured on B200 they cut
            # the sweeps' DRAM traffic (bottom four forward levels 500 -> 200 MB) but run slower than the pull-form launches
            # they replace; cluster_rows / FCB_CLUSTER_ROWS > 0 turns them on (512 rows fit one CTA per SM)
            import os

            crow = int(os.environ.get("FCB_CLUSTER_ROWS", 0 if cluster_rows is None else cluster_rows))
            chgt = int(os.environ.get("FCB_CLUSTER_HEIGHT", 6 if cluster_height is None else cluster_height))
            # in-place leaves (build_plan: leaf_inplace): on by default -- 138 MB less DRAM traffic per solve of the cylinder
            # at 256 trajectories, 0.915 -> 0.900 ms per step, bit-identical results; FCB_LEAF_INPLACE=0 turns it off
            self.plans[order] = build_plan(fac, top_levels=top_levels, cluster_rows=crow, cluster_height=chgt,
                                           presum_height=int(os.environ.get("FCB_PRESUM", 0 if presum_height is None else presum_height)),
                                           leaf_inplace=bool(int(os.environ.get("FCB_LEAF_INPLACE", 1))))
            self._leaf_inplace = bool(int(os.environ.get("FCB_LEAF_INPLACE", 1)))
            self._plan_args = (top_levels, cluster_rows is None and "FCB_CLUSTER_ROWS" not in os.environ)
            # rhs contribution per unit u_ctrl_k in solver row order: (F_k - A[:,Gamma] shape_k)[perm]
            lift = (A @ G.T).toarray().T if na else np.zeros((0, tab.N))
            self.ctrl_rhs[order] = np.ascontiguousarray((half_force * force - lift)[:, self.sym.perm])
        if time_scheme == "cn":
            for d in (self.A_raw, self.factors, self.plans, self.ctrl_rhs):
                d[1] = d[2]
        self.sensor_ptr, self.sensor_idx, self.sensor_val = sensor_matrix(tab, self.sensors)

    def plans_for_batch(self, B: int) -> dict:
        """Solve plans for an ensemble of ``B`` trajectories.  Up to 32 trajectories (one 32-wide slab: the single-trajectory
        latency case, BASELINE configs[0]) every launch of the sweeps is pure latency, and sweeping the bottom four levels of
        the tree as shared-memory subtree clusters (one launch per sweep instead of four) is faster (measured on B200,
        cylinder B=1: 0.337 vs 0.359 ms per step); wider ensembles use the pull-form launches throughout.  An explicit
        ``cluster_rows`` / FCB_CLUSTER_ROWS setting is always respected."""
        top_levels, auto = self._plan_args
        if not auto or B > 32:
            return self.plans
        if getattr(self, "_plans_small", None) is None:
            built = {o: build_plan(self.factors[o], top_levels=top_levels, cluster_rows=512, cluster_height=3,
                                     leaf_inplace=self._leaf_inplace)
                     for o in sorted(set(self.factors))}
            self._plans_small = {o: built[o] for o in self.plans} if self.time_scheme != "cn" else {1: built[2], 2: built[2]}
        return self._plans_small

    @property
    def na(self) -> int:
        return len(self.actuators)

    @property
    def ns(self) -> int:
        return len(self.sensors)

    def host_step_rhs(self, order: int, u_n, u_nn, u_ctrl, u_ctrl_prev=None) -> np.ndarray:
        """Host restatement of the device right-hand side for one trajectory (setup/tests only)."""
        b = self.blocks
        if self.time_scheme == "cn":
            rv = self.E_cn @ u_n - (b.convection(u_n) if self.nonlinear else 0.0)
            full = np.concatenate([rv, np.zeros(self.tab.nV)])
            prev = np.zeros(self.na) if u_ctrl_prev is None else np.asarray(u_ctrl_prev, dtype=np.float64)
            return (full[self.sym.perm] + np.asarray(u_ctrl, dtype=np.float64) @ self.ctrl_rhs[2] + prev @ self.ctrl_rhs_prev)
        if order == 1:
            rv = b.Mv @ u_n / self.dt - (b.convection(u_n) if self.nonlinear else 0.0)
        else:
            rv = b.Mv @ (4 * u_n - u_nn) / (2 * self.dt)
            if self.nonlinear:
                rv = rv - 2 * b.convection(u_n) + b.convection(u_nn)
        full = np.concatenate([rv, np.zeros(self.tab.nV)])
        return full[self.sym.perm] + np.asarray(u_ctrl, dtype=np.float64) @ self.ctrl_rhs[order]
