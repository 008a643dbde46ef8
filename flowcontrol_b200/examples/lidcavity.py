"""Lid-driven cavity (Re=8000 proposed): uniform lid actuator, two probes.

Restates /root/reference/src/examples/lidcavity/lidcavityflowsolver.py:22-148.
The flow is enclosed (all-Dirichlet velocity), so the discrete pressure has a
constant null space; the reference leaves that to MUMPS, this build pins the
first pressure dof to zero.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np
import pandas as pd

from .. import flowsolverparameters as fsp
from ..actuator import ActuatorBCUniformU
from ..flowfield import BoundaryConditions
from ..flowsolver import FlowSolver, SubDomain, near
from ..problem import DirichletBC
from ..sensor import SENSOR_TYPE, SensorPoint

DATA = Path(__file__).resolve().parents[2] / "data" / "meshes"


class LidCavityFlowSolver(FlowSolver):
    def _make_boundaries(self) -> pd.DataFrame:
        ud = self.params_mesh.user_data
        subs = {
            "lid": lambda x, y: near(y, ud["yup"]),
            "leftwall": lambda x, y: near(x, ud["xle"]),
            "rightwall": lambda x, y: near(x, ud["xri"]),
            "bottomwall": lambda x, y: near(y, ud["ylo"]),
        }
        return pd.DataFrame(index=list(subs), data={"subdomain": [SubDomain(f) for f in subs.values()]})

    def _walls(self):
        sub = lambda n: self.get_subdomain(n).inside  # noqa: E731
        return [DirichletBC(sub(n), (0, 1), (0.0, 0.0)) for n in ("leftwall", "rightwall", "bottomwall")]

    def _make_bcs(self) -> BoundaryConditions:
        lid = DirichletBC(self.get_subdomain("lid").inside, (0, 1), self.params_control.actuator_list[0])
        return BoundaryConditions(bcu=[lid] + self._walls(), bcp=[])

    def _make_BCs(self) -> BoundaryConditions:
        lid = DirichletBC(self.get_subdomain("lid").inside, (0, 1), (self.params_flow.uinf, 0.0))
        return BoundaryConditions(bcu=[lid] + self._walls(), bcp=[])

    def _pin_pressure(self) -> bool:
        return True

    def _default_steady_state_initial_guess(self, x, y):
        return np.zeros_like(x), np.zeros_like(x)

    @classmethod
    def make_default(cls, Re: float = 8000, path_out=None, num_steps: int = 10, save_every: int = 0,
                     Tstart: float = 0.0, verbose: int = 0, meshpath=None, batch: int = 1, device: int = 0):
        path_out = Path(path_out) if path_out is not None else Path.cwd() / "data_output"
        params_flow = fsp.ParamFlow(Re=Re, uinf=1.0)
        params_flow.user_data["D"] = 1.0
        params_mesh = fsp.ParamMesh(meshpath=Path(meshpath) if meshpath else DATA / "lidcavity_mesh64.npz")
        params_mesh.user_data.update({"yup": 1, "ylo": 0, "xri": 1, "xle": 0})
        params_control = fsp.ParamControl(
            sensor_list=[
                SensorPoint(sensor_type=SENSOR_TYPE.V, position=np.array([0.05, 0.5])),
                SensorPoint(sensor_type=SENSOR_TYPE.U, position=np.array([0.5, 0.95])),
            ],
            actuator_list=[ActuatorBCUniformU(boundary_name="lid")],
        )
        return cls(
            params_flow=params_flow,
            params_time=fsp.ParamTime(num_steps=num_steps, dt=0.005, Tstart=Tstart),
            params_save=fsp.ParamSave(save_every=save_every, path_out=path_out),
            params_solver=fsp.ParamSolver(throw_error=True, is_eq_nonlinear=True, shift=0.0),
            params_mesh=params_mesh,
            params_control=params_control,
            params_ic=fsp.ParamIC(),
            verbose=verbose,
            params_ensemble=fsp.ParamEnsemble(batch=batch, device=device),
        )


def make_problem(Re: float = 1000.0, UP0=None, **kw):
    """FlowProblem of the default lid cavity for a given base flow (tests / smoke)."""
    import tempfile

    fs = LidCavityFlowSolver.make_default(Re=Re, path_out=Path(tempfile.gettempdir()) / "fcb200_lid", **kw)
    from ..flowfield import Field

    tab = fs.tables
    fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
    from ..problem import FlowProblem

    pe = fs.params_ensemble
    return FlowProblem(tab, fs.blocks, Re, fs.params_time.dt, fs.bc.bcu, fs.params_control.actuator_list,
                       fs.params_control.sensor_list, UP0, pin_pressure=True, leaf_cells=pe.leaf_cells, top_levels=pe.top_levels)
