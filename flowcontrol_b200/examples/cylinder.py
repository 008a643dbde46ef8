"""Flow past a cylinder (Re=100): two parabolic blowing/suction slots, three wake probes.

Restates /root/reference/src/examples/cylinder/cylinderflowsolver.py:17-186
(boundaries :20-88, perturbation BCs :90-108, defaults :128-186).
"""

from __future__ import annotations

from pathlib import Path

import numpy as np
import pandas as pd

from .. import flowsolverparameters as fsp
from ..actuator import ActuatorBCParabolicV
from ..flowfield import BoundaryConditions
from ..flowsolver import FlowSolver, SubDomain, between, near
from ..problem import DirichletBC
from ..sensor import SENSOR_TYPE, SensorPoint

DATA = Path(__file__).resolve().parents[2] / "data" / "meshes"


class CylinderFlowSolver(FlowSolver):
    def _make_boundaries(self) -> pd.DataFrame:
        ud = self.params_mesh.user_data
        xinfa, xinf, yinf = ud["xinfa"], ud["xinf"], ud["yinf"]
        r = self.params_flow.user_data["D"] / 2
        L = self.params_control.actuator_list[0].width

        def close(x, y):
            return between(x, -r, r) & between(y, -r, r)

        subs = {
            "inlet": lambda x, y: near(x, xinfa),
            "outlet": lambda x, y: near(x, xinf),
            "walls": lambda x, y: near(y, -yinf) | near(y, yinf),
            "cylinder": lambda x, y: close(x, y) & (between(x, -r, -L) | between(x, L, r)),
            "actuator_up": lambda x, y: close(x, y) & between(x, -L, L, 0.01) & between(y, 0.0, r),
            "actuator_lo": lambda x, y: close(x, y) & between(x, -L, L, 0.01) & between(y, -r, 0.0),
        }
        return pd.DataFrame(index=list(subs), data={"subdomain": [SubDomain(f) for f in subs.values()]})

    def _make_bcs(self) -> BoundaryConditions:
        acts = self.params_control.actuator_list
        sub = lambda n: self.get_subdomain(n).inside  # noqa: E731
        return BoundaryConditions(
            bcu=[
                DirichletBC(sub("inlet"), (0, 1), (0.0, 0.0)),
                DirichletBC(sub("walls"), (1,), (0.0,)),
                DirichletBC(sub("cylinder"), (0, 1), (0.0, 0.0)),
                DirichletBC(sub("actuator_up"), (0, 1), acts[0]),
                DirichletBC(sub("actuator_lo"), (0, 1), acts[1]),
            ],
            bcp=[],
        )

    @classmethod
    def make_default(cls, Re: float = 100, path_out=None, num_steps: int = 10, save_every: int = 0,
                     Tstart: float = 0.0, verbose: int = 0, meshpath=None, batch: int = 1, device: int = 0):
        path_out = Path(path_out) if path_out is not None else Path.cwd() / "data_output"
        params_flow = fsp.ParamFlow(Re=Re, uinf=1.0)
        params_flow.user_data["D"] = 1.0
        params_mesh = fsp.ParamMesh(meshpath=Path(meshpath) if meshpath else DATA / "cylinder_O1.npz")
        params_mesh.user_data.update({"xinf": 20, "xinfa": -10, "yinf": 10})
        width = ActuatorBCParabolicV.angular_size_deg_to_width(10, 0.5)
        params_control = fsp.ParamControl(
            sensor_list=[
                SensorPoint(sensor_type=SENSOR_TYPE.V, position=np.array([3.0, 0.0])),
                SensorPoint(sensor_type=SENSOR_TYPE.V, position=np.array([3.1, 1.0])),
                SensorPoint(sensor_type=SENSOR_TYPE.V, position=np.array([3.1, -1.0])),
            ],
            actuator_list=[
                ActuatorBCParabolicV(width=width, position_x=0.0, boundary_name="actuator_up"),
                ActuatorBCParabolicV(width=width, position_x=0.0, boundary_name="actuator_lo"),
            ],
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
