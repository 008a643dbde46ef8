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
from ..sensor import SENSOR_TYPE, SensorForceCoefficient, SensorPoint

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

    def force_sensors(self) -> list:
        """[lift, drag] measurement rows of the cylinder (surfaces cylinder + actuator_up + actuator_lo,
        cylinderflowsolver.py:115-126); append them to ``params_control.sensor_list`` to log cl, cd every step."""
        subs = [self.get_subdomain(n).inside for n in ("cylinder", "actuator_up", "actuator_lo")]
        D = self.params_flow.user_data["D"]
        nu = self.params_flow.uinf * D / self.params_flow.Re

        def body(x, y):
            return subs[0](x, y) | subs[1](x, y) | subs[2](x, y)

        return [SensorForceCoefficient(sensor_type=SENSOR_TYPE.OTHER, inside=body, component=c, nu=nu,
                                       uinf=self.params_flow.uinf, D=D) for c in (1, 0)]

    def compute_force_coefficients(self, u, p) -> tuple[float, float]:
        """(cl, cd) of the fields (u, p) (numpy arrays or Field objects), cylinderflowsolver.py:115-126."""
        vec = np.concatenate([np.asarray(u.vector().get_local() if hasattr(u, "vector") else u),
                              np.asarray(p.vector().get_local() if hasattr(p, "vector") else p)])
        out = []
        for s in self.force_sensors():
            idx, val = s.row(self.tables)
            out.append(float(val @ vec[idx]))
        return out[0], out[1]

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
