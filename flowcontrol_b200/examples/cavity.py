"""Open cavity (Re=7500): Gaussian body-force actuator, wall-shear + point sensors.

Restates /root/reference/src/examples/cavity/cavityflowsolver.py:19-280.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np
import pandas as pd

from .. import flowsolverparameters as fsp
from ..actuator import ActuatorForceGaussianV
from ..flowfield import BoundaryConditions
from ..flowsolver import DOLFIN_EPS, FlowSolver, SubDomain, between, near
from ..problem import DirichletBC
from ..sensor import SENSOR_TYPE, SensorHorizontalWallShear, SensorPoint

DATA = Path(__file__).resolve().parents[2] / "data" / "meshes"


class CavityFlowSolver(FlowSolver):
    def _make_boundaries(self) -> pd.DataFrame:
        L, D = self.params_flow.user_data["L"], self.params_flow.user_data["D"]
        ud = self.params_mesh.user_data
        xinfa, xinf, yinf, x0l, x0r = ud["xinfa"], ud["xinf"], ud["yinf"], ud["x0ns_left"], ud["x0ns_right"]
        T = DOLFIN_EPS
        subs = {
            "inlet": lambda x, y: near(x, xinfa),
            "outlet": lambda x, y: near(x, xinf),
            "upper_wall": lambda x, y: near(y, yinf),
            "cavity_left": lambda x, y: near(x, 0.0) & between(y, -D, 0.0),
            "cavity_botm": lambda x, y: near(y, -D) & between(x, 0.0, L),
            "cavity_right": lambda x, y: near(x, L) & between(y, -D, 0.0),
            "lower_wall_left_sf": lambda x, y: (x >= xinfa) & (x <= x0l + 10 * T) & near(y, 0.0),
            "lower_wall_left_ns": lambda x, y: (x >= x0l - 10 * T) & (x <= 0.0) & near(y, 0.0),
            "lower_wall_right_ns": lambda x, y: near(y, 0.0) & between(x, L, x0r),
            "lower_wall_right_sf": lambda x, y: near(y, 0.0) & between(x, x0r, xinf),
        }
        return pd.DataFrame(index=list(subs), data={"subdomain": [SubDomain(f) for f in subs.values()]})

    def _make_bcs(self) -> BoundaryConditions:
        sub = lambda n: self.get_subdomain(n).inside  # noqa: E731
        noslip = lambda n: DirichletBC(sub(n), (0, 1), (0.0, 0.0))  # noqa: E731
        slip = lambda n: DirichletBC(sub(n), (1,), (0.0,))  # noqa: E731
        return BoundaryConditions(
            bcu=[
                noslip("inlet"), slip("upper_wall"), slip("lower_wall_left_sf"), noslip("lower_wall_left_ns"),
                noslip("lower_wall_right_ns"), slip("lower_wall_right_sf"), noslip("cavity_left"),
                noslip("cavity_botm"), noslip("cavity_right"),
            ],
            bcp=[],
        )

    def _default_steady_state_initial_guess(self, x, y):
        return np.where(y >= 0.0, 1.0, 0.0), np.zeros_like(x)

    @classmethod
    def make_default(cls, Re: float = 7500, path_out=None, num_steps: int = 10, save_every: int = 0,
                     Tstart: float = 0.0, verbose: int = 0, meshpath=None, batch: int = 1, device: int = 0):
        path_out = Path(path_out) if path_out is not None else Path.cwd() / "data_output"
        params_flow = fsp.ParamFlow(Re=Re, uinf=1.0)
        params_flow.user_data.update({"L": 1.0, "D": 1.0})
        params_mesh = fsp.ParamMesh(meshpath=Path(meshpath) if meshpath else DATA / "cavity_coarse.npz")
        params_mesh.user_data.update({"xinf": 2.5, "xinfa": -1.2, "yinf": 0.5, "x0ns_left": -0.4, "x0ns_right": 1.75})
        params_control = fsp.ParamControl(
            sensor_list=[
                SensorHorizontalWallShear(sensor_index=100, x_sensor_left=1.0, x_sensor_right=1.1, y_sensor=0.0,
                                          sensor_type=SENSOR_TYPE.OTHER),
                SensorPoint(sensor_type=SENSOR_TYPE.U, position=np.array([0.1, 0.1])),
            ],
            actuator_list=[ActuatorForceGaussianV(sigma=0.0849, position=np.array([-0.1, 0.02]))],
        )
        return cls(
            params_flow=params_flow,
            params_time=fsp.ParamTime(num_steps=num_steps, dt=0.0004, Tstart=Tstart),
            params_save=fsp.ParamSave(save_every=save_every, path_out=path_out),
            params_solver=fsp.ParamSolver(throw_error=True, is_eq_nonlinear=True, shift=0.0),
            params_mesh=params_mesh,
            params_control=params_control,
            params_ic=fsp.ParamIC(),
            verbose=verbose,
            params_ensemble=fsp.ParamEnsemble(batch=batch, device=device),
        )
