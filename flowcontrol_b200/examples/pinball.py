"""Fluidic pinball (three cylinders): rotation or suction actuation, three wake probes.

Restates /root/reference/src/examples/pinball/pinballflowsolver.py:22-325.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np
import pandas as pd

from .. import flowsolverparameters as fsp
from ..actuator import CYLINDER_ACTUATION_MODE, ActuatorBCParabolicV, ActuatorBCRotation
from ..flowfield import BoundaryConditions
from ..flowsolver import FlowSolver, SubDomain, between, near
from ..problem import DirichletBC
from ..sensor import SENSOR_TYPE, SensorForceCoefficient, SensorPoint

DATA = Path(__file__).resolve().parents[2] / "data" / "meshes"
C30 = 1.5 * np.cos(np.pi / 6)


class PinballFlowSolver(FlowSolver):
    def _make_boundaries(self) -> pd.DataFrame:
        ud = self.params_mesh.user_data
        xinfa, xinf, yinf = ud["xinfa"], ud["xinf"], ud["yinf"]
        r = self.params_flow.user_data["D"] / 2
        mode = self.params_control.user_data["mode_actuation"]

        def top(x, y):
            return between(x, -r, r) & between(y, r / 2, 5 * r / 2)

        def bot(x, y):
            return between(x, -r, r) & between(y, -5 * r / 2, -r / 2)

        def mid(x, y):
            return between(x, -r - C30, r - C30) & between(y, -r, r)

        subs = {
            "inlet": lambda x, y: near(x, xinfa),
            "outlet": lambda x, y: near(x, xinf),
            "walls": lambda x, y: near(y, -yinf) | near(y, yinf),
        }
        if mode == CYLINDER_ACTUATION_MODE.SUCTION:
            L = self.params_control.actuator_list[0].width
            subs.update({
                "cylinder_top": top, "cylinder_bot": bot, "cylinder_mid": mid,
                "actuator_mid": lambda x, y: mid(x, y) & between(x, -L - C30, -C30 + L),
                "actuator_top": lambda x, y: top(x, y) & between(x, -L, L),
                "actuator_bot": lambda x, y: bot(x, y) & between(x, -L, L),
            })
        else:
            subs.update({"actuator_mid": mid, "actuator_top": top, "actuator_bot": bot})
        return pd.DataFrame(index=list(subs), data={"subdomain": [SubDomain(f) for f in subs.values()]})

    def _tail_bcs(self):
        sub = lambda n: self.get_subdomain(n).inside  # noqa: E731
        acts = self.params_control.actuator_list
        out = []
        if self.params_control.user_data["mode_actuation"] == CYLINDER_ACTUATION_MODE.SUCTION:
            out += [DirichletBC(sub(n), (0, 1), (0.0, 0.0)) for n in ("cylinder_top", "cylinder_bot", "cylinder_mid")]
        out += [DirichletBC(sub("actuator_mid"), (0, 1), acts[0]), DirichletBC(sub("actuator_top"), (0, 1), acts[1]),
                DirichletBC(sub("actuator_bot"), (0, 1), acts[2])]
        return out

    def _make_bcs(self) -> BoundaryConditions:
        sub = lambda n: self.get_subdomain(n).inside  # noqa: E731
        return BoundaryConditions(
            bcu=[DirichletBC(sub("inlet"), (0, 1), (0.0, 0.0)), DirichletBC(sub("walls"), (1,), (0.0,))] + self._tail_bcs(),
            bcp=[],
        )

    def _make_BCs(self) -> BoundaryConditions:
        """Uniform flow on inlet AND walls (pinballflowsolver.py:186-192)."""
        sub = lambda n: self.get_subdomain(n).inside  # noqa: E731
        uni = (self.params_flow.uinf, 0.0)
        return BoundaryConditions(
            bcu=[DirichletBC(sub("inlet"), (0, 1), uni), DirichletBC(sub("walls"), (0, 1), uni)] + self._tail_bcs(), bcp=[]
        )

    def compute_force_coefficients(self, u, p) -> dict:
        """{surface: (cl, cd)} for each cylinder surface (pinballflowsolver.py:202-232)."""
        vec = np.concatenate([np.asarray(u.vector().get_local() if hasattr(u, "vector") else u),
                              np.asarray(p.vector().get_local() if hasattr(p, "vector") else p)])
        D = self.params_flow.user_data["D"]
        nu = self.params_flow.uinf * D / self.params_flow.Re
        if self.params_control.user_data["mode_actuation"] == CYLINDER_ACTUATION_MODE.SUCTION:
            names = ["cylinder_mid", "actuator_mid", "cylinder_top", "actuator_top", "cylinder_bot", "actuator_bot"]
        else:
            names = ["actuator_mid", "actuator_top", "actuator_bot"]
        out = {}
        for name in names:
            vals = []
            for c in (1, 0):
                idx, val = SensorForceCoefficient(sensor_type=SENSOR_TYPE.OTHER, inside=self.get_subdomain(name).inside,
                                                  component=c, nu=nu, uinf=self.params_flow.uinf, D=D).row(self.tables)
                vals.append(float(val @ vec[idx]))
            out[name] = (vals[0], vals[1])
        return out

    @classmethod
    def make_default(cls, Re: float = 50, mode_actuation=None, path_out=None, num_steps: int = 10, save_every: int = 0,
                     Tstart: float = 0.0, verbose: int = 0, meshpath=None, batch: int = 1, device: int = 0):
        path_out = Path(path_out) if path_out is not None else Path.cwd() / "data_output"
        mode = mode_actuation or CYLINDER_ACTUATION_MODE.ROTATION
        params_flow = fsp.ParamFlow(Re=Re, uinf=1.0)
        params_flow.user_data["D"] = 1.0
        params_mesh = fsp.ParamMesh(meshpath=Path(meshpath) if meshpath else DATA / "pinball_middle.npz")
        params_mesh.user_data.update({"xinf": 20, "xinfa": -6, "yinf": 6})
        if mode == CYLINDER_ACTUATION_MODE.SUCTION:
            w = ActuatorBCParabolicV.angular_size_deg_to_width(10, 0.5)
            acts = [ActuatorBCParabolicV(width=w, position_x=-C30), ActuatorBCParabolicV(width=w, position_x=0.0),
                    ActuatorBCParabolicV(width=w, position_x=0.0)]
        else:
            acts = [ActuatorBCRotation(position_x=-C30, position_y=0.0, diameter=1.0),
                    ActuatorBCRotation(position_x=0.0, position_y=0.75, diameter=1.0),
                    ActuatorBCRotation(position_x=0.0, position_y=-0.75, diameter=1.0)]
        params_control = fsp.ParamControl(
            sensor_list=[SensorPoint(sensor_type=SENSOR_TYPE.V, position=np.array([x, 0.0])) for x in (8.0, 10.0, 12.0)],
            actuator_list=acts,
            user_data={"mode_actuation": mode},
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
