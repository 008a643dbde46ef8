"""Parameter dataclasses of FlowSolver (configuration system).

Same names, fields and defaults as the reference's
/root/reference/src/flowcontrol/flowsolverparameters.py:26-217 so that user
scripts keep working; ``ParamEnsemble`` is new (ensemble width and devices).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path


@dataclass(kw_only=True)
class ParamFlowSolver:
    user_data: dict = field(default_factory=dict)


@dataclass
class ParamFlow(ParamFlowSolver):
    Re: float
    uinf: float = 1.0


@dataclass
class ParamMesh(ParamFlowSolver):
    meshpath: Path


@dataclass
class ParamControl(ParamFlowSolver):
    sensor_list: list
    sensor_number: int = field(init=False)
    actuator_list: list
    actuator_number: int = field(init=False)

    def __post_init__(self) -> None:
        self.sensor_number = len(self.sensor_list)
        self.actuator_number = len(self.actuator_list)


@dataclass
class ParamTime(ParamFlowSolver):
    num_steps: int
    dt: float
    Tstart: float
    Tfinal: float = field(init=False)

    def __post_init__(self) -> None:
        self.Tfinal = self.num_steps * self.dt


@dataclass
class ParamRestart(ParamFlowSolver):
    save_every_old: int = 0
    restart_order: int = 2
    dt_old: float = 0.0
    Trestartfrom: float = 0.0


@dataclass
class ParamSave(ParamFlowSolver):
    path_out: Path
    save_every: int
    energy_every: int = 1


@dataclass
class ParamSolver(ParamFlowSolver):
    throw_error: bool = True
    shift: float = 0.0
    is_eq_nonlinear: bool = True
    time_scheme: str = "bdf"


@dataclass
class ParamIC(ParamFlowSolver):
    xloc: float = 0.0
    yloc: float = 0.0
    radius: float = 1.0
    amplitude: float = 1.0


@dataclass
class ParamEnsemble(ParamFlowSolver):
    """New in this build: width of the trajectory ensemble and its placement."""

    batch: int = 1
    device: int = 0
    top_levels: int = 2  # levels at the top of the elimination tree merged into one dense inverse (multifrontal.py)
    leaf_cells: int = 16  # nested-dissection leaf size (ordering.py); 16 measured best on B200 (fewer, fatter bottom levels)
