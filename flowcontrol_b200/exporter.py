"""Timeseries logging and checkpoint export (host side).

Mirrors /root/reference/src/flowcontrol/exporter.py:42-290: same record columns
(``time, dE, runtime, u_ctrl_i, y_meas_i``), same CSV and JSON-sidecar contents, and the same field files: the
reference writes fields through dolfin's XDMF/HDF5 checkpoint writer (utils/io.py:21-39); xdmf_checkpoint.py writes and
reads that layout (``<name>.xdmf`` + ``<name>.h5`` with ``/<fn>/<fn>_<k>/{vector,cell_dofs,x_cell_dofs,cells,mesh/...}``)
without dolfin or libhdf5.  An ensemble (batch > 1) stores one function per trajectory in the same files: trajectory 0
under the reference's function name, trajectory b as ``<name>_traj<b:04d>``, so every trajectory can be restarted on its
own (by this package or by the reference) and the whole ensemble can be restarted trajectory by trajectory.
"""

from __future__ import annotations

import json
import logging

import numpy as np
import pandas as pd

from . import xdmf_checkpoint as xc
from .flowfield import Field, FlowFieldCollection, SimPaths

logger = logging.getLogger(__name__)


def _kind(tab, size: int) -> str:
    if size == tab.Nv:
        return "V"
    if size == tab.nV:
        return "P"
    raise ValueError(f"a field of {size} dofs is neither a velocity ({tab.Nv}) nor a pressure ({tab.nV}) field")


def write_checkpoint(path, name: str, data: np.ndarray, time: float, append: bool, tab) -> None:
    """``write_xdmf`` of utils/io.py:21-39.  ``data`` [n] is one function, [n, B] one function per trajectory."""
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        xc.write_checkpoint(path, name, tab, data, _kind(tab, data.shape[0]), time, append=append)
        return
    names = [xc.trajectory_name(name, b) for b in range(data.shape[1])]
    xc.write_checkpoints(path, names, tab, data, _kind(tab, data.shape[0]), time, append=append)


def checkpoint_function_name(path) -> str:
    """Name of the first function stored in an XDMF checkpoint file."""
    import re
    from pathlib import Path

    m = re.search(r'<Grid Name="([^"]+)" GridType="Collection"', Path(path).read_text())
    if not m:
        raise ValueError(f"{path}: not an XDMF function checkpoint")
    return m.group(1)


def read_checkpoint(path, counter: int = -1, tab=None, name: str | None = None, batch: int = 1) -> np.ndarray:
    """``read_xdmf`` of utils/io.py:42-50: canonical dof vector [n]; with ``batch`` > 1 every trajectory, [n, batch]
    (a file that holds a single function is broadcast)."""
    import re
    from pathlib import Path

    if tab is None:
        raise ValueError("read_checkpoint needs the mesh tables")
    name = name or checkpoint_function_name(path)
    text = Path(path).read_text()
    kind = "V" if re.search(rf'Name="{re.escape(name)}" Center="Other" AttributeType="Vector"', text) else "P"
    if batch <= 1:
        return xc.read_checkpoint(path, name, tab, kind, counter)
    names = [xc.trajectory_name(name, b) for b in range(batch)]
    have = [fn for fn in names if f'<Grid Name="{fn}" GridType="Collection"' in text]
    got = xc.read_checkpoints(path, have, tab, kind, counter)  # the file is opened once for all trajectories
    col = {fn: k for k, fn in enumerate(have)}
    return np.stack([got[:, col.get(fn, 0)] for fn in names], axis=1)  # a file of a single run is broadcast


class FlowExporter:
    def __init__(self, paths: SimPaths, fields: FlowFieldCollection, V=None, P=None, Tstart: float = 0.0,
                 dt: float = 0.0, save_every: int = 0, tab=None) -> None:
        self.paths, self.fields, self.V, self.P, self.tab = paths, fields, V, P, tab
        self._Tstart, self._dt, self._save_every = Tstart, dt, save_every
        self._records: list[dict] = []
        self._checkpoints_written = 0
        self._u_cols = None
        self._y_cols = None

    def export_xdmf(self, u_n: Field, u_nn: Field, p_n: Field, time: float, append: bool = True,
                    write_mesh: bool = False, adjust_baseflow: float = 0.0) -> None:
        """Write (U, Uprev, P) snapshots, optionally as full fields (exporter.py:85-165).  The arguments are Fields
        (one trajectory) or arrays [n, B] (every trajectory of an ensemble: one function per trajectory in the same files)."""
        U0v, P0v = self.fields.U0.vector()[:], self.fields.P0.vector()[:]
        arr = lambda f: f.vector()[:] if isinstance(f, Field) else np.asarray(f, dtype=np.float64)  # noqa: E731
        base = lambda v, ref: ref if v.ndim == 1 else ref[:, None]  # noqa: E731
        U, Un, Pn = arr(u_n), arr(u_nn), arr(p_n)
        U, Un, Pn = U + adjust_baseflow * base(U, U0v), Un + adjust_baseflow * base(Un, U0v), Pn + adjust_baseflow * base(Pn, P0v)
        first = lambda v: v if v.ndim == 1 else v[:, 0]  # noqa: E731
        self.fields.Usave, self.fields.Usave_n, self.fields.Psave = Field(first(U)), Field(first(Un)), Field(first(Pn))
        self._checkpoints_written += 1
        write_checkpoint(self.paths.U_restart, "U", U, time, append, self.tab)
        write_checkpoint(self.paths.Uprev_restart, "U_n", Un, time, append, self.tab)
        write_checkpoint(self.paths.P_restart, "P", Pn, time, append, self.tab)

    def log_ic(self, t: float, y_meas, dE: float) -> None:
        row = {"time": t, "dE": dE, "runtime": 0.0}
        for i, v in enumerate(y_meas):
            row[f"y_meas_{i + 1}"] = float(v)
        self._records.append(row)

    def log(self, u_ctrl, y_meas, dE: float, t: float, runtime: float) -> None:
        if self._u_cols is None:
            self._u_cols = [f"u_ctrl_{i + 1}" for i in range(len(u_ctrl))]
            self._y_cols = [f"y_meas_{i + 1}" for i in range(len(y_meas))]
        row = {"time": t, "dE": dE, "runtime": runtime}
        row.update(zip(self._u_cols, (float(v) for v in u_ctrl)))
        row.update(zip(self._y_cols, (float(v) for v in y_meas)))
        self._records.append(row)

    def to_dataframe(self) -> pd.DataFrame:
        return pd.DataFrame(self._records)

    def write_metadata(self, restart_order=2) -> None:
        meta = {
            "Tstart": self._Tstart,
            "dt": self._dt,
            "save_every": self._save_every,
            "checkpoints_written": self._checkpoints_written,
            "restart_order": restart_order,
            "files": {
                "U": self.paths.U_restart.name,
                "Uprev": self.paths.Uprev_restart.name,
                "P": self.paths.P_restart.name,
            },
        }
        self.paths.metadata.parent.mkdir(parents=True, exist_ok=True)
        self.paths.metadata.write_text(json.dumps(meta, indent=2))

    def write_timeseries(self) -> None:
        self.paths.timeseries.parent.mkdir(parents=True, exist_ok=True)
        self.to_dataframe().to_csv(self.paths.timeseries, sep=",", index=False)

    def log_progress(self, iter: int, num_steps: int, t: float, t_end: float, runtime: float) -> None:
        logger.info("--- iter: %5d/%5d --- time: %3.3f/%3.3f --- elapsed %5.5f ---", iter, num_steps, t, t_end, runtime)

    def reset(self) -> None:
        self._records.clear()
        self._checkpoints_written = 0
