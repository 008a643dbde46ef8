"""Timeseries logging and checkpoint export (host side).

Mirrors /root/reference/src/flowcontrol/exporter.py:42-290: same record columns
(``time, dE, runtime, u_ctrl_i, y_meas_i``), same CSV and JSON-sidecar contents.
The reference writes fields through dolfin's XDMF/HDF5 checkpoint writer
(utils/io.py:21-39); no HDF5 library exists here, so each checkpoint is one
``.npz`` per field file name (``U_restartT.npz`` next to where the ``.xdmf`` would
be) holding the appended snapshots in canonical numbering.  SURVEY.md section
8(f) row f2 tracks the byte-level XDMF writer.
"""

from __future__ import annotations

import json
import logging

import numpy as np
import pandas as pd

from .flowfield import Field, FlowFieldCollection, SimPaths

logger = logging.getLogger(__name__)


def _npz_path(path):
    return path.with_suffix(".npz")


def write_checkpoint(path, name: str, data: np.ndarray, time: float, append: bool) -> None:
    p = _npz_path(path)
    p.parent.mkdir(parents=True, exist_ok=True)
    if append and p.exists():
        old = np.load(p)
        snaps = np.concatenate([old["snapshots"], data[None]], axis=0)
        times = np.concatenate([old["times"], [time]])
    else:
        snaps, times = data[None], np.array([time])
    np.savez(p, snapshots=snaps, times=times, name=name)


def read_checkpoint(path, counter: int = -1) -> np.ndarray:
    return np.load(_npz_path(path))["snapshots"][counter]


class FlowExporter:
    def __init__(self, paths: SimPaths, fields: FlowFieldCollection, V=None, P=None, Tstart: float = 0.0,
                 dt: float = 0.0, save_every: int = 0) -> None:
        self.paths, self.fields, self.V, self.P = paths, fields, V, P
        self._Tstart, self._dt, self._save_every = Tstart, dt, save_every
        self._records: list[dict] = []
        self._checkpoints_written = 0
        self._u_cols = None
        self._y_cols = None

    def export_xdmf(self, u_n: Field, u_nn: Field, p_n: Field, time: float, append: bool = True,
                    write_mesh: bool = False, adjust_baseflow: float = 0.0) -> None:
        """Write (U, Uprev, P) snapshots, optionally as full fields (exporter.py:85-165)."""
        U0v, P0v = self.fields.U0.vector()[:], self.fields.P0.vector()[:]
        self.fields.Usave = Field(u_n.vector()[:] + adjust_baseflow * U0v)
        self.fields.Usave_n = Field(u_nn.vector()[:] + adjust_baseflow * U0v)
        self.fields.Psave = Field(p_n.vector()[:] + adjust_baseflow * P0v)
        self._checkpoints_written += 1
        write_checkpoint(self.paths.U_restart, "U", self.fields.Usave.array, time, append)
        write_checkpoint(self.paths.Uprev_restart, "U_n", self.fields.Usave_n.array, time, append)
        write_checkpoint(self.paths.P_restart, "P", self.fields.Psave.array, time, append)

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
