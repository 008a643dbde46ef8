"""Cost functions of the reference's optimisation loops (src/utils/optim.py:231-288), for one trajectory or an ensemble.

``compute_signal_cost`` / ``compute_control_cost`` keep the reference's names, arguments and error behaviour; they accept
a 1-D series (one trajectory, as the reference) or an array ``[nsteps, B]`` / ``[nsteps, na, B]`` (one column per
trajectory) and then return one cost per trajectory.  ``Ensemble.costs`` returns the same numbers accumulated on the
device during ``run_closed_loop`` (``fcb_get_costs``), so a controller sweep moves 24 bytes per trajectory to the host
instead of the time series."""
from __future__ import annotations

from typing import Callable

import numpy as np


def compute_signal_cost(signal, Tnorm: float, criterion: str, scaling: Callable | None = None):
    """Integral (time-averaged) or terminal cost of a time series (optim.py:231-270)."""
    if criterion not in ("integral", "terminal"):
        raise ValueError(f"Unknown criterion {criterion!r}: expected 'integral' or 'terminal'.")
    if scaling is None:
        def scaling(x):
            return x
    s = np.asarray(signal, dtype=np.float64)
    out = np.sum(scaling(s), axis=0) * Tnorm if criterion == "integral" else scaling(s[-1])
    return float(out) if np.ndim(out) == 0 else np.asarray(out)


def compute_control_cost(u_ctrl, Tnorm: float):
    """Time-normalised control effort, all actuator channels summed (optim.py:273-288).
    ``u_ctrl``: ``[nsteps]``, ``[nsteps, na]`` (one trajectory, as the reference) or ``[nsteps, na, B]``."""
    u = np.asarray(u_ctrl, dtype=np.float64)
    if u.ndim <= 2:
        return float(np.sum(u**2) * Tnorm)
    return np.sum(u**2, axis=(0, 1)) * Tnorm
