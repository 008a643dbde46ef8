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


def fun_array(x, fun: Callable[..., float], **kwargs) -> np.ndarray:
    """The reference's batch evaluation (optim.py:48-66): ``fun`` on every row of ``x`` [n_points, dim] -> [n_points, 1].
    When ``fun`` is an :class:`EnsembleCost` the whole batch runs as ensembles on the GPU instead of one simulation per
    point; any other callable is evaluated point by point as in the reference."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    if isinstance(fun, EnsembleCost):
        return fun.evaluate(x, **kwargs)[:, None]
    out = np.zeros((x.shape[0], 1))
    for i in range(x.shape[0]):
        out[i, :] = fun(x[i, :], **kwargs)
    return out


class EnsembleCost:
    """Closed-loop cost ``J(x) = xQx + u_penalty * uRu`` of a parametrised controller (the body of the reference's
    optimisation loops, optim.py:1-25), evaluated for MANY parameter vectors at once: every point becomes one trajectory
    of an ensemble (same mesh, base flow, initial condition), ``fcb_run_closed_loop`` runs the simulations on the device,
    and the cost sums come back as 24 bytes per trajectory (``fcb_get_costs``).

        cost = EnsembleCost(ensemble, make_controller, u_n, nsteps, Ky=[[-1, 0, 0]], Fu=[[1], [1]], u_penalty=1e-2)
        J = fun_array(X, cost)              # drop-in for fun_array(X, fun) of the reference
        j = cost(x)                         # a single point (scalar signature of the reference's `fun`)

    ``make_controller(x) -> Controller`` maps a parameter vector to a continuous-time controller (e.g.
    ``controller_residues_wrapper`` of utils/lticontrol.py); ``criterion`` is 'integral' or 'terminal'
    (compute_signal_cost); a diverged trajectory costs ``diverged_cost``.  Populations larger than the ensemble are
    evaluated in chunks; a last, partial chunk is padded with its last point."""

    def __init__(self, ensemble, make_controller: Callable, u_n, nsteps: int, Ky, Fu, u_penalty: float = 0.0,
                 criterion: str = "integral", diverged_cost: float = 1e10, u_nn=None, p_n=None, order: int = 1):
        if criterion not in ("integral", "terminal"):
            raise ValueError(f"Unknown criterion {criterion!r}: expected 'integral' or 'terminal'.")
        self.ens, self.make_controller, self.nsteps = ensemble, make_controller, int(nsteps)
        self.state = (u_n, u_nn, p_n, order)
        self.Ky, self.Fu = np.atleast_2d(np.asarray(Ky, dtype=np.float64)), np.atleast_2d(np.asarray(Fu, dtype=np.float64))
        self.u_penalty, self.criterion, self.diverged_cost = float(u_penalty), criterion, float(diverged_cost)
        self.last = None  # per-point cost components of the last evaluate()

    def evaluate(self, X) -> np.ndarray:
        from .controller import ControllerBank

        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        npts, B = X.shape[0], self.ens.B
        dt = self.ens.problem.dt
        Tnorm = 1.0 / self.nsteps  # fs.dt / (fs.t - fs.Tc) for a run of nsteps from Tc
        J = np.empty(npts)
        parts = {"xQx": np.empty(npts), "uRu": np.empty(npts), "diverged": np.zeros(npts, dtype=bool)}
        for c0 in range(0, npts, B):
            idx = np.minimum(np.arange(c0, c0 + B), npts - 1)  # pad a partial chunk with its last point
            bank = ControllerBank([self.make_controller(X[i]) for i in idx], dt, self.Ky, self.Fu)
            u_n, u_nn, p_n, order = self.state
            self.ens.set_state(u_n, u_nn, p_n, order=order)
            self.ens.set_controllers(bank)
            self.ens.run_closed_loop(self.nsteps, log=False)
            c = self.ens.costs(Tnorm)
            self.ens.measurement()
            n = min(B, npts - c0)
            xqx = c["energy_integral"] if self.criterion == "integral" else c["energy_terminal"]
            bad = (self.ens.diverged != 0) | ~np.isfinite(xqx) | ~np.isfinite(c["control"])
            cost = np.where(bad, self.diverged_cost, xqx + self.u_penalty * c["control"])
            J[c0 : c0 + n] = cost[:n]
            parts["xQx"][c0 : c0 + n], parts["uRu"][c0 : c0 + n], parts["diverged"][c0 : c0 + n] = xqx[:n], c["control"][:n], bad[:n]
        self.last = parts
        return J

    def __call__(self, x, **kwargs) -> float:
        return float(self.evaluate(np.atleast_2d(x))[0])
