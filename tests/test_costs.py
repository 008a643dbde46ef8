"""Host cost helpers (flowcontrol_b200/costs.py) keep the semantics of the reference's
src/utils/optim.py:231-288 for one trajectory and extend them column-wise to ensembles."""
import numpy as np
import pytest

from flowcontrol_b200.costs import compute_control_cost, compute_signal_cost


def test_signal_cost_integral_terminal_and_scaling():
    rng = np.random.default_rng(0)
    sig = rng.random(50)
    Tnorm = 0.005 / 0.25
    assert compute_signal_cost(sig, Tnorm, "integral") == pytest.approx(float(np.sum(sig) * Tnorm), rel=1e-15)
    assert compute_signal_cost(sig, Tnorm, "terminal") == sig[-1]
    assert compute_signal_cost(sig, Tnorm, "integral", scaling=np.sqrt) == pytest.approx(float(np.sum(np.sqrt(sig)) * Tnorm))
    with pytest.raises(ValueError, match="Unknown criterion"):
        compute_signal_cost(sig, Tnorm, "mean")
    ens = rng.random((50, 7))
    c = compute_signal_cost(ens, Tnorm, "integral")
    assert c.shape == (7,) and np.allclose(c, [compute_signal_cost(ens[:, b], Tnorm, "integral") for b in range(7)])
    assert np.array_equal(compute_signal_cost(ens, Tnorm, "terminal"), ens[-1])


def test_control_cost_sums_all_channels():
    rng = np.random.default_rng(1)
    u = rng.standard_normal((40, 2))
    assert compute_control_cost(u, 0.1) == pytest.approx(float(np.sum(u**2) * 0.1))
    assert compute_control_cost(u[:, 0], 0.1) == pytest.approx(float(np.sum(u[:, 0] ** 2) * 0.1))
    ub = rng.standard_normal((40, 2, 5))
    c = compute_control_cost(ub, 0.1)
    assert c.shape == (5,) and np.allclose(c, [compute_control_cost(ub[:, :, b], 0.1) for b in range(5)])


def test_fun_array_point_by_point_fallback():
    """utils/optim.py:48-66: any callable is evaluated row by row, output [n_points, 1]."""
    from flowcontrol_b200.costs import fun_array

    X = np.arange(12.0).reshape(4, 3)
    out = fun_array(X, lambda x, p=1.0: p * float(x @ x), p=2.0)
    assert out.shape == (4, 1) and np.allclose(out[:, 0], 2.0 * (X * X).sum(axis=1))
