"""Crank-Nicolson in the oracle (nsforms.py:191-236).  The reference holds no numeric fixture for time_scheme='cn'
(tests/test_nsforms.py:62-68 only checks that the form builds), so this row's parity is pinned through properties of
the scheme itself: the step satisfies the theta=1/2 weak form row by row, it is second-order accurate in time, and it
converges to the same trajectory as the (golden-pinned) BDF2 path."""
import numpy as np

from oracle import flow_oracle as fo
from util import unit_square_mesh


def _case(dt):
    return fo.CaseSpec(
        name="t", mesh_file="", Re=60.0, dt=dt, uinf=1.0,
        bcs_pert=[fo.DirichletSpec(lambda x, y: fo.near(y, 1.0), (0, 1), ("actuator", 0)),
                  fo.DirichletSpec(lambda x, y: fo.near(x, 0.0), (0, 1), (0.0, 0.0)),
                  fo.DirichletSpec(lambda x, y: fo.near(y, 0.0), (1,), (0.0,))],
        bcs_full=[], actuators=[fo.ActuatorSpec("bc", fo.uniform_u()), fo.ActuatorSpec("force", fo.gaussian_v(0.15, (0.4, 0.5)))],
        sensors=[fo.SensorSpec("point", comp=1, position=(0.31, 0.52))], initial_guess=None,
    )


def _setup(dt, scheme, seed=3, nonlinear=True):
    xy, tri = unit_square_mesh(8, jitter=0.1, seed=5)
    orc = fo.FlowOracle(_case(dt), xy, tri, time_scheme=scheme, nonlinear=nonlinear)
    rng = np.random.default_rng(seed)
    m = orc.mesh
    # smooth divergence-free-ish base flow and initial perturbation
    x, y = m.node_xy[:, 0], m.node_xy[:, 1]
    U0 = np.concatenate([np.sin(np.pi * y) * 0.4, 0.2 * np.sin(np.pi * x), np.zeros(m.nV)])
    orc.set_base_flow(U0)
    ic = np.concatenate([0.2 * np.sin(2 * np.pi * y) * x * (1 - y), 0.2 * x * np.sin(np.pi * y) * (1 - y), np.zeros(m.nV)])
    ic[orc.bc_pert.dofs] = 0.0
    orc.init_time_stepping(ic=ic)
    return orc, rng


def _uc(t):
    return np.array([0.3 * np.sin(3.0 * t), 0.5 * np.sin(2.0 * t) + 0.2 * np.sin(5.0 * t)])


def _run(dt, scheme, T=0.32, nonlinear=True):
    """From rest, driven by actuation that vanishes at t = 0 (the cached force f_n starts at zero, flowsolver.py:681-686)."""
    orc, _ = _setup(dt, scheme, nonlinear=nonlinear)
    orc.init_time_stepping(ic=np.zeros(orc.mesh.N))
    for k in range(int(round(T / dt))):
        orc.step(_uc((k + 1) * dt))
    return orc.up[: orc.mesh.Nv].copy()


def test_cn_step_satisfies_theta_half_weak_form():
    dt = 0.01
    orc, _ = _setup(dt, "cn")
    m, o = orc.mesh, orc.ops
    Re = orc.case.Re
    U0v = orc.UP0[: m.Nv]
    # linear velocity operator L = C(U0) + D(U0) + K/Re from the oracle's own blocks at c = 0, shift = 0
    L = o.lhs(0.0, Re, U0v, newton_terms=True).tocsr()
    Lvv, G = L[: m.Nv, : m.Nv], L[: m.Nv, m.Nv :]
    Bdiv = L[m.Nv :, : m.Nv]
    free = np.setdiff1d(np.arange(m.Nv), orc.bc_pert.dofs)
    prev = np.zeros(2)
    for k in range(3):
        u0 = orc.u_n.copy()
        uc = _uc((k + 1) * dt)
        orc.step(uc)
        u1, p1 = orc.up[: m.Nv], orc.up[m.Nv :]
        r = (o.Mv @ (u1 - u0) / dt + 0.5 * Lvv @ (u1 + u0) + o.convection(u0) + G @ p1
             - 0.5 * (orc.force_rhs(uc) + orc.force_rhs(prev)))
        scale = np.abs(o.Mv @ u1 / dt).max()
        assert np.abs(r[free]).max() < 1e-10 * scale          # momentum rows, theta = 1/2, force averaged
        assert np.abs(Bdiv @ u1)[1 if orc.case.pin_pressure else 0:].max() < 1e-11   # continuity fully implicit
        g = orc.bc_pert.values(uc)
        assert np.array_equal(u1[orc.bc_pert.dofs[orc.bc_pert.dofs < m.Nv]], g[orc.bc_pert.dofs < m.Nv])
        prev = uc


def test_cn_order_of_accuracy_and_convergence_to_bdf2():
    """theta = 1/2 with the force averaged is second order for the linearised equations; the perturbation advection
    is explicit at t^n (nsforms.py:229), which makes the nonlinear scheme first order asymptotically (between 1 and 2 at these amplitudes).
    Either way it approaches the trajectory of the golden-pinned BDF path as dt -> 0."""
    for nonlinear, lo, hi in ((False, 3.4, 4.6), (True, 1.8, 4.6)):
        fine = _run(0.32 / 256, "cn", nonlinear=nonlinear)
        errs = [np.linalg.norm(_run(0.32 / n, "cn", nonlinear=nonlinear) - fine) for n in (4, 8, 16)]
        assert lo < errs[0] / errs[1] < hi and lo < errs[1] / errs[2] < hi, (nonlinear, errs)
        bdf = _run(0.32 / 256, "bdf", nonlinear=nonlinear)
        assert np.linalg.norm(bdf - fine) < errs[2], (nonlinear, errs)
