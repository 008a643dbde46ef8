"""CPU ORACLE (test infrastructure): the four benchmark configurations restated.

Each factory mirrors ``<Example>FlowSolver._make_boundaries/_make_bcs/_make_BCs/
make_default`` of the reference (files cited per function).  See flow_oracle.py
for the rules about who may import this package.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np

from .flow_oracle import (
    ActuatorSpec,
    CaseSpec,
    DirichletSpec,
    SensorSpec,
    between,
    gaussian_v,
    near,
    parabolic_slot,
    rotation_profile,
    uniform_u,
)

MESH_DIR = Path(__file__).resolve().parent.parent / "data" / "meshes"


def load_mesh(name: str):
    d = np.load(MESH_DIR / f"{name}.npz")
    return d["vertices"], d["triangles"]


def cylinder(Re: float = 100.0) -> CaseSpec:
    """examples/cylinder/cylinderflowsolver.py:20-108 (boundaries, BCs), :128-186 (defaults)."""
    xinfa, xinf, yinf = -10.0, 20.0, 10.0
    radius = 0.5
    L = radius * np.sin(0.5 * 10.0 * np.pi / 180.0)  # actuator.py:221

    def inlet(x, y):
        return near(x, xinfa)

    def walls(x, y):
        return near(y, -yinf) | near(y, yinf)

    def close(x, y):
        return between(x, -radius, radius) & between(y, -radius, radius)

    def cyl(x, y):
        return close(x, y) & (between(x, -radius, -L) | between(x, L, radius))

    def act_up(x, y):
        return close(x, y) & between(x, -L, L, 0.01) & between(y, 0.0, radius)

    def act_lo(x, y):
        return close(x, y) & between(x, -L, L, 0.01) & between(y, -radius, 0.0)

    acts = [ActuatorSpec("bc", parabolic_slot(L, 0.0)), ActuatorSpec("bc", parabolic_slot(L, 0.0))]
    tail = [
        DirichletSpec(walls, (1,), (0.0,)),
        DirichletSpec(cyl, (0, 1), (0.0, 0.0)),
        DirichletSpec(act_up, (0, 1), ("actuator", 0)),
        DirichletSpec(act_lo, (0, 1), ("actuator", 1)),
    ]
    return CaseSpec(
        name="cylinder",
        mesh_file="cylinder_O1",
        Re=Re,
        dt=0.005,
        uinf=1.0,
        bcs_pert=[DirichletSpec(inlet, (0, 1), (0.0, 0.0))] + tail,
        bcs_full=[DirichletSpec(inlet, (0, 1), (1.0, 0.0))] + tail,
        actuators=acts,
        sensors=[
            SensorSpec("point", comp=1, position=(3.0, 0.0)),
            SensorSpec("point", comp=1, position=(3.1, 1.0)),
            SensorSpec("point", comp=1, position=(3.1, -1.0)),
        ],
        initial_guess=lambda x, y: (np.ones_like(x), np.zeros_like(x)),
    )


def cavity(Re: float = 7500.0) -> CaseSpec:
    """examples/cavity/cavityflowsolver.py:22-193 (boundaries, BCs), :195-280 (guess, defaults)."""
    L, D = 1.0, 1.0
    xinfa, xinf, yinf = -1.2, 2.5, 0.5
    x0l, x0r = -0.4, 1.75
    T = 3.0e-16

    def inlet(x, y):
        return near(x, xinfa)

    def upper(x, y):
        return near(y, yinf)

    def cav_left(x, y):
        return near(x, 0.0) & between(y, -D, 0.0)

    def cav_botm(x, y):
        return near(y, -D) & between(x, 0.0, L)

    def cav_right(x, y):
        return near(x, L) & between(y, -D, 0.0)

    def ll_sf(x, y):
        return (x >= xinfa) & (x <= x0l + 10 * T) & near(y, 0.0)

    def ll_ns(x, y):
        return (x >= x0l - 10 * T) & (x <= 0.0) & near(y, 0.0)

    def lr_ns(x, y):
        return near(y, 0.0) & between(x, L, x0r)

    def lr_sf(x, y):
        return near(y, 0.0) & between(x, x0r, xinf)

    tail = [
        DirichletSpec(upper, (1,), (0.0,)),
        DirichletSpec(ll_sf, (1,), (0.0,)),
        DirichletSpec(ll_ns, (0, 1), (0.0, 0.0)),
        DirichletSpec(lr_ns, (0, 1), (0.0, 0.0)),
        DirichletSpec(lr_sf, (1,), (0.0,)),
        DirichletSpec(cav_left, (0, 1), (0.0, 0.0)),
        DirichletSpec(cav_botm, (0, 1), (0.0, 0.0)),
        DirichletSpec(cav_right, (0, 1), (0.0, 0.0)),
    ]
    return CaseSpec(
        name="cavity",
        mesh_file="cavity_coarse",
        Re=Re,
        dt=0.0004,
        uinf=1.0,
        bcs_pert=[DirichletSpec(inlet, (0, 1), (0.0, 0.0))] + tail,
        bcs_full=[DirichletSpec(inlet, (0, 1), (1.0, 0.0))] + tail,
        actuators=[ActuatorSpec("force", gaussian_v(0.0849, (-0.1, 0.02)))],
        sensors=[
            SensorSpec("wall_shear", x_left=1.0, x_right=1.1, y=0.0),
            SensorSpec("point", comp=0, position=(0.1, 0.1)),
        ],
        initial_guess=lambda x, y: (np.where(y >= 0.0, 1.0, 0.0), np.zeros_like(x)),
    )


def lidcavity(Re: float = 8000.0) -> CaseSpec:
    """examples/lidcavity/lidcavityflowsolver.py:25-95 (boundaries, BCs, guess), :97-148."""

    def lid(x, y):
        return near(y, 1.0)

    def left(x, y):
        return near(x, 0.0)

    def right(x, y):
        return near(x, 1.0)

    def bottom(x, y):
        return near(y, 0.0)

    tail = [
        DirichletSpec(left, (0, 1), (0.0, 0.0)),
        DirichletSpec(right, (0, 1), (0.0, 0.0)),
        DirichletSpec(bottom, (0, 1), (0.0, 0.0)),
    ]
    return CaseSpec(
        name="lidcavity",
        mesh_file="lidcavity_mesh64",
        Re=Re,
        dt=0.005,
        uinf=1.0,
        bcs_pert=[DirichletSpec(lid, (0, 1), ("actuator", 0))] + tail,
        bcs_full=[DirichletSpec(lid, (0, 1), (1.0, 0.0))] + tail,
        actuators=[ActuatorSpec("bc", uniform_u())],
        sensors=[
            SensorSpec("point", comp=1, position=(0.05, 0.5)),
            SensorSpec("point", comp=0, position=(0.5, 0.95)),
        ],
        initial_guess=lambda x, y: (np.zeros_like(x), np.zeros_like(x)),
        pin_pressure=True,
    )


def pinball(Re: float = 50.0, mode: str = "rotation") -> CaseSpec:
    """examples/pinball/pinballflowsolver.py:25-192 (boundaries, BCs), :234-325 (defaults)."""
    xinfa, xinf, yinf = -6.0, 20.0, 6.0
    radius = 0.5
    c30 = 1.5 * np.cos(np.pi / 6)

    def inlet(x, y):
        return near(x, xinfa)

    def walls(x, y):
        return near(y, -yinf) | near(y, yinf)

    def close_top(x, y):
        return between(x, -radius, radius) & between(y, radius / 2, 5 * radius / 2)

    def close_bot(x, y):
        return between(x, -radius, radius) & between(y, -5 * radius / 2, -radius / 2)

    def close_mid(x, y):
        return between(x, -radius - c30, radius - c30) & between(y, -radius, radius)

    pert = [DirichletSpec(inlet, (0, 1), (0.0, 0.0)), DirichletSpec(walls, (1,), (0.0,))]
    full = [DirichletSpec(inlet, (0, 1), (1.0, 0.0)), DirichletSpec(walls, (0, 1), (1.0, 0.0))]
    if mode == "suction":
        L = radius * np.sin(0.5 * 10.0 * np.pi / 180.0)
        acts = [
            ActuatorSpec("bc", parabolic_slot(L, -c30)),
            ActuatorSpec("bc", parabolic_slot(L, 0.0)),
            ActuatorSpec("bc", parabolic_slot(L, 0.0)),
        ]

        def act_mid(x, y):
            return close_mid(x, y) & between(x, -L - c30, -c30 + L)

        def act_top(x, y):
            return close_top(x, y) & between(x, -L, L)

        def act_bot(x, y):
            return close_bot(x, y) & between(x, -L, L)

        tail = [
            DirichletSpec(close_top, (0, 1), (0.0, 0.0)),
            DirichletSpec(close_bot, (0, 1), (0.0, 0.0)),
            DirichletSpec(close_mid, (0, 1), (0.0, 0.0)),
            DirichletSpec(act_mid, (0, 1), ("actuator", 0)),
            DirichletSpec(act_top, (0, 1), ("actuator", 1)),
            DirichletSpec(act_bot, (0, 1), ("actuator", 2)),
        ]
    else:
        acts = [
            ActuatorSpec("bc", rotation_profile(-c30, 0.0, 1.0)),
            ActuatorSpec("bc", rotation_profile(0.0, 0.75, 1.0)),
            ActuatorSpec("bc", rotation_profile(0.0, -0.75, 1.0)),
        ]
        tail = [
            DirichletSpec(close_mid, (0, 1), ("actuator", 0)),
            DirichletSpec(close_top, (0, 1), ("actuator", 1)),
            DirichletSpec(close_bot, (0, 1), ("actuator", 2)),
        ]
    return CaseSpec(
        name=f"pinball_{mode}",
        mesh_file="pinball_middle",
        Re=Re,
        dt=0.005,
        uinf=1.0,
        bcs_pert=pert + tail,
        bcs_full=full + tail,
        actuators=acts,
        sensors=[
            SensorSpec("point", comp=1, position=(8.0, 0.0)),
            SensorSpec("point", comp=1, position=(10.0, 0.0)),
            SensorSpec("point", comp=1, position=(12.0, 0.0)),
        ],
        initial_guess=lambda x, y: (np.ones_like(x), np.zeros_like(x)),
    )
