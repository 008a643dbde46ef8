"""BASELINE.json configs[2..4] (and the B=1 latency case) at their full ensemble widths on one B200.

    python tools/config_runs.py [out.json]

For each configuration: build the problem from the committed mesh + base-flow fixtures, run the named scenario
(device-resident inputs, CUDA-graph replay), time the steady BDF2 phase with CUDA events, and check the
size-independent properties the domain offers at that width: all trajectories finite, trajectories with identical
inputs bit-identical, and a trajectory of the wide ensemble equal to the same trajectory stepped in a 32-wide ensemble.
The fixtures' base flows are the regression scenarios' (pinball Re=30, lid cavity Re=1000): the cost of a step does not
depend on the Reynolds number (same sparsity), the base-flow computation at the configs' Re is setup-time work.
"""
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from flowcontrol_b200.controller import Controller, ControllerBank  # noqa: E402
from flowcontrol_b200.ensemble import Ensemble  # noqa: E402
from flowcontrol_b200.flowfield import Field  # noqa: E402
from flowcontrol_b200.problem import FlowProblem  # noqa: E402

GOLD = ROOT / "tests" / "golden"


def run_open_loop(prob, B, ic, u_of_step, nsteps, label, same_pair):
    """u_of_step(k) -> [na, B] numpy; returns the record for the JSON table."""
    tab = prob.tab
    t0 = time.time()
    ens = Ensemble(prob, B)
    create_s = time.time() - t0
    stream = torch.cuda.ExternalStream(ens.stream)
    ens.set_state(ic[: tab.Nv] if ic.ndim == 1 else ic[: tab.Nv, :], None, ic[tab.Nv :] if ic.ndim == 1 else ic[tab.Nv :, :], order=1)
    us = [torch.as_tensor(np.ascontiguousarray(u_of_step(k)), device="cuda") for k in range(nsteps)]
    for k in range(5):  # BDF1 start-up + graph capture
        ens.step_device(us[k])
    ens.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for k in range(5, nsteps):
            ens.step_device(us[k])
        e1.record(stream)
    ens.synchronize()
    ms = e0.elapsed_time(e1) / (nsteps - 5)
    up = ens.fields(0)
    y = ens.measurement().copy()
    prof = [ens.profile_step(u_of_step(nsteps - 1)) for _ in range(4)][-1]
    rec = {
        "config": label, "dofs": int(tab.N), "cells": int(tab.nT), "B": int(B), "steps_timed": nsteps - 5, "ms_per_step": ms,
        "trajectory_steps_per_s": B / (ms * 1e-3), "finite": bool(np.isfinite(up).all() and np.isfinite(y).all()),
        "diverged": int(ens.diverged.sum()), "create_s": create_s,
        "phase_ms": {k: v["ms"] for k, v in prof.items()}, "factor_entries": int(prob.sym.factor_entries()),
    }
    if same_pair is not None:
        a, b = same_pair
        rec["identical_inputs_bit_identical"] = bool(np.array_equal(up[:, a], up[:, b]))
    ens.close()
    # the same trajectories in a 32-wide ensemble
    pick = np.linspace(0, B - 1, 32).astype(int) if B > 32 else None
    if pick is not None:
        e32 = Ensemble(prob, 32)
        ic32 = ic if ic.ndim == 1 else ic[:, pick]
        e32.set_state(ic32[: tab.Nv], None, ic32[tab.Nv :], order=1)
        for k in range(nsteps):
            e32.step(np.ascontiguousarray(u_of_step(k)[:, pick]))
        up32 = e32.fields(0)
        # the profile steps above advanced the wide ensemble further: re-run it to the same step count
        ens = Ensemble(prob, B)
        ens.set_state(ic[: tab.Nv] if ic.ndim == 1 else ic[: tab.Nv, :], None, ic[tab.Nv :] if ic.ndim == 1 else ic[tab.Nv :, :], order=1)
        for k in range(nsteps):
            ens.step_device(us[k])
        upw = ens.fields(0)[:, pick]
        ens.close()
        e32.close()
        rec["wide_vs_32_rel_err"] = float(np.linalg.norm(upw - up32) / np.linalg.norm(up32))
    print(json.dumps(rec), flush=True)
    return rec


def pinball(B=512, nsteps=40):
    from flowcontrol_b200.actuator import CYLINDER_ACTUATION_MODE
    from flowcontrol_b200.examples.pinball import PinballFlowSolver

    UP0 = np.load(GOLD / "pinball_baseflow.npz")["UP0"]
    fs = PinballFlowSolver.make_default(Re=30.0, mode_actuation=CYLINDER_ACTUATION_MODE.ROTATION, path_out=Path(tempfile.mkdtemp()))
    tab = fs.tables
    fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
    prob = FlowProblem(tab, fs.blocks, 30.0, 0.005, fs.bc.bcu, fs.params_control.actuator_list, fs.params_control.sensor_list, UP0)
    rng = np.random.default_rng(0)
    amp = rng.uniform(-2, 2, size=(3, B))  # SURVEY config 3: a_bk ~ U(-2,2), seed 0
    amp[:, 1] = amp[:, 0]                   # two trajectories with identical inputs
    tk = np.array([0.05, 0.1, 0.15])        # pulse centres inside the timed window (the config's 0.25/0.5/0.75 s scaled)

    def u(k):
        t = (k + 1) * 0.005
        return amp * np.exp(-0.5 * (t - tk[:, None]) ** 2 / 0.02**2)

    ic = fs._default_initial_perturbation()
    return run_open_loop(prob, B, ic, u, nsteps, "configs[2] fluidic pinball (mesh_middle), 3 rotation actuators, Gaussian pulses", (0, 1))


def cavity(B=256, nsteps=40):
    from flowcontrol_b200.examples.cavity import CavityFlowSolver

    UP0 = np.load(GOLD / "cavity_baseflow.npz")["UP0"]
    fs = CavityFlowSolver.make_default(path_out=Path(tempfile.mkdtemp()))
    tab = fs.tables
    fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
    prob = FlowProblem(tab, fs.blocks, 7500.0, 0.0004, fs.bc.bcu, fs.params_control.actuator_list, fs.params_control.sensor_list, UP0)
    ic = fs._default_initial_perturbation()
    # closed loop with static gains u = -k_b y_1, k_b log-spaced 1e-3..1e-1 (SURVEY config 4), on the device
    gains = np.logspace(-3, -1, B)
    gains[1] = gains[0]
    ctrls = [Controller(np.array([[-1.0]]), np.zeros((1, 1)), np.zeros((1, 1)), np.array([[-g]])) for g in gains]
    Ky = np.zeros((1, prob.ns)); Ky[0, 0] = 1.0
    ens = Ensemble(prob, B)
    ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    ens.set_controllers(ControllerBank(ctrls, prob.dt, Ky, np.ones((prob.na, 1))))
    stream = torch.cuda.ExternalStream(ens.stream)
    ens.run_closed_loop(5, log=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        series = ens.run_closed_loop(nsteps)
        e1.record(stream)
    ens.synchronize()
    ms = e0.elapsed_time(e1) / nsteps
    up = ens.fields(0)
    prof = [ens.profile_step(np.zeros((prob.na, B))) for _ in range(4)][-1]
    rec = {
        "config": "configs[3] open cavity Re=7500 (cavity_coarse), force actuator + wall-shear sensor, static-gain closed loop on the device",
        "dofs": int(tab.N), "cells": int(tab.nT), "B": int(B), "steps_timed": nsteps, "ms_per_step": ms,
        "trajectory_steps_per_s": B / (ms * 1e-3), "finite": bool(np.isfinite(series).all() and np.isfinite(up).all()),
        "diverged": int(ens.diverged.sum()), "phase_ms": {k: v["ms"] for k, v in prof.items()},
        "factor_entries": int(prob.sym.factor_entries()),
        "identical_inputs_bit_identical": bool(np.array_equal(series[:, :, 0], series[:, :, 1])),
        # u = -k y_1 with the pre-update measurement: column 1 is u_ctrl, column 2 is y_1 of the same step's end
        "static_gain_law_holds": bool(np.allclose(series[1:, 1, :], -gains[None, :] * series[:-1, 2, :], rtol=1e-12, atol=0)),
    }
    ens.close()
    print(json.dumps(rec), flush=True)
    return rec


def lidcavity(B=1024, nsteps=60):
    from flowcontrol_b200.examples import lidcavity as ex

    UP0 = np.load(GOLD / "lidcavity_baseflow.npz")["UP0"]
    prob = ex.make_problem(Re=1000.0, UP0=UP0)
    tab = prob.tab
    fs = ex.LidCavityFlowSolver.make_default(Re=1000.0, path_out=Path(tempfile.mkdtemp()))
    fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
    rng = np.random.default_rng(0)  # SURVEY config 5: ICs ParamIC(xloc, yloc ~ U(0.2,0.8), radius 0.1, amplitude 0.1)
    loc = rng.uniform(0.2, 0.8, size=(B, 2))
    loc[1] = loc[0]
    ic = np.stack([0.1 * fs._default_initial_perturbation(xloc=x, yloc=y, radius=0.1) for x, y in loc], axis=1)
    ic[tab.Nv :, :] = 0.0
    u0 = np.zeros((prob.na, B))
    return run_open_loop(prob, B, ic, lambda k: u0, nsteps, "configs[4] lid-driven cavity (mesh64), open loop, random Gaussian-vortex ICs", (0, 1))


def cylinder_single(nsteps=200):
    sys.path.insert(0, str(ROOT / "tools"))
    from gpu_check import build_cylinder_problem, default_ic

    prob, UP0 = build_cylinder_problem()
    ic = default_ic(prob.tab, UP0, xloc=2.0, yloc=0.0, radius=0.5, amp=1.0)
    u0 = np.zeros((2, 1))
    return run_open_loop(prob, 1, ic, lambda k: u0, nsteps, "configs[0] cylinder Re=100, single open-loop trajectory (latency case)", None)


if __name__ == "__main__":
    out = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "gpurun_out" / "r01_configs.json"
    which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["cylinder1", "lidcavity", "cavity", "pinball"]
    fns = {"cylinder1": cylinder_single, "lidcavity": lidcavity, "cavity": cavity, "pinball": pinball}
    recs = []
    for name in which:
        t0 = time.time()
        r = fns[name]()
        r["wall_s_total"] = time.time() - t0
        recs.append(r)
        out.parent.mkdir(exist_ok=True)
        out.write_text(json.dumps({"device": torch.cuda.get_device_name(0), "runs": recs}, indent=1))
