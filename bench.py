"""Headline benchmark: trajectory-steps/s of the closed-loop cylinder ensemble (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...   # CPU stand-in for the reference path

One "step" advances every trajectory of the ensemble by one closed-loop time step:
controller update -> RHS assembly -> sparse direct solve -> sensors/energy -> log.
Weak scaling (default): every GPU owns 256 trajectories (its own gain-swept controller family); `--scaling strong
--trajectories T` shards a FIXED ensemble of T trajectories over the GPUs instead (BASELINE configs[2]/[4] ask for 512 /
1024 trajectories across 1/2/4/8 GPUs).  There is no collective inside the step, only an all-gather of the time series
at the end (timed inside `value`, reported separately as `allgather_ms`).  Without torchrun, `--gpus N` (N > 1)
re-launches this script under `python -m torch.distributed.run` with N ranks.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

B_PER_GPU = 256
METRIC = "trajectory-steps/s (cylinder Re=100 closed-loop ensemble)"
UNIT = "trajectory-steps/s"
WORKLOAD = "cylinder O1 Re=100 dt=0.005, closed loop, 256 gain-swept 13-state LTI controllers per GPU (BASELINE configs[1])"
FP64_PEAK_TFLOPS = 37.1  # measured on this pool's B200 with tools/bench_src/fp64_peak.cu (DMMA m8n8k4; DFMA: 33.6)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None  # perf_counter bounds of the timed region (samples are stamped on arrival)

    def window(self, t0: float, t1: float):
        self.t0, self.t1 = t0, t1

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(2)
        # the sampler runs from before the warm-up (so that no process is forked near the timed region); the samples that
        # arrived inside the timed window (+- one period) are the ones reported, the rest only back them up when the
        # window was shorter than two sampling periods
        rows = [r for t, r in self.rows if self.t0 is None or (self.t0 - 0.03 <= t <= self.t1 + 0.03)]
        scope = "timed region"
        if len(rows) < 2:
            rows = [r for t, r in self.rows if self.t0 is None or t >= self.t0 - 2.0]
            scope = "timed region and the 2 s of warm-up load before it (the region was shorter than two 20 ms samples)"
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "scope": scope}


def build_problem(time_scheme="bdf"):
    import tempfile

    from flowcontrol_b200.examples.cylinder import CylinderFlowSolver
    from flowcontrol_b200.flowfield import Field
    from flowcontrol_b200.problem import FlowProblem

    UP0 = np.load(ROOT / "tests/golden/cylinder_baseflow.npz")["UP0"]
    fs = CylinderFlowSolver.make_default(path_out=Path(tempfile.mkdtemp()))
    tab = fs.tables
    fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
    prob = FlowProblem(tab, fs.blocks, 100.0, 0.005, fs.bc.bcu, fs.params_control.actuator_list,
                       fs.params_control.sensor_list, UP0, time_scheme=time_scheme)
    return fs, prob


def build_workload(name: str):
    """Other BASELINE configurations as strong-scaling workloads (open loop): returns (problem, ic [N] or fn(lo, hi) -> [N, B],
    u_of_step(k, lo, hi) -> [na, B], description)."""
    import tempfile

    from flowcontrol_b200.flowfield import Field
    from flowcontrol_b200.problem import FlowProblem

    sys.path.insert(0, str(ROOT / "tools"))
    import make_goldens_r2 as mk  # amplitudes / initial conditions of the configurations (shared with the parity tests)

    if name == "pinball":
        from flowcontrol_b200.actuator import CYLINDER_ACTUATION_MODE
        from flowcontrol_b200.examples.pinball import PinballFlowSolver

        UP0 = np.load(ROOT / "tests/golden/pinball_Re100_baseflow.npz")["UP0"]
        fs = PinballFlowSolver.make_default(Re=100.0, mode_actuation=CYLINDER_ACTUATION_MODE.ROTATION, path_out=Path(tempfile.mkdtemp()))
        tab = fs.tables
        fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))
        prob = FlowProblem(tab, fs.blocks, 100.0, 0.005, fs.bc.bcu, fs.params_control.actuator_list, fs.params_control.sensor_list, UP0)
        ic = fs._default_initial_perturbation(2.0, 0.0, 0.5)
        return prob, (lambda lo, hi, total: ic), (lambda k, lo, hi, total: mk.pinball_u((k + 1) * 0.005, mk.pinball_amplitudes(total)[:, lo:hi])), \
            "fluidic pinball Re=100 (mesh_middle), three rotation actuators with Gaussian pulses, open loop (BASELINE configs[2])"
    if name == "lidcavity":
        from flowcontrol_b200.examples import lidcavity as ex

        UP0 = np.load(ROOT / "tests/golden/lidcavity_Re8000_baseflow.npz")["UP0"]
        prob = ex.make_problem(Re=8000.0, UP0=UP0)
        tab = prob.tab
        fs = ex.LidCavityFlowSolver.make_default(Re=8000.0, path_out=Path(tempfile.mkdtemp()))
        fs._assign_steady_state(Field(UP0[: tab.Nv]), Field(UP0[tab.Nv :]))

        def ic(lo, hi, total):
            loc = mk.lid_ics(total)[lo:hi]
            return np.stack([0.1 * fs._default_initial_perturbation(xloc=x, yloc=y, radius=0.1) for x, y in loc], axis=1)

        return prob, ic, (lambda k, lo, hi, total: np.zeros((1, hi - lo))), \
            "lid-driven cavity Re=8000 (mesh64), random Gaussian-vortex initial conditions, open loop (BASELINE configs[4])"
    raise ValueError(name)


def run_open_workload(args):
    """Strong / weak scaling of the open-loop configurations (pinball B=512, lid cavity B=1024): same timing rules as the
    headline benchmark, the device-resident loop is fcb_run_open_loop."""
    import torch

    import __graft_entry__ as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        g.build()
    if world > 1:
        dist.barrier()
    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.sharding import SeriesGatherer, shard_bounds

    prob, ic_of, u_of, desc = build_workload(args.workload)
    tab = prob.tab
    total = args.trajectories if args.scaling == "strong" else args.trajectories * world
    lo, hi = shard_bounds(total, rank, world)
    B = hi - lo
    K, W = args.steps, max(args.warmup, 3)
    ens = Ensemble(prob, B, device=local_rank)
    ic = ic_of(lo, hi, total)
    ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    useries = torch.as_tensor(np.stack([u_of(k, lo, hi, total) for k in range(W + 2 * K)]), device="cuda").contiguous()
    stream = torch.cuda.ExternalStream(ens.stream, device=torch.device("cuda", local_rank))
    ncol = 1 + prob.na + prob.ns

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    series_dev = torch.empty((K, ncol, B), dtype=torch.float64, device="cuda")
    gatherer = SeriesGatherer(K, ncol, total, torch.float64, torch.device("cuda", local_rank)) if world > 1 else None
    with torch.cuda.stream(stream):
        ens.run_open_loop(useries[:W], log=False)
        ens.run_open_loop(useries[W : W + K], log=True, out=series_dev)
        if gatherer:
            gatherer(series_dev)
    barrier()
    l0 = ens.launch_count()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t_begin = time.perf_counter()
    with torch.cuda.stream(stream):
        e0.record(stream)
        ens.run_open_loop(useries[W + K :], log=True, out=series_dev)
        e1.record(stream)
        gathered = gatherer(series_dev) if gatherer else series_dev
        e2.record(stream)
    barrier()
    sampler.window(t_begin, time.perf_counter())
    t = torch.tensor([e0.elapsed_time(e2), e1.elapsed_time(e2)], dtype=torch.float64, device="cuda")
    launches = ens.launch_count() - l0
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, allgather_ms = (float(v) for v in t.tolist())
    finite = bool(torch.isfinite(gathered).all().item())
    prof = [ens.profile_step(np.zeros((prob.na, B))) for _ in range(4)][1:]
    phase_ms = {k: float(np.mean([p[k]["ms"] for p in prof])) for k in prof[0]}
    # end to end with host buffers
    u_host = [np.ascontiguousarray(u_of(k, lo, hi, total)) for k in range(K)]
    for k in range(3):
        ens.step(u_host[k])
    barrier()
    t0 = time.perf_counter()
    for k in range(K):
        ens.step(u_host[k])
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    ens.close()
    if rank == 0:
        print(json.dumps({
            "metric": f"trajectory-steps/s ({args.workload} open-loop ensemble)", "value": total * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (shipped mesh, committed base-flow fixture, seeded amplitudes / initial conditions of the configuration)",
            "config": {"workload": desc, "trajectories_total": total, "trajectories_per_gpu": B, "dofs_per_trajectory": int(tab.N),
                       "l2": "working set >> L2; no flush needed", "parallelism": f"ensemble-sharded x{world}, time-series all-gather only"},
            "clocks": clocks, "e2e": {"value": total * K / float(t_e2e.item()), "unit": UNIT, "h2d_bytes_per_step": int(total * prob.na * 8),
                                      "d2h_bytes_per_step": int(total * (prob.ns * 8 + 12))},
            "gpu_launches": int(launches * world), "allgather_ms": allgather_ms, "phase_ms": phase_ms, "all_finite": finite}))
    if world > 1:
        dist.destroy_process_group()


def controller_bank(prob, lo: int, hi: int, total: int):
    from flowcontrol_b200.controller import Controller, ControllerBank
    from flowcontrol_b200.sharding import controller_gain_sweep

    k = np.load(ROOT / "tests/golden/Kopt_reduced13.npz")
    gains = controller_gain_sweep(total)[lo:hi]
    ctrls = [Controller(k["A"], g * k["B"], k["C"], g * k["D"]) for g in gains]
    return ControllerBank(ctrls, prob.dt, Ky=np.array([[-1.0, 0.0, 0.0]]), Fu=np.array([[1.0], [1.0]]))


class HostBank:
    """Vectorised host-side controller bank for the end-to-end (host buffers) measurement: u = Cd x + Dd v, x <- Ad x + Bd v
    for every trajectory (controller.py:157-158, pre-update state in the output) as ONE contraction with the stacked matrix
    [[Ad, Bd], [Cd, Dd]], stored trajectory-innermost like the bank itself."""

    def __init__(self, bank):
        B, nx, ny, nu = bank.B, bank.nx, bank.ny, bank.nu
        M = np.zeros((nx + nu, nx + ny, B))  # M[i, j, b]
        M[:nx, :nx] = bank.Ad.reshape(nx, nx, B)
        M[:nx, nx:] = bank.Bd.reshape(nx, ny, B)
        M[nx:, :nx] = bank.Cd.reshape(nu, nx, B)
        M[nx:, nx:] = bank.Dd.reshape(nu, ny, B)
        self.T = np.ascontiguousarray(M.transpose(1, 0, 2))  # [j, i, b]
        self.z = np.zeros((nx + ny, B))
        self.z[:nx] = bank.x0
        self.nx = nx
        self.Ky, self.Fu = bank.Ky, bank.Fu

    def step(self, y_meas):  # y_meas [ns, B] -> u_ctrl [na, B]
        self.z[self.nx :] = self.Ky @ y_meas
        out = np.einsum("jib,jb->ib", self.T, self.z)
        self.z[: self.nx] = out[: self.nx]
        return np.ascontiguousarray(self.Fu @ out[self.nx :])


def newest_kernel_summary():
    """Latest offline ncu summary (tools/kernel_summary.py) with the measured DRAM traffic of the sweep launches."""
    for path in sorted((ROOT / "profiles").glob("r*_kernel_summary.json"), reverse=True):
        try:
            k = json.loads(path.read_text())["kernels"]["k_front_sweep"]
            return float(k["dram_bytes"]), f"profiles/{path.name}"
        except Exception:
            continue
    return None, None


def run_ours(args):
    import torch

    import __graft_entry__ as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    # the clock sampler (a forked nvidia-smi) starts here, long before the timed region, so that no rank forks a process
    # or does anything the others do not between the barrier and the first event
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.sharding import SeriesGatherer, shard_bounds

    fs, prob = build_problem(args.time_scheme)
    tab = prob.tab
    strong = args.scaling == "strong"
    total = args.trajectories if strong else B_PER_GPU * world
    lo, hi = shard_bounds(total, rank, world)
    B = hi - lo
    bank = controller_bank(prob, lo, hi, total)
    ens = Ensemble(prob, B, device=local_rank)
    ic = fs._default_initial_perturbation()
    ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    ens.set_controllers(bank)
    stream = torch.cuda.ExternalStream(ens.stream, device=torch.device("cuda", local_rank))
    ncol = 1 + prob.na + prob.ns
    K, W = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident closed loop (value) -------------------------------------------------
    ens.run_closed_loop(W, log=False)  # warm-up: BDF1 start-up step + graph capture + steady BDF2
    series_dev = torch.empty((K, ncol, B), dtype=torch.float64, device="cuda")
    gatherer = SeriesGatherer(K, ncol, total, torch.float64, torch.device("cuda", local_rank)) if world > 1 else None
    # one untimed logged pass of the same length (the library allocates its series chunk buffer on first use) and one
    # untimed all-gather (NCCL communicator, channels and the gatherer's buffers exist before the timed region)
    with torch.cuda.stream(stream):
        ens.run_closed_loop(K, log=True, out=series_dev)
        if gatherer:
            gatherer(series_dev)
    barrier()
    l0 = ens.launch_count()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t_begin = time.perf_counter()
    with torch.cuda.stream(stream):
        e0.record(stream)
        ens.run_closed_loop(K, log=True, out=series_dev)
        e1.record(stream)
        gathered = gatherer(series_dev) if gatherer else series_dev
        e2.record(stream)
    barrier()
    sampler.window(t_begin, time.perf_counter())
    t = torch.tensor([e0.elapsed_time(e2), e1.elapsed_time(e2), e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    launches = ens.launch_count() - l0
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, allgather_ms, loop_ms = (float(v) for v in t.tolist())
    value = total * K / (ms * 1e-3)
    finite = bool(torch.isfinite(gathered).all().item())

    # ---- per-phase device times for the roofline of the dominant kernel -------------------------
    prof_runs = [ens.profile_step(np.zeros((prob.na, B))) for _ in range(6)][1:]
    phase_ms = {k: float(np.mean([p[k]["ms"] for p in prof_runs])) for k in prof_runs[0]}
    plan = prob.plans[2]
    n = plan.n
    solve_ms = phase_ms["forward"] + phase_ms["backward"]
    solve_launches = prof_runs[0]["forward"]["launches"] + prof_runs[0]["backward"]["launches"]
    ldb = 32 if B <= 32 else (64 if B <= 64 else (B + 127) // 128 * 128)
    # ALGORITHMIC bytes of one constant-LHS solve, SURVEY.md section 8(d) "Factor solve" (unique-touch model):
    #   12 bytes per factor entry (value + index; the dense-block factor stores no per-entry index, the formula is kept as
    #   the survey states it) read once per step + 8 * 2N per trajectory (right-hand side read, solution written).
    # The update vectors of the multifrontal sweeps and the once-per-slab re-reads of the factor are NOT algorithmic:
    # they show up in `traffic` (ncu DRAM bytes of the same launches) and in `traffic_over_algorithmic`.
    solve_bytes = 12 * plan.nnz + 16 * tab.N * B
    solve_flops = 2.0 * plan.nnz * ldb
    hbm_peak, peak_src = peaks()
    achieved = solve_bytes / (solve_ms * 1e-3) / 1e9
    traffic, traffic_src = newest_kernel_summary() if B == B_PER_GPU else (None, None)
    # internal byte model of the implementation (factor + gather lists once, rhs/y/x and the update vectors once each way)
    impl_bytes = 8 * plan.nnz + 4 * len(plan.i0) + 8 * (4 * n + 2 * plan.nU) * ldb
    # element kernel (second largest), SURVEY 8(d) rhs_assemble: 8 (4 Nv + N) state bytes per trajectory + geometry/index bytes
    elem_bytes = 8 * (4 * tab.Nv + tab.N) * B + tab.nT * (24 + 40)
    elem_flops = 2.0 * 600 * tab.nT * ldb
    roofline = {
        "kernel": "k_front_sweep (multifrontal forward+backward sweeps on FP64 tensor cores, all launches of one solve)",
        "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
        "peak_source": peak_src,
        "bytes_model": "SURVEY.md 8(d) factor solve: 12*nnz(factor) + 16*N*B per step (unique touch)",
        "algorithmic_bytes_per_step": solve_bytes,
        "traffic": traffic, "traffic_source": (f"offline ncu capture ({traffic_src}: dram__bytes_read+write over the sweep launches of "
                                               "one step at B=256), not measured in this run") if traffic else None,
        "traffic_over_algorithmic": (traffic / solve_bytes) if traffic else None,
        "implementation_bytes_per_step": impl_bytes,
        "implementation_gbs": impl_bytes / (solve_ms * 1e-3) / 1e9,
        "timing": "CUDA events between the phases of un-graphed profile steps on the handle's stream (mean of 5, after the timed region)",
        "element_kernel": {"ms_per_step": phase_ms["element"], "achieved_gbs": elem_bytes / (phase_ms["element"] * 1e-3) / 1e9,
                           "hbm_frac": elem_bytes / (phase_ms["element"] * 1e-3) / 1e9 / hbm_peak,
                           "fp64_tflops": elem_flops / (phase_ms["element"] * 1e-3) / 1e12,
                           "fp64_frac_of_dfma_peak": elem_flops / (phase_ms["element"] * 1e-3) / 1e12 / 33.6,
                           "algorithmic_bytes_per_step": elem_bytes,
                           "bytes_model": "SURVEY.md 8(d) rhs_assemble: 8*(4*Nv + N)*B + nT*64"},
        "launches_per_step": solve_launches, "ms_per_launch": solve_ms / solve_launches, "ms_per_step": solve_ms,
        "fp64": {"achieved_tflops": solve_flops / (solve_ms * 1e-3) / 1e12, "peak_tflops": FP64_PEAK_TFLOPS,
                 "frac": solve_flops / (solve_ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
                 "note": "the sweeps run on the FP64 tensor pipe (mma.sync m8n8k4); peak = measured DMMA rate (profiles/r01_fp64_peak.log); "
                         "at B=256 the DMMA floor of one solve is 2*nnz*B / peak"},
        "phase_ms": phase_ms,
    }
    if prob.time_scheme == "cn":
        # k_spmm_mma, SURVEY 8(d) SpMM: (12 nnz + 4 (n+1)) + 8 (n_in + n_out) B
        sp_bytes = 12 * prob.E_cn.nnz + 4 * (n + 1) + 8 * (tab.Nv + n) * B
        roofline["spmm_kernel"] = {"ms_per_step": phase_ms["spmm"], "algorithmic_bytes_per_step": sp_bytes,
                                   "achieved_gbs": sp_bytes / (phase_ms["spmm"] * 1e-3) / 1e9,
                                   "hbm_frac": sp_bytes / (phase_ms["spmm"] * 1e-3) / 1e9 / hbm_peak}

    # ---- end to end through the public API with host buffers ---------------------------------------
    host_bank = HostBank(bank)
    y = ens.measurement().copy()
    for _ in range(3):
        y = ens.step(host_bank.step(y)).copy()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        y = ens.step(host_bank.step(y)).copy()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e = {"value": total * K / float(t_e2e.item()), "unit": UNIT, "h2d_bytes_per_step": int(total * prob.na * 8),
           "d2h_bytes_per_step": int(total * (prob.ns * 8 + 8 + 4)),
           "note": "FlowSolver/Ensemble.step with host numpy buffers, controller bank stepped on the host"}

    clocks = sampler.stop() if rank == 0 else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_bench import time_oracle

        cpu = time_oracle(nsteps=20, warmup=3)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "threads_per_worker", "per_core_steps_per_s")}
    ens.close()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (shipped mesh O1, cached base flow, gain-swept controllers, default ParamIC perturbation)",
            "config": {"workload": WORKLOAD + (" [time_scheme=cn variant]" if args.time_scheme == "cn" else "")
                       + (f" [strong scaling: {total} trajectories in total]" if strong else ""),
                       "trajectories_per_gpu": B if strong else B_PER_GPU, "trajectories_total": total,
                       "dofs_per_trajectory": int(tab.N), "warmup_note": "W steps + one untimed logged pass of K steps and one untimed all-gather (buffers, graph capture, NCCL channels)", "l2": "working set >> L2 (solve buffer 560 MB + packed factors 130 MB + state 430 MB per GPU vs 126 MB L2); no flush needed",
                       "parallelism": f"ensemble-sharded x{world}, time-series all-gather only"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches * world), "roofline": roofline,
            "allgather_ms": allgather_ms, "loop_ms": loop_ms,
            "timed_region": "K graph-replayed closed-loop steps + the all-gather of the [K, ncol, B] series (allgather_ms, a fixed cost per run: "
                            "it is inside `value`, so `value` at small K understates the steady step rate loop_ms / K)",
            "all_finite": finite,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """CPU stand-in for the reference path on all host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.cpu_bench import time_oracle

    r = time_oracle(nsteps=max(args.steps, 1), warmup=max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", args.gpus)),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["wall_s"] / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (same mesh, base flow, controllers and IC as the CUDA arm)",
        "config": {"workload": WORKLOAD, "note": "FEniCS/PETSc/MUMPS cannot be installed here; numpy/scipy (SuperLU) port of the same step, "
                   "one trajectory per host core"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "threads_per_worker", "per_core_steps_per_s")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--time-scheme", default="bdf", choices=["bdf", "cn"],
                    help="bdf = the reference's default BDF1->BDF2 (the benchmark); cn = Crank-Nicolson variant (adds the SpMM kernel's roofline)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 256 trajectories per GPU (the benchmark); strong: a fixed ensemble of --trajectories sharded over the GPUs")
    ap.add_argument("--trajectories", type=int, default=512, help="total ensemble width for --scaling strong (per GPU for weak scaling of --workload pinball/lidcavity)")
    ap.add_argument("--workload", default="cylinder", choices=["cylinder", "pinball", "lidcavity"],
                    help="cylinder = the headline benchmark (BASELINE configs[1]); pinball / lidcavity = configs[2] / configs[4] as open-loop scaling runs")
    a = ap.parse_args()
    if a.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # launched by hand without torchrun: spawn the ranks ourselves (one process per GPU, NCCL rendezvous on 127.0.0.1)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", os.environ.get("MASTER_PORT", "29517"), str(Path(__file__).resolve())] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) != a.gpus:
        sys.exit(f"bench.py: --gpus {a.gpus} does not match WORLD_SIZE={os.environ['WORLD_SIZE']}")
    if a.impl == "reference":
        run_reference(a)
    elif a.workload != "cylinder":
        run_open_workload(a)
    else:
        run_ours(a)
