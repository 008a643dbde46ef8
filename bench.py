"""Headline benchmark: trajectory-steps/s of the closed-loop cylinder ensemble (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...   # CPU stand-in for the reference path

One "step" advances every trajectory of the ensemble by one closed-loop time step:
controller update -> RHS assembly -> sparse direct solve -> sensors/energy -> log.
Weak scaling: every GPU owns 256 trajectories (its own gain-swept controller family);
there is no collective inside the step, only an all-gather of the time series at the end.
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
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


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


def controller_bank(prob, lo: int, hi: int, total: int):
    from flowcontrol_b200.controller import Controller, ControllerBank
    from flowcontrol_b200.sharding import controller_gain_sweep

    k = np.load(ROOT / "tests/golden/Kopt_reduced13.npz")
    gains = controller_gain_sweep(total)[lo:hi]
    ctrls = [Controller(k["A"], g * k["B"], k["C"], g * k["D"]) for g in gains]
    return ControllerBank(ctrls, prob.dt, Ky=np.array([[-1.0, 0.0, 0.0]]), Fu=np.array([[1.0], [1.0]]))


class HostBank:
    """Vectorised host-side controller bank for the end-to-end (host buffers) measurement."""

    def __init__(self, bank):
        B = bank.B
        self.Ad = bank.Ad.T.reshape(B, bank.nx, bank.nx).copy()
        self.Bd = bank.Bd.T.reshape(B, bank.nx, bank.ny).copy()
        self.Cd = bank.Cd.T.reshape(B, bank.nu, bank.nx).copy()
        self.Dd = bank.Dd.T.reshape(B, bank.nu, bank.ny).copy()
        self.x = bank.x0.T.copy()
        self.Ky, self.Fu = bank.Ky, bank.Fu

    def step(self, y_meas):  # y_meas [ns, B] -> u_ctrl [na, B]
        v = (self.Ky @ y_meas).T  # [B, ny]
        u = np.einsum("bij,bj->bi", self.Cd, self.x) + np.einsum("bij,bj->bi", self.Dd, v)
        self.x = np.einsum("bij,bj->bi", self.Ad, self.x) + np.einsum("bij,bj->bi", self.Bd, v)
        return np.ascontiguousarray((self.Fu @ u.T))


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
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    from flowcontrol_b200.ensemble import Ensemble
    from flowcontrol_b200.sharding import gather_series, shard_bounds

    fs, prob = build_problem(args.time_scheme)
    tab = prob.tab
    total = B_PER_GPU * world
    lo, hi = shard_bounds(total, rank, world)
    B = hi - lo
    bank = controller_bank(prob, lo, hi, total)
    ens = Ensemble(prob, B, device=local_rank)
    ic = fs._default_initial_perturbation()
    ens.set_state(ic[: tab.Nv], None, ic[tab.Nv :], order=1)
    ens.set_controllers(bank)
    stream = torch.cuda.ExternalStream(ens.stream, device=torch.device("cuda", local_rank))
    ncol = 1 + prob.na + prob.ns
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident closed loop (value) -------------------------------------------------
    ens.run_closed_loop(max(W, 3), log=False)  # warm-up: BDF1 start-up step + graph capture + steady BDF2
    series_dev = torch.empty((K, ncol, B), dtype=torch.float64, device="cuda")
    # one untimed logged pass of the same length: the library sizes its device-side series buffer and re-captures the
    # step graphs for it on first use; neither belongs in the timed region
    ens.run_closed_loop(K, log=True, out=series_dev)
    if world > 1:
        gather_series(torch.zeros_like(series_dev), total)  # warm-up: NCCL communicator + buffers exist before the timed region
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ens.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        ens.run_closed_loop(K, log=True, out=series_dev)
        gathered = gather_series(series_dev, total) if world > 1 else series_dev
        e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    launches = ens.launch_count() - l0
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    ms = float(ms.item())
    value = total * K / (ms * 1e-3)
    finite = bool(torch.isfinite(gathered).all().item())

    # ---- per-phase device times for the roofline of the dominant kernel -------------------------
    prof_runs = [ens.profile_step(np.zeros((prob.na, B))) for _ in range(6)][1:]
    phase_ms = {k: float(np.mean([p[k]["ms"] for p in prof_runs])) for k in prof_runs[0]}
    plan = prob.plans[2]
    n = plan.n
    solve_ms = phase_ms["forward"] + phase_ms["backward"]
    solve_launches = prof_runs[0]["forward"]["launches"] + prof_runs[0]["backward"]["launches"]
    ldb = (B + 31) // 32 * 32
    # algorithmic bytes of one solve (all launches of both sweeps): factor values + gather indices read
    # once; b read, y written, update vectors written and read once (forward); y read, x written (backward)
    solve_bytes = 8 * plan.nnz + 4 * len(plan.i0) + 8 * (4 * n + 2 * plan.nU) * ldb
    solve_flops = 2.0 * plan.nnz * ldb
    hbm_peak, peak_src = peaks()
    achieved = solve_bytes / (solve_ms * 1e-3) / 1e9
    # measured DRAM traffic of the same launches (ncu, profiles/r01_kernel_summary.json written by tools/kernel_summary.py)
    traffic = None
    ks = ROOT / "profiles" / "r01_kernel_summary.json"
    if ks.exists():
        try:
            traffic = float(json.loads(ks.read_text())["kernels"]["k_front_sweep"]["dram_bytes"])
        except Exception:
            traffic = None
    # element kernel (second largest): reads u and b_{n-1}, writes the next rhs rows and b_n; ~600 FP64 FMA per cell
    elem_bytes = 8 * (3 * tab.Nv + n) * ldb
    elem_flops = 2.0 * 600 * tab.nT * ldb
    roofline = {
        "kernel": "k_front_sweep (multifrontal forward+backward sweeps on FP64 tensor cores, all launches of one solve)",
        "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
        "peak_source": peak_src, "traffic": traffic,
        "traffic_note": "dram__bytes_read+write summed over the sweep launches of one step (ncu); algorithmic bytes assume the factor is read once",
        "element_kernel": {"ms_per_step": phase_ms["element"], "achieved_gbs": elem_bytes / (phase_ms["element"] * 1e-3) / 1e9,
                           "hbm_frac": elem_bytes / (phase_ms["element"] * 1e-3) / 1e9 / hbm_peak,
                           "fp64_tflops": elem_flops / (phase_ms["element"] * 1e-3) / 1e12,
                           "fp64_frac_of_dfma_peak": elem_flops / (phase_ms["element"] * 1e-3) / 1e12 / 33.6,
                           "algorithmic_bytes_per_step": elem_bytes},
        "launches_per_step": solve_launches, "ms_per_launch": solve_ms / solve_launches, "ms_per_step": solve_ms,
        "algorithmic_bytes_per_step": solve_bytes,
        "fp64": {"achieved_tflops": solve_flops / (solve_ms * 1e-3) / 1e12, "peak_tflops": FP64_PEAK_TFLOPS,
                 "frac": solve_flops / (solve_ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
                 "note": "the sweeps run on the FP64 tensor pipe (mma.sync m8n8k4); peak = measured DMMA rate (profiles/r01_fp64_peak.log)"},
        "phase_ms": phase_ms,
    }
    if prob.time_scheme == "cn":
        # k_spmm_mma: packed operator read once + one 8-byte read per input row and one write per output row and trajectory
        sp_bytes = 12 * prob.E_cn.nnz + 4 * (n + 1) + 8 * (tab.Nv + n) * ldb
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
    e2e = {"value": total * K / float(t_e2e.item()), "unit": UNIT, "h2d_bytes_per_step": int(world * prob.na * B * 8),
           "d2h_bytes_per_step": int(world * (prob.ns * B * 8 + B * 8 + B * 4)),
           "note": "FlowSolver/Ensemble.step with host numpy buffers, controller bank stepped on the host"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_bench import time_oracle

        cpu = time_oracle(nsteps=20, warmup=3)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    ens.close()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (shipped mesh O1, cached base flow, gain-swept controllers, default ParamIC perturbation)",
            "config": {"workload": WORKLOAD + (" [time_scheme=cn variant]" if args.time_scheme == "cn" else ""), "trajectories_per_gpu": B_PER_GPU, "trajectories_total": total,
                       "dofs_per_trajectory": int(tab.N), "warmup_note": "W steps + one untimed logged pass of K steps (series buffer allocation, graph capture)", "l2": "working set >> L2 (solve buffer 560 MB + packed factors 130 MB + state 430 MB per GPU vs 126 MB L2); no flush needed",
                       "parallelism": f"ensemble-sharded x{world}, time-series all-gather only"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches * world), "roofline": roofline,
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
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
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
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
