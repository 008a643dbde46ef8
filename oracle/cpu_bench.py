"""CPU ORACLE (test infrastructure): timing harness for the numpy/scipy restatement.

Used only by bench.py's ``cpu_baseline`` leg and by ``bench.py --impl reference``.  The real
reference path (FEniCS 2019.1.0 + PETSc + MUMPS, optionally under mpirun) cannot be installed in
this image, so this "port" is the stand-in: SuperLU factorised once per process, then per step a
vectorised RHS assembly, one ``lu.solve``, sparse sensor rows, the energy and the ZOH controller —
one independent trajectory per worker process, one process per host core.
"""

from __future__ import annotations

import multiprocessing as mp
import os
import time
from pathlib import Path

# One trajectory per worker process, one BLAS/OpenMP thread per worker: the thread caps must be in the environment BEFORE
# numpy (and its BLAS) is imported, in the parent too, because spawned children import numpy while unpickling the worker
# function -- setting them inside the worker is too late and 16-32 workers x all-core BLAS pools oversubscribe the host.
THREAD_VARS = ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS", "VECLIB_MAXIMUM_THREADS")
for _v in THREAD_VARS:
    os.environ[_v] = "1"

import numpy as np  # noqa: E402

ROOT = Path(__file__).resolve().parent.parent


def _worker(idx: int, nsteps: int, warmup: int, gain: float, ready, go, out):
    try:  # belt and braces: cap an already-loaded BLAS pool as well
        from threadpoolctl import threadpool_limits

        threadpool_limits(1)
    except Exception:
        pass
    from oracle import cases
    from oracle.flow_oracle import FlowOracle, ZOHController

    case = cases.cylinder(100.0)
    xy, tri = cases.load_mesh(case.mesh_file)
    fo = FlowOracle(case, xy, tri)
    fo.set_base_flow(np.load(ROOT / "tests/golden/cylinder_baseflow.npz")["UP0"])
    fo.init_time_stepping()
    k = np.load(ROOT / "tests/golden/Kopt_reduced13.npz")
    K = ZOHController(k["A"], gain * k["B"], k["C"], gain * k["D"])

    def one():
        u = K.step(-fo.y_meas[0], case.dt)
        fo.step([u[0], u[0]])

    for _ in range(warmup):
        one()
    ready.put(idx)
    go.wait()
    t0 = time.perf_counter()
    for _ in range(nsteps):
        one()
    out.put((idx, time.perf_counter() - t0, float(fo.y_meas[0])))


def time_oracle(nsteps: int = 20, warmup: int = 3, workers: int | None = None) -> dict:
    """Closed-loop cylinder steps on ``workers`` independent trajectories (one per process).

    Returns trajectory-steps/s over all workers, timed from a common start to the slowest finish."""
    ncpu = os.cpu_count() or 1
    try:
        ncpu = min(ncpu, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    workers = workers or min(ncpu, 64)
    for v in THREAD_VARS:  # inherited by the spawned children before they import numpy
        os.environ[v] = "1"
    ctx = mp.get_context("spawn")
    ready, out = ctx.Queue(), ctx.Queue()
    go = ctx.Event()
    gains = 0.5 + np.arange(workers) / max(workers - 1, 1)
    procs = [ctx.Process(target=_worker, args=(i, nsteps, warmup, float(gains[i]), ready, go, out)) for i in range(workers)]
    for p in procs:
        p.start()
    for _ in procs:
        ready.get(timeout=900)
    t0 = time.perf_counter()
    go.set()
    res = [out.get(timeout=900) for _ in procs]
    wall = time.perf_counter() - t0
    for p in procs:
        p.join(30)
    per = [r[1] for r in res]
    return {
        "value": workers * nsteps / wall,
        "unit": "trajectory-steps/s",
        "cores": workers,
        "kind": "port",
        "threads_per_worker": 1,
        "sample": f"{workers} trajectories x {nsteps} closed-loop steps of the cylinder Re=100 config, one process per core "
                  f"(host has {ncpu} logical CPUs); numpy/scipy SuperLU stand-in for the FEniCS/MUMPS path",
        "wall_s": wall,
        "per_core_steps_per_s": nsteps / float(np.mean(per)),
    }
