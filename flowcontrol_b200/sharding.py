"""Ensemble sharding across the GPUs of one node.

The reference scales by MPI domain decomposition inside dolfin/PETSc/MUMPS
(README.md:48); its hand-written collectives on the step path are an
``Allreduce(MIN)`` per sensor (utils/mpi.py:22-37) and an ``MPI.max`` for the
divergence flag (flowsolver.py:816-819).  Here trajectories are independent, so
each rank owns a contiguous block of the ensemble, holds a full replica of the
constant operators, and the only collective is an all-gather of the time-series
blocks (NCCL on GPUs, gloo in CPU tests).  No exchange inside the step.
"""

from __future__ import annotations

import numpy as np


def shard_bounds(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of ``total`` trajectories owned by ``rank``."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_series(local, total: int, group=None):
    """All-gather per-rank series blocks ``[nsteps, ncol, B_local]`` (torch tensor, trajectory
    innermost) into ``[nsteps, ncol, total]`` on every rank.  Ragged shards are padded to the
    widest shard for the collective and cropped afterwards."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    widths = [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]
    wmax = max(widths)
    nsteps, ncol, bl = local.shape
    assert bl == widths[rank]
    pad = torch.zeros((nsteps, ncol, wmax), dtype=local.dtype, device=local.device)
    pad[:, :, :bl] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous(), group=group)
    return torch.cat([o[:, :, :w] for o, w in zip(out, widths)], dim=2)


def gather_costs(local: dict, total: int, device=None, group=None) -> dict:
    """All-gather the per-trajectory cost sums of every shard (``Ensemble.costs()``: dict of ``[B_local]`` arrays) into
    dicts of ``[total]`` arrays on every rank: what an optimiser rank needs from a sharded controller sweep
    (24 bytes per trajectory instead of the time series).  Implemented on top of ``gather_series``."""
    import torch

    keys = sorted(local)
    block = torch.as_tensor(np.stack([np.asarray(local[k], dtype=np.float64) for k in keys])[None], device=device)
    full = gather_series(block, total, group=group)[0].cpu().numpy()
    return {k: full[i] for i, k in enumerate(keys)}


def controller_gain_sweep(B: int, lo: float = 0.5, hi: float = 1.5) -> np.ndarray:
    """Deterministic gain family of config 2: g_b = lo + (hi-lo) b/(B-1)."""
    if B == 1:
        return np.array([1.0])
    return lo + (hi - lo) * np.arange(B) / (B - 1)
