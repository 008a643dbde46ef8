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


class SeriesGatherer:
    """All-gather of per-rank series blocks ``[nsteps, ncol, B_local]`` (torch tensors, trajectory innermost) into
    ``[nsteps, ncol, total]`` on every rank, with every buffer allocated ONCE: the collective is a single
    ``all_gather_into_tensor`` into a preallocated ``[world, nsteps, ncol, wmax]`` block followed by one strided copy into
    the preallocated result (no per-call allocation, no list-form all_gather, no torch.cat).  Ragged shards are staged
    through a preallocated pad of the widest shard; even shards are sent in place."""

    def __init__(self, nsteps: int, ncol: int, total: int, dtype, device, group=None):
        import torch
        import torch.distributed as dist

        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.total = int(total)
        bounds = [shard_bounds(total, r, self.world) for r in range(self.world)]
        self.widths = [hi - lo for lo, hi in bounds]
        self.offsets = [lo for lo, _ in bounds]
        self.wmax = max(self.widths)
        self.even = min(self.widths) == self.wmax
        self.shape = (int(nsteps), int(ncol))
        # [world * nsteps, ncol, wmax]: the concatenation-along-dim-0 form every backend (NCCL, gloo) accepts
        self.recv = torch.empty((self.world * nsteps, ncol, self.wmax), dtype=dtype, device=device)
        self.pad = None if self.even else torch.zeros((nsteps, ncol, self.wmax), dtype=dtype, device=device)
        self.out = torch.empty((nsteps, ncol, self.total), dtype=dtype, device=device)

    def __call__(self, local):
        import torch.distributed as dist

        nsteps, ncol = self.shape
        if tuple(local.shape) != (nsteps, ncol, self.widths[self.rank]):
            raise ValueError(f"expected a [{nsteps}, {ncol}, {self.widths[self.rank]}] block, got {tuple(local.shape)}")
        send = local
        if not self.even:
            self.pad[:, :, : local.shape[2]].copy_(local)
            send = self.pad
        if not send.is_contiguous():
            send = send.contiguous()
        dist.all_gather_into_tensor(self.recv, send, group=self.group)
        recv = self.recv.view(self.world, nsteps, ncol, self.wmax)
        if self.even:
            self.out.view(nsteps, ncol, self.world, self.wmax).copy_(recv.permute(1, 2, 0, 3))
        else:
            for r, (o, w) in enumerate(zip(self.offsets, self.widths)):
                self.out[:, :, o : o + w].copy_(recv[r, :, :, :w])
        return self.out


def gather_series(local, total: int, group=None):
    """One-shot form of :class:`SeriesGatherer` (allocates its buffers for this call)."""
    nsteps, ncol, _ = local.shape
    return SeriesGatherer(nsteps, ncol, total, local.dtype, local.device, group=group)(local)


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
