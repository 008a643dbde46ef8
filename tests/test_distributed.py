"""Ensemble sharding + the all-gather of time-series blocks on world_size 2 (gloo, CPU)."""
import os
import socket

import numpy as np
import pytest

from flowcontrol_b200.sharding import controller_gain_sweep, shard_bounds


def test_shard_bounds_cover_and_balance():
    for total in (1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            w = [b - a for a, b in spans]
            assert max(w) - min(w) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_gain_sweep():
    g = controller_gain_sweep(256)
    assert g[0] == 0.5 and np.isclose(g[-1], 1.5) and len(g) == 256


def _worker(rank, world, port, total, q):
    import torch
    import torch.distributed as dist

    from flowcontrol_b200.sharding import gather_series, shard_bounds

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(total, rank, world)
    full = torch.arange(5 * 3 * total, dtype=torch.float64).reshape(5, 3, total)
    out = gather_series(full[:, :, lo:hi].contiguous(), total)
    ok = bool(torch.equal(out, full))
    # the preallocated form used by bench.py: same buffers on every call, non-contiguous shard views accepted
    from flowcontrol_b200.sharding import SeriesGatherer

    gat = SeriesGatherer(5, 3, total, torch.float64, "cpu")
    o1 = gat(full[:, :, lo:hi])
    p1 = o1.data_ptr()
    ok = ok and bool(torch.equal(o1, full))
    o2 = gat((2.0 * full)[:, :, lo:hi])
    ok = ok and o2.data_ptr() == p1 and bool(torch.equal(o2, 2.0 * full))
    from flowcontrol_b200.sharding import gather_costs

    costs = {"energy_integral": full[0, 0, lo:hi].numpy(), "control": full[1, 1, lo:hi].numpy(), "energy_terminal": full[2, 2, lo:hi].numpy()}
    g = gather_costs(costs, total)
    ok = ok and np.array_equal(g["energy_integral"], full[0, 0].numpy()) and np.array_equal(g["control"], full[1, 1].numpy()) \
        and np.array_equal(g["energy_terminal"], full[2, 2].numpy())
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_series_world2_gloo(total):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]
