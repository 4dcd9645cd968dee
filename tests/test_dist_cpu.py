"""Host-side logic of the multi-GPU path (pynmr/dist.py) on CPU: row ownership, packing, the gather collective and the
reassembly, with world_size 2 and 3 on the gloo backend (the GPU path uses the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _import_dist():
    import importlib.util
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("pynmr_dist", os.path.join(here, "nerf-glasses_b200", "pynmr", "dist.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_owned_rows_partition():
    D = _import_dist()
    for H, world, band in [(1080, 8, 8), (108, 3, 8), (7, 2, 4), (64, 1, 8), (2160, 8, 32), (5, 8, 1)]:
        seen = np.zeros(H, dtype=int)
        for r in range(world):
            rows = D.owned_rows(H, r, world, band)
            assert np.all(np.diff(rows) > 0)
            seen[rows] += 1
            assert np.all((rows // band) % world == r)
        assert np.all(seen == 1)
        assert D.max_owned_rows(H, world, band) >= (H + world - 1) // world
    with pytest.raises(ValueError):
        D.owned_rows(10, 2, 2, 8)


def test_view_slice_partition():
    D = _import_dist()
    for n, world in [(64, 8), (63, 8), (3, 8), (0, 2), (95, 4)]:
        allv = [v for r in range(world) for v in D.view_slice(n, r, world)]
        assert allv == list(range(n))
        sizes = [len(D.view_slice(n, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, H, W, band, n_views, q):
    import torch
    import torch.distributed as dist
    D = _import_dist()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        truth = torch.rand((H, W, 4), generator=g)
        # a sharded render leaves rows it does not own untouched: fill them with a poison value
        local = torch.full((H, W, 4), float("nan"))
        rows = torch.as_tensor(D.owned_rows(H, rank, world, band), dtype=torch.long)
        local[rows] = truth[rows]
        full = D.gather_frame(local, rank, world, band, dst=0)
        ok_frame = (full is not None and torch.equal(full, truth)) if rank == 0 else (full is None)
        views = torch.rand((n_views, 4, 5, 4), generator=g)
        sl = D.view_slice(n_views, rank, world)
        got = D.gather_views(views[sl.start:sl.stop], n_views, rank, world, dst=0)
        ok_views = (got is not None and torch.equal(got, views)) if rank == 0 else (got is None)
        q.put((rank, bool(ok_frame), bool(ok_views)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,H,band,n_views", [(2, 108, 8, 7), (3, 50, 4, 5)])
def test_gather_frame_and_views_gloo(world, H, band, n_views):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, 24, band, n_views, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(r, True, True) for r in range(world)]
