"""Multi-GPU path on real devices (needs >= 2 GPUs, NCCL): one process per GPU, row-band sharded hybrid frame gathered with
one NCCL collective must equal the single-GPU frame bit for bit; views dealt to ranks must equal the same
views rendered on one GPU."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, HH = 320, 180


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _scene(rank, snap_path, gltf):
    import pynmr
    import synth
    r = pynmr.NerfMeshRenderer(W, HH, rank)
    nerf = r.load_nerf(snap_path)
    assert nerf is not None
    assert r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is not None
    r.remove_floaties()
    r.orbit(0.35, -0.2, 4.0)
    # every rank must take the same surface-insertion decision as the full frame would (the auto rule looks at the
    # fraction of live pixels of the rows it renders): pin it
    r.set_surface_insertion(pynmr.NerfMeshRenderer.SURFACE_BATCH8)
    return r, nerf


def _worker(rank, world, port, snap_path, gltf, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tools"), os.path.join(root, "nerf-glasses_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        from pynmr import dist as D
        r, nerf = _scene(rank, snap_path, gltf)
        # tiles: one frame, rows dealt to ranks in bands of 8
        sr = D.ShardedRenderer(r, rank, world, band=8)
        full = sr.render_frame(dst=0)
        ok_tiles = None
        if rank == 0:
            r.set_shard(0, 1, 8)
            r.frame()
            single = torch.from_numpy(np.asarray(r.read_frame()).copy()).cuda()
            ok_tiles = bool(torch.equal(full, single))
        # tiles, fused: every rank's kernels store their rows into rank 0's image over NVLink, device-side flags, no collective
        r.set_shard(0, 1, 8)
        ps = D.PeerShardedRenderer(r, rank, world, band=8, dst=0)
        ok_peer = None
        for k in range(3):                      # several frames: the consumed / written handshake is exercised
            r.orbit(0.02, 0.01, 0)
            img = ps.render_frame()
            if rank == 0:
                peer_img = img.clone()
        ps.close()
        if rank == 0:
            r.set_shard(0, 1, 8)
            r.frame()
            single = torch.from_numpy(np.asarray(r.read_frame()).copy()).cuda()
            ok_peer = bool(torch.equal(peer_img, single))
        # views: 5 cameras dealt to ranks
        cams = []
        r.set_shard(0, 1, 8)
        for k in range(5):
            r.orbit(0.05, 0.01, 0)
            cams.append(r.view_projection_mat)
        cams = np.stack(cams)
        sl = D.view_slice(len(cams), rank, world)
        mine = np.asarray(r.render_views(nerf, cams[sl.start:sl.stop], 96, 54)) if len(sl) else np.zeros((0, 54, 96, 4), np.float32)
        got = D.gather_views(torch.from_numpy(mine.copy()).cuda(), len(cams), rank, world, dst=0)
        ok_views = None
        if rank == 0:
            want = torch.from_numpy(np.asarray(r.render_views(nerf, cams, 96, 54)).copy()).cuda()
            ok_views = bool(torch.equal(got, want))
        q.put((rank, ok_tiles, ok_views, ok_peer))
    finally:
        dist.destroy_process_group()


def test_sharded_frame_and_views_nccl(small_snapshot, glasses_gltf):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, small_snapshot[0], glasses_gltf, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == (0, True, True, True), res


def test_two_contexts_on_two_devices_in_one_process(small_snapshot, glasses_gltf):
    """One process driving two GPUs through two contexts ('single process driving 8 devices is sufficient'):
    both devices render the same frame bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    imgs = []
    for dev in (0, 1):
        r, nerf = _scene(dev, small_snapshot[0], glasses_gltf)
        imgs.append(np.asarray(nerf.render(W, HH, 1, linear=False)).copy())
        assert r.frame()
    assert np.array_equal(imgs[0].view(np.uint32), imgs[1].view(np.uint32))


def test_shared_frame_target_single_rank(small_snapshot, glasses_gltf):
    """nmr_gather_* with a world of one (runs on any GPU box): the shared image, its sequence flags and the signal / wait kernels
    are exercised over several frames; the image equals the ordinary frame."""
    import torch
    from pynmr import dist as D
    torch.cuda.set_device(0)
    r, nerf = _scene(0, small_snapshot[0], glasses_gltf)
    ps = D.PeerShardedRenderer(r, 0, 1, band=8, dst=0)
    for _ in range(3):
        r.orbit(0.02, 0.01, 0)
        got = ps.render_frame().clone()
    ps.close()
    assert r.frame()
    want = torch.from_numpy(np.asarray(r.read_frame()).copy()).cuda()
    assert torch.equal(got, want)
