"""Random poses, one frame split over the GPUs of a box against the same frame on one GPU - bit for bit:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/fuzz_multi.py [n_poses] [seed]
Both assembly paths of pynmr.dist (rows stored into rank 0's image by the render kernels over NVLink; NCCL gather), lens scene,
row bands of 8 / 16 / 32, poses from far outside to inside the head."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "nerf-glasses_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

W, HH = 384, 216


def main():
    n_poses = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    import pynmr, synth
    from pynmr import dist as D
    rng = np.random.default_rng(seed)          # the same poses on every rank
    tmp = os.path.join(tempfile.gettempdir(), f"nmr_fuzz_multi_{seed}")
    if rank == 0:
        os.makedirs(tmp, exist_ok=True)
        synth.write_snapshot(os.path.join(tmp, "s.msgpack"), seed=1337, log2_hashmap_size=15)
        synth.write_lens_glasses_gltf(os.path.join(tmp, "lens"))
    dist.barrier()
    r = pynmr.NerfMeshRenderer(W, HH, local)
    nerf = r.load_nerf(os.path.join(tmp, "s.msgpack"))
    r.load_mesh(os.path.join(tmp, "lens", "glasses.gltf"), t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ)
    r.set_surface_insertion(pynmr.NerfMeshRenderer.SURFACE_BATCH8)
    base = r.view_projection_mat.copy()
    bad = []
    for band in (8, 16, 32):
        cams = []
        for k in range(n_poses):
            r.view_projection_mat = base
            r.orbit(float(rng.uniform(-3, 3)), float(rng.uniform(-1.2, 1.2)), float(rng.uniform(-2, 5.3)))
            m = r.view_projection_mat
            if rng.random() < 0.6:
                m[:, 3] += float(rng.uniform(0.0, 1.1)) * m[:, 2] * float(np.linalg.norm(m[:, 3])); r.view_projection_mat = m
            cams.append(r.view_projection_mat.copy())
        # rows stored straight into rank 0's image by every rank's render kernels
        ps = D.PeerShardedRenderer(r, rank, world, band=band, dst=0)
        peer = []
        for cam in cams:
            r.view_projection_mat = cam
            img = ps.render_frame()
            if rank == 0: peer.append(img.clone())
        ps.close()
        # packed rows, one NCCL gather
        sr = D.ShardedRenderer(r, rank, world, band=band)
        gathered = []
        for cam in cams:
            r.view_projection_mat = cam
            img = sr.render_frame(dst=0)
            if rank == 0: gathered.append(img.clone())
        r.set_shard(0, 1, band)
        if rank == 0:
            for k, cam in enumerate(cams):
                r.view_projection_mat = cam
                assert r.frame()
                single = torch.from_numpy(np.asarray(r.read_frame()).copy()).cuda()
                if not torch.equal(peer[k], single): bad.append((band, k, "peer"))
                if not torch.equal(gathered[k], single): bad.append((band, k, "nccl"))
        dist.barrier()
    if rank == 0:
        print(f"{world} GPUs, {3 * n_poses} poses (bands 8 / 16 / 32), seed {seed}: {len(bad)} frames differ from the single-GPU frame {bad[:6]}", flush=True)
    dist.barrier(); dist.destroy_process_group()
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
