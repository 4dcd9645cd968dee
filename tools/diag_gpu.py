"""GPU-vs-oracle diagnostic: renders the smoke scene in both MLP modes, reports where pixels differ, dumps arrays to gpurun_out/."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "nerf-glasses_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import tempfile
import numpy as np
import helpers, pynmr, synth

W, H = 128, 72
out_dir = os.path.join(ROOT, "gpurun_out"); os.makedirs(out_dir, exist_ok=True)
with tempfile.TemporaryDirectory() as d:
    snap_path = os.path.join(d, "s.msgpack"); synth.write_snapshot(snap_path, seed=1337, log2_hashmap_size=15)
    gltf = synth.write_glasses_gltf(os.path.join(d, "mesh"))
    snap = synth.read_snapshot(snap_path)
    g = {"path": gltf, "t": synth.GLASSES_T, "s": synth.GLASSES_S, "r": synth.GLASSES_R_WXYZ,
             "texture": np.tile(np.array([128, 128, 128, 255], dtype=np.uint8), (4, 4, 1))}
    for mesh_on in (False, True):
        for flags in (1, 0):
            r = pynmr.NerfMeshRenderer(W, H, 0)
            nerf = r.load_nerf(snap_path)
            r.orbit(0.35, -0.2, 4.0)
            if mesh_on:
                r.load_mesh(gltf, t=g["t"], s=g["s"], r=g["r"])
            helpers.set_flags(r, flags)
            cam12 = np.ascontiguousarray(r.view_projection_mat.T.reshape(-1))
            want, wframe, wns, wst, (osurf, ots) = helpers.oracle_scene(snap, W, H, cam12, glasses=g if mesh_on else None)
            for api in ("frame", "render"):
                if api == "frame":
                    r.frame(); img = np.asarray(r.read_frame()).copy()
                else:
                    img = np.asarray(nerf.render(W, H, 1, linear=False)).copy()
                fr, dp, ns = helpers.debug_last_frame(r, W, H)
                err = np.abs(img - want).max(axis=2)
                bad = np.argwhere(err > 2 / 255)
                st = r.stats()
                print(f"mesh={mesh_on} flags={flags} api={api}: max err {err.max():.4f}, bad px {len(bad)}, samples gpu {st['samples']} oracle {wst['samples']}, "
                      f"alive gpu {st['rays_alive']} oracle {wst['alive_after_first_hit']}, ns equal {np.mean(ns == wns):.4f}, gpu_ms {st['gpu_ms']:.3f} march_ms {st['march_ms']:.3f}")
                for (y, x) in bad[:6]:
                    print("   px", x, y, "gpu", img[y, x], "want", want[y, x], "ns", ns[y, x], wns[y, x], "tsurf", None if ots is None else ots[y, x])
                np.savez_compressed(os.path.join(out_dir, f"diag_m{int(mesh_on)}_f{flags}_{api}.npz"), img=img, want=want, ns=ns, wns=wns)
            del r
