"""Random poses for the features around the single-NeRF frame, each against the oracle (GPU box):
    python tools/fuzz_scene.py [n_poses] [seed]
 * accumulation: frame() called 1 - 4 times on an unchanged camera (spp index 0..3, the reference's sample jitter) vs the oracle
   accumulating the same samples;
 * two NeRFs in one frame (the second one shifted / rotated by a random model transform), merged by depth, vs two oracle renders
   merged with combineBuffersKernel's rule."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "nerf-glasses_b200")):
    sys.path.insert(0, p)
import numpy as np
import pynmr, synth
import helpers as H

W, HH = 160, 96
TOL = 2.0 / 255.0


def random_camera(r, base, rng):
    r.view_projection_mat = base
    r.orbit(float(rng.uniform(-3, 3)), float(rng.uniform(-1.0, 1.0)), float(rng.uniform(0, 5.3)))
    if rng.random() < 0.5:
        m = r.view_projection_mat; m[:, 3] += float(rng.uniform(0.0, 0.8)) * m[:, 2] * float(np.linalg.norm(m[:, 3])); r.view_projection_mat = m
    return np.ascontiguousarray(r.view_projection_mat.T.reshape(-1))


def run(n_poses: int = 40, seed: int = 0, verbose: bool = True):
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    say = print if verbose else (lambda *a, **k: None)
    failures = []
    with tempfile.TemporaryDirectory() as d:
        pa = os.path.join(d, "a.msgpack"); synth.write_snapshot(pa, seed=1337, log2_hashmap_size=15)
        pb = os.path.join(d, "b.msgpack"); synth.write_snapshot(pb, seed=4242, log2_hashmap_size=15)
        sa, sb = synth.read_snapshot(pa), synth.read_snapshot(pb)
        gltf = synth.write_glasses_gltf(os.path.join(d, "mesh"))
        g = {"path": gltf, "t": synth.GLASSES_T, "s": synth.GLASSES_S, "r": synth.GLASSES_R_WXYZ,
             "texture": np.tile(np.array([128, 128, 128, 255], dtype=np.uint8), (4, 4, 1))}
        # ---- accumulation ----
        r = pynmr.NerfMeshRenderer(W, HH, 0)
        nerf = r.load_nerf(pa)
        r.load_mesh(gltf, t=g["t"], s=g["s"], r=g["r"])
        base = r.view_projection_mat.copy()
        for k in range(n_poses):
            c12 = random_camera(r, base, rng)
            n_spp = int(rng.integers(1, 5))
            acc = None
            for i in range(n_spp):
                assert r.frame()
                _, frame, _, _, _ = H.oracle_scene(sa, W, HH, c12, glasses=g, spp_index=i, n_steps_mode=1)
                want, acc = O.accumulate_tonemap(frame, acc, i, to_srgb=True)
            img = np.asarray(r.read_frame())
            dmax = float(np.abs(img - want).max())
            if dmax > TOL or H.psnr(img, want) < 45.0:
                failures.append(("spp", k, n_spp, dmax)); say(f"accumulation pose {k}: {n_spp} samples, max |d| {dmax:.4f}", flush=True)
        # ---- two NeRFs ----
        r2 = pynmr.NerfMeshRenderer(W, HH, 0)
        a = r2.load_nerf(pa); b = r2.load_nerf(pb)
        r2.load_mesh(gltf, t=g["t"], s=g["s"], r=g["r"])
        base2 = r2.view_projection_mat.copy()
        ma, mb = O.Model.from_snapshot(sa), O.Model.from_snapshot(sb)
        for k in range(n_poses):
            tb = rng.uniform(-0.3, 0.3, 3).astype(np.float32); rb = rng.uniform(-0.15, 0.15, 3).astype(np.float32)
            b.model_translation = tb; b.model_rotation = rb
            Rb = b.model_matrix
            c12 = random_camera(r2, base2, rng)
            assert r2.frame()
            img = np.asarray(r2.read_frame()).copy()
            _, _, _, _, (surf, ts) = H.oracle_scene(sa, W, HH, c12, glasses=g, n_steps_mode=1)
            Pa = ma.params_struct(W, HH, c12, aabb_min=sa["render_aabb_min"], aabb_max=sa["render_aabb_max"], n_steps_mode=1)
            fa, da, _, _ = ma.render_frame(Pa, surf, ts)
            Pb = mb.params_struct(W, HH, c12, aabb_min=sb["render_aabb_min"], aabb_max=sb["render_aabb_max"], n_steps_mode=1, model_rot=Rb, model_trans=tb)
            fb, db, _, _ = mb.render_frame(Pb, None, None)
            take_b = db < da
            want, _ = O.accumulate_tonemap(np.where(take_b[..., None], fb, fa), None, 0, to_srgb=True)
            close = (np.abs(da - db) <= 1e-3 * np.minimum(da, db)) & ((da < 1e9) | (db < 1e9))      # either renderer may own such a pixel
            dmax = float(np.abs(img - want)[~close].max()) if (~close).any() else 0.0
            if dmax > TOL or close.mean() > 0.02:
                failures.append(("two nerfs", k, float(close.mean()), dmax)); say(f"two-NeRF pose {k}: max |d| {dmax:.4f}, undecided pixels {close.mean():.3%}", flush=True)
    say(f"{n_poses} + {n_poses} poses, seed {seed}: {len(failures)} failures {failures[:6]}")
    return failures


if __name__ == "__main__":
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 40, int(sys.argv[2]) if len(sys.argv) > 2 else 0) else 0)
