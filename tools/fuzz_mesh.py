"""Random poses AND random mesh transforms of the textured glTF (every texture slot of the reference's closest-hit program in use),
mesh stage against the oracle (GPU box):
    python tools/fuzz_mesh.py [n_poses] [seed]
Per pose: the visibility buffer (triangle ids, hit distances bit for bit), the shaded colours, and the hybrid frame."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "nerf-glasses_b200")):
    sys.path.insert(0, p)
import numpy as np
import pynmr, synth
import helpers as H
from test_gpu_mesh_textures import oracle_mesh

W, HH = 160, 96
TOL = 2.0 / 255.0


def run(n_poses: int = 40, seed: int = 0, verbose: bool = True):
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    say = print if verbose else (lambda *a, **k: None)
    failures = []
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "s.msgpack"); synth.write_snapshot(path, seed=1337, log2_hashmap_size=15)
        snap = synth.read_snapshot(path)
        gltf = synth.write_textured_glasses_gltf(os.path.join(d, "tex"))
        model = O.Model.from_snapshot(snap)
        hits = []
        for k in range(n_poses):
            # a fresh renderer per pose: the mesh transform goes in at load time, as in render.py
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            if rng.random() < 0.5:
                q = np.array(synth.GLASSES_R_WXYZ, dtype=np.float64) + 0.15 * rng.normal(size=4); q /= np.linalg.norm(q)
            s = tuple(float(v) for v in np.array(synth.GLASSES_S) * rng.uniform(0.6, 1.6, 3))
            t = tuple(float(v) for v in np.array(synth.GLASSES_T) + rng.uniform(-0.08, 0.08, 3))
            rq = tuple(float(v) for v in q)
            r = pynmr.NerfMeshRenderer(W, HH, 0)
            nerf = r.load_nerf(path)
            assert r.load_mesh(gltf, t=t, s=s, r=rq) is not None
            r.orbit(float(rng.uniform(-3, 3)), float(rng.uniform(-1.0, 1.0)), float(rng.uniform(0, 5.3)))
            if rng.random() < 0.5:
                m = r.view_projection_mat; m[:, 3] += float(rng.uniform(0.0, 0.9)) * m[:, 2] * float(np.linalg.norm(m[:, 3])); r.view_projection_mat = m
            c12 = np.ascontiguousarray(r.view_projection_mat.T.reshape(-1))
            rgba2, d2, tri2, surf, ts = H.debug_mesh(r, W, HH)
            mesh, g = oracle_mesh(gltf, t, s, rq)
            want_rgba2, want_d2, want_tri = mesh.render(c12, 2 * W, 2 * HH)
            hit = want_tri >= 0
            hits.append(float(hit.mean()))
            bad = []
            if not np.array_equal(tri2, want_tri): bad.append(f"triangle ids differ at {int((tri2 != want_tri).sum())} sub-pixels")
            elif hit.any() and not np.array_equal(d2[hit].view(np.uint32), want_d2[hit].view(np.uint32)): bad.append("hit distances differ")
            if float(np.abs(rgba2 - want_rgba2).max()) > 2e-4: bad.append(f"shading differs by {float(np.abs(rgba2 - want_rgba2).max()):.5f}")
            want_surf, want_ts = O.mesh_resolve(want_rgba2, want_d2, W, HH, 2)
            P = model.params_struct(W, HH, c12, aabb_min=snap["render_aabb_min"], aabb_max=snap["render_aabb_max"], n_steps_mode=1)
            frame, _, _, _ = model.render_frame(P, want_surf, want_ts)
            want_img, _ = O.accumulate_tonemap(frame, None, 0, to_srgb=True)
            img = np.asarray(nerf.render(W, HH, 1, linear=False))
            dmax = float(np.abs(img - want_img).max())
            if dmax > TOL: bad.append(f"hybrid frame differs by {dmax:.4f} at {int((np.abs(img - want_img).max(axis=2) > TOL).sum())} pixels")
            if bad:
                failures.append((k, bad)); say(f"pose {k} (mesh covers {hits[-1]:.3%}): " + "; ".join(bad), flush=True)
        say(f"{n_poses} poses, seed {seed}: {len(failures)} failures; mesh coverage of the 2x buffer: median {np.median(hits):.3%} max {np.max(hits):.3%}, poses without a hit {int((np.array(hits) == 0).sum())}")
    return failures


if __name__ == "__main__":
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 40, int(sys.argv[2]) if len(sys.argv) > 2 else 0) else 0)
