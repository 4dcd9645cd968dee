"""Random camera poses through every way libnmr can produce the same picture - they must agree bit for bit:
    python tools/fuzz_paths.py [n_poses] [seed]
frame() with and without the set-up / march overlap, Testbed.render() in float32 / sRGB8 / float16 (against the conversions of
the float image), render_update() into a kept buffer, render_views() of this pose and the previous one, and three row shards
(nmr_set_shard, surface rule pinned to the ray-local one as pynmr.dist does) reassembled.  Lens scene, small frames."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "nerf-glasses_b200")):
    sys.path.insert(0, p)
import numpy as np
import pynmr, synth

W, HH = 320, 288          # (>= 256 rows: render() takes its copy-overlapped path)


def same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def run(n_poses: int = 60, seed: int = 0, verbose: bool = True, size=None):
    W, HH = size if size else (globals()["W"], globals()["HH"])
    rng = np.random.default_rng(seed)
    say = print if verbose else (lambda *a, **k: None)
    failures = []
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "s.msgpack"); synth.write_snapshot(path, seed=1337, log2_hashmap_size=15)
        gltf = synth.write_lens_glasses_gltf(os.path.join(d, "lens"))
        r = pynmr.NerfMeshRenderer(W, HH, 0)
        nerf = r.load_nerf(path)
        r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ)
        base = r.view_projection_mat.copy()
        kept = {np.float32: None, np.uint8: None}
        prev_cam, prev_img = None, None
        for k in range(n_poses):
            r.view_projection_mat = base
            r.orbit(float(rng.uniform(-3, 3)), float(rng.uniform(-1.2, 1.2)), float(rng.uniform(-2, 5.3)))
            m = r.view_projection_mat
            if rng.random() < 0.6:
                m[:, 3] += float(rng.uniform(0.0, 1.1)) * m[:, 2] * float(np.linalg.norm(m[:, 3])); r.view_projection_mat = m
            if rng.random() < 0.4:
                m[:, 3] += float(rng.uniform(-1.0, 1.0)) * m[:, 0] + float(rng.uniform(-0.6, 0.6)) * m[:, 1]; r.view_projection_mat = m
            cam = r.view_projection_mat.copy()

            def check(name, ok):
                if not ok:
                    failures.append((k, name)); say(f"pose {k}: {name} differs", flush=True)

            r.set_overlap(True); assert r.frame(); A = np.asarray(r.read_frame()).copy()
            alive = r.stats()["rays_alive"]
            r.view_projection_mat = cam        # (setting the camera restarts the accumulation: a second frame() alone would be sample 2)
            r.set_overlap(False); assert r.frame(); check("frame() without overlap", same(np.asarray(r.read_frame()), A)); r.set_overlap(True)
            f32 = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
            check("render() float32", same(f32, A))
            check("render() uint8", same(np.asarray(nerf.render(W, HH, 1, linear=False, dtype=np.uint8)), np.uint8(np.clip(A, 0.0, 1.0) * np.float32(255.0))))
            check("render() float16", same(np.asarray(nerf.render(W, HH, 1, linear=False, dtype=np.float16)), A.astype(np.float16)))
            for dt in (np.float32, np.uint8):
                kept[dt] = nerf.render_update(kept[dt], W, HH, linear=False, dtype=dt)
                check(f"render_update() {np.dtype(dt).name}", same(kept[dt], A if dt is np.float32 else np.uint8(np.clip(A, 0.0, 1.0) * np.float32(255.0))))
            if prev_cam is not None:
                v = np.asarray(r.render_views(nerf, np.stack([cam, prev_cam, cam]), W, HH))
                check("render_views()", same(v[0], A) and same(v[1], prev_img) and same(v[2], A))
            r.view_projection_mat = cam
            # shards under the ray-local surface rule
            r.set_surface_insertion(pynmr.NerfMeshRenderer.SURFACE_BATCH8)
            assert r.frame(); B = np.asarray(r.read_frame()).copy()
            merged = np.zeros_like(B)
            for rank in range(3):
                r.set_shard(rank, 3, 16); r.view_projection_mat = cam; assert r.frame()
                part = np.asarray(r.read_frame())
                rows = [y for y in range(HH) if (y // 16) % 3 == rank]
                merged[rows] = part[rows]
            r.set_shard(0, 1, 16); r.view_projection_mat = cam
            r.set_surface_insertion(pynmr.NerfMeshRenderer.SURFACE_AUTO)
            check("three row shards", same(merged, B))
            prev_cam, prev_img = cam, A
            if verbose and k % 20 == 0:
                say(f"pose {k}: live rays {alive} of {W * HH}", flush=True)
    say(f"{n_poses} poses, seed {seed}, {W}x{HH}: {len(failures)} disagreements {failures[:8]}")
    return failures


if __name__ == "__main__":
    kw = dict(a.split("=") for a in sys.argv[3:])        # size=333x217
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 60, int(sys.argv[2]) if len(sys.argv) > 2 else 0,
                      size=tuple(int(v) for v in kw["size"].split("x")) if "size" in kw else None) else 0)
