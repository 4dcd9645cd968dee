"""Small frames through every render path (plain, hybrid, lens, close-up two-pass, overlapped and serial, views, formats, probes,
multi-NeRF, model transform) for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize.py
    compute-sanitizer --tool racecheck python tools/sanitize.py
    compute-sanitizer --tool initcheck python tools/sanitize.py"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "nerf-glasses_b200")):
    sys.path.insert(0, p)
import numpy as np
import pynmr, synth

W, H = 160, 96
with tempfile.TemporaryDirectory() as d:
    snap = os.path.join(d, "s.msgpack"); synth.write_snapshot(snap, seed=1337, log2_hashmap_size=15)
    snap2 = os.path.join(d, "s2.msgpack"); synth.write_snapshot(snap2, seed=7, log2_hashmap_size=15)
    gltf = synth.write_lens_glasses_gltf(os.path.join(d, "lens"))
    tex = synth.write_textured_glasses_gltf(os.path.join(d, "tex"))
    r = pynmr.NerfMeshRenderer(W, H, 0)
    nerf = r.load_nerf(snap)
    r.orbit(0.3, -0.1, 4.0)
    print("plain", float(np.asarray(nerf.render(W, H, 1, linear=False)).mean()))
    r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ)
    r.remove_floaties()
    for overlap in (True, False):
        r.set_overlap(overlap)
        r.frame(); print("lens frame overlap", overlap, float(np.asarray(r.read_frame()).mean()))
    r.set_lens_model(1, 0.02); r.frame(); r.set_lens_model(0, 0.0)
    print("u8", int(np.asarray(nerf.render(W, H, 1, linear=False, dtype=np.uint8)).sum()), "f16", float(np.asarray(nerf.render(W, H, 2, linear=True, dtype=np.float16)).mean()))
    print("banded", float(np.asarray(nerf.render(320, 288, 1, linear=False)).mean()))
    m = r.view_projection_mat; m[:, 3] += 0.55 * m[:, 2]; r.view_projection_mat = m          # close-up: two-pass schedule
    r.frame(); print("close-up", r.stats()["rays_alive"], float(np.asarray(r.read_frame()).mean()))
    cams = []
    for k in range(5):
        r.orbit(0.1, 0.01, 0); cams.append(r.view_projection_mat)
    print("views", float(np.asarray(r.render_views(nerf, np.stack(cams), 96, 64)).mean()), int(np.asarray(r.render_views(nerf, np.stack(cams), 96, 64, dtype=np.uint8)).sum()))
    for rank in range(2):
        r.set_shard(rank, 2, 8); r.frame()
    r.set_shard(0, 1, 8)
    pts = np.random.default_rng(0).uniform(-0.3, 0.3, (500, 3)).astype(np.float32)
    print("probes", float(nerf.probe_points(pts, [0, -1, 0]).sum()), float(nerf.probe_rays(pts, [0, -1, 0]).sum()))
    nerf.model_translation = (0.05, 0.0, 0.02); nerf.model_rotation = (0.02, 0.1, 0.0)
    r.frame()
    b = r.load_nerf(snap2); b.model_translation = (0.2, 0, 0)
    r.frame(); fr, dp = r.read_combined(); print("two nerfs", float(fr.mean()), float((dp < 1e9).mean()))
    cells = nerf.dump_density_grid(); nerf.load_density_grid(cells)
    r2 = pynmr.NerfMeshRenderer(W, H, 0); n2 = r2.load_nerf(snap); r2.load_mesh(tex, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ)
    r2.orbit(0.3, -0.1, 4.0); r2.frame(); print("textured", float(np.asarray(r2.read_frame()).mean()))
    print("l2", r2.measure_l2(16 << 20, False) > 0, r2.measure_l2(16 << 20, True) > 0)
print("done")
