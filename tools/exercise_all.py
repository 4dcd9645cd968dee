"""Small end-to-end exercise of every kernel family in one process (tiny sizes, so it also suits a run under
compute-sanitizer where that tool is available - it is closed on the authoring pool):

    python tools/exercise_all.py

Hybrid frame with lens surfaces, banded render(), render_views with lanes, floatie removal, density probes, tonemap curve,
shared frame target (one rank)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "nerf-glasses_b200")):
    sys.path.insert(0, p)
import numpy as np
import pynmr, synth

W, H = 160, 96
with tempfile.TemporaryDirectory() as d:
    snap = os.path.join(d, "s.msgpack"); synth.write_snapshot(snap, seed=1337, log2_hashmap_size=14)
    gltf = synth.write_lens_glasses_gltf(os.path.join(d, "mesh"))
    r = pynmr.NerfMeshRenderer(W, H, 0)
    nerf = r.load_nerf(snap)
    assert nerf is not None and r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is not None
print("floaties", r.remove_floaties())
r.orbit(0.3, -0.1, 4.0)
assert r.frame(); a = np.asarray(r.read_frame()).copy()
print("frame", r.stats()["samples"], float(a.mean()))
nerf.tonemap_curve = pynmr.TonemapCurve.ACES
b = np.asarray(nerf.render(W, 288, 1, linear=False)).copy()      # >= 256 rows: the banded path
nerf.tonemap_curve = 0
cams = []
for _ in range(5):
    r.orbit(0.05, 0.01, 0); cams.append(r.view_projection_mat)
v = np.asarray(r.render_views(nerf, np.stack(cams), 96, 64)).copy()
print("views", v.shape, float(v.mean()))
pts = np.random.default_rng(0).uniform(-0.3, 0.3, (300, 3)).astype(np.float32)
print("probes", float(nerf.probe_points(pts, [0, -1, 0]).sum()), float(nerf.probe_rays(pts, [0, -1, 0]).sum()))
h, ptr = r.gather_create()
for _ in range(2):
    r.orbit(0.02, 0.0, 0); r.frame()
r.gather_detach()
print("ok")
