import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "nerf-glasses_b200")):
    sys.path.insert(0, p)
import numpy as np
import pynmr, synth
with tempfile.TemporaryDirectory() as d:
    snap = os.path.join(d, "s.msgpack"); synth.write_snapshot(snap, seed=1337, log2_hashmap_size=19)
    gltf = synth.write_glasses_gltf(os.path.join(d, "mesh"))
    def mk(mesh, W=1920, H=1080):
        r = pynmr.NerfMeshRenderer(W, H, 0)
        nerf = r.load_nerf(snap)
        if mesh: r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ)
        r.remove_floaties()
        return r, nerf
    def run(tag, r, n=8):
        v = []
        for i in range(n):
            r.orbit(0.01, 0.002, 0); r.frame(); st = r.stats()
            if i >= 3: v.append((st['gpu_ms'] - st['march_ms'], st['march_ms'], st['rays_alive'], st['samples']))
        a = np.median(np.array(v), axis=0)
        print(f"{tag:40s} setup_ms {a[0]:.4f} march_ms {a[1]:.4f} alive {int(a[2])} samples {int(a[3])}", flush=True)
    r, nerf = mk(True); run("hybrid 1080p", r)
    r, nerf = mk(False); run("nerf only 1080p", r)
    r, nerf = mk(False); nerf.render_aabb.min = [0.95, 0.95, 0.95]; nerf.render_aabb.max = [1, 1, 1]; run("nerf only, render box in a corner", r)
    r, nerf = mk(False); r.orbit(0, 0, -30.0); run("nerf only, zoomed far out", r)
    r, nerf = mk(False); r.orbit(3.14159, 0, 0); r.orbit(0, 0, 0); run("nerf only, half orbit", r)
    r, nerf = mk(False, 1920, 540); run("nerf only 1920x540", r)
    r, nerf = mk(False, 3840, 2160); run("nerf only 4K", r)
