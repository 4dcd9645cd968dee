"""Soak: a few thousand hybrid frames along the orbit path (render(), frame(), render_views, lens on/off, close-ups) - checks
that nothing hangs, leaks device memory or produces non-finite pixels.  (development aid)"""
import os, sys, tempfile, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/tools", ROOT + "/nerf-glasses_b200", ROOT + "/tests"): sys.path.insert(0, p)
import pynmr, synth
W, H = 1280, 720
with tempfile.TemporaryDirectory() as d:
    sp = os.path.join(d, "s.msgpack"); synth.write_snapshot(sp, seed=1337, log2_hashmap_size=19)
    gl = synth.write_lens_glasses_gltf(os.path.join(d, "m"))
    r = pynmr.NerfMeshRenderer(W, H); nerf = r.load_nerf(sp); r.load_mesh(gl, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ); r.remove_floaties()
free0 = torch.cuda.mem_get_info()[0]
t0 = time.time(); a = 0.0; n = 0
for i in range(3000):
    a += 0.03
    r.orbit(-np.sin(a * 1.733) / 100, np.cos(a * 1.733) / 200, 0.02 * np.sin(a * 0.37))
    if i % 500 == 250: r.set_lens(i % 1000 == 250)
    if i % 7 == 0:
        img = nerf.render(W, H, 1, linear=False)
        assert np.isfinite(img).all()
    elif i % 11 == 0:
        m = r.view_projection_mat; m2 = m.copy(); m2[:, 3] += 0.45 * m2[:, 2]
        out = r.render_views(nerf, np.stack([m, m2]), 640, 360)
        assert np.isfinite(out).all()
    elif i % 13 == 0:
        d = nerf.probe_rays(np.random.default_rng(i).uniform(-0.3, 0.3, (257, 3)).astype(np.float32), [0.1, -0.95, 0.1])
        assert np.isfinite(d).all()
    elif i % 17 == 0:
        nerf.tonemap_curve = (i // 17) % 4
    elif i == 1000:
        h, ptr = r.gather_create()               # shared frame target, one rank: the flag kernels in the loop for a while
    elif i == 2000:
        r.gather_detach()
    else:
        assert r.frame()
    n += 1
r.synchronize()
free1 = torch.cuda.mem_get_info()[0]
print(f"{n} iterations in {time.time() - t0:.1f} s; device memory delta {(free0 - free1) / 2**20:.1f} MiB; last stats {r.stats()}")
