"""Soak (GPU box): thousands of frames, resolution changes and context create / destroy cycles - device memory, pinned-pool size and
host RSS must come to rest, and the picture of a fixed camera must be the same at the end as at the start.
    python tools/soak.py [frames]"""
import math, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "nerf-glasses_b200")):
    sys.path.insert(0, p)
import numpy as np
import psutil, torch
import pynmr, synth

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
proc = psutil.Process()


def mem():
    free, total = torch.cuda.mem_get_info(0)
    return (total - free) / 2 ** 20, proc.memory_info().rss / 2 ** 20


torch.cuda.init()
with tempfile.TemporaryDirectory() as d:
    snap = os.path.join(d, "s.msgpack"); synth.write_snapshot(snap, seed=1337, log2_hashmap_size=19)
    gltf = synth.write_lens_glasses_gltf(os.path.join(d, "lens"))
    r = pynmr.NerfMeshRenderer(1920, 1080, 0)
    nerf = r.load_nerf(snap); r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ); r.remove_floaties()
    ref_cam = r.view_projection_mat.copy()
    first = np.asarray(nerf.render(1920, 1080, 1, linear=False)).copy()
    a = 0.0
    kept = None
    marks = []
    t0 = time.time()
    for k in range(n_frames):
        a += 0.03; r.orbit(-math.sin(a * 1.733) / 100.0, math.cos(a * 1.733) / 200.0, 0.0)
        m = k % 6
        if m == 0: r.frame()
        elif m == 1: nerf.render(1920, 1080, 1, linear=False)
        elif m == 2: nerf.render(1920, 1080, 1, linear=False, dtype=np.uint8)
        elif m == 3: kept = nerf.render_update(kept, 1920, 1080, linear=False)
        elif m == 4: r.render_views(nerf, np.stack([r.view_projection_mat] * 3), 512, 512)
        else: nerf.render(int(200 + (k * 37) % 900), int(100 + (k * 53) % 700), 1, linear=False)       # a different size every time
        if k % (n_frames // 8) == 0:
            marks.append(mem()); print(f"frame {k}: device {marks[-1][0]:.0f} MiB, host RSS {marks[-1][1]:.0f} MiB", flush=True)
    dt = time.time() - t0
    r.view_projection_mat = ref_cam
    last = np.asarray(nerf.render(1920, 1080, 1, linear=False)).copy()
    same = np.array_equal(first.view(np.uint32), last.view(np.uint32))
    print(f"{n_frames} mixed calls in {dt:.1f} s; the fixed camera renders the same bits as at the start: {same}")
    # contexts come and go
    base = mem()
    for k in range(60):
        r2 = pynmr.NerfMeshRenderer(640, 360, 0); n2 = r2.load_nerf(snap); r2.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ)
        r2.frame(); n2.render(640, 360, 1, linear=False)
        del n2, r2
    end = mem()
    print(f"60 contexts created and destroyed: device {base[0]:.0f} -> {end[0]:.0f} MiB, host RSS {base[1]:.0f} -> {end[1]:.0f} MiB")
    grow_dev = marks[-1][0] - marks[2][0]; grow_host = marks[-1][1] - marks[2][1]
    ok = same and grow_dev < 64 and grow_host < 256 and end[0] - base[0] < 64 and end[1] - base[1] < 256
    print(f"growth over the last three quarters of the run: device {grow_dev:.0f} MiB, host {grow_host:.0f} MiB -> {'OK' if ok else 'LEAK?'}")
    sys.exit(0 if ok else 1)
