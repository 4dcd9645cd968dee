"""Where do hybrid frames differ from the reference's renderer?  Classifies differing pixels by mesh coverage and checks
whether the oracle replaying the reference's n_steps batching (n_steps_mode 1) closes the gap."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "nerf-glasses_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers as H
import pynmr
import synth
from oracle import oracle as O
from oracle import refgpu


def cmp(x, y):
    d = np.abs(x - y)
    return f"max {d.max():.5f} psnr {H.psnr(x, y):.1f} over2/255 {np.mean(d.max(axis=2) > 2 / 255) * 100:.4f}% ({int((d.max(axis=2) > 2 / 255).sum())} px)"


def classify(got, want, surf, ts):
    bad = np.abs(got - want).max(axis=2) > 2 / 255
    w = surf[..., 3]
    print(f"   bad pixels: {int(bad.sum())}; by coverage  w==0: {int((bad & (w == 0)).sum())}  0<w<1: {int((bad & (w > 0) & (w < 1)).sum())}  w==1: {int((bad & (w == 1)).sum())}"
          f"   (pixels with 0<w<1: {int(((w > 0) & (w < 1)).sum())}, w==1: {int((w == 1).sum())})")


with tempfile.TemporaryDirectory() as d:
    gltf = synth.write_glasses_gltf(os.path.join(d, "mesh"))
    g = {"path": gltf, "t": synth.GLASSES_T, "s": synth.GLASSES_S, "r": synth.GLASSES_R_WXYZ,
         "texture": np.tile(np.array([128, 128, 128, 255], dtype=np.uint8), (4, 4, 1))}
    for (W, HH, log2T, zoom, with_oracle) in [(192, 108, 15, 4.0, True), (480, 270, 15, 4.0, True), (1920, 1080, 19, 0.0, False), (1920, 1080, 19, 4.0, False)]:
        snap_path = os.path.join(d, f"s{log2T}.msgpack")
        if not os.path.exists(snap_path):
            synth.write_snapshot(snap_path, seed=1337, log2_hashmap_size=log2T)
        snap = synth.read_snapshot(snap_path)
        ref = refgpu.ReferenceRenderer(snap_path)
        r = pynmr.NerfMeshRenderer(W, HH)
        nerf = r.load_nerf(snap_path)
        r.orbit(0.35, -0.2, zoom)
        cam12 = np.ascontiguousarray(r.view_projection_mat.T.reshape(-1))
        print(f"=== {W}x{HH} log2T {log2T} zoom {zoom}")
        img_ref, _ = ref.render(cam12, W, HH, 1, False)
        img_gpu = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
        print("  NeRF only   gpu vs ref:", cmp(img_gpu, img_ref), " alive", r.stats()["rays_alive"])
        r.load_mesh(gltf, t=g["t"], s=g["s"], r=g["r"])
        _, _, _, surf, ts = H.debug_mesh(r, W, HH)
        img_ref, _ = ref.render(cam12, W, HH, 1, False, surf=surf, ts=ts)
        img_gpu = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
        print("  hybrid      gpu vs ref:", cmp(img_gpu, img_ref))
        classify(img_gpu, img_ref, surf, ts)
        if with_oracle:
            for mode in (0, 1):
                img_orc = H.oracle_scene(snap, W, HH, cam12, glasses=g, n_steps_mode=mode)[0]
                print(f"  hybrid oracle n_steps_mode {mode} vs ref:", cmp(img_orc, img_ref))
                classify(img_orc, img_ref, surf, ts)
        ref.close()
