"""Exploratory comparison: libnmr vs the C oracle vs the reference's own renderer (oracle/_ref/libnmr_refgpu.so) on the
GPU box.  Prints the figures the assertions of tests/test_gpu_vs_reference.py are set from; writes a JSON summary."""
import argparse
import json
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

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=192); ap.add_argument("--height", type=int, default=108)
ap.add_argument("--log2T", type=int, default=15); ap.add_argument("--regime", default="opaque")
ap.add_argument("--zoom", type=float, default=4.0)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "compare_reference_gpu.json"))
ap.add_argument("--time-1080p", action="store_true")
a = ap.parse_args()
W, HH = a.width, a.height
res = {}


def psnr(x, y):
    return H.psnr(x, y)


with tempfile.TemporaryDirectory() as d:
    snap_path = os.path.join(d, "s.msgpack")
    synth.write_snapshot(snap_path, seed=1337, log2_hashmap_size=a.log2T, regime=a.regime)
    gltf = synth.write_glasses_gltf(os.path.join(d, "mesh"))
    snap = synth.read_snapshot(snap_path)
    ref = refgpu.ReferenceRenderer(snap_path)
    r = pynmr.NerfMeshRenderer(W, HH)
    nerf = r.load_nerf(snap_path)
    r.orbit(0.35, -0.2, a.zoom)
    cam12 = np.ascontiguousarray(r.view_projection_mat.T.reshape(-1))
    m = O.Model.from_snapshot(snap)
    print("reference render_aabb", ref.render_aabb(), "snapshot", snap["render_aabb_min"], snap["render_aabb_max"])

    # occupancy
    rb, ob, gb = ref.bitfield(), m.bitfield(), H.get_bitfield(r, nerf)
    res["bitfield"] = {"ref_eq_oracle": bool(np.array_equal(rb, ob)), "ref_eq_gpu": bool(np.array_equal(rb, gb)), "set_bits": int(np.unpackbits(rb).sum())}
    print("bitfield", res["bitfield"])

    # encoding
    rng = np.random.default_rng(5)
    pos = rng.uniform(0, 1, size=(20000, 3)).astype(np.float32)
    e_ref = ref.encode(pos); e_gpu = H.debug_encode(r, nerf, pos); e_orc = m.encode(pos).view(np.uint16)
    f_ref, f_gpu = e_ref.view(np.float16).astype(np.float32), e_gpu.view(np.float16).astype(np.float32)
    res["encode"] = {"gpu_eq_oracle": bool(np.array_equal(e_gpu, e_orc)), "frac_bits_equal_ref": float(np.mean(e_ref == e_gpu)),
                     "max_abs_vs_ref": float(np.abs(f_ref - f_gpu).max()), "rows_equal_ref": float(np.mean(np.all(e_ref == e_gpu, axis=1)))}
    print("encode", res["encode"])

    # network
    n = 128 * 37 + 5
    pos = rng.uniform(0.3, 0.7, size=(n, 3)).astype(np.float32)
    dd = rng.normal(size=(n, 3)); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    d01 = ((dd + 1) * 0.5).astype(np.float32)
    n_ref = ref.network(pos, d01).astype(np.float32)[:, :4]
    n_gpu = H.debug_network(r, nerf, pos, d01).astype(np.float32)
    n_orc = m.network(pos, d01).astype(np.float32)
    res["network"] = {"max_abs_gpu_vs_ref": float(np.abs(n_ref - n_gpu).max()), "max_abs_oracle_vs_ref": float(np.abs(n_ref - n_orc).max()),
                      "max_abs_gpu_vs_oracle": float(np.abs(n_gpu - n_orc).max()), "mean_abs_gpu_vs_ref": float(np.abs(n_ref - n_gpu).mean()),
                      "value_range": [float(n_ref.min()), float(n_ref.max())]}
    print("network", res["network"])

    # traversal
    MS = 48
    t_ref = ref.trace(cam12, W, HH, MS)
    pixels = np.arange(W * HH, dtype=np.uint32)
    t_gpu = H.debug_trace(r, nerf, W, HH, pixels, MS)
    alive_ref, alive_gpu = t_ref["ray"][:, 7] > 0, t_gpu["ray"][:, 7] > 0
    same_cnt = t_ref["count"] == t_gpu["count"]
    both = alive_ref & alive_gpu & same_cnt
    kmax = np.arange(MS)[None, :] < t_gpu["count"][:, None]
    dpos = np.abs(t_ref["pos"] - t_gpu["pos"]).max(axis=2)
    res["trace"] = {"alive_equal": float(np.mean(alive_ref == alive_gpu)), "alive_ref": int(alive_ref.sum()), "alive_gpu": int(alive_gpu.sum()),
                    "count_equal_frac_of_alive": float(np.mean(same_cnt[alive_ref | alive_gpu])),
                    "dir_bits_equal": float(np.mean(t_ref["ray"][:, 3:6].view(np.uint32) == t_gpu["ray"][:, 3:6].view(np.uint32))),
                    "origin_bits_equal": float(np.mean(t_ref["ray"][:, 0:3].view(np.uint32) == t_gpu["ray"][:, 0:3].view(np.uint32))),
                    "t_first_bits_equal_alive": float(np.mean(t_ref["ray"][both, 6].view(np.uint32) == t_gpu["ray"][both, 6].view(np.uint32))),
                    "max_pos_diff_same_count": float(dpos[both][kmax[both]].max()) if both.any() else None,
                    "pos_bits_equal_same_count": float(np.mean((t_ref["pos"].view(np.uint32) == t_gpu["pos"].view(np.uint32))[both][kmax[both]])) if both.any() else None,
                    "samples_ref": int(t_ref["count"].sum()), "samples_gpu": int(t_gpu["count"].sum())}
    print("trace", res["trace"])

    # pixels, NeRF only
    img_ref, ms_ref = ref.render(cam12, W, HH, 1, False)
    img_gpu = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
    img_orc = H.oracle_scene(snap, W, HH, cam12)[0]
    def cmp(x, y):
        dd_ = np.abs(x - y)
        return {"max_abs": float(dd_.max()), "psnr": float(psnr(x, y)), "frac_over_2_255": float(np.mean(dd_.max(axis=2) > 2 / 255)), "p999": float(np.quantile(dd_, 0.999))}
    res["pixels_nerf"] = {"gpu_vs_ref": cmp(img_gpu, img_ref), "oracle_vs_ref": cmp(img_orc, img_ref), "gpu_vs_oracle": cmp(img_gpu, img_orc), "ref_ms": ms_ref}
    print("pixels_nerf", res["pixels_nerf"])

    # pixels, hybrid: the mesh stage output (surface colour, t_surface) comes from libnmr's own mesh stage
    g = {"path": gltf, "t": synth.GLASSES_T, "s": synth.GLASSES_S, "r": synth.GLASSES_R_WXYZ,
         "texture": np.tile(np.array([128, 128, 128, 255], dtype=np.uint8), (4, 4, 1))}
    r.load_mesh(gltf, t=g["t"], s=g["s"], r=g["r"])
    _, _, _, surf, ts = H.debug_mesh(r, W, HH)
    img_ref, ms_ref = ref.render(cam12, W, HH, 1, False, surf=surf, ts=ts)
    img_gpu = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
    img_orc = H.oracle_scene(snap, W, HH, cam12, glasses=g)[0]
    res["pixels_hybrid"] = {"gpu_vs_ref": cmp(img_gpu, img_ref), "oracle_vs_ref": cmp(img_orc, img_ref), "gpu_vs_oracle": cmp(img_gpu, img_orc),
                            "covered": float((ts > 0).mean()), "ref_ms": ms_ref}
    print("pixels_hybrid", res["pixels_hybrid"])
    np.savez_compressed(os.path.join(os.path.dirname(a.out), "compare_reference_gpu_images.npz"), ref=img_ref.astype(np.float16), gpu=img_gpu.astype(np.float16), orc=img_orc.astype(np.float16))

    if a.time_1080p:
        snap19 = os.path.join(d, "s19.msgpack")
        synth.write_snapshot(snap19, seed=1337, log2_hashmap_size=19, regime=a.regime)
        ref2 = refgpu.ReferenceRenderer(snap19)
        r2 = pynmr.NerfMeshRenderer(1920, 1080)
        nerf2 = r2.load_nerf(snap19)
        r2.load_mesh(gltf, t=g["t"], s=g["s"], r=g["r"])
        for zoom in (0.0, 4.0):
            if zoom:
                r2.orbit(0, 0, zoom)
            c12 = np.ascontiguousarray(r2.view_projection_mat.T.reshape(-1))
            _, _, _, surf, ts = H.debug_mesh(r2, 1920, 1080)
            img_ref, ms_ref = ref2.render(c12, 1920, 1080, 1, False, surf=surf, ts=ts, repeat=5)
            for _ in range(3):
                r2.frame()
            st = r2.stats()
            img_gpu = np.asarray(r2.read_frame())
            res[f"time_1080p_zoom{zoom:g}"] = {"ref_render_frame_ms": ms_ref, "ours_gpu_ms": st["gpu_ms"], "ours_samples": st["samples"], "cmp": cmp(img_gpu, img_ref)}
            print("1080p zoom", zoom, res[f"time_1080p_zoom{zoom:g}"])

os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)
