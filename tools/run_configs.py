"""Measures the BASELINE.json configs that are not the bench.py headline (configs[1]) on one GPU and prints one JSON line
per config (committed under profiles/).  Parity for the same configs lives in tests/; this is throughput only.

  C1  512x512, one view, no mesh (+ the CPU oracle on the same full frame, and pixel parity)
  C3  render.py's multi-view landmark pass: 64 views at 512x512 in one nmr_render_views call (poses: tests/golden/alice_views64.npy)
  C4  3840x2160 hybrid frame
  C5  1080p, log2_hashmap_size 19..24 (hash table 23 MiB .. 507 MiB: leaves the L2), opaque and translucent medium
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "nerf-glasses_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers as H
import pynmr
import synth

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="1,3,4,5")
ap.add_argument("--max-log2T", type=int, default=24)
a = ap.parse_args()
todo = set(a.configs.split(","))


def frames(r, n, flush=True):
    ms, mms, smp = [], [], 0
    for i in range(n + 3):
        r.orbit(0.01, 0.002, 0)
        if flush:
            r.flush_l2()
        r.frame_async(); st = r.stats()
        if i >= 3:
            ms.append(st["gpu_ms"]); mms.append(st["march_ms"]); smp += st["samples"]
    return float(np.mean(ms)), float(np.mean(mms)), smp / n


with tempfile.TemporaryDirectory() as d:
    gltf = synth.write_glasses_gltf(os.path.join(d, "mesh"))
    snap19 = os.path.join(d, "s19.msgpack")
    synth.write_snapshot(snap19, seed=1337, log2_hashmap_size=19)

    if "1" in todo:
        W = HH = 512
        r = pynmr.NerfMeshRenderer(W, HH)
        nerf = r.load_nerf(snap19)
        t0 = time.perf_counter(); n_frames = 20
        for _ in range(n_frames):
            img = nerf.render(W, HH, 1, linear=False)
        e2e_ms = (time.perf_counter() - t0) / n_frames * 1e3
        st = r.stats()
        cam12 = np.ascontiguousarray(r.view_projection_mat.T.reshape(-1))
        t0 = time.perf_counter()
        want = H.oracle_scene(synth.read_snapshot(snap19), W, HH, cam12)[0]
        cpu_s = time.perf_counter() - t0
        from oracle import oracle as O
        dd = np.abs(np.asarray(img) - want)
        print(json.dumps({"config": "C1 512x512 one view, no mesh", "gpu_ms": st["gpu_ms"], "mrays_per_s": W * HH / st["gpu_ms"] / 1e3, "e2e_ms_render_call": e2e_ms,
                          "samples": st["samples"], "cpu_oracle_s": cpu_s, "cpu_oracle_mrays_per_s": W * HH / cpu_s / 1e6, "cpu_threads": O.lib().orc_num_threads(),
                          "max_abs_vs_oracle": float(dd.max()), "psnr_vs_oracle": H.psnr(np.asarray(img), want)}), flush=True)

    if "3" in todo:
        W = HH = 512
        cams = np.load(os.path.join(ROOT, "tests", "golden", "alice_views64.npy"))
        r = pynmr.NerfMeshRenderer(W, HH)
        nerf = r.load_nerf(snap19)
        r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ)
        r.remove_floaties()
        out = r.render_views(nerf, cams, W, HH)          # warm-up (allocations, pinned pool)
        t0 = time.perf_counter(); reps = 3
        for _ in range(reps):
            out = None                                   # hand the pinned block back to the pool before asking for the next one
            out = r.render_views(nerf, cams, W, HH)
        dt = (time.perf_counter() - t0) / reps
        alive = float(np.mean(np.asarray(out)[..., :3].min(axis=-1) < 0.999))
        print(json.dumps({"config": "C3 64 views at 512x512 in one render_views call (hybrid), images copied to pinned host memory", "seconds": dt,
                          "views_per_s": len(cams) / dt, "mrays_per_s_e2e": len(cams) * W * HH / dt / 1e6, "d2h_bytes": int(np.asarray(out).nbytes),
                          "non_background_fraction": alive}), flush=True)

    if "4" in todo:
        W, HH = 3840, 2160
        r = pynmr.NerfMeshRenderer(W, HH)
        nerf = r.load_nerf(snap19)
        r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ)
        r.remove_floaties()
        for zoom in (0.0, 4.0):
            if zoom:
                r.orbit(0, 0, zoom)
            ms, mms, smp = frames(r, 10)
            print(json.dumps({"config": f"C4 3840x2160 hybrid frame, orbit zoom {zoom:g}", "gpu_ms": ms, "march_ms": mms, "fps": 1e3 / ms, "mrays_per_s": W * HH / ms / 1e3,
                              "samples_per_frame": smp, "msamples_per_s": smp / ms / 1e3}), flush=True)

    if "4" in todo:     # the same 4K frame with lens panes: mirror + transmitted segments on the covered pixels
        W, HH = 3840, 2160
        lens_gltf = synth.write_lens_glasses_gltf(os.path.join(d, "lensmesh"))
        r = pynmr.NerfMeshRenderer(W, HH)
        nerf = r.load_nerf(snap19)
        r.load_mesh(lens_gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ)
        r.remove_floaties()
        for zoom in (0.0, 4.0):
            if zoom:
                r.orbit(0, 0, zoom)
            ms, mms, smp = frames(r, 10)
            print(json.dumps({"config": f"C4 3840x2160 hybrid frame with lens secondary rays, orbit zoom {zoom:g}", "gpu_ms": ms, "march_ms": mms, "fps": 1e3 / ms,
                              "mrays_per_s": W * HH / ms / 1e3, "samples_per_frame": smp, "msamples_per_s": smp / ms / 1e3}), flush=True)

    if "5" in todo:
        W, HH = 1920, 1080
        for regime in ("opaque", "translucent"):
            for log2T in range(19, a.max_log2T + 1):
                sp = os.path.join(d, f"s{log2T}{regime}.msgpack")
                synth.write_snapshot(sp, seed=1337, log2_hashmap_size=log2T, regime=regime)
                r = pynmr.NerfMeshRenderer(W, HH)
                nerf = r.load_nerf(sp)
                os.remove(sp)
                r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ)
                r.remove_floaties()
                r.orbit(0, 0, 4.0)
                ms, mms, smp = frames(r, 6)
                table_mib = (4096 + 12168 + 29792 + 79512 + 205384 + 11 * 2 ** log2T) * 4 / 2 ** 20 if log2T <= 19 else None
                print(json.dumps({"config": f"C5 1080p hybrid zoom 4, {regime}, log2_hashmap_size {log2T}", "gpu_ms": ms, "march_ms": mms, "samples_per_frame": smp,
                                  "march_msamples_per_s": smp / mms / 1e3, "march_gbs_algorithmic": 512 * smp / mms / 1e6}), flush=True)
                del r, nerf
