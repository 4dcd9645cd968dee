#!/usr/bin/env python
"""bench.py - headline benchmark of the hybrid NeRF + mesh render path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one hybrid frame (mesh stage at 2x -> hand-off -> fused march/encode/MLP/composite -> tonemap) of the
workload BASELINE.json quotes the metric on: configs[1], "bundled NeRF + glasses.gltf hybrid render at 1920x1080,
floatie removal on".  The bundled model and texture are git-LFS pointers in the reference, so the inputs are the
synthetic stand-ins of tools/synth.py (seed 1337, stock network, log2_hashmap_size 19) - see SURVEY.md 8d.
The camera follows render.py's orbit loop (orbit(-sin(1.733a)/100, cos(1.733a)/200, 0), a += 0.03 per frame).

Prints ONE JSON line (rank 0).  `value` = whole-job Mrays/s with model and mesh resident in HBM and the image left in
HBM (device time, CUDA events on the renderer's stream, L2 flushed between steps); `e2e` = the same metric through the
public API call Testbed.render() with the image copied to pinned host memory every step.
With N > 1 each rank renders its own views (weak scaling by view, no data-path collective); torch.distributed is used
for the barrier and the max-over-ranks only.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "nerf-glasses_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

ALGO_BYTES_PER_SAMPLE = 512          # 16 levels x 8 corners x 4 B table gathers
ALGO_FLOP_PER_SAMPLE = 18816


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of one march_kernel launch of this workload, from the committed
    `ncu --set full` capture (profiles/r1_march_opaque_summary.json); None when the summary is missing."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_march_opaque_summary.json")) as f:
            d = json.load(f)
        return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.samples, self.proc, self.thread = [], None, None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.samples.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(tmpdir: str, log2T: int, regime: str):
    import synth
    snap = os.path.join(tmpdir, f"synthetic_T{log2T}_{regime}.msgpack")
    synth.write_snapshot(snap, seed=1337, log2_hashmap_size=log2T, regime=regime)
    gltf = synth.write_glasses_gltf(os.path.join(tmpdir, "mesh"))
    return snap, gltf


def orbit_step(a: float):
    return -math.sin(a * 1.733) / 100.0, math.cos(a * 1.733) / 200.0, 0.0


def run_ours(args, rank: int, world: int, local_rank: int, dist):
    import pynmr
    import synth
    W, H = args.width, args.height
    with tempfile.TemporaryDirectory() as tmp:
        snap, gltf = make_inputs(tmp, args.log2_hashmap_size, args.regime)
        r = pynmr.NerfMeshRenderer(W, H, local_rank)
        nerf = r.load_nerf(snap)
        if nerf is None:
            raise RuntimeError("snapshot failed to load")
        if r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is None:
            raise RuntimeError("mesh failed to load")
        clusters, kept = r.remove_floaties()          # "floatie removal on"
    # each rank renders its own views: phase-shift the orbit so ranks do not render identical frames
    a = 0.03 * 1000 * rank
    if args.zoom:
        r.orbit(0.0, 0.0, args.zoom)

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- device-resident throughput (value) ----
    for _ in range(args.warmup):
        a += 0.03; r.orbit(*orbit_step(a)); r.frame()
    r.synchronize(); barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, march_ms, samples, alive, launches = [], [], 0, 0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        a += 0.03; r.orbit(*orbit_step(a))
        r.flush_l2()
        r.frame_async()
        st = r.stats()                                  # synchronises this step; events bracket the step's kernels only
        dev_ms.append(st["gpu_ms"]); march_ms.append(st["march_ms"]); samples += st["samples"]; alive += st["rays_alive"]; launches += st["kernel_launches"]
    r.synchronize(); barrier()
    wall_value = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    total_dev_ms = float(np.sum(dev_ms))

    # ---- end to end through the public API (e2e): Testbed.render() -> pinned host image ----
    img = None
    for _ in range(max(3, args.warmup // 2)):
        # (the result is kept while the next call runs, exactly as in the timed loop, so that both page-locked image buffers of
        # the pool exist before the clock starts: page-locking 33 MB takes ~15 ms, once)
        a += 0.03; r.orbit(*orbit_step(a)); img = nerf.render(W, H, 1, linear=False)
    r.synchronize(); barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for _ in range(args.steps):
        a += 0.03; r.orbit(*orbit_step(a))
        img = nerf.render(W, H, 1, linear=False)
        checksum += float(img[H // 2, W // 2, 0])
    r.synchronize(); barrier()
    e2e_s = time.perf_counter() - t0

    # max over ranks
    if dist is not None:
        import torch
        t = torch.tensor([total_dev_ms, e2e_s, float(samples), float(np.sum(march_ms))], dtype=torch.float64, device=f"cuda:{local_rank}")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_dev_ms, e2e_s = float(tmax[0]), float(tmax[1])
        samples_all, march_ms_all = float(tsum[2]), float(tmax[3])
    else:
        samples_all, march_ms_all = float(samples), float(np.sum(march_ms))
    if rank != 0:
        return None
    rays_all = float(W) * H * args.steps * world
    peak, peak_src = measured_peaks()
    march_s = march_ms_all / 1e3
    samples_per_launch = samples / max(1, args.steps)
    achieved = ALGO_BYTES_PER_SAMPLE * samples / max(march_s, 1e-12) / 1e9 if dist is None else ALGO_BYTES_PER_SAMPLE * samples / max(float(np.sum(march_ms)) / 1e3, 1e-12) / 1e9
    out = {
        "metric": "Mrays/s", "value": rays_all / (total_dev_ms / 1e3) / 1e6, "unit": "Mrays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": f"hybrid NeRF + glasses mesh render, {W}x{H}, 1 spp, floatie removal on (BASELINE configs[1])",
                   "model": f"synthetic iNGP snapshot seed 1337 ({args.regime}), 16-level hash grid log2_hashmap_size={args.log2_hashmap_size}, 64-wide MLPs, SH4",
                   "mesh": "glasses.gltf geometry (2952 triangles), constant stand-in texture", "camera": "render.py orbit loop from cam_pos=(0,0,2)",
                   "parallelism": "one process per GPU, views dealt to ranks, no data-path collective" if world > 1 else "single GPU",
                   "l2": "256 MiB memset between timed steps (L2 flushed)", "zoom": args.zoom},
        "fps": args.steps * world / (total_dev_ms / 1e3),
        "msamples_per_s": samples_all / (total_dev_ms / 1e3) / 1e6,
        "samples_per_frame": samples_per_launch, "rays_alive_per_frame": alive / max(1, args.steps),
        "floaties": {"clusters": clusters, "kept_cells": kept},
        "wall_s_value_loop": wall_value,
        "clocks": clocks,
        "e2e": {"value": rays_all / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 48, "d2h_bytes_per_step": W * H * 16,
                "fps": args.steps * world / e2e_s, "api": "pynmr.Testbed.render(width, height, 1, linear=False) -> pinned float32[H,W,4]"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic_bytes(),
                     "kernel": "march_kernel<tcgen05>", "peak_source": peak_src + " copy bandwidth (MEASURED_PEAKS.json)",
                     "algorithmic_bytes_per_sample": ALGO_BYTES_PER_SAMPLE, "samples_per_launch": samples_per_launch,
                     "kernel_ms_per_launch": float(np.mean(march_ms)),
                     "tensor_tflops_achieved": ALGO_FLOP_PER_SAMPLE * samples / max(float(np.sum(march_ms)) / 1e3, 1e-12) / 1e12,
                     "kernel_share_of_step": float(np.sum(march_ms)) / max(float(np.sum(dev_ms)), 1e-12)},
        "checksum": checksum,
    }
    return out


def stress_leg(args, local_rank: int, regime: str, zoom: float, steps: int = 12):
    """The sample-bound regimes (not the headline): head filling the frame, opaque or translucent medium."""
    import pynmr
    import synth
    W, H = args.width, args.height
    with tempfile.TemporaryDirectory() as tmp:
        snap, gltf = make_inputs(tmp, args.log2_hashmap_size, regime)
        r = pynmr.NerfMeshRenderer(W, H, local_rank)
        if r.load_nerf(snap) is None or r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is None:
            raise RuntimeError("stress inputs failed to load")
        r.remove_floaties()
    r.orbit(0.0, 0.0, zoom)
    a, ms, mms, smp = 0.0, [], [], 0
    for i in range(steps + 3):
        a += 0.03; r.orbit(*orbit_step(a)); r.flush_l2(); r.frame_async(); st = r.stats()
        if i >= 3:
            ms.append(st["gpu_ms"]); mms.append(st["march_ms"]); smp += st["samples"]
    tot, mtot = float(np.sum(ms)) / 1e3, float(np.sum(mms)) / 1e3
    return {"workload": f"{W}x{H} hybrid, {regime} medium, orbit zoom {zoom:g}", "steps": steps, "ms_per_frame": tot / steps * 1e3,
            "mrays_per_s": W * H * steps / tot / 1e6, "msamples_per_s": smp / tot / 1e6, "samples_per_frame": smp / steps,
            "march_gbs_algorithmic": ALGO_BYTES_PER_SAMPLE * smp / mtot / 1e9, "march_tflops_algorithmic": ALGO_FLOP_PER_SAMPLE * smp / mtot / 1e12}


def pipelined_leg(args, local_rank: int, n_frames: int = 64, repeats: int = 5):
    """Informational (not the headline): independent frames of the same workload submitted together - nmr_render_views keeps
    several in flight on the GPU, images left in HBM.  Throughput of offline / multi-view rendering; a single frame's
    latency is `value`."""
    import pynmr
    import synth
    W, H = args.width, args.height
    with tempfile.TemporaryDirectory() as tmp:
        snap, gltf = make_inputs(tmp, args.log2_hashmap_size, args.regime)
        r = pynmr.NerfMeshRenderer(W, H, local_rank)
        nerf = r.load_nerf(snap)
        if nerf is None or r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is None:
            raise RuntimeError("inputs failed to load")
        r.remove_floaties()
    a, cams = 0.0, []
    for _ in range(n_frames):
        a += 0.03; r.orbit(*orbit_step(a)); cams.append(r.view_projection_mat)
    cams = np.stack(cams)
    r.render_views(nerf, cams, W, H, to_host=False)                      # warm-up (lane allocations)
    best = None
    for _ in range(repeats):
        r.flush_l2(); r.synchronize()
        t0 = time.perf_counter()
        r.render_views(nerf, cams, W, H, to_host=False)                  # returns when every view has been rendered
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"what": f"{n_frames} frames of the orbit path in one nmr_render_views call, images left on the device, best of {repeats} (host clock around the call, L2 flushed before it)",
            "ms_per_frame": best / n_frames * 1e3, "mrays_per_s": W * H * n_frames / best / 1e6, "fps": n_frames / best}


class quiet_stdout:
    """The reference's C++ code prints to stdout (e.g. "aabb_scale: 1" in load_snapshot); bench.py's stdout carries exactly one
    JSON line, so file descriptor 1 points at /dev/null while the reference library runs."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)
        return self

    def __exit__(self, *exc):
        os.dup2(self._saved, 1)
        os.close(self._saved); os.close(self._null)
        return False


def reference_gpu_leg(args, local_rank: int):
    """Informational: the reference's OWN renderer (ngp::Testbed + tiny-cuda-nn recompiled for sm_100,
    oracle/_ref/libnmr_refgpu.so) on this GPU, same snapshot / camera / mesh buffers; device time of Testbed::render_frame."""
    from oracle import refgpu
    if not refgpu.available():
        return None
    import helpers
    import pynmr
    import synth
    W, H = args.width, args.height
    with tempfile.TemporaryDirectory() as tmp:
        snap, gltf = make_inputs(tmp, args.log2_hashmap_size, args.regime)
        ref = refgpu.ReferenceRenderer(snap)
        r = pynmr.NerfMeshRenderer(W, H, local_rank)
        nerf = r.load_nerf(snap)
        r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ)
    if args.zoom:
        r.orbit(0.0, 0.0, args.zoom)
    cam12 = np.ascontiguousarray(r.view_projection_mat.T.reshape(-1))
    _, _, _, surf, ts = helpers.debug_mesh(r, W, H)
    img_ref, ms = ref.render(cam12, W, H, 1, False, surf=surf, ts=ts, repeat=5)
    ours = np.asarray(nerf.render(W, H, 1, linear=False)).copy()
    cam = r.view_projection_mat
    for _ in range(4):                      # warm device time of the same frame, image left on the device (like the reference's figure)
        r.view_projection_mat = cam         # restarts the accumulation: every frame() is sample 0
        r.frame_async(); st = r.stats()
    d = np.abs(ours - img_ref)
    out = {"what": "reference NeRF renderer (Testbed::render_frame, mesh hand-off buffers supplied) on the same GPU, floatie removal off, best of 5",
           "ms_per_frame": ms, "mrays_per_s": W * H / ms / 1e3, "ours_ms_same_frame": st["gpu_ms"],
           "max_abs_pixel_diff": float(d.max()), "psnr_db": float(helpers.psnr(ours, img_ref)), "pixels_over_2_255": int((d.max(axis=2) > 2 / 255).sum())}
    ref.close()
    return out


def run_tiles(args, rank: int, world: int, local_rank: int, dist):
    """--mode tiles (BASELINE configs[3]): ONE frame per step, its rows dealt to the ranks in bands, gathered to rank 0 with one
    NCCL collective (pynmr/dist.py).  Strong scaling: total work per step is fixed.  Timed on the device (CUDA events on the
    stream the gather runs on, after libnmr's stream has been joined), max over ranks."""
    import torch
    import pynmr
    import synth
    from pynmr import dist as D
    W, H = args.width, args.height
    with tempfile.TemporaryDirectory() as tmp:
        snap, _ = make_inputs(tmp, args.log2_hashmap_size, args.regime)
        gltf = synth.write_lens_glasses_gltf(os.path.join(tmp, "lensmesh")) if args.lens else synth.write_glasses_gltf(os.path.join(tmp, "mesh2"))
        r = pynmr.NerfMeshRenderer(W, H, local_rank)
        if r.load_nerf(snap) is None or r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is None:
            raise RuntimeError("inputs failed to load")
        r.remove_floaties()
    if args.zoom:
        r.orbit(0.0, 0.0, args.zoom)
    r.set_surface_insertion(pynmr.NerfMeshRenderer.SURFACE_BATCH8)     # one decision for all shards
    a = 0.0
    if args.gather == "peer":
        # render + gather fused: every rank's kernels store their rows into rank 0's image over NVLink; device-side sequence
        # flags order the frames, so the timed loop holds no collective and no host synchronisation.  Timed with events on the
        # renderer's own stream (rank 0's stream ends each frame with the wait for all ranks' rows).
        ps = D.PeerShardedRenderer(r, rank, world, band=args.band, dst=0)
        stream = torch.cuda.ExternalStream(r.stream_ptr(), device=torch.device("cuda", local_rank))
        for _ in range(args.warmup):
            a += 0.03; r.orbit(*orbit_step(a)); ps.render_frame()
        r.synchronize()
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            a += 0.03; r.orbit(*orbit_step(a))
            full = ps.render_frame(sync=False)
        e1.record(stream); r.synchronize(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if full is not None:
            full = full.clone()
        ps.close()
    else:
        sr = D.ShardedRenderer(r, rank, world, band=args.band)
        for _ in range(args.warmup):
            a += 0.03; r.orbit(*orbit_step(a)); sr.render_frame(dst=0)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            a += 0.03; r.orbit(*orbit_step(a))
            full = sr.render_frame(dst=0)           # copy_device_image joins libnmr's stream before the gather is enqueued
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t[0]); dist.barrier()
    if rank != 0:
        return None
    return {"metric": "Mrays/s", "value": W * H * args.steps / (ms / 1e3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": f"one {W}x{H} hybrid frame per step{' with lens secondary rays' if args.lens else ''}, rows dealt to ranks in bands of {args.band}, {'peer stores into rank 0 image over NVLink, device-side flags' if args.gather == 'peer' else 'NCCL gather to rank 0'} (BASELINE configs[3])",
                       "model": f"synthetic iNGP snapshot seed 1337 ({args.regime}), log2_hashmap_size={args.log2_hashmap_size}", "zoom": args.zoom,
                       "parallelism": f"tiles: {world} ranks, one process per GPU, " + ("no collective (fused peer stores)" if args.gather == "peer" else "one NCCL gather per frame")},
            "fps": args.steps / (ms / 1e3), "checksum": float(full[H // 2, W // 2, 0]) if full is not None else None}


def oracle_sample(args, steps: int, warmup: int):
    """Times the CPU oracle (the reference has no CPU renderer; this is the restated reference algorithm) on a bounded
    sample of the same workload: a crop of the 1080p hybrid frame around the head, same model / mesh / camera path."""
    import synth
    from oracle import oracle as O
    O.lib().orc_set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))   # torchrun pins OMP_NUM_THREADS=1
    W, H = args.width, args.height
    cw, ch = min(W, args.cpu_crop[0]), min(H, args.cpu_crop[1])
    x0, y0 = (W - cw) // 2, (H - ch) // 2
    with tempfile.TemporaryDirectory() as tmp:
        snap_path, gltf = make_inputs(tmp, args.log2_hashmap_size, args.regime)
        snap = synth.read_snapshot(snap_path)
        g = synth.read_gltf(gltf)
        tex = np.tile(np.array([128, 128, 128, 255], dtype=np.uint8), (4, 4, 1))
    m = O.Model.from_snapshot(snap)
    m.set_bitfield(O.remove_floaties_bitfield(m.bitfield())[0])
    mesh = O.Mesh(g["positions"], g["normals"], g["texcoords"], g["indices"], synth.GLASSES_T, synth.GLASSES_S, synth.GLASSES_R_WXYZ,
                  g["base_color"], g["metallic"], g["roughness"], (0, 0, 0), tex)
    cam = O.OrbitCamera(W, H)
    if args.zoom:
        cam.orbit(0.0, 0.0, args.zoom)
    a = 0.0
    times, samples = [], 0

    def one_step():
        nonlocal a, samples
        a += 0.03; cam.orbit(*orbit_step(a))
        c12 = cam.matrix()
        rgba2, d2, _ = mesh.render(c12, 2 * W, 2 * H, window=(2 * x0, 2 * y0, 2 * (x0 + cw), 2 * (y0 + ch)))
        surf, ts = O.mesh_resolve(rgba2, d2, W, H, 2)
        P = m.params_struct(W, H, c12, aabb_min=snap["render_aabb_min"], aabb_max=snap["render_aabb_max"], window=(x0, y0, x0 + cw, y0 + ch))
        frame, _, _, st = m.render_frame(P, surf, ts)
        O.accumulate_tonemap(frame[y0:y0 + ch, x0:x0 + cw].copy(), None, 0)
        return st["samples"]

    for _ in range(warmup):
        one_step()
    for _ in range(steps):
        t0 = time.perf_counter()
        s = one_step()
        times.append(time.perf_counter() - t0); samples += s
    total = float(np.sum(times))
    return {"rays": cw * ch * steps, "samples": samples, "seconds": total, "cores": O.lib().orc_num_threads(),
            "sample": f"{cw}x{ch} centre crop of the {W}x{H} hybrid frame (mesh at 2x over the crop), {steps} steps of the orbit path"}


def run_reference(args):
    """--impl reference: the reference algorithm's CPU implementation (oracle port; the reference itself has no CPU
    renderer and its GPU build needs OptiX/GLFW) on all host threads, same config and metric."""
    res = oracle_sample(args, args.steps, min(args.warmup, 1))
    v = res["rays"] / res["seconds"] / 1e6
    return {
        "impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": res["seconds"] / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": f"hybrid NeRF + glasses mesh render, {args.width}x{args.height}, 1 spp, floatie removal on (BASELINE configs[1])",
                   "model": f"synthetic iNGP snapshot seed 1337 ({args.regime}), log2_hashmap_size={args.log2_hashmap_size}", "zoom": args.zoom},
        "msamples_per_s": res["samples"] / res["seconds"] / 1e6,
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": res["cores"], "kind": "port", "sample": res["sample"]},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--log2-hashmap-size", type=int, default=19)
    ap.add_argument("--regime", default="opaque", choices=["opaque", "translucent"])
    ap.add_argument("--zoom", type=float, default=0.0, help="orbit zoom applied before the run (0 = render.py start pose)")
    ap.add_argument("--cpu-crop", type=int, nargs=2, default=[256, 144])
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the stress and reference-on-GPU legs")
    ap.add_argument("--mode", default="views", choices=["views", "tiles"], help="views (default, the headline): every rank renders its own frames; tiles: one frame split over the ranks and gathered")
    ap.add_argument("--band", type=int, default=16)
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"], help="tiles mode: rows stored straight into rank 0's image by the render kernels (peer), or packed and gathered with NCCL")
    ap.add_argument("--lens", action="store_true", help="tiles mode: glasses with lens panes (secondary rays)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # stdout carries exactly ONE line, the JSON: whatever libraries print meanwhile (NCCL's version banner, the reference's C++
    # code) goes to stderr - file descriptor 1 points at stderr until the line is written
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        if rank == 0:
            emit(run_reference(args))
        return 0

    dist = None
    if world > 1:
        import torch
        import torch.distributed as td
        torch.cuda.set_device(local_rank)
        td.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        dist = td
    if args.mode == "tiles":
        if world == 1:
            import torch
            torch.cuda.set_device(local_rank)
        out = run_tiles(args, rank, world, local_rank, dist)
        if rank == 0:
            emit(out)
        if dist is not None:
            dist.barrier(); dist.destroy_process_group()
        return 0
    out = run_ours(args, rank, world, local_rank, dist)
    if rank == 0:
        if world == 1 and not args.no_extras:
            out["stress"] = [stress_leg(args, local_rank, "opaque", 4.0), stress_leg(args, local_rank, "translucent", 4.0)]
            try:
                out["pipelined_frames"] = pipelined_leg(args, local_rank)
            except Exception as e:
                out["pipelined_frames"] = {"error": str(e)[:200]}
            try:
                with quiet_stdout():
                    out["reference_gpu"] = reference_gpu_leg(args, local_rank)
            except Exception as e:   # test infrastructure; never fails the bench
                out["reference_gpu"] = {"error": str(e)[:200]}
        if not args.no_cpu_baseline and world == 1:      # cpu_baseline: rank 0 at N = 1 only
            res = oracle_sample(args, args.cpu_steps, 1)
            out["cpu_baseline"] = {"value": res["rays"] / res["seconds"] / 1e6, "unit": "Mrays/s", "cores": res["cores"], "kind": "port",
                                   "sample": res["sample"], "msamples_per_s": res["samples"] / res["seconds"] / 1e6}
        emit(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
