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
public API call Testbed.render() with the whole image copied to pinned host memory every step (float32; `e2e_u8` = the same
call returning sRGB8).  `e2e_update` (extra) = Testbed.render_update(): the caller keeps the host image and only the screen
rectangle of head + mesh is moved - same pixels, `d2h_bytes_per_step` = what was actually copied.
`roofline` = the march kernel alone (K further frames with the set-up / march overlap off, CUDA events around the kernel)
against the L2 rate measured in this run, the sustained tensor rate and the HBM rate; `cpu_baseline` / `--impl reference` = the
oracle port on the host cores over the same full frame.
With N > 1 each rank renders its own views (weak scaling by view, no data-path collective); torch.distributed is used
for the barrier and the max-over-ranks only.  Every N also reports `tiles_4k` (BASELINE configs[3]: one 3840x2160 lens frame
split over the ranks, fused peer stores and NCCL gather, bit-identity to the single-GPU frame) and `views_c3` (configs[2]: 64
views at 512x512 dealt to the ranks); N = 1 adds the stress frames, the hash-map sweep of configs[4], pipelined frames and the
reference's own renderer on this GPU (`reference_gpu`).
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
    """(HBM copy GB/s, sustained dense bf16 TFLOP/s, source) from the driver-written MEASURED_PEAKS.json, else the profiling
    recipe's fallbacks."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", 1418.0)), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1418.0, "fallback (B200_PROFILING.md)"


NCU_SUMMARIES = ("r2_march_opaque_summary.json", "r1_march_opaque_summary.json")     # newest first


def ncu_capture():
    """Counters of one march_kernel launch of this workload from the committed `ncu --set full` capture under profiles/ (they
    cannot be read live: ncu replays a kernel ~40 times).  -> (dict, file name) or (None, None)."""
    for name in NCU_SUMMARIES:
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return json.load(f), "profiles/" + name
        except Exception:
            continue
    return None, None


def bind_to_gpu_numa(gpu_index: int):
    """N > 1: run this rank (and allocate its page-locked image buffers, first touch) on the CPU cores next to its GPU, so that
    eight ranks' device->host copies do not all land in one socket's memory.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
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


def workload_config(args, world: int) -> dict:
    """The `config` object both arms print: identical keys and values for the same workload."""
    W, H = args.width, args.height
    return {"workload": f"hybrid NeRF + glasses mesh render, {W}x{H}, 1 spp, floatie removal on (BASELINE configs[1])",
            "model": f"synthetic iNGP snapshot seed 1337 ({args.regime}), 16-level hash grid log2_hashmap_size={args.log2_hashmap_size}, 64-wide MLPs, SH4",
            "mesh": "glasses.gltf geometry (2952 triangles), constant stand-in texture", "camera": "render.py orbit loop from cam_pos=(0,0,2)",
            "zoom": args.zoom,
            "parallelism": "one process per GPU, views dealt to ranks, no data-path collective" if world > 1 else "single GPU",
            "timing": "GPU arm: CUDA events around each frame's kernels, 256 MiB memset between timed steps (L2 flushed); reference arm: host clock around each full frame of the CPU port (rank 0, all host threads)",
            "parity": "pixels <= 2/255 and >= 45 dB vs the CPU oracle and vs the reference's own renderer recompiled for sm_100 (tests/); "
                      "the mesh stage (OptiX in the reference) and lens secondary rays are pinned on the oracle only"}


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
    span_ms = list(march_ms)      # overlapped frames: the march kernel's span by device timestamps, waiting for rays included

    # ---- the dominant kernel alone (roofline): the same K frames with the set-up / march overlap off, so that CUDA events on the
    # renderer's stream bracket the march kernel and nothing else ----
    r.set_overlap(False)
    march_ms, samples_k, serial_ms = [], 0, []
    for _ in range(3):
        a += 0.03; r.orbit(*orbit_step(a)); r.frame()
    for _ in range(args.steps):
        a += 0.03; r.orbit(*orbit_step(a))
        r.flush_l2()
        r.frame_async()
        st = r.stats()
        march_ms.append(st["march_ms"]); samples_k += st["samples"]; serial_ms.append(st["gpu_ms"])
    r.set_overlap(True)

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

    # ---- the same call returning what render.py makes of the image right away, np.uint8(img * 255), converted on the device ----
    for _ in range(3):
        a += 0.03; r.orbit(*orbit_step(a)); img = nerf.render(W, H, 1, linear=False, dtype=np.uint8)
    r.synchronize(); barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        a += 0.03; r.orbit(*orbit_step(a))
        img = nerf.render(W, H, 1, linear=False, dtype=np.uint8)
        checksum += float(img[H // 2, W // 2, 0]) / 255.0
    r.synchronize(); barrier()
    e2e_u8_s = time.perf_counter() - t0

    # ---- the image kept by the caller and updated in place (Testbed.render_update, an addition): the same complete image in host
    # memory after every call, but only the screen rectangle of head + mesh crosses PCIe - reported beside e2e, not instead of it ----
    upd = {}
    for name, dt in (("float32", np.float32), ("uint8", np.uint8)):
        held = None
        for _ in range(3):
            a += 0.03; r.orbit(*orbit_step(a)); held = nerf.render_update(held, W, H, linear=False, dtype=dt)
        r.synchronize(); barrier()
        t0 = time.perf_counter()
        moved = 0
        for _ in range(args.steps):
            a += 0.03; r.orbit(*orbit_step(a))
            held = nerf.render_update(held, W, H, linear=False, dtype=dt)
            checksum += float(held[H // 2, W // 2, 0]) / (255.0 if dt is np.uint8 else 1.0)
            moved += nerf.last_update_bytes
        r.synchronize(); barrier()
        upd[name] = (time.perf_counter() - t0, moved / args.steps)
        del held

    # what the L2 of this GPU delivers (roofline denominators; rank 0)
    l2_copy_gbs = l2_gather_gbs = None
    if rank == 0:
        l2_copy_gbs, l2_gather_gbs = r.measure_l2(32 << 20, gather=False), r.measure_l2(32 << 20, gather=True)

    # max over ranks
    if dist is not None:
        import torch
        t = torch.tensor([total_dev_ms, e2e_s, float(samples), float(np.sum(march_ms)), e2e_u8_s, upd["float32"][0], upd["uint8"][0]], dtype=torch.float64, device=f"cuda:{local_rank}")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_dev_ms, e2e_s, e2e_u8_s = float(tmax[0]), float(tmax[1]), float(tmax[4])
        upd = {"float32": (float(tmax[5]), upd["float32"][1]), "uint8": (float(tmax[6]), upd["uint8"][1])}
        samples_all = float(tsum[2])
    else:
        samples_all = float(samples)
    if rank != 0:
        return None
    rays_all = float(W) * H * args.steps * world
    hbm_peak, tensor_peak, peak_src = measured_peaks()
    march_s = float(np.sum(march_ms)) / 1e3                   # rank 0's own launches: its samples over its kernel time
    samples_per_launch = samples_k / max(1, args.steps)
    achieved = ALGO_BYTES_PER_SAMPLE * samples_k / max(march_s, 1e-12) / 1e9
    tflops = ALGO_FLOP_PER_SAMPLE * samples_k / max(march_s, 1e-12) / 1e12
    cap, cap_file = ncu_capture()
    out = {
        "metric": "Mrays/s", "value": rays_all / (total_dev_ms / 1e3) / 1e6, "unit": "Mrays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": workload_config(args, world),
        "fps": args.steps * world / (total_dev_ms / 1e3),
        "ms_per_step_median": float(np.median(dev_ms)), "ms_per_step_max": float(np.max(dev_ms)),
        "msamples_per_s": samples_all / (total_dev_ms / 1e3) / 1e6,
        "samples_per_frame": samples_per_launch, "rays_alive_per_frame": alive / max(1, args.steps),
        "floaties": {"clusters": clusters, "kept_cells": kept},
        "wall_s_value_loop": wall_value,
        "clocks": clocks,
        "e2e": {"value": rays_all / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 48, "d2h_bytes_per_step": W * H * 16,
                "fps": args.steps * world / e2e_s, "api": "pynmr.Testbed.render(width, height, 1, linear=False) -> pinned float32[H,W,4]"},
        "e2e_u8": {"value": rays_all / e2e_u8_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 48, "d2h_bytes_per_step": W * H * 4,
                   "fps": args.steps * world / e2e_u8_s,
                   "api": "pynmr.Testbed.render(width, height, 1, linear=False, dtype=np.uint8) -> pinned uint8[H,W,4] == np.uint8(float_image * 255), render.py's next line"},
        "e2e_update": {k: {"value": rays_all / v[0] / 1e6, "unit": "Mrays/s", "fps": args.steps * world / v[0], "h2d_bytes_per_step": 48,
                           "d2h_bytes_per_step": v[1], "full_image_bytes": W * H * (16 if k == "float32" else 4)} for k, v in upd.items()}
                      | {"api": "img = pynmr.Testbed.render_update(img, width, height, linear=False, dtype=...): the caller keeps the pinned image, the call "
                                "moves the screen rectangle of head + mesh (old and new) and leaves the background pixels, which are already there; "
                                "bit-identical to render() after every call (tests/test_gpu_update.py); d2h_bytes_per_step = bytes actually moved (rank 0)"},
        "gpu_launches": int(launches),
        # The dominant kernel gathers from a hash table that is L2-resident (23 MiB at log2_hashmap_size 19, DRAM traffic per launch
        # ~0.1x the algorithmic bytes): the bound is the L2 -> SM path, measured on this GPU by nmr_measure_l2 just now.
        "roofline": {"bound": "l2", "achieved": achieved, "peak": l2_copy_gbs, "unit": "GB/s", "frac": achieved / l2_copy_gbs if l2_copy_gbs else None,
                     "frac_l2": achieved / l2_copy_gbs if l2_copy_gbs else None,
                     "frac_l2_gather": achieved / l2_gather_gbs if l2_gather_gbs else None,
                     "frac_hbm": achieved / hbm_peak, "frac_tensor": tflops / tensor_peak,
                     "peak_l2_gbs": l2_copy_gbs, "peak_l2_gather_4B_gbs": l2_gather_gbs, "peak_hbm_gbs": hbm_peak, "peak_tensor_tflops": tensor_peak,
                     "peak_source": "L2: nmr_measure_l2 in this run (32 MiB resident buffer: coalesced 16-byte loads / independent 4-byte gathers, best of 5); "
                                    "HBM and dense bf16 (fp16 proxy): " + peak_src,
                     "traffic": (float(cap["dram_bytes_read"]) + float(cap["dram_bytes_write"])) if cap else None, "traffic_from": cap_file,
                     "l1_sectors_per_sample": (float(cap["l1_global_load_sectors"]) / max(1.0, float(cap.get("samples", 0)))) if cap and cap.get("samples") else None,
                     "tensor_pipe_active_pct_ncu": cap.get("tensor_pipe_active_pct") if cap else None,
                     "l2_hit_rate_pct_ncu": cap.get("l2_hit_rate_pct") if cap else None,
                     # what actually limits the kernel: instruction issue - ~4100 warp instructions per 32 samples
                     # at ~55 % of the issue slots; none of the memory roofs above is near
                     "issue_slot_utilisation_pct_ncu": cap.get("issue_slot_utilisation_pct") if cap else None,
                     "warp_instructions_per_32_samples_ncu": cap.get("warp_instructions_per_32_samples") if cap else None,
                     "kernel": "march_kernel<tcgen05>",
                     "algorithmic_bytes_per_sample": ALGO_BYTES_PER_SAMPLE, "algorithmic_flop_per_sample": ALGO_FLOP_PER_SAMPLE,
                     "samples_per_launch": samples_per_launch, "kernel_ms_per_launch": float(np.mean(march_ms)),
                     "tensor_tflops_achieved": tflops,
                     "measured_on": f"{args.steps} further frames of the same path with the set-up / march overlap switched off (nmr_set_overlap 0), "
                                    "CUDA events on the renderer's stream around the march kernel, L2 flushed between frames",
                     "kernel_share_of_step": float(np.sum(march_ms)) / max(float(np.sum(serial_ms)), 1e-12),
                     "serial_ms_per_step": float(np.mean(serial_ms)),
                     "overlapped_kernel_span_ms": float(np.mean(span_ms))},
        "checksum": checksum,
    }
    return out


def tiles_block(args, rank: int, world: int, local_rank: int, dist, steps: int = 10, warmup: int = 3):
    """BASELINE configs[3] inside the default bench line: ONE 3840x2160 hybrid frame with lens secondary rays per step, rows dealt
    to the ranks in bands.  Strong scaling.  Measures the single-GPU frame (every rank, unsharded), the fused path (render kernels
    store straight into rank 0's image over NVLink, device-side sequence flags) and the NCCL-gather path, device-timed with CUDA
    events, max over ranks; rank 0 compares both assembled frames with its own single-GPU frame of the same camera bit for bit."""
    import torch
    import pynmr
    import synth
    from pynmr import dist as D
    FW, FH, band = 3840, 2160, 16
    dev = torch.device("cuda", local_rank)
    with tempfile.TemporaryDirectory() as tmp:
        snap, _ = make_inputs(tmp, args.log2_hashmap_size, args.regime)
        gltf = synth.write_lens_glasses_gltf(os.path.join(tmp, "lensmesh"))
        r = pynmr.NerfMeshRenderer(FW, FH, local_rank)
        if r.load_nerf(snap) is None or r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is None:
            raise RuntimeError("inputs failed to load")
        r.remove_floaties()
    r.set_surface_insertion(pynmr.NerfMeshRenderer.SURFACE_BATCH8)     # one ray-local rule for the single-GPU frame and all shards
    stream = torch.cuda.ExternalStream(r.stream_ptr(), device=dev)
    a, path = 0.0, []                                     # the camera path, the same list on every rank and for every leg
    for _ in range(warmup + steps):
        a += 0.03; r.orbit(*orbit_step(a)); path.append(r.view_projection_mat)

    def barrier():
        if dist is not None:
            dist.barrier()

    def reduce_max(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def timed(frame_fn, sync_fn) -> float:
        """warm-up, then `steps` frames of the path between two events on the renderer's stream -> ms per frame, max over ranks"""
        for k in range(warmup):
            r.view_projection_mat = path[k]; frame_fn()
        sync_fn(); barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(warmup, warmup + steps):
            r.view_projection_mat = path[k]; frame_fn()
        e1.record(stream); sync_fn(); torch.cuda.synchronize(dev)
        return reduce_max(e0.elapsed_time(e1) / steps)

    # single GPU, unsharded (every rank does the same work; rank 0 keeps the last frame of the path)
    n1_ms = timed(r.frame_async, r.synchronize)
    single = torch.from_numpy(np.asarray(r.read_frame()).copy()) if rank == 0 else None
    out = {"workload": f"one {FW}x{FH} hybrid frame with lens secondary rays per step, rows dealt to {world} rank(s) in bands of {band} (BASELINE configs[3]); "
                       f"{steps} steps after {warmup} warm-up, CUDA events on the renderer's stream, max over ranks",
           "n1_ms_per_frame": n1_ms, "n1_mrays_per_s": FW * FH / n1_ms / 1e3}
    if world == 1:
        return out
    for name in ("peer", "nccl"):
        if name == "peer":
            sr = D.PeerShardedRenderer(r, rank, world, band=band, dst=0)
            last = [None]

            def frame_fn():
                last[0] = sr.render_frame(sync=False)
            sync_fn = r.synchronize
        else:
            sr = D.ShardedRenderer(r, rank, world, band=band)
            last = [None]

            def frame_fn():
                last[0] = sr.render_frame(dst=0)          # copy_device_image joins libnmr's stream before the gather is enqueued
            sync_fn = lambda: torch.cuda.synchronize(dev)
        try:
            if name == "peer":
                ms = timed(frame_fn, sync_fn)
            else:                                         # the gather runs on torch's stream: events there
                for k in range(warmup):
                    r.view_projection_mat = path[k]; frame_fn()
                sync_fn(); barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for k in range(warmup, warmup + steps):
                    r.view_projection_mat = path[k]; frame_fn()
                e1.record(); sync_fn()
                ms = reduce_max(e0.elapsed_time(e1) / steps)
            full = last[0].clone().cpu() if (rank == 0 and last[0] is not None) else None
        finally:
            barrier()
            if name == "peer":
                sr.close()
            else:
                r.set_shard(0, 1, band)
        same = bool(torch.equal(full.view(torch.int32), single.view(torch.int32))) if rank == 0 and full is not None else None
        out[name] = {"ms_per_frame": ms, "mrays_per_s": FW * FH / ms / 1e3, "speedup_vs_n1": n1_ms / ms, "bit_identical_to_single_gpu": same,
                     "how": "render kernels store their rows into rank 0's image over NVLink (CUDA IPC), device-side sequence flags, no collective" if name == "peer"
                            else "owned rows packed, one NCCL gather to rank 0, scattered back"}
    return out


def views_block(args, rank: int, world: int, local_rank: int, dist):
    """BASELINE configs[2] inside the default bench line: render.py's landmark pass - the 64 camera poses of
    tests/golden/alice_views64.npy at 512x512 (hybrid) - dealt to the ranks by view, each rank's block rendered by ONE
    nmr_render_views call straight into device memory, then gathered to rank 0 with one NCCL gather.  Device-timed (events around
    render + gather), max over ranks; rank 0 then renders all 64 views alone and compares bit for bit."""
    import torch
    import pynmr
    import synth
    from pynmr import dist as D
    w = h = 512
    dev = torch.device("cuda", local_rank)
    cams = np.load(os.path.join(ROOT, "tests", "golden", "alice_views64.npy"))
    n_views = len(cams)
    with tempfile.TemporaryDirectory() as tmp:
        snap, gltf = make_inputs(tmp, args.log2_hashmap_size, args.regime)
        r = pynmr.NerfMeshRenderer(w, h, local_rank)
        nerf = r.load_nerf(snap)
        if nerf is None or r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is None:
            raise RuntimeError("inputs failed to load")
        r.remove_floaties()
    sl = D.view_slice(n_views, rank, world)
    mine = torch.empty((len(sl), h, w, 4), dtype=torch.float32, device=dev)

    def one_pass():
        if len(sl):
            r.render_views(nerf, cams[sl.start:sl.stop], w, h, linear=False, out_device_ptr=mine.data_ptr())
        return D.gather_views(mine, n_views, rank, world, dst=0) if world > 1 else mine

    one_pass(); torch.cuda.synchronize(dev)               # warm-up (lane allocations, NCCL channels)
    best, full = None, None
    for _ in range(3):
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        full = one_pass()
        e1.record(); torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t[0])
        best = ms if best is None else min(best, ms)
    same = None
    if rank == 0:
        alone = torch.empty((n_views, h, w, 4), dtype=torch.float32, device=dev)
        r.render_views(nerf, cams, w, h, linear=False, out_device_ptr=alone.data_ptr())
        torch.cuda.synchronize(dev)
        same = bool(torch.equal(full.view(torch.int32), alone.view(torch.int32)))
    if dist is not None:
        dist.barrier()
    return {"workload": f"{n_views} views at {w}x{h} (hybrid, poses of the reference's bundled dataset), dealt to {world} rank(s) by view, images gathered to rank 0 on the device (BASELINE configs[2]); best of 3, CUDA events around render + gather, max over ranks",
            "seconds": best / 1e3, "views_per_s": n_views / (best / 1e3), "mrays_per_s": n_views * w * h / best / 1e3,
            "bit_identical_to_single_gpu": same, "gathered_bytes": int(n_views * w * h * 16) if world > 1 else 0}


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


def hashmap_sweep_leg(args, local_rank: int, sizes=(19, 22, 24), steps: int = 6):
    """BASELINE configs[4]: the encoding-bound stress - synthetic models with log2_hashmap_size 19 / 22 / 24 (hash tables of 23 /
    151 / 507 MiB: the larger ones leave the L2), 1080p hybrid frame with the head filling it and a translucent medium (13 M
    samples per frame).  Device time of the march kernel, L2 flushed between frames."""
    import pynmr
    import synth
    W, H = args.width, args.height
    out = []
    for log2T in sizes:
        with tempfile.TemporaryDirectory() as tmp:
            snap, gltf = make_inputs(tmp, log2T, "translucent")
            r = pynmr.NerfMeshRenderer(W, H, local_rank)
            nerf = r.load_nerf(snap)
            if nerf is None or r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is None:
                raise RuntimeError("inputs failed to load")
            os.remove(snap)
            r.remove_floaties()
        table_mib = (nerf.n_params - 10240) * 2 / 2 ** 20
        r.orbit(0.0, 0.0, 4.0)
        a, ms, mms, smp = 0.0, [], [], 0
        for i in range(steps + 3):
            a += 0.03; r.orbit(*orbit_step(a)); r.flush_l2(); r.frame_async(); st = r.stats()
            if i >= 3:
                ms.append(st["gpu_ms"]); mms.append(st["march_ms"]); smp += st["samples"]
        tot, mtot = float(np.sum(ms)) / 1e3, float(np.sum(mms)) / 1e3
        out.append({"log2_hashmap_size": log2T, "table_mib": round(table_mib, 1), "ms_per_frame": tot / steps * 1e3, "samples_per_frame": smp / steps,
                    "march_msamples_per_s": smp / mtot / 1e6, "march_gbs_algorithmic": ALGO_BYTES_PER_SAMPLE * smp / mtot / 1e9})
        del r, nerf
    return {"workload": f"{W}x{H} hybrid, translucent medium, orbit zoom 4, log2_hashmap_size sweep (BASELINE configs[4]); {steps} frames each, L2 flushed between frames", "sizes": out}


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


def reference_gpu_leg(args, local_rank: int, regime=None, zoom=None):
    """Informational: the reference's OWN renderer (ngp::Testbed + tiny-cuda-nn recompiled for sm_100,
    oracle/_ref/libnmr_refgpu.so) on this GPU, same snapshot / camera / mesh buffers; device time of Testbed::render_frame.
    regime / zoom select one of the sample-bound stress workloads instead of the bench frame."""
    from oracle import refgpu
    if not refgpu.available():
        return None
    import helpers
    import pynmr
    import synth
    W, H = args.width, args.height
    regime = regime or args.regime
    zoom = args.zoom if zoom is None else zoom
    with tempfile.TemporaryDirectory() as tmp:
        snap, gltf = make_inputs(tmp, args.log2_hashmap_size, regime)
        ref = refgpu.ReferenceRenderer(snap)
        r = pynmr.NerfMeshRenderer(W, H, local_rank)
        nerf = r.load_nerf(snap)
        r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ)
    if zoom:
        r.orbit(0.0, 0.0, zoom)
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
           "workload": f"{W}x{H} hybrid, {regime} medium, orbit zoom {zoom:g}", "samples_ours": int(st["samples"]),
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


def mesh_window(world_positions: np.ndarray, c12, W2: int, H2: int):
    """Screen bounding box (supersampled pixels, padded) of the mesh: outside it no mesh ray can hit a triangle, so the oracle's
    brute-force mesh stage (every pixel against every triangle) is only run there - the image is the full frame's."""
    M = np.array(c12[:9], dtype=np.float64).reshape(3, 3).T            # columns U, V, W
    q = world_positions.astype(np.float64) - np.array(c12[9:12], dtype=np.float64)
    abw = np.linalg.solve(M, q.T).T
    if not (abw[:, 2] > 1e-3).all():
        return None
    px = (abw[:, 0] / abw[:, 2] + 1.0) * 0.5 * W2 - 0.5
    py = (abw[:, 1] / abw[:, 2] + 1.0) * 0.5 * H2 - 0.5
    x0, y0 = max(0, int(np.floor(px.min())) - 3), max(0, int(np.floor(py.min())) - 3)
    x1, y1 = min(W2, int(np.ceil(px.max())) + 4), min(H2, int(np.ceil(py.max())) + 4)
    return (x0, y0, x1, y1) if x1 > x0 and y1 > y0 else (0, 0, 1, 1)


def oracle_sample(args, steps: int, warmup: int):
    """Times the CPU oracle (the reference has no CPU renderer; this is the restated reference algorithm, its wavefront loop
    included) on the SAME workload as the GPU arm: the full 1920x1080 hybrid frame, same model / mesh / floatie removal / camera
    path, `steps` frames after `warmup`.  Bounded by the number of steps, not by a crop."""
    import synth
    from oracle import oracle as O
    O.lib().orc_set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))   # torchrun pins OMP_NUM_THREADS=1
    W, H = args.width, args.height
    with tempfile.TemporaryDirectory() as tmp:
        snap_path, gltf = make_inputs(tmp, args.log2_hashmap_size, args.regime)
        snap = synth.read_snapshot(snap_path)
        g = synth.read_gltf(gltf)
        tex = np.tile(np.array([128, 128, 128, 255], dtype=np.uint8), (4, 4, 1))
    m = O.Model.from_snapshot(snap)
    m.set_bitfield(O.remove_floaties_bitfield(m.bitfield())[0])
    mesh = O.Mesh(g["positions"], g["normals"], g["texcoords"], g["indices"], synth.GLASSES_T, synth.GLASSES_S, synth.GLASSES_R_WXYZ,
                  g["base_color"], g["metallic"], g["roughness"], (0, 0, 0), tex)
    wpos = mesh.world_positions()
    cam = O.OrbitCamera(W, H)
    if args.zoom:
        cam.orbit(0.0, 0.0, args.zoom)
    a = 0.0
    times, samples = [], 0

    def one_step():
        nonlocal a
        a += 0.03; cam.orbit(*orbit_step(a))
        c12 = cam.matrix()
        rgba2, d2, _ = mesh.render(c12, 2 * W, 2 * H, window=mesh_window(wpos, c12, 2 * W, 2 * H))
        surf, ts = O.mesh_resolve(rgba2, d2, W, H, 2)
        P = m.params_struct(W, H, c12, aabb_min=snap["render_aabb_min"], aabb_max=snap["render_aabb_max"], n_steps_mode=1)
        frame, _, _, st = m.render_frame(P, surf, ts)
        O.accumulate_tonemap(frame, None, 0)
        return st["samples"]

    for _ in range(warmup):
        one_step()
    for _ in range(steps):
        t0 = time.perf_counter()
        s = one_step()
        times.append(time.perf_counter() - t0); samples += s
    total = float(np.sum(times))
    return {"rays": W * H * steps, "samples": samples, "seconds": total, "cores": O.lib().orc_num_threads(),
            "sample": f"the full {W}x{H} hybrid frame (mesh stage at 2x inside the mesh's screen rectangle, NeRF wavefront loop over all {W * H} rays, accumulate + tonemap), {steps} steps of the orbit path"}


def run_reference(args):
    """--impl reference: the reference algorithm's CPU implementation (oracle port; the reference itself has no CPU
    renderer and its GPU build needs OptiX/GLFW) on all host threads, same config, metric and warm-up."""
    res = oracle_sample(args, args.steps, args.warmup)
    v = res["rays"] / res["seconds"] / 1e6
    return {
        "impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["seconds"] / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": workload_config(args, int(os.environ.get("WORLD_SIZE", "1"))),
        "msamples_per_s": res["samples"] / res["seconds"] / 1e6,
        "fps": args.steps / res["seconds"],
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
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the stress, reference-on-GPU, 4K-tiles and 64-views legs")
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
        bind_to_gpu_numa(local_rank)
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
    # BASELINE configs[3] and [2] at this N, inside the same line (every rank takes part): the tile-sharded 4K frame with its
    # gather paths and the view-sharded landmark pass, each with a bit-identity check against one GPU on rank 0
    tiles = views = None
    if not args.no_extras:
        import torch
        torch.cuda.set_device(local_rank)
        try:
            tiles = tiles_block(args, rank, world, local_rank, dist)
        except Exception as e:
            tiles = {"error": f"{type(e).__name__}: {e}"[:300]}
        try:
            views = views_block(args, rank, world, local_rank, dist)
        except Exception as e:
            views = {"error": f"{type(e).__name__}: {e}"[:300]}
    if rank == 0:
        if tiles is not None:
            out["tiles_4k"] = tiles
        if views is not None:
            out["views_c3"] = views
        if world == 1 and not args.no_extras:
            out["stress"] = [stress_leg(args, local_rank, "opaque", 4.0), stress_leg(args, local_rank, "translucent", 4.0)]
            try:
                out["hashmap_sweep_c5"] = hashmap_sweep_leg(args, local_rank)
            except Exception as e:
                out["hashmap_sweep_c5"] = {"error": str(e)[:200]}
            try:
                out["pipelined_frames"] = pipelined_leg(args, local_rank)
            except Exception as e:
                out["pipelined_frames"] = {"error": str(e)[:200]}
            try:
                with quiet_stdout():
                    out["reference_gpu"] = reference_gpu_leg(args, local_rank)
                    if out["reference_gpu"] is not None:      # where kernel quality, not host synchronisation, decides: the sample-bound frames
                        out["reference_gpu"]["stress"] = [reference_gpu_leg(args, local_rank, "opaque", 4.0), reference_gpu_leg(args, local_rank, "translucent", 4.0)]
            except Exception as e:   # test infrastructure; never fails the bench
                out["reference_gpu"] = {"error": str(e)[:200]} if not isinstance(out.get("reference_gpu"), dict) else dict(out["reference_gpu"], stress_error=str(e)[:200])
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
