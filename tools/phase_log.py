"""Where a march tile iteration's time goes (measurement aid).  Needs the instrumented build:

    python nerf-glasses_b200/build.py -DNMR_PHASE_LOG_BUILD --out=$PWD/build_tmp/libnmr_phase.so
    NMR_LIB=$PWD/build_tmp/libnmr_phase.so NMR_PHASE_LOG=/tmp/phase.bin python tools/phase_log.py [--regime translucent --zoom 4]

Per warpgroup and tile iteration the kernel stores clock64 at: iteration start, after batch generation, after the encoding (+ the
tile barrier), after the network, after compositing.  Prints medians per iteration index in microseconds (SM clock 1.965 GHz)."""
import argparse, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "nerf-glasses_b200")):
    sys.path.insert(0, p)
import numpy as np
import pynmr, synth
ap = argparse.ArgumentParser()
ap.add_argument("--regime", default="opaque"); ap.add_argument("--zoom", type=float, default=0.0); ap.add_argument("--mhz", type=float, default=1965.0)
ap.add_argument("--serial", action="store_true", help="set-up / march overlap off")
a = ap.parse_args()
path = os.environ["NMR_PHASE_LOG"]
with tempfile.TemporaryDirectory() as d:
    snap = os.path.join(d, "s.msgpack"); synth.write_snapshot(snap, seed=1337, log2_hashmap_size=19, regime=a.regime)
    gltf = synth.write_glasses_gltf(os.path.join(d, "mesh"))
    r = pynmr.NerfMeshRenderer(1920, 1080, 0)
    nerf = r.load_nerf(snap); r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ); r.remove_floaties()
if a.zoom:
    r.orbit(0, 0, a.zoom)
if a.serial:
    r.set_overlap(False)
for i in range(5):
    r.orbit(0.01, 0.002, 0); r.flush_l2(); r.frame_async(); st = r.stats()
print(f"frame: gpu_ms {st['gpu_ms']:.3f} march_ms {st['march_ms']:.3f} samples {st['samples']} alive {st['rays_alive']} batches {st['batches']}")
IT = 32
raw6 = np.fromfile(path, dtype=np.uint64).reshape(-1, IT, 6).astype(np.int64)
raw, gt = raw6[:, :, :5], raw6[:, :, 5]
us = 1.0 / a.mhz
start = raw[:, 0, 0]
ok = start > 0
print(f"warpgroups logged: {int(ok.sum())} of {len(raw)}")
t00 = start[ok].min()       # (clock64 is per SM: offsets between SMs are only roughly comparable)
print("iter  n_wg   start(us, vs first)   generate   encode+barrier   network   composite   total")
for it in range(IT):
    v = raw[:, it, :]
    m = (v[:, 0] > 0) & (v[:, 4] > v[:, 0])
    if m.sum() == 0:
        break
    v = v[m]
    d = np.diff(v, axis=1) * us
    print(f"{it:4d} {int(m.sum()):5d}   {np.median((v[:, 0] - t00)) * us:10.1f}      {np.median(d[:, 0]):8.2f}   {np.median(d[:, 1]):8.2f}   {np.median(d[:, 2]):10.2f}   {np.median(d[:, 3]):8.2f}   {np.median((v[:, 4] - v[:, 0])) * us:8.2f}")
tot = (raw[:, :, 4] - raw[:, :, 0]) * us
valid = (raw[:, :, 0] > 0) & (raw[:, :, 4] > raw[:, :, 0])
per_wg_iters = valid.sum(axis=1)
print("iterations per warpgroup: median", np.median(per_wg_iters[ok]), "max", per_wg_iters.max(), " busy time per warpgroup (us): median %.1f max %.1f" % (np.median((tot * valid).sum(axis=1)[ok]), (tot * valid).sum(axis=1).max()))
# the frame on one clock (globaltimer): when each warpgroup starts its first tile iteration and when it leaves the loop
n_it = (gt > 0).sum(axis=1) - 1                       # (the last stamp of a warpgroup is its exit)
live = n_it >= 0
g0 = gt[live, 0].min()
first = (gt[live, 0] - g0) * 1e-3
last = (gt[live, np.maximum(n_it[live], 0)] - g0) * 1e-3
print("globaltimer, us after the first warpgroup's start:  first iteration starts  median %.1f  p99 %.1f  max %.1f   |  loop exits  p1 %.1f  median %.1f  p90 %.1f  p99 %.1f  max %.1f"
      % (np.median(first), np.percentile(first, 99), first.max(), np.percentile(last, 1), np.median(last), np.percentile(last, 90), np.percentile(last, 99), last.max()))
print("tile iterations per warpgroup: histogram", np.bincount(np.maximum(n_it[live], 0))[:IT].tolist())
for k in range(1, 8):
    m = n_it >= k
    if m.sum():
        print("  start of iteration %d: median %.1f us  p99 %.1f us  (%d warpgroups)" % (k, np.median((gt[m, k] - g0) * 1e-3), np.percentile((gt[m, k] - g0) * 1e-3, 99), int(m.sum())))
