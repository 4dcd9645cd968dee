"""Per-call time distribution of Testbed.render() on the bench workload (host clock), to see what the e2e figure is made of.
    python tools/e2e_jitter.py [uint8|float16|float32]
Prints host time per call (orbit(), render(), reading one pixel) and the device time of the frame inside the call."""
import math, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "nerf-glasses_b200")):
    sys.path.insert(0, p)
import numpy as np
import pynmr, synth
dtype = np.dtype(sys.argv[1]) if len(sys.argv) > 1 else np.dtype(np.float32)
W, H = 1920, 1080
with tempfile.TemporaryDirectory() as d:
    snap = os.path.join(d, "s.msgpack"); synth.write_snapshot(snap, seed=1337, log2_hashmap_size=19)
    gltf = synth.write_glasses_gltf(os.path.join(d, "mesh"))
    r = pynmr.NerfMeshRenderer(W, H, 0)
    nerf = r.load_nerf(snap); r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ); r.remove_floaties()
a = 0.0
def step():
    global a
    a += 0.03
    r.orbit(-math.sin(a * 1.733) / 100.0, math.cos(a * 1.733) / 200.0, 0.0)
for _ in range(10):
    step(); img = nerf.render(W, H, 1, linear=False, dtype=dtype)
ts, to, tr, tg = [], [], [], []
for _ in range(400):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter()
    img = nerf.render(W, H, 1, linear=False, dtype=dtype); t2 = time.perf_counter()
    c = float(img[H // 2, W // 2, 0]); t3 = time.perf_counter()
    ts.append(t3 - t0); to.append(t1 - t0); tr.append(t2 - t1); tg.append(r.stats()["gpu_ms"])
ts, to, tr, tg = np.array(ts) * 1e3, np.array(to) * 1e3, np.array(tr) * 1e3, np.array(tg)
print(f"dtype {dtype}, env: " + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("NMR_") and k != "NMR_LIB"))
for name, v in (("step total", ts), ("orbit()", to), ("render()", tr), ("frame on GPU", tg)):
    print(f"{name:12s} ms: mean {v.mean():.3f} p5 {np.percentile(v, 5):.3f} p50 {np.percentile(v, 50):.3f} p95 {np.percentile(v, 95):.3f} p99 {np.percentile(v, 99):.3f} max {v.max():.3f}")
print("calls over 1.0 ms:", int((tr > 1.0).sum()), " fps from mean:", 1e3 / ts.mean())
