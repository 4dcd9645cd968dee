"""Renders a few frames of the bench workload (no oracle, no host copies) - the command profiled under ncu."""
import argparse, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "nerf-glasses_b200")):
    sys.path.insert(0, p)
import pynmr, synth
ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--width", type=int, default=1920); ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--log2T", type=int, default=19); ap.add_argument("--regime", default="opaque"); ap.add_argument("--zoom", type=float, default=0.0)
ap.add_argument("--no-mesh", action="store_true"); ap.add_argument("--surface-mode", type=int, default=0)
a = ap.parse_args()
with tempfile.TemporaryDirectory() as d:
    snap = os.path.join(d, "s.msgpack"); synth.write_snapshot(snap, seed=1337, log2_hashmap_size=a.log2T, regime=a.regime)
    gltf = synth.write_glasses_gltf(os.path.join(d, "mesh"))
    r = pynmr.NerfMeshRenderer(a.width, a.height, 0)
    nerf = r.load_nerf(snap)
    if not a.no_mesh:
        r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ)
    r.remove_floaties()
    r.set_surface_insertion(a.surface_mode)
if a.zoom:
    r.orbit(0, 0, a.zoom)
for i in range(a.frames):
    r.orbit(0.01, 0.002, 0)
    r.frame()
    st = r.stats()
    print(f"frame {i}: gpu_ms {st['gpu_ms']:.3f} march_ms {st['march_ms']:.3f} samples {st['samples']} alive {st['rays_alive']} Msamples/s(march) {st['samples']/st['march_ms']/1e3:.1f} batches {st['batches']} passes {st['batch_passes']}")
