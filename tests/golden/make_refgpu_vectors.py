"""Generates tests/golden/refgpu_vectors.npz from the REFERENCE'S OWN renderer running on a GPU
(oracle/_ref/libnmr_refgpu.so = ngp::Testbed + tiny-cuda-nn compiled from /root/reference by oracle/Makefile.refgpu):

    gpurun -- 'python tests/golden/make_refgpu_vectors.py gpurun_out/refgpu_vectors.npz'     # then copy it into tests/golden/

Inputs are the seeded synthetic snapshot of the test-suite (tools/synth.py, seed 1337, log2_hashmap_size 15) and seeded
positions; outputs are what the reference's kernels return for them: kernel_grid features, the full network, the per-ray
sample sequence of advance_pos_nerf + generate_next_nerf_network_inputs, a NeRF-only frame (Identity and ACES tonemap) and the
collision tool's probes (NerfTracer::intersects / collide).  tests/test_oracle_golden.py replays the same inputs through the C
oracle on the CPU, so the oracle's device-only parts stay pinned to the reference where no GPU is present."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import refgpu  # noqa: E402

W, H = 96, 54


def inputs():
    """The seeded inputs, shared with the replaying test."""
    rng = np.random.default_rng(2024)
    pos = rng.uniform(0, 1, size=(512, 3)).astype(np.float32)
    npos = rng.uniform(0.3, 0.7, size=(512, 3)).astype(np.float32)
    d = rng.normal(size=(512, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d01 = ((d + 1) * 0.5).astype(np.float32)
    pts = rng.uniform(-0.35, 0.35, size=(1024, 3)).astype(np.float32)
    org = np.stack([rng.uniform(-0.3, 0.3, 512), np.full(512, 0.45), rng.uniform(-0.3, 0.3, 512)], axis=1).astype(np.float32)
    cam = O.OrbitCamera(W, H)
    cam.orbit(0.35, -0.2, 4.0)
    return {"pos": pos, "npos": npos, "d01": d01, "pts": pts, "org": org, "cam12": cam.matrix(), "probe_dir": np.array([0.3, -0.9, 0.2], np.float32)}


def main(out_path):
    if not refgpu.available():
        raise SystemExit("oracle/_ref/libnmr_refgpu.so not built")
    I = inputs()
    with tempfile.TemporaryDirectory() as d:
        snap = os.path.join(d, "small.msgpack")
        synth.write_snapshot(snap, seed=1337, log2_hashmap_size=15)
        ref = refgpu.ReferenceRenderer(snap)
    out = {"enc": ref.encode(I["pos"]), "net": ref.network(I["npos"], I["d01"]).view(np.uint16)}
    tr = ref.trace(I["cam12"], W, H, 16)
    out.update({"trace_ray": tr["ray"], "trace_pos": tr["pos"], "trace_count": tr["count"]})
    out["img_identity"], _ = ref.render(I["cam12"], W, H, 1, False)
    ref.set_tonemap_curve(1)
    out["img_aces"], _ = ref.render(I["cam12"], W, H, 1, False)
    ref.set_tonemap_curve(0)
    out["probe_points"] = ref.probe(0, I["pts"], I["probe_dir"])
    out["probe_rays"] = ref.probe(1, I["org"], I["probe_dir"])
    ref.close()
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "refgpu_vectors.npz"))
