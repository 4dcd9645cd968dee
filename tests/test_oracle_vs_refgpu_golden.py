"""Pins the C oracle's DEVICE-ONLY parts (hash-grid encoding, fused MLPs + SH, first-hit walk and sample generation,
compositing + tonemap incl. the ACES curve, the collision tool's probes) against what the REFERENCE'S OWN kernels returned on a
B200: tests/golden/refgpu_vectors.npz, written by tests/golden/make_refgpu_vectors.py through oracle/_ref/libnmr_refgpu.so
(ngp::Testbed + tiny-cuda-nn compiled from /root/reference).  CPU only; the inputs are re-created from the same seeds.

The reference binary contracts multiply-adds into FMAs and accumulates its MLPs in fp16, so agreement is asserted within the
tolerances the GPU-side tests use for the product (tests/test_gpu_vs_reference.py), not bit for bit."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "refgpu_vectors.npz"))
_spec = importlib.util.spec_from_file_location("make_refgpu_vectors", os.path.join(HERE, "golden", "make_refgpu_vectors.py"))
_gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_gen)
W, H = _gen.W, _gen.H
PIX_TOL = 2.0 / 255.0


@pytest.fixture(scope="module")
def model(small_snapshot):
    _, snap = small_snapshot
    m = O.Model.from_snapshot(snap)
    I = _gen.inputs()
    P = m.params_struct(W, H, I["cam12"], aabb_min=snap["render_aabb_min"], aabb_max=snap["render_aabb_max"], n_steps_mode=1)
    return m, I, P


def test_encoding_vs_reference_kernel_grid(model):
    m, I, _ = model
    got = m.encode(I["pos"]).view(np.uint16)
    want = G["enc"]
    assert np.mean(got == want) > 0.97
    d = np.abs(got.view(np.float16).astype(np.float32) - want.view(np.float16).astype(np.float32))
    assert float(d.max()) <= 2.0 ** -9


def test_network_vs_reference_mlp(model):
    m, I, _ = model
    got = m.network(I["npos"], I["d01"]).astype(np.float32)
    want = G["net"].view(np.float16).astype(np.float32)[:, :4]
    err = np.abs(got - want)
    tol = 2.0 ** -6 * np.maximum(np.abs(want), 1.0)          # fp16-accumulated reference: ~1e-2 relative
    assert np.mean(err <= tol) > 0.999, float(err.max())
    assert float(err.mean()) < 2e-3


def test_traversal_vs_reference_kernels(model):
    m, I, P = model
    got = m.trace_samples(P, np.arange(W * H, dtype=np.uint32), 16)
    aw, ag = G["trace_ray"][:, 7] > 0, got["ray"][:, 7] > 0
    assert aw.sum() > 200
    assert np.mean(aw == ag) > 0.999
    assert np.allclose(G["trace_ray"][:, :6], got["ray"][:, :6], atol=2e-7, rtol=0)
    live = aw & ag
    assert np.mean(G["trace_count"][live] == got["count"][live]) > 0.98
    both = live & (G["trace_count"] == got["count"])
    valid = np.arange(16)[None, :] < got["count"][:, None]
    dpos = np.abs(G["trace_pos"] - got["pos"]).max(axis=2)[both][valid[both]]
    assert np.mean(dpos <= 1e-6) > 0.98                      # a ray whose FMA-rounded start lands one step off stays one step off
    assert float(dpos.max()) <= 2.0 * 1.7320508 / 1024.0


@pytest.mark.parametrize("curve,key", [(0, "img_identity"), (1, "img_aces")])
def test_pixels_vs_reference_render(model, curve, key):
    m, I, P = model
    frame, _, _, _ = m.render_frame(P, None, None)
    got, _ = O.accumulate_tonemap(frame, None, 0, to_srgb=True, curve=curve)
    want = G[key]
    d = np.abs(got - want)
    mse = float(np.mean((got.astype(np.float64) - want.astype(np.float64)) ** 2))
    assert 10.0 * np.log10(1.0 / max(mse, 1e-12)) >= 45.0
    assert float(np.mean(d.max(axis=2) > PIX_TOL)) <= 0.002, float(d.max())


def test_probes_vs_reference_tracer(model):
    m, I, P = model
    a_got, a_ref = m.probe_points(P, I["pts"], I["probe_dir"]), G["probe_points"]
    assert (a_ref > 0).sum() > 50
    assert np.mean((a_ref > 0) == (a_got > 0)) >= 0.999
    both = (a_ref > 0) & (a_got > 0)
    assert float(np.max(np.abs(a_ref[both] - a_got[both]))) <= 5e-3
    d_got, d_ref = m.probe_rays(P, I["org"], I["probe_dir"]), G["probe_rays"]
    assert (d_ref > 0).sum() > 50 and (d_ref == 0).sum() > 50
    assert np.mean((d_ref > 0) == (d_got > 0)) >= 0.995
    both = (d_ref > 0) & (d_got > 0)
    step = 1.7320508 / 1024 * float(np.linalg.norm(I["probe_dir"]))
    assert np.mean(np.abs(d_ref[both] - d_got[both]) <= 1.01 * step) >= 0.995
    assert float(np.median(np.abs(d_ref[both] - d_got[both]))) <= 1e-6
