"""GPU parity of the collision tool's density probes (NerfTracer::intersects / collide + check_collision,
S/ngp/testbed.cu:721-782, 1814-1935) against the oracle, and first-hit traversal on whole frames (the walk's exact
empty-space jumps must not move a single first sample)."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene(small_snapshot):
    import pynmr
    path, snap = small_snapshot
    r = pynmr.NerfMeshRenderer(480, 270)
    nerf = r.load_nerf(path)
    assert nerf is not None
    return r, nerf, snap


def _params(m, snap, r, W=480, HH=270):
    cam12 = np.ascontiguousarray(r.view_projection_mat.T.reshape(-1))
    return m.params_struct(W, HH, cam12, aabb_min=snap["render_aabb_min"], aabb_max=snap["render_aabb_max"])


def test_first_hit_bit_exact_on_whole_frames(scene):
    from oracle import oracle as O
    r, nerf, snap = scene
    m = O.Model.from_snapshot(snap)
    W, HH = 480, 270
    pixels = np.arange(W * HH, dtype=np.uint32)
    total_live = 0
    for daz, dpol, dz in [(0.0, 0.0, 0.0), (0.9, -0.3, 2.0), (-2.2, 0.45, 5.0), (3.0, 0.1, -6.0)]:
        r.orbit(daz, dpol, dz)
        P = _params(m, snap, r)
        want = m.trace_samples(P, pixels, 2)
        got = H.debug_trace(r, nerf, W, HH, pixels, 2)
        gr, wr = got["ray"].view(np.uint32), want["ray"].view(np.uint32)
        assert np.array_equal(gr[:, 7], wr[:, 7])                     # alive masks
        live = want["ray"][:, 7] > 0
        total_live += int(live.sum())
        assert np.array_equal(gr[live, 6], wr[live, 6])               # t of the first occupied sample, bit for bit
        assert np.array_equal(got["count"], want["count"])
        assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
        assert np.array_equal(got["cell"], want["cell"])
    assert total_live > 20000


def test_probe_points_match_oracle(scene):
    from oracle import oracle as O
    r, nerf, snap = scene
    m = O.Model.from_snapshot(snap)
    P = _params(m, snap, r)
    rng = np.random.default_rng(11)
    pts = rng.uniform(-0.35, 0.35, size=(6000, 3)).astype(np.float32)          # world space: NeRF space minus 0.5
    d = np.array([0.0, -1.0, 0.0], dtype=np.float32)
    got = nerf.probe_points(pts, d)
    want = m.probe_points(P, pts, d)
    inside = want > 0
    assert inside.sum() > 300 and (~inside).sum() > 300
    assert np.array_equal(got > 0, inside)
    assert np.max(np.abs(got - want)) <= 2e-3            # alpha from fp16 densities: fp16 rounding of the network + __expf


def test_probe_rays_match_oracle(scene):
    from oracle import oracle as O
    r, nerf, snap = scene
    m = O.Model.from_snapshot(snap)
    P = _params(m, snap, r)
    rng = np.random.default_rng(12)
    n = 4096
    org = np.stack([rng.uniform(-0.3, 0.3, n), np.full(n, 0.45), rng.uniform(-0.3, 0.3, n)], axis=1).astype(np.float32)   # a plane above the head
    org[:64] = rng.uniform(-0.05, 0.05, size=(64, 3))                                                                       # and some origins inside it
    for d in ([0.0, -1.0, 0.0], [0.3, -0.9, 0.2]):
        d = np.asarray(d, dtype=np.float32)
        got = nerf.probe_rays(org, d)
        want = m.probe_rays(P, org, d)
        hit = want > 0
        assert hit.sum() > 200 and (~hit).sum() > 200
        # the collision sample is the first with alpha > 0: an occupancy walk (bit-exact) plus a density threshold that only a
        # density within fp16 rounding of exp(-inf) could flip - none on this model
        assert np.array_equal(got > 0, hit)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_probe_edge_cases(scene):
    r, nerf, _ = scene
    assert nerf.probe_points(np.zeros((0, 3), np.float32), [0, -1, 0]).shape == (0,)
    assert nerf.probe_rays(np.zeros((0, 3), np.float32), [0, -1, 0]).shape == (0,)
    far = np.array([[5.0, 5.0, 5.0], [-3.0, 0.0, 0.0]], dtype=np.float32)          # outside the render box: no sample at all
    assert np.array_equal(nerf.probe_rays(far, [0, -1, 0]), np.zeros(2, np.float32))
    one = nerf.probe_rays(np.array([[0.0, 0.45, 0.0]], np.float32), [0, -1, 0])    # a single ray (one lane of one tile)
    assert one.shape == (1,) and one[0] > 0
