"""GPU test of nmr_render_update (include/nmr.h; addition): an image kept by the caller and updated in place equals what
Testbed.render() returns for every frame of a camera path - rectangle moving, jumping, leaving the picture and coming back,
background colour changing - while only the rectangle's bytes cross PCIe.  (The reference copies the whole frame per call,
S/python_api.cu:83-111.)"""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

W, HH = 640, 360


@pytest.mark.parametrize("dtype", [np.float32, np.uint8, np.float16])
def test_updated_image_equals_full_render(small_snapshot, glasses_gltf, dtype):
    import pynmr
    import synth
    path, _ = small_snapshot
    r = pynmr.NerfMeshRenderer(W, HH)
    nerf = r.load_nerf(path)
    assert r.load_mesh(glasses_gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is not None
    full_bytes = W * HH * 4 * np.dtype(dtype).itemsize
    img = None
    moved = []
    # a path that drifts, zooms, jumps sideways until the head has left the picture, and comes back
    steps = [(0.02, 0.01, 0.0)] * 4 + [(0.0, 0.0, 1.0)] * 2 + [(0.6, 0.0, 0.0), (0.02, 0.0, 0.0), (-0.6, 0.0, 0.0)] + [(0.0, 0.0, -1.5)] * 2
    for k, (dx, dy, dz) in enumerate(steps):
        r.orbit(dx, dy, dz)
        if k == 6:
            m = r.view_projection_mat; m[:, 3] += 5.0 * m[:, 0]; r.view_projection_mat = m     # far to the side: nothing but background
        if k == 8:
            m = r.view_projection_mat; m[:, 3] -= 5.0 * m[:, 0]; r.view_projection_mat = m
        want = np.asarray(nerf.render(W, HH, 1, linear=False, dtype=dtype)).copy()
        img = nerf.render_update(img, W, HH, linear=False, dtype=dtype)
        assert not img.flags.writeable
        assert np.array_equal(img.view(np.uint8), want.view(np.uint8)), f"step {k}"
        moved.append(nerf.last_update_bytes)
    assert moved[0] == full_bytes                      # a fresh buffer is filled completely
    assert max(moved[1:4]) < 0.6 * full_bytes          # the head fills a good part of this small frame, but never all of it
    assert min(moved[1:]) < 0.25 * full_bytes
    # a new background colour: the whole image changes, and the library notices by itself
    nerf.background_color = [0.2, 0.4, 0.1, 1.0]
    want = np.asarray(nerf.render(W, HH, 1, linear=False, dtype=dtype)).copy()
    img = nerf.render_update(img, W, HH, linear=False, dtype=dtype)
    assert nerf.last_update_bytes == full_bytes
    assert np.array_equal(img.view(np.uint8), want.view(np.uint8))
    # a plain render() into a pooled buffer does not confuse a later update of the kept image
    r.orbit(0.05, 0.0, 0.0)
    want = np.asarray(nerf.render(W, HH, 1, linear=False, dtype=dtype)).copy()
    img = nerf.render_update(img, W, HH, linear=False, dtype=dtype)
    assert nerf.last_update_bytes < full_bytes
    assert np.array_equal(img.view(np.uint8), want.view(np.uint8))
    with pytest.raises(ValueError):
        nerf.render_update(img, W // 2, HH, linear=False, dtype=dtype)


def test_update_at_1080p_moves_a_fraction(small_snapshot, glasses_gltf):
    import pynmr
    import synth
    path, _ = small_snapshot
    r = pynmr.NerfMeshRenderer(1920, 1080)
    nerf = r.load_nerf(path)
    assert r.load_mesh(glasses_gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is not None
    img = None
    for k in range(6):
        r.orbit(0.01, 0.005, 0.0)
        want = np.asarray(nerf.render(1920, 1080, 1, linear=False)).copy()
        img = nerf.render_update(img, 1920, 1080, linear=False)
        assert np.array_equal(img.view(np.uint32), want.view(np.uint32))
    assert nerf.last_update_bytes < 0.2 * 1920 * 1080 * 16          # render.py's framing: the head is a small part of the picture
