"""The C ABI with arguments it must refuse (GPU box): every call returns an error code and a message - no crash, no CUDA error left
behind - and the context renders normally afterwards.  (The reference's bindings throw Python exceptions or print and return None
for the same mistakes, S/python_api.cu:286-331.)"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, HH = 128, 72


def test_bad_arguments_are_refused_and_leave_the_context_usable(small_snapshot, glasses_gltf, tmp_path):
    import pynmr
    import synth
    L = pynmr.lib()
    path, _ = small_snapshot
    r = pynmr.NerfMeshRenderer(W, HH)
    nerf = r.load_nerf(path)
    assert r.load_mesh(glasses_gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is not None
    r.orbit(0.3, -0.1, 4.0)
    good = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
    h = r._h
    out = np.zeros((HH, W, 4), np.float32)
    p = out.ctypes.data_as(C.c_void_p)
    cam = np.zeros(12, np.float32)
    f3 = (C.c_float * 3)(0, 0, 0)
    pts = np.zeros((4, 3), np.float32)
    res = np.zeros(4, np.float32)
    calls = {
        "render: zero width": lambda: L.nmr_render(h, nerf._id, 0, HH, 1, 0, p),
        "render: negative height": lambda: L.nmr_render(h, nerf._id, W, -5, 1, 0, p),
        "render: zero samples": lambda: L.nmr_render(h, nerf._id, W, HH, 0, 0, p),
        "render: null image": lambda: L.nmr_render(h, nerf._id, W, HH, 1, 0, None),
        "render: unknown nerf": lambda: L.nmr_render(h, 99, W, HH, 1, 0, p),
        "render: negative nerf": lambda: L.nmr_render(h, -1, W, HH, 1, 0, p),
        "render_format: bad format": lambda: L.nmr_render_format(h, nerf._id, W, HH, 1, 0, 7, p),
        "render_update: null image": lambda: L.nmr_render_update(h, nerf._id, W, HH, 0, 0, None, 0, None),
        "render_views: no views": lambda: L.nmr_render_views_format(h, nerf._id, 0, cam.ctypes.data_as(C.c_void_p), W, HH, 0, 0, p),
        "render_views: null cameras": lambda: L.nmr_render_views_format(h, nerf._id, 2, None, W, HH, 0, 0, p),
        "set_shard: rank >= world": lambda: L.nmr_set_shard(h, 3, 3, 8),
        "set_shard: zero world": lambda: L.nmr_set_shard(h, 0, 0, 8),
        "set_shard: zero band": lambda: L.nmr_set_shard(h, 0, 2, 0),
        "lens model: unknown": lambda: L.nmr_set_lens_model(h, 5, 0.01),
        "lens model: negative thickness": lambda: L.nmr_set_lens_model(h, 1, -1.0),
        "surface rule: unknown": lambda: L.nmr_set_surface_insertion(h, 9),
        "tonemap curve: unknown": lambda: L.nmr_set_tonemap_curve(h, nerf._id, 11),
        "render_aabb: unknown nerf": lambda: L.nmr_set_render_aabb(h, 42, f3, f3),
        "probe_points: unknown nerf": lambda: L.nmr_probe_points(h, 42, 4, pts.ctypes.data_as(C.c_void_p), f3, res.ctypes.data_as(C.c_void_p)),
        "load_nerf: no such file": lambda: L.nmr_load_nerf(h, str(tmp_path / "nope.msgpack").encode(), C.byref(C.c_int())),
        "load_mesh: no such file": lambda: L.nmr_load_mesh(h, str(tmp_path / "nope.gltf").encode(), f3, f3, (C.c_float * 4)(1, 0, 0, 0), C.byref(C.c_int())),
        "load_density_grid: no such file": lambda: L.nmr_load_density_grid(h, nerf._id, str(tmp_path / "nope.bin").encode(), None),
        "read_combined: nothing merged": lambda: L.nmr_read_combined(h, p, None),
        "camera: NaN": lambda: L.nmr_set_camera(h, np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, float("nan"), 2], np.float32).ctypes.data_as(C.POINTER(C.c_float))),
        "camera: inf": lambda: L.nmr_set_camera(h, np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, float("inf"), 0, 2], np.float32).ctypes.data_as(C.POINTER(C.c_float))),
        "orbit: NaN": lambda: L.nmr_orbit(h, float("nan"), 0.0, 0.0),
        "render_aabb: NaN": lambda: L.nmr_set_render_aabb(h, nerf._id, (C.c_float * 3)(float("nan"), 0, 0), f3),
        "background: inf": lambda: L.nmr_set_background(h, nerf._id, (C.c_float * 4)(float("inf"), 0, 0, 1)),
        "min transmittance: 2": lambda: L.nmr_set_min_transmittance(h, nerf._id, 2.0),
        "model transform: NaN": lambda: L.nmr_set_model_transform(h, nerf._id, (C.c_float * 3)(0, float("nan"), 0), None),
        "mesh transform: inf scale": lambda: L.nmr_set_mesh_transform(h, 0, None, (C.c_float * 3)(1, float("inf"), 1), None),
        "render_views: NaN camera": lambda: L.nmr_render_views_format(h, nerf._id, 1, np.full(12, np.nan, np.float32).ctypes.data_as(C.c_void_p), W, HH, 0, 0, p),
        "trajectory pose: inf": lambda: L.nmr_trajectory_pose(h, float("inf"), 1.1, 0.1, None),
    }
    for name, call in calls.items():
        rc = call()
        assert rc != 0, name
        assert L.nmr_last_error(h), name
        # the context still renders, and renders the same picture
        again = np.asarray(nerf.render(W, HH, 1, linear=False))
        assert np.array_equal(again.view(np.uint32), good.view(np.uint32)), name
    # a garbage file under a real name
    bad = tmp_path / "garbage.msgpack"; bad.write_bytes(b"\x93\x01\x02")
    assert r.load_nerf(str(bad)) is None
    badg = tmp_path / "garbage.gltf"; badg.write_text("{\"asset\": 1")
    assert r.load_mesh(str(badg)) is None
    assert np.array_equal(np.asarray(nerf.render(W, HH, 1, linear=False)).view(np.uint32), good.view(np.uint32))


def test_absurd_but_finite_cameras_return(small_snapshot, glasses_gltf):
    """A camera in the wrong units (1e4 .. 3e38 away), one looking exactly along an axis, one whose direction block is zero: the frame
    comes back, finite, at once.  (At |t| beyond ~4e4 a step of the reference's walk no longer changes t and its loops would never
    end; such rays are declared dead at set-up, device_common.cuh: init_ray.)"""
    import time
    import pynmr
    import synth
    path, _ = small_snapshot
    r = pynmr.NerfMeshRenderer(W, HH)
    r.load_nerf(path)
    assert r.load_mesh(glasses_gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is not None
    base = r.view_projection_mat.copy()
    cams = []
    for dist in (1e2, 1e4, 4e4, 1e5, 1e9, 1e20, 3e38):
        m = base.copy(); m[:, 3] = -m[:, 2] * dist
        cams.append(m)
    d0 = float(np.linalg.norm(base[:, 3]))
    for dist in (3e4, 3.5e4, 4e4, 1e5, 1e6):          # telephoto views of the same framing: rays that DO aim at the head from that far
        m = base.copy(); k = dist / d0
        m[:, 3] = base[:, 3] * k; m[:, 0] = base[:, 0] / k; m[:, 1] = base[:, 1] / k
        cams.append(m)
    cams.append(np.array([[1, 0, 0, 0.5], [0, 1, 0, 0.5], [0, 0, -1, 3.0]], np.float32))
    cams.append(np.array([[0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 2.0]], np.float32))
    for m in cams:
        t0 = time.time()
        r.view_projection_mat = m
        assert r.frame()
        img = np.asarray(r.read_frame())
        assert np.isfinite(img).all() and time.time() - t0 < 5.0
