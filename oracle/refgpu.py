"""ctypes front-end of oracle/_ref/libnmr_refgpu.so: the REFERENCE's own NeRF renderer (ngp::Testbed + tiny-cuda-nn),
compiled for sm_100 from /root/reference by oracle/Makefile.refgpu.

TEST INFRASTRUCTURE ONLY.  Used by the -m gpu tests (pinning libnmr and the C oracle against the reference's kernels
on the same B200) and by bench.py's informational "reference kernels on this GPU" figure.  The product never imports it.
The library only exists where it was built (this container) and on GPU boxes that received the built file via gpurun.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libnmr_refgpu.so")
MAKEFILE = os.path.join(HERE, "Makefile.refgpu")


def available() -> bool:
    return os.path.exists(LIB)


def build(jobs: int = 8) -> str | None:
    """make -f oracle/Makefile.refgpu (about 25 minutes from scratch; incremental afterwards).  No-op without the
    reference tree or nvcc."""
    if not os.path.isdir("/root/reference/nerf_mesh_renderer") or shutil.which("nvcc") is None:
        return LIB if available() else None
    subprocess.check_call(["make", "-f", MAKEFILE, f"-j{jobs}"], stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        vp = C.c_void_p
        L.refgpu_create.restype = vp
        L.refgpu_last_error.restype = C.c_char_p
        L.refgpu_last_error.argtypes = [vp]
        L.refgpu_destroy.argtypes = [vp]
        L.refgpu_load_snapshot.argtypes = [vp, C.c_char_p]
        L.refgpu_set_render_aabb.argtypes = [vp, vp, vp]
        L.refgpu_get_render_aabb.argtypes = [vp, vp, vp]
        L.refgpu_set_background.argtypes = [vp, vp]
        L.refgpu_get_bitfield.argtypes = [vp, vp]
        L.refgpu_set_bitfield.argtypes = [vp, vp]
        L.refgpu_bitfield_bytes.restype = C.c_int64
        L.refgpu_bitfield_bytes.argtypes = [vp]
        L.refgpu_network.argtypes = [vp, vp, vp, C.c_int64, vp]
        L.refgpu_encode.argtypes = [vp, vp, C.c_int64, vp]
        L.refgpu_trace.argtypes = [vp, vp, C.c_int, C.c_int, C.c_uint32, vp, vp, C.c_uint32, vp, vp, vp, vp, vp]
        if hasattr(L, "refgpu_set_tonemap_curve"):
            L.refgpu_set_tonemap_curve.argtypes = [vp, C.c_int]
        if hasattr(L, "refgpu_probe"):
            L.refgpu_probe.argtypes = [vp, C.c_int, vp, vp, C.c_int64, vp]
        L.refgpu_render.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp]
        if hasattr(L, "refgpu_get_buffers"):
            L.refgpu_get_buffers.argtypes = [vp, C.c_int, C.c_int, vp, vp]
            L.refgpu_set_model_transform.argtypes = [vp, vp, vp]
        if hasattr(L, "refgpu_set_crop_box"):
            L.refgpu_set_crop_box.argtypes = [vp, vp, C.c_int, vp, vp, vp]
            L.refgpu_crop_box.argtypes = [vp, C.c_int, vp, vp]
            L.refgpu_camera_ops.argtypes = [vp, vp, C.c_float, vp, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class ReferenceRenderer:
    """The reference's Testbed behind NerfMeshRenderer::loadNerf, headless."""

    def __init__(self, snapshot_path: str):
        self._L = lib()
        self._h = self._L.refgpu_create()
        if not self._h:
            raise RuntimeError("refgpu_create failed: " + self._L.refgpu_last_error(None).decode())
        self._ck(self._L.refgpu_load_snapshot(self._h, snapshot_path.encode()))

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(self._L.refgpu_last_error(self._h).decode())

    def close(self):
        if self._h:
            self._L.refgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_render_aabb(self, mn, mx):
        a = np.ascontiguousarray(mn, dtype=np.float32); b = np.ascontiguousarray(mx, dtype=np.float32)
        self._ck(self._L.refgpu_set_render_aabb(self._h, _p(a), _p(b)))

    def render_aabb(self):
        a = np.zeros(3, np.float32); b = np.zeros(3, np.float32)
        self._ck(self._L.refgpu_get_render_aabb(self._h, _p(a), _p(b)))
        return a, b

    def bitfield(self) -> np.ndarray:
        out = np.zeros(int(self._L.refgpu_bitfield_bytes(self._h)), dtype=np.uint8)
        self._ck(self._L.refgpu_get_bitfield(self._h, _p(out)))
        return out

    def set_bitfield(self, bits: np.ndarray):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        assert bits.size == int(self._L.refgpu_bitfield_bytes(self._h))
        self._ck(self._L.refgpu_set_bitfield(self._h, _p(bits)))

    def encode(self, pos: np.ndarray) -> np.ndarray:
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        out = np.zeros((pos.shape[0], 32), dtype=np.uint16)
        self._ck(self._L.refgpu_encode(self._h, _p(pos), pos.shape[0], _p(out)))
        return out

    def network(self, pos: np.ndarray, dir01: np.ndarray) -> np.ndarray:
        """-> float16 [n, 16]: r, g, b raw, density raw, 12 padding channels."""
        pos = np.ascontiguousarray(pos, dtype=np.float32); d = np.ascontiguousarray(dir01, dtype=np.float32)
        out = np.zeros((pos.shape[0], 16), dtype=np.uint16)
        self._ck(self._L.refgpu_network(self._h, _p(pos), _p(d), pos.shape[0], _p(out)))
        return out.view(np.float16)

    def set_tonemap_curve(self, curve: int):
        self._ck(self._L.refgpu_set_tonemap_curve(self._h, int(curve)))

    def probe(self, mode: int, points_world: np.ndarray, direction) -> np.ndarray:
        """The reference's NerfTracer::intersects (mode 0) / collide (mode 1) over world-space points, as NerfMeshRenderer::collide calls them."""
        pts = np.ascontiguousarray(points_world, dtype=np.float32).reshape(-1, 3); d = np.ascontiguousarray(direction, dtype=np.float32)
        out = np.zeros(len(pts), np.float32)
        self._ck(self._L.refgpu_probe(self._h, int(mode), _p(pts), _p(d), len(pts), _p(out)))
        return out

    def trace(self, cam12, W, H, max_samples, spp_index=0, surf=None, ts=None):
        cam12 = np.ascontiguousarray(cam12, dtype=np.float32)
        n = W * H
        ray = np.zeros((n, 10), np.float32); pos = np.zeros((n, max_samples, 3), np.float32)
        dt = np.zeros((n, max_samples), np.float32); ta = np.zeros((n, max_samples), np.float32); cnt = np.zeros(n, np.uint32)
        s = None if surf is None else np.ascontiguousarray(surf, dtype=np.float32)
        t = None if ts is None else np.ascontiguousarray(ts, dtype=np.float32)
        self._ck(self._L.refgpu_trace(self._h, _p(cam12), W, H, spp_index, _p(s), _p(t), max_samples, _p(ray), _p(pos), _p(dt), _p(ta), _p(cnt)))
        return {"ray": ray, "pos": pos, "dt": dt, "t_after": ta, "count": cnt}

    def render(self, cam12, W, H, spp=1, linear=False, surf=None, ts=None, repeat=1):
        """-> (float32 [H, W, 4] bottom-up like Testbed.render, best device ms of Testbed::render_frame)."""
        cam12 = np.ascontiguousarray(cam12, dtype=np.float32)
        out = np.zeros((H, W, 4), np.float32)
        s = None if surf is None else np.ascontiguousarray(surf, dtype=np.float32)
        t = None if ts is None else np.ascontiguousarray(ts, dtype=np.float32)
        ms = C.c_float(0)
        self._ck(self._L.refgpu_render(self._h, _p(cam12), W, H, spp, 1 if linear else 0, _p(s), _p(t), _p(out), repeat, C.byref(ms)))
        return out, float(ms.value)

    def set_model_transform(self, translation, rotation_pi):
        """Testbed::m_model_translation / m_model_rotation (angles in units of pi about X, Y, Z)."""
        t = np.ascontiguousarray(translation, dtype=np.float32).reshape(3); r = np.ascontiguousarray(rotation_pi, dtype=np.float32).reshape(3)
        self._ck(self._L.refgpu_set_model_transform(self._h, _p(t), _p(r)))

    def set_crop_box(self, matrix34, nerf_space=True):
        """Testbed::set_crop_box -> (render_aabb_to_local 3x3, render_aabb min, max) as the reference leaves them."""
        m = np.ascontiguousarray(matrix34, dtype=np.float32).reshape(3, 4)
        r2l = np.zeros((3, 3), np.float32); mn = np.zeros(3, np.float32); mx = np.zeros(3, np.float32)
        self._ck(self._L.refgpu_set_crop_box(self._h, _p(m), int(bool(nerf_space)), _p(r2l), _p(mn), _p(mx)))
        return r2l, mn, mx

    def crop_box(self, nerf_space=True):
        """Testbed::crop_box and crop_box_corners -> (3x4 matrix, corners [8, 3])."""
        m = np.zeros((3, 4), np.float32); c = np.zeros((8, 3), np.float32)
        self._ck(self._L.refgpu_crop_box(self._h, int(bool(nerf_space)), _p(m), _p(c)))
        return m, c

    def camera_ops(self, cam34, new_scale, look_at, view_dir, up):
        """set_scale, set_look_at, set_view_dir on the given camera -> (camera afterwards 3x4, look_at(), scale() before)."""
        cam = np.ascontiguousarray(cam34, dtype=np.float32).reshape(3, 4)
        la = np.ascontiguousarray(look_at, dtype=np.float32).reshape(3); vd = np.ascontiguousarray(view_dir, dtype=np.float32).reshape(3)
        u = np.ascontiguousarray(up, dtype=np.float32).reshape(3)
        out = np.zeros((3, 4), np.float32); lo = np.zeros(3, np.float32); sb = C.c_float(0)
        self._ck(self._L.refgpu_camera_ops(self._h, _p(cam), float(new_scale), _p(la), _p(vd), _p(u), _p(out), _p(lo), C.byref(sb)))
        return out, lo, float(sb.value)

    def buffers(self, W, H):
        """Linear frame buffer [H, W, 4] and depth buffer [H, W] of the render surface after the last render() - what
        NerfMeshRenderer::render_frame copies / z-merges per NeRF (S/nerf_mesh_renderer.cu:582-597)."""
        frame = np.zeros((H, W, 4), np.float32); depth = np.zeros((H, W), np.float32)
        self._ck(self._L.refgpu_get_buffers(self._h, W, H, _p(frame), _p(depth)))
        return frame, depth
