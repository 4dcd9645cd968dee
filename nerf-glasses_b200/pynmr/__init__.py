"""pynmr - drop-in Python surface of the reference's pybind11 module of the same name
(nerf_mesh_renderer/src/python_api.cu:156-623), implemented as a thin ctypes shim over the C ABI of libnmr.so
(include/nmr.h).  `volume/render.py` of the reference runs unchanged against this module:

    import pynmr as nmr
    renderer = nmr.NerfMeshRenderer(1280, 720)
    renderer.envmap("sunflowers_puresky_1k.png")           # accepted and ignored (missing in the reference module)
    nerf = renderer.load_nerf("nerf.msgpack")
    nerf.render_aabb.min = np.array([-0.2, 0.15, -0.2])
    mesh = renderer.load_mesh("glasses.gltf", t=..., s=..., r=[w, x, y, z])
    while renderer.frame():
        renderer.orbit(daz, dpolar, 0)
        im = nerf.render(1280, 720, linear=False)          # float32[H, W, 4], row 0 = bottom

There is no CPU fallback: importing works anywhere, but constructing a NerfMeshRenderer needs libnmr.so and an
sm_100 GPU and raises RuntimeError otherwise.
"""
from __future__ import annotations

import ctypes as C
import enum
import os
import sys
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NMR_LIB") or os.path.join(os.path.dirname(_HERE), "libnmr.so")   # NMR_LIB: tuning builds

NMR_OK = 0


class Stats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("rays_alive", C.c_uint64), ("samples", C.c_uint64), ("mesh_rays", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("gpu_ms", C.c_float), ("march_ms", C.c_float), ("batches", C.c_uint64), ("batch_passes", C.c_uint64)]


_lib = None


class NerfInfo(C.Structure):
    _fields_ = [("training_step", C.c_int64), ("loss", C.c_float), ("aabb_scale", C.c_int), ("max_cascade", C.c_int),
                ("cone_angle_constant", C.c_float), ("rgb_activation", C.c_int), ("density_activation", C.c_int),
                ("render_aabb_to_local", C.c_float * 9), ("n_levels", C.c_int), ("n_features_per_level", C.c_int),
                ("log2_hashmap_size", C.c_int), ("base_resolution", C.c_int), ("per_level_scale", C.c_float), ("n_params", C.c_uint64)]


class NerfDataset(C.Structure):
    _fields_ = [("scale", C.c_float), ("offset", C.c_float * 3), ("up", C.c_float * 3), ("from_mitsuba", C.c_int),
                ("bounding_radius", C.c_float), ("raw_aabb_min", C.c_float * 3), ("raw_aabb_max", C.c_float * 3)]


def lib():
    """Loads libnmr.so (raises RuntimeError with the build hint when it is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found - build it with `python nerf-glasses_b200/build.py` (nvcc, sm_100a)")
    L = C.CDLL(LIB_PATH)
    vp, ip, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)
    sig = {
        "nmr_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
        "nmr_destroy": (None, [vp]),
        "nmr_last_error": (C.c_char_p, [vp]),
        "nmr_load_nerf": (C.c_int, [vp, C.c_char_p, ip]),
        "nmr_reload_nerf": (C.c_int, [vp, C.c_int, C.c_char_p]),
        "nmr_load_mesh": (C.c_int, [vp, C.c_char_p, fp, fp, fp, ip]),
        "nmr_set_mesh_transform": (C.c_int, [vp, C.c_int, fp, fp, fp]),
        "nmr_get_mesh_transform": (C.c_int, [vp, C.c_int, fp, fp, fp]),
        "nmr_set_envmap": (C.c_int, [vp, C.c_char_p]),
        "nmr_remove_floaties": (C.c_int, [vp, ip, C.POINTER(C.c_int64)]),
        "nmr_get_render_aabb": (C.c_int, [vp, C.c_int, fp, fp]),
        "nmr_set_render_aabb": (C.c_int, [vp, C.c_int, fp, fp]),
        "nmr_get_aabb": (C.c_int, [vp, C.c_int, fp, fp]),
        "nmr_get_background": (C.c_int, [vp, C.c_int, fp]),
        "nmr_set_background": (C.c_int, [vp, C.c_int, fp]),
        "nmr_set_min_transmittance": (C.c_int, [vp, C.c_int, C.c_float]),
        "nmr_orbit": (C.c_int, [vp, C.c_float, C.c_float, C.c_float]),
        "nmr_get_camera": (C.c_int, [vp, fp]),
        "nmr_set_camera": (C.c_int, [vp, fp]),
        "nmr_frame": (C.c_int, [vp, ip]),
        "nmr_frame_async": (C.c_int, [vp]),
        "nmr_read_frame": (C.c_int, [vp, vp]),
        "nmr_render": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
        "nmr_render_views": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp]),
        "nmr_render_format": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
        "nmr_render_views_format": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
        "nmr_debug_parse_gltf": (C.c_int, [C.c_char_p, C.POINTER(C.c_int64), C.c_char_p, C.c_size_t, vp, C.c_size_t]),
        "nmr_get_nerf_dataset": (C.c_int, [vp, C.c_int, C.POINTER(NerfDataset)]),
        "nmr_set_render_aabb_to_local": (C.c_int, [vp, C.c_int, fp]),
        "nmr_mikk_tangents": (C.c_int, [vp, vp, vp, C.c_int64, vp, C.c_int64, vp]),
        "nmr_set_shard": (C.c_int, [vp, C.c_int, C.c_int, C.c_int]),
        "nmr_set_surface_insertion": (C.c_int, [vp, C.c_int]),
        "nmr_set_overlap": (C.c_int, [vp, C.c_int]),
        "nmr_set_model_transform": (C.c_int, [vp, C.c_int, fp, fp]),
        "nmr_get_model_transform": (C.c_int, [vp, C.c_int, fp, fp, fp]),
        "nmr_dump_density_grid": (C.c_int, [vp, C.c_int, C.c_char_p, vp]),
        "nmr_load_density_grid": (C.c_int, [vp, C.c_int, C.c_char_p, vp]),
        "nmr_read_combined": (C.c_int, [vp, vp, vp]),
        "nmr_set_lens": (C.c_int, [vp, C.c_int, C.c_float, C.c_float, fp]),
        "nmr_set_lens_model": (C.c_int, [vp, C.c_int, C.c_float]),
        "nmr_get_nerf_info": (C.c_int, [vp, C.c_int, C.POINTER(NerfInfo)]),
        "nmr_get_stream": (C.c_int, [vp, C.POINTER(vp)]),
        "nmr_set_tonemap_curve": (C.c_int, [vp, C.c_int, C.c_int]),
        "nmr_get_tonemap_curve": (C.c_int, [vp, C.c_int, ip]),
        "nmr_probe_points": (C.c_int, [vp, C.c_int, C.c_int64, vp, fp, vp]),
        "nmr_probe_rays": (C.c_int, [vp, C.c_int, C.c_int64, vp, fp, vp]),
        "nmr_gather_create": (C.c_int, [vp, vp, C.POINTER(vp)]),
        "nmr_gather_attach": (C.c_int, [vp, vp]),
        "nmr_gather_detach": (C.c_int, [vp]),
        "nmr_debug_lens": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp]),
        "nmr_get_device_image": (C.c_int, [vp, C.POINTER(vp), ip, ip]),
        "nmr_copy_device_image": (C.c_int, [vp, vp, C.c_size_t]),
        "nmr_flush_l2": (C.c_int, [vp]),
        "nmr_measure_l2": (C.c_int, [vp, C.c_size_t, C.c_int, fp]),
        "nmr_get_stats": (C.c_int, [vp, C.POINTER(Stats)]),
        "nmr_synchronize": (C.c_int, [vp]),
        "nmr_trajectory_pose": (C.c_int, [vp, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float)]),
        "nmr_render_update": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.POINTER(C.c_size_t)]),
        "nmr_host_alloc": (vp, [C.c_size_t]),
        "nmr_host_free": (None, [vp]),
        "nmr_get_density_bitfield": (C.c_int, [vp, C.c_int, vp]),
        "nmr_set_density_bitfield": (C.c_int, [vp, C.c_int, vp]),
        "nmr_debug_encode": (C.c_int, [vp, C.c_int, vp, C.c_int64, vp]),
        "nmr_debug_network": (C.c_int, [vp, C.c_int, vp, vp, C.c_int64, vp]),
        "nmr_debug_trace": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int64, C.c_uint32, vp, vp, vp, vp, vp, vp]),
        "nmr_debug_mesh": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]),
        "nmr_debug_last_frame": (C.c_int, [vp, vp, vp, vp]),
        "nmr_debug_set_flags": (C.c_int, [vp, C.c_uint32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "nmr_create", "nmr_destroy", "nmr_last_error", "nmr_load_nerf", "nmr_load_mesh", "nmr_set_mesh_transform",
    "nmr_get_mesh_transform", "nmr_set_envmap", "nmr_remove_floaties", "nmr_get_render_aabb", "nmr_set_render_aabb",
    "nmr_get_aabb", "nmr_get_background", "nmr_set_background", "nmr_set_min_transmittance", "nmr_orbit", "nmr_get_camera",
    "nmr_set_camera", "nmr_frame", "nmr_frame_async", "nmr_read_frame", "nmr_render", "nmr_render_views", "nmr_set_shard", "nmr_set_surface_insertion", "nmr_set_lens", "nmr_debug_lens", "nmr_get_nerf_info", "nmr_gather_create", "nmr_gather_attach", "nmr_gather_detach", "nmr_get_stream", "nmr_probe_points", "nmr_probe_rays", "nmr_set_tonemap_curve", "nmr_get_tonemap_curve",
    "nmr_get_device_image", "nmr_copy_device_image", "nmr_flush_l2", "nmr_get_stats", "nmr_synchronize", "nmr_host_alloc", "nmr_host_free", "nmr_render_update", "nmr_trajectory_pose", "nmr_get_density_bitfield",
    "nmr_set_density_bitfield", "nmr_debug_encode", "nmr_debug_network", "nmr_debug_trace", "nmr_debug_mesh",
    "nmr_debug_last_frame", "nmr_debug_set_flags", "nmr_render_format", "nmr_render_views_format", "nmr_debug_parse_gltf", "nmr_measure_l2", "nmr_set_overlap", "nmr_set_model_transform", "nmr_get_model_transform",
    "nmr_dump_density_grid", "nmr_load_density_grid", "nmr_read_combined", "nmr_set_lens_model", "nmr_mikk_tangents", "nmr_get_nerf_dataset", "nmr_set_render_aabb_to_local", "nmr_reload_nerf",
]


PIXEL_F32, PIXEL_F16, PIXEL_U8 = 0, 1, 2


def _pixel_format(dtype) -> int:
    """numpy dtype of an output image -> nmr_pixel_format (include/nmr.h)."""
    dt = np.dtype(dtype)
    if dt == np.float32:
        return PIXEL_F32
    if dt == np.float16:
        return PIXEL_F16
    if dt == np.uint8:
        return PIXEL_U8
    raise ValueError(f"unsupported image dtype {dt}: float32, float16 or uint8")


def parse_gltf(path: str, tangents: bool = False) -> dict:
    """Host-only check of a .gltf / .glb through the loader behind load_mesh (no GPU needed).  Raises RuntimeError with the loader's
    message on malformed input."""
    counts = (C.c_int64 * 5)()
    err = C.create_string_buffer(512)
    rc = lib().nmr_debug_parse_gltf(os.fsencode(path), counts, err, len(err), None, 0)
    if rc != NMR_OK:
        raise RuntimeError(f"libnmr error {rc}: {err.value.decode(errors='replace')}")
    out = {"vertices": counts[0], "triangles": counts[1], "lens_triangles": counts[2], "texture": (counts[3], counts[4]), "warning": err.value.decode(errors="replace")}
    if tangents:
        t = np.zeros((counts[0], 4), dtype=np.float32)
        lib().nmr_debug_parse_gltf(os.fsencode(path), counts, err, len(err), _ptr(t), t.size)
        out["tangents"] = t
    return out


def mikk_tangents(positions, normals, texcoords, indices) -> np.ndarray:
    """Host-only: per-vertex tangents (xyz + handedness) of an indexed triangle list by the reference's tangent generator for
    primitives without a TANGENT attribute (S/gltf_scene.cpp:150-155 -> mikktspace.c); what load_mesh computes for such a file."""
    p = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
    n = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(texcoords, dtype=np.float32).reshape(-1, 2)
    i = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
    if not (len(p) == len(n) == len(t)):
        raise ValueError("positions, normals and texcoords need one entry per vertex")
    out = np.zeros((len(p), 4), dtype=np.float32)
    rc = lib().nmr_mikk_tangents(_ptr(p), _ptr(n), _ptr(t), len(p), _ptr(i), i.size, _ptr(out))
    if rc != NMR_OK:
        raise RuntimeError(f"libnmr error {rc}")
    return out


def _f3(v):
    a = np.asarray(v, dtype=np.float32).reshape(-1)
    return (C.c_float * len(a))(*[float(x) for x in a])


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


_envmap_warned = False
_pinned_pool: dict = {}     # size in bytes -> [device-registered host pointers ready for reuse]; insertion order = age of a size
_PINNED_POOL_CAP = 512 << 20   # bytes kept for reuse; beyond it the sizes not used for longest go back to the driver


def _pinned_release(nbytes: int, p: int):
    try:
        blocks = _pinned_pool.pop(nbytes, [])
        blocks.append(p)
        _pinned_pool[nbytes] = blocks                      # most recently used size last
        total = sum(n * len(b) for n, b in _pinned_pool.items())
        while total > _PINNED_POOL_CAP and _pinned_pool:
            n0 = next(iter(_pinned_pool))                  # the size that has not been released to for longest
            b0 = _pinned_pool[n0]
            if b0:
                lib().nmr_host_free(b0.pop())
                total -= n0
            if not b0:
                del _pinned_pool[n0]
    except Exception:       # (interpreter shutdown)
        pass


def _pinned_array(shape, dtype=np.float32) -> np.ndarray:
    """numpy array over page-locked memory from nmr_host_alloc (fast, asynchronous device->host copies).
    The ctypes buffer at the root of the array's .base chain owns the allocation; when the last view dies the block goes
    back to a per-size pool (page-locking tens of MB costs milliseconds, so a render loop must not allocate per frame)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    free = _pinned_pool.get(n)
    p = free.pop() if free else lib().nmr_host_alloc(n)
    if free is not None and not free:
        del _pinned_pool[n]                                # (no empty lists: the pool's first key is always a size with a block to give back)
    if not p:
        raise MemoryError("nmr_host_alloc failed")
    buf = (C.c_char * n).from_address(p)
    weakref.finalize(buf, _pinned_release, n, p)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def format_transform(mat) -> str:
    """A 3 x 4 float matrix as the reference's trajectory tool writes it to `transform_N` (S/nerf_mesh_renderer.cu:639-642:
    `viewProjectionMat.format(Eigen::IOFormat(Eigen::FullPrecision, 0, ", ", ",\\n", "[", "]", "[", "]"))`): six significant digits
    (%g), every entry right-aligned to the widest one, rows "[...]" joined by ",\\n" with one space of indent after the first, the
    whole in "[...]".  Pinned on strings produced by the reference's own Eigen (tests/golden/ref_trajectory.npz)."""
    m = np.asarray(mat, dtype=np.float32).reshape(3, 4)
    cells = [["%.6g" % float(v) for v in row] for row in m]
    width = max(len(c) for row in cells for c in row)
    rows = ["[" + ", ".join(c.rjust(width) for c in row) + "]" for row in cells]
    return "[" + ",\n ".join(rows) + "]"


def _write_png_rgb8(path: str, rgb: np.ndarray):
    """uint8 [H, W, 3], top row first -> PNG (stdlib zlib only)."""
    import struct, zlib
    h, w, _ = rgb.shape
    raw = b"".join(b"\x00" + rgb[y].tobytes() for y in range(h))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def free_pinned_pool():
    """Returns every pooled page-locked block to the driver."""
    for blocks in _pinned_pool.values():
        while blocks:
            lib().nmr_host_free(blocks.pop())


class Vec3:
    """glm::vec3 stand-in (S/python_api.cu:263-270)."""

    def __init__(self, x=0.0, y=0.0, z=0.0, _on_change=None):
        if not np.isscalar(x):
            x, y, z = [float(v) for v in np.asarray(x).reshape(3)]
        self._v = [float(x), float(y), float(z)]
        self._cb = _on_change

    def _set(self, i, val):
        self._v[i] = float(val)
        if self._cb:
            self._cb(self)

    x = property(lambda s: s._v[0], lambda s, v: s._set(0, v))
    y = property(lambda s: s._v[1], lambda s, v: s._set(1, v))
    z = property(lambda s: s._v[2], lambda s, v: s._set(2, v))

    def __iter__(self):
        return iter(self._v)

    def __array__(self, dtype=None, copy=None):
        return np.array(self._v, dtype=dtype or np.float32)

    def __repr__(self):
        return f"Vec3({self._v[0]}, {self._v[1]}, {self._v[2]})"


class BoundingBox:
    """ngp::BoundingBox (S/python_api.cu:242-261).  When obtained from Testbed.render_aabb it is a live view:
    assigning .min / .max updates the renderer's crop box, like the reference's def_readwrite on the C++ object."""

    def __init__(self, a=None, b=None, _owner=None):
        self._owner = _owner
        self._min = np.full(3, np.inf, dtype=np.float32) if a is None else np.asarray(a, dtype=np.float32).copy()
        self._max = np.full(3, -np.inf, dtype=np.float32) if b is None else np.asarray(b, dtype=np.float32).copy()

    def _pull(self):
        if self._owner is not None:
            self._min, self._max = self._owner._get_render_aabb()

    def _push(self):
        if self._owner is not None:
            self._owner._set_render_aabb(self._min, self._max)

    @property
    def min(self):
        self._pull()
        return self._min.copy()

    @min.setter
    def min(self, v):
        self._pull()
        self._min = np.asarray(v, dtype=np.float32).reshape(3).copy()
        self._push()

    @property
    def max(self):
        self._pull()
        return self._max.copy()

    @max.setter
    def max(self, v):
        self._pull()
        self._max = np.asarray(v, dtype=np.float32).reshape(3).copy()
        self._push()

    def center(self):
        return 0.5 * (self.max + self.min)

    def diag(self):
        return self.max - self.min

    def contains(self, p):
        p = np.asarray(p, dtype=np.float32)
        return bool(np.all(p >= self.min) and np.all(p <= self.max))

    def relative_pos(self, p):
        return (np.asarray(p, dtype=np.float32) - self.min) / self.diag()

    def inflate(self, amount):
        self._pull()
        self._min = self._min - np.float32(amount)
        self._max = self._max + np.float32(amount)
        self._push()

    def enlarge(self, other):
        self._pull()
        if isinstance(other, BoundingBox):
            self._min = np.minimum(self._min, other.min); self._max = np.maximum(self._max, other.max)
        else:
            p = np.asarray(other, dtype=np.float32)
            self._min = np.minimum(self._min, p); self._max = np.maximum(self._max, p)
        self._push()

    def ray_intersect(self, pos, dir):
        """BoundingBox::ray_intersect (S/ngp/bounding_box.cuh:106-147): the slab test in fp32, axis by axis -> (tmin, tmax), both FLT_MAX
        when the ray misses the box."""
        f = np.float32
        fmax = np.finfo(np.float32).max
        p = np.asarray(pos, f).reshape(3); d = np.asarray(dir, f).reshape(3)
        mn, mx = self.min.astype(f), self.max.astype(f)
        with np.errstate(divide="ignore", invalid="ignore"):
            lo = (mn - p) / d; hi = (mx - p) / d
        tmin, tmax = (hi[0], lo[0]) if lo[0] > hi[0] else (lo[0], hi[0])
        for k in (1, 2):
            a, b = (hi[k], lo[k]) if lo[k] > hi[k] else (lo[k], hi[k])
            if tmin > b or a > tmax:
                return fmax, fmax
            if a > tmin:
                tmin = a
            if b < tmax:
                tmax = b
        return f(tmin), f(tmax)

    def intersection(self, other):
        return BoundingBox(np.maximum(self.min, other.min), np.minimum(self.max, other.max))

    def intersects(self, other):
        i = self.intersection(other)
        return not bool(np.any(i.max < i.min))

    def __repr__(self):
        return f"[min={self.min.tolist()}, max={self.max.tolist()}]"


class _NerfSettings:
    """Testbed.nerf (S/python_api.cu:470-496): the members that affect pixels."""

    ACTIVATIONS = ("None", "ReLU", "Logistic", "Exponential")

    @property
    def cone_angle_constant(self) -> float:
        return float(self._tb._info().cone_angle_constant)

    @property
    def rgb_activation(self) -> str:
        return self.ACTIVATIONS[self._tb._info().rgb_activation]

    @property
    def density_activation(self) -> str:
        return self.ACTIVATIONS[self._tb._info().density_activation]

    def __init__(self, tb):
        self._tb = tb            # (created per access by Testbed.nerf: no reference cycle that would keep a renderer alive until the GC runs)

    @property
    def render_min_transmittance(self):
        return self._tb._min_t

    @render_min_transmittance.setter
    def render_min_transmittance(self, v):
        self._tb._min_t = float(v)
        self._tb._r._ck(lib().nmr_set_min_transmittance(self._tb._r._h, self._tb._id, float(v)))


class TonemapCurve(enum.IntEnum):
    """pynmr.TonemapCurve (S/python_api.cu:228-232)."""
    Identity = 0
    ACES = 1
    Hable = 2
    Reinhard = 3


class ColorSpace(enum.IntEnum):
    """pynmr.ColorSpace (S/python_api.cu:222-226)."""
    Linear = 0
    SRGB = 1


class Testbed:
    """ngp::Testbed as returned by NerfMeshRenderer.load_nerf (S/python_api.cu:299-496)."""

    def __init__(self, renderer: "NerfMeshRenderer", nerf_id: int):
        self._r, self._id = renderer, nerf_id
        self._min_t = 0.01
        self._scale = 1.5                                          # Testbed::m_scale after the constructor's reset_camera() (S/ngp/testbed.cu:96, 1388)
        self._up_dir = None                                        # the dataset's up vector until set (S/ngp/testbed.cu:1117)
        self._screen_center = np.array([0.5, 0.5], np.float32)
        self._sun_dir = (np.ones(3) / np.sqrt(3.0)).astype(np.float32)

    @property
    def nerf(self) -> "_NerfSettings":
        return _NerfSettings(self)

    # -- crop box
    def _get_render_aabb(self):
        mn = (C.c_float * 3)(); mx = (C.c_float * 3)()
        self._r._ck(lib().nmr_get_render_aabb(self._r._h, self._id, mn, mx))
        return np.array(list(mn), dtype=np.float32), np.array(list(mx), dtype=np.float32)

    def _set_render_aabb(self, mn, mx):
        self._r._ck(lib().nmr_set_render_aabb(self._r._h, self._id, _f3(mn), _f3(mx)))

    @property
    def render_aabb(self) -> BoundingBox:
        return BoundingBox(_owner=self)

    @render_aabb.setter
    def render_aabb(self, box: BoundingBox):
        self._set_render_aabb(box.min, box.max)

    @property
    def aabb(self) -> BoundingBox:
        mn = (C.c_float * 3)(); mx = (C.c_float * 3)()
        self._r._ck(lib().nmr_get_aabb(self._r._h, self._id, mn, mx))
        return BoundingBox(list(mn), list(mx))

    raw_aabb = aabb

    @property
    def background_color(self):
        c = (C.c_float * 4)()
        self._r._ck(lib().nmr_get_background(self._r._h, self._id, c))
        return np.array(list(c), dtype=np.float32)

    @background_color.setter
    def background_color(self, rgba):
        self._r._ck(lib().nmr_set_background(self._r._h, self._id, _f3(rgba)))

    @property
    def tonemap_curve(self) -> "TonemapCurve":
        c = C.c_int()
        self._r._ck(lib().nmr_get_tonemap_curve(self._r._h, self._id, C.byref(c)))
        return TonemapCurve(c.value)

    @tonemap_curve.setter
    def tonemap_curve(self, curve):
        self._r._ck(lib().nmr_set_tonemap_curve(self._r._h, self._id, int(curve)))

    # -- read-only snapshot facts and the camera (S/python_api.cu:301-496: secondary properties of the reference's Testbed)
    def _info(self) -> NerfInfo:
        i = NerfInfo()
        self._r._ck(lib().nmr_get_nerf_info(self._r._h, self._id, C.byref(i)))
        return i

    @property
    def training_step(self) -> int:
        return int(self._info().training_step)

    @property
    def loss(self) -> float:
        return float(self._info().loss)

    @property
    def render_aabb_to_local(self) -> np.ndarray:
        return np.array(list(self._info().render_aabb_to_local), dtype=np.float32).reshape(3, 3)

    @render_aabb_to_local.setter
    def render_aabb_to_local(self, m):
        a = np.ascontiguousarray(m, dtype=np.float32).reshape(3, 3)
        self._r._ck(lib().nmr_set_render_aabb_to_local(self._r._h, self._id, a.ctypes.data_as(C.POINTER(C.c_float))))

    @property
    def n_params(self) -> int:
        return int(self._info().n_params)

    @property
    def camera_matrix(self) -> np.ndarray:
        """Testbed.camera_matrix: kept equal to NerfMeshRenderer.view_projection_mat by frame() (S/nerf_mesh_renderer.cu:567)."""
        return self._r.view_projection_mat

    @camera_matrix.setter
    def camera_matrix(self, mat):
        self._r.view_projection_mat = mat

    # -- Testbed::m_model_translation / m_model_rotation (GUI sliders in the reference; S/ngp/testbed.cuh:508-509)
    def _model(self):
        t = (C.c_float * 3)(); r = (C.c_float * 3)(); m = (C.c_float * 9)()
        self._r._ck(lib().nmr_get_model_transform(self._r._h, self._id, t, r, m))
        return np.array(list(t), np.float32), np.array(list(r), np.float32), np.array(list(m), np.float32).reshape(3, 3)

    @property
    def model_translation(self) -> np.ndarray:
        return self._model()[0]

    @model_translation.setter
    def model_translation(self, t):
        self._r._ck(lib().nmr_set_model_transform(self._r._h, self._id, _f3(t), None))

    @property
    def model_rotation(self) -> np.ndarray:
        """three angles in units of pi about X, Y, Z (the reference's m_model_rotation)"""
        return self._model()[1]

    @model_rotation.setter
    def model_rotation(self, r):
        self._r._ck(lib().nmr_set_model_transform(self._r._h, self._id, None, _f3(r)))

    @property
    def model_matrix(self) -> np.ndarray:
        """the 3x3 rotation in use (AngleAxis(rx pi, X) * AngleAxis(ry pi, Y) * AngleAxis(rz pi, Z))"""
        return self._model()[2]

    def dump_density_grid(self, path: str | None = None) -> np.ndarray:
        """NerfMeshRenderer::dumpDensityGrid: uint8 [8, 128, 128, 128] (cascade, z, y, x), optionally written to `path`."""
        cells = np.zeros((8, 128, 128, 128), dtype=np.uint8)
        self._r._ck(lib().nmr_dump_density_grid(self._r._h, self._id, os.fsencode(path) if path else None, _ptr(cells)))
        return cells

    def load_density_grid(self, path_or_cells):
        """NerfMeshRenderer::loadDensityGrid: from a file written by dump_density_grid (or the reference), or from an array."""
        if isinstance(path_or_cells, (str, bytes, os.PathLike)):
            self._r._ck(lib().nmr_load_density_grid(self._r._h, self._id, os.fsencode(path_or_cells), None))
        else:
            cells = np.ascontiguousarray(path_or_cells, dtype=np.uint8).reshape(8, 128, 128, 128)
            self._r._ck(lib().nmr_load_density_grid(self._r._h, self._id, None, _ptr(cells)))

    def load_snapshot(self, path: str):
        """Testbed.load_snapshot (S/python_api.cu:319): another snapshot into this Testbed; raises RuntimeError when it does not load
        (the reference throws from load_snapshot as well) and keeps the old model in that case."""
        self._r._ck(lib().nmr_reload_nerf(self._r._h, self._id, os.fsencode(path)))
        self._up_dir = None

    # -- crop box as a 3x4 matrix (Testbed::crop_box / set_crop_box / crop_box_corners, S/ngp/testbed.cu:1421-1477): columns = the
    #    box's half axes and its centre; nerf_space=True converts to / from the dataset's coordinates (S/ngp/nerf_loader.cuh:115-153)
    def _dataset(self) -> NerfDataset:
        d = NerfDataset()
        self._r._ck(lib().nmr_get_nerf_dataset(self._r._h, self._id, C.byref(d)))
        return d

    def _ngp_matrix_to_nerf(self, m, scale_columns):
        d = self._dataset(); r = np.array(m, dtype=np.float32).reshape(3, 4)
        if d.from_mitsuba:
            r[:, 0] *= -1; r[:, 2] *= -1
        else:
            r = r[[2, 0, 1], :]                                    # rows: x <- z, y <- x, z <- y
        k = np.float32(1.0) / np.float32(d.scale) if scale_columns else np.float32(1.0)
        r[:, 0] *= k; r[:, 1] *= -k; r[:, 2] *= -k
        r[:, 3] = (r[:, 3] - np.array(list(d.offset), np.float32)) / np.float32(d.scale)
        return r

    def _nerf_matrix_to_ngp(self, m, scale_columns):
        d = self._dataset(); r = np.array(m, dtype=np.float32).reshape(3, 4)
        k = np.float32(d.scale) if scale_columns else np.float32(1.0)
        r[:, 0] *= k; r[:, 1] *= -k; r[:, 2] *= -k
        r[:, 3] = r[:, 3] * np.float32(d.scale) + np.array(list(d.offset), np.float32)
        if d.from_mitsuba:
            r[:, 0] *= -1; r[:, 2] *= -1
        else:
            r = r[[1, 2, 0], :]                                    # rows: x <- y, y <- z, z <- x
        return r

    def crop_box(self, nerf_space: bool = True) -> np.ndarray:
        mn, mx = self._get_render_aabb()
        r2l = self.render_aabb_to_local
        radius = (mx - mn) * np.float32(0.5)
        rv = np.zeros((3, 4), np.float32)
        for k in range(3):
            rv[:, k] = r2l[k, :] * radius[k]
        rv[:, 3] = r2l.T @ ((mx + mn) * np.float32(0.5))
        return self._ngp_matrix_to_nerf(rv, True) if nerf_space else rv

    def set_crop_box(self, matrix, nerf_space: bool = True):
        if isinstance(matrix, BoundingBox):                        # (earlier versions of this shim took the box itself)
            self._set_render_aabb(matrix.min, matrix.max)
            return
        m = np.array(matrix, dtype=np.float32).reshape(3, 4)
        if nerf_space:
            m = self._nerf_matrix_to_ngp(m, True)
        radius = np.linalg.norm(m[:, :3], axis=0).astype(np.float32)
        r2l = np.stack([m[:, k] / radius[k] for k in range(3)]).astype(np.float32)
        cen = r2l @ m[:, 3]
        self.render_aabb_to_local = r2l
        self._set_render_aabb(cen - radius, cen + radius)

    def crop_box_corners(self, nerf_space: bool = True) -> list:
        m = self.crop_box(nerf_space)
        return [m @ np.array([1 if i & 1 else -1, 1 if i & 2 else -1, 1 if i & 4 else -1, 1], np.float32) for i in range(8)]

    # -- the camera helpers of Testbed (S/ngp/testbed.cu:1319-1349) on camera_matrix: columns right, up, view direction, position
    @property
    def bounding_radius(self) -> float:
        return float(self._dataset().bounding_radius)

    @property
    def raw_aabb(self) -> BoundingBox:
        d = self._dataset()
        return BoundingBox(list(d.raw_aabb_min), list(d.raw_aabb_max))

    @property
    def up_dir(self) -> np.ndarray:
        if self._up_dir is None:
            self._up_dir = np.array(list(self._dataset().up), np.float32)
        return self._up_dir

    @up_dir.setter
    def up_dir(self, v):
        self._up_dir = np.asarray(v, np.float32).reshape(3).copy()

    def view_pos(self) -> np.ndarray:
        return self.camera_matrix[:, 3].copy()

    @property
    def view_dir(self) -> np.ndarray:
        return self.camera_matrix[:, 2].copy()

    @view_dir.setter
    def view_dir(self, d):
        d = np.asarray(d, np.float32).reshape(3)
        old = self.look_at
        cam = self.camera_matrix
        n = lambda v: v / np.float32(np.linalg.norm(v))
        cam[:, 0] = n(np.cross(d, self.up_dir)); cam[:, 1] = n(np.cross(d, cam[:, 0])); cam[:, 2] = n(d)
        self.camera_matrix = cam
        self.look_at = old

    @property
    def look_at(self) -> np.ndarray:
        cam = self.camera_matrix
        return cam[:, 3] + cam[:, 2] * np.float32(self._scale)

    @look_at.setter
    def look_at(self, pos):
        cam = self.camera_matrix
        cam[:, 3] += np.asarray(pos, np.float32).reshape(3) - self.look_at
        self.camera_matrix = cam

    @property
    def scale(self) -> float:
        return self._scale

    @scale.setter
    def scale(self, v):
        prev = self.look_at
        cam = self.camera_matrix
        cam[:, 3] = (cam[:, 3] - prev) * np.float32(float(v) / self._scale) + prev
        self.camera_matrix = cam
        self._scale = float(v)

    def translate_camera(self, rel):
        cam = self.camera_matrix
        cam[:, 3] += cam[:, :3] @ np.asarray(rel, np.float32).reshape(3) * np.float32(self.bounding_radius)
        self.camera_matrix = cam

    # -- members the reference exposes and this fork's ray generation never reads (pixel_to_ray derives the direction from the pixel
    #    centre alone, S/ngp/ngp_common.cuh:361-366): kept as plain values so that scripts setting them keep running
    zoom = 1.0
    camera_smoothing = False
    snap_to_pixel_centers = False
    display_gui = False
    visualize_unit_cube = False
    fixed_res_factor = 1
    max_level_rand_training = False
    visualized_dimension = -1
    visualized_layer = 0

    @property
    def screen_center(self) -> np.ndarray:
        return self._screen_center

    @screen_center.setter
    def screen_center(self, v):
        self._screen_center = np.asarray(v, np.float32).reshape(2).copy()

    @property
    def sun_dir(self) -> np.ndarray:
        return self._sun_dir

    @sun_dir.setter
    def sun_dir(self, v):
        self._sun_dir = np.asarray(v, np.float32).reshape(3).copy()

    # -- members that WOULD change the picture in the reference and are not built: anything but the default is refused, not ignored
    @property
    def parallax_shift(self) -> np.ndarray:
        return np.zeros(3, np.float32)

    @parallax_shift.setter
    def parallax_shift(self, v):
        if np.any(np.asarray(v, np.float32).reshape(3)[:2] != 0):
            raise NotImplementedError("parallax_shift: only (0, 0, z) - the viewer's origin is not shifted by this renderer")

    @property
    def color_space(self) -> ColorSpace:
        return ColorSpace.Linear

    @color_space.setter
    def color_space(self, v):
        if ColorSpace(int(v)) != ColorSpace.Linear:
            raise NotImplementedError("color_space: this renderer accumulates in linear colour (ColorSpace.Linear) only")

    def render(self, width: int = 1920, height: int = 1080, spp: int = 1, linear: bool = True, dtype=np.float32) -> np.ndarray:
        """float32[H, W, 4]; row 0 is the bottom of the picture; sRGB when linear=False (S/python_api.cu:83-111).
        dtype (addition): np.uint8 returns what render.py computes from the float image right after the call,
        np.uint8(img * 255) (V/render.py:62-66), converted on the device - a quarter of the bytes cross PCIe; np.float16: half."""
        fmt = _pixel_format(dtype)
        out = _pinned_array((height, width, 4), dtype=np.dtype(dtype))
        if fmt == PIXEL_F32:
            self._r._ck(lib().nmr_render(self._r._h, self._id, int(width), int(height), int(spp), int(bool(linear)), _ptr(out)))
        else:
            self._r._ck(lib().nmr_render_format(self._r._h, self._id, int(width), int(height), int(spp), int(bool(linear)), fmt, _ptr(out)))
        return out

    def render_update(self, out=None, width: int = 1920, height: int = 1080, linear: bool = True, dtype=np.float32) -> np.ndarray:
        """Addition (include/nmr.h: nmr_render_update): render(spp=1) into an image the caller keeps.  Pass the array the previous
        call returned as `out` and only the screen rectangle of head and mesh (old and new) crosses PCIe - the rest of the image
        is the background colour and already there.  The result equals render()'s; `last_update_bytes` says how much was moved.
        The array is returned read-only: whoever wants to draw into it copies it first (the next update relies on its contents)."""
        fmt = _pixel_format(dtype)
        fresh = out is None
        if fresh:
            out = _pinned_array((height, width, 4), dtype=np.dtype(dtype))
        elif out.shape != (height, width, 4) or out.dtype != np.dtype(dtype) or not out.flags.c_contiguous:
            raise ValueError("render_update: `out` must be the array a previous call returned for this size and dtype")
        moved = C.c_size_t(0)
        self._r._ck(lib().nmr_render_update(self._r._h, self._id, int(width), int(height), int(bool(linear)), fmt, C.c_void_p(out.ctypes.data), 0 if fresh else 1, C.byref(moved)))
        self.last_update_bytes = int(moved.value)
        out.flags.writeable = False
        return out

    def probe_points(self, points_world, direction) -> np.ndarray:
        """NerfTracer::intersects over world-space points (include/nmr.h: nmr_probe_points) -> alpha per point."""
        pts = np.ascontiguousarray(points_world, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(len(pts), dtype=np.float32)
        self._r._ck(lib().nmr_probe_points(self._r._h, self._id, len(pts), _ptr(pts), _f3(direction), _ptr(out)))
        return out

    def probe_rays(self, origins_world, direction) -> np.ndarray:
        """NerfTracer::collide over world-space origins (include/nmr.h: nmr_probe_rays) -> distance to the first dense sample, 0 = none."""
        pts = np.ascontiguousarray(origins_world, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(len(pts), dtype=np.float32)
        self._r._ck(lib().nmr_probe_rays(self._r._h, self._id, len(pts), _ptr(pts), _f3(direction), _ptr(out)))
        return out

    def reset_accumulation(self, due_to_camera_movement: bool = False, immediate_redraw: bool = True):
        cam = self._r.view_projection_mat
        self._r.view_projection_mat = cam   # resets the sample counter


class GltfNode:
    """GltfNode (S/python_api.cu:272-276): scale / translation are writable and re-transform the mesh."""

    def __init__(self, renderer: "NerfMeshRenderer", mesh_id: int):
        self._r, self._id = renderer, mesh_id

    def _get(self):
        t = (C.c_float * 3)(); s = (C.c_float * 3)(); r = (C.c_float * 4)()
        self._r._ck(lib().nmr_get_mesh_transform(self._r._h, self._id, t, s, r))
        return list(t), list(s), list(r)

    @property
    def translation(self):
        return Vec3(*self._get()[0], _on_change=lambda v: setattr(self, "translation", v))

    @translation.setter
    def translation(self, v):
        self._r._ck(lib().nmr_set_mesh_transform(self._r._h, self._id, _f3(list(v)), None, None))

    @property
    def scale(self):
        return Vec3(*self._get()[1], _on_change=lambda v: setattr(self, "scale", v))

    @scale.setter
    def scale(self, v):
        self._r._ck(lib().nmr_set_mesh_transform(self._r._h, self._id, None, _f3(list(v)), None))


class GltfScene:
    def __init__(self, renderer, mesh_id):
        self.nodes = [GltfNode(renderer, mesh_id)]


class NerfMeshRenderer:
    """NerfMeshRenderer (S/python_api.cu:284-298), headless."""

    def __init__(self, width: int, height: int, device: int = -1):
        h = C.c_void_p()
        rc = lib().nmr_create(int(width), int(height), int(device), C.byref(h))
        if rc != NMR_OK:
            raise RuntimeError("nmr_create failed: " + lib().nmr_last_error(None).decode())
        self._h = h
        self.width, self.height = int(width), int(height)
        self._frame_buf = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().nmr_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _err(self) -> str:
        return lib().nmr_last_error(self._h).decode()

    def _ck(self, rc: int):
        if rc != NMR_OK:
            raise RuntimeError(f"libnmr error {rc}: {self._err()}")

    # -- reference surface ---------------------------------------------------------------------------------------
    def frame(self) -> bool:
        keep = C.c_int(1)
        self._ck(lib().nmr_frame(self._h, C.byref(keep)))
        return bool(keep.value)

    def load_nerf(self, path: str):
        nid = C.c_int(-1)
        rc = lib().nmr_load_nerf(self._h, os.fsencode(path), C.byref(nid))
        if rc != NMR_OK:   # the reference logs and returns None (S/nerf_mesh_renderer.cu:995-999)
            print(f"[error] Failed to load NeRF {path}: {self._err()}", file=sys.stderr)
            return None
        return Testbed(self, nid.value)

    def load_mesh(self, path: str, t=(0.0, 0.0, 0.0), s=(1.0, 1.0, 1.0), r=(0.0, 0.0, 0.0, 1.0)):
        mid = C.c_int(-1)
        rc = lib().nmr_load_mesh(self._h, os.fsencode(path), _f3(t), _f3(s), _f3(r), C.byref(mid))
        if rc != NMR_OK:
            print(f"[error] Failed to load mesh {path}: {self._err()}", file=sys.stderr)
            return None
        warn = self._err()
        if warn:
            print(f"[warn] {path}: {warn}", file=sys.stderr)
        return GltfScene(self, mid.value)

    def remove_floaties(self):
        n = C.c_int(0); kept = C.c_int64(0)
        self._ck(lib().nmr_remove_floaties(self._h, C.byref(n), C.byref(kept)))
        return n.value, kept.value

    def orbit(self, delta_azimuth: float, delta_polar: float, delta_zoom: float):
        self._ck(lib().nmr_orbit(self._h, float(delta_azimuth), float(delta_polar), float(delta_zoom)))

    @property
    def view_projection_mat(self) -> np.ndarray:
        m = (C.c_float * 12)()
        self._ck(lib().nmr_get_camera(self._h, m))
        return np.array(list(m), dtype=np.float32).reshape(4, 3).T.copy()   # 3x4, like Eigen -> numpy

    @view_projection_mat.setter
    def view_projection_mat(self, mat):
        a = np.asarray(mat, dtype=np.float32).reshape(3, 4)
        self._ck(lib().nmr_set_camera(self._h, _f3(a.T.reshape(-1))))

    # -- the GUI's trajectory tool ("Run Trajectory", S/nerf_mesh_renderer.cu:604-659, 806-826), without the GUI
    def trajectory_pose(self, angle: float, distance: float = 1.1, height: float = 0.1, lookat=(0.0, 0.0, 0.0)):
        """The camera of one step of the tool's arc: eye = (cos(angle) distance, height, sin(angle) distance), looking at `lookat`."""
        self._ck(lib().nmr_trajectory_pose(self._h, float(angle), float(distance), float(height), _f3(lookat)))

    def export_trajectory(self, out_dir: str = ".", start_angle: float = 0.5, end_angle: float = 2.5, num_images: int = 10,
                          distance: float = 1.1, height: float = 0.1, lookat=(0.0, 0.0, 0.0)) -> int:
        """Walks the arc like the reference's GUI loop (defaults = its sliders' initial values) and writes, per step N = 1, 2, ...:
        `transform_N` - view_projection_mat in the text layout of the reference (Eigen::IOFormat(FullPrecision, 0, ", ", ",\\n",
        "[", "]", "[", "]"), see format_transform) - and the displayed frame as `trajectory_N.png` (the reference writes a JPEG of
        the same picture through stb_image_write; no JPEG encoder here, PNG is lossless).  Returns the number of steps written.
        As in the reference the pose of the step that reaches end_angle is set but not written."""
        os.makedirs(out_dir, exist_ok=True)
        f32 = np.float32
        angle, idx, running = f32(start_angle), 0, True
        step = f32(f32(f32(end_angle) - f32(start_angle)) / f32(num_images))
        while running:
            if idx > 0:
                self._ck_frame()
                img = np.asarray(self.read_frame())
                rgb = np.uint8(np.clip(img[::-1, :, :3], 0.0, 1.0) * f32(255.0))       # row 0 of a frame is the bottom of the picture
                _write_png_rgb8(os.path.join(out_dir, f"trajectory_{idx}.png"), rgb)
                with open(os.path.join(out_dir, f"transform_{idx}"), "w") as fh:
                    fh.write(format_transform(self.view_projection_mat))
            angle = f32(angle + step)
            idx += 1
            if angle >= f32(end_angle):
                running = False
            self.trajectory_pose(float(angle), distance, height, lookat)
        return idx - 1

    def _ck_frame(self):
        if not self.frame():
            raise RuntimeError("frame() asked to stop")

    def envmap(self, path: str):
        """V/render.py:228 calls this, but the reference module has no such method and its renderer no environment map
        (primary rays over a constant background): accepted so that the script runs, ignored, and said so once."""
        global _envmap_warned
        if not _envmap_warned:
            import warnings
            warnings.warn("pynmr: envmap() has no effect - the renderer composites over the constant background colour, like the reference", stacklevel=2)
            _envmap_warned = True
        self._ck(lib().nmr_set_envmap(self._h, os.fsencode(path)))

    # -- additions -----------------------------------------------------------------------------------------------
    def read_frame(self) -> np.ndarray:
        """Image of the last frame(): float32[H, W, 4], row 0 = bottom (replaces the reference's on-screen blit)."""
        out = _pinned_array((self.height, self.width, 4))
        self._ck(lib().nmr_read_frame(self._h, _ptr(out)))
        return out

    def read_combined(self):
        """(linear premultiplied frame float32[H, W, 4], depth float32[H, W]) merged over all NeRFs of the last frame()
        (include/nmr.h: nmr_read_combined - the reference's _dFramebuffer / _dDepthBuffer)."""
        fr = np.zeros((self.height, self.width, 4), dtype=np.float32); dp = np.zeros((self.height, self.width), dtype=np.float32)
        self._ck(lib().nmr_read_combined(self._h, _ptr(fr), _ptr(dp)))
        return fr, dp

    def frame_async(self):
        self._ck(lib().nmr_frame_async(self._h))

    def synchronize(self):
        self._ck(lib().nmr_synchronize(self._h))

    def render_views(self, nerf: Testbed, cameras, width: int, height: int, linear: bool = False, to_host: bool = True, dtype=np.float32, out_device_ptr: int = 0):
        """cameras: [n, 3, 4] -> dtype[n, H, W, 4] (render.py's landmark pass in one call; up to 8 views in flight on the GPU).
        to_host=False leaves the images on the device (returns None; device_image() is the last view).  dtype as in Testbed.render."""
        cams = np.ascontiguousarray(np.asarray(cameras, dtype=np.float32).reshape(-1, 3, 4).transpose(0, 2, 1)).reshape(-1, 12)
        fmt = _pixel_format(dtype)
        if out_device_ptr:       # images go to caller-owned device memory (n x H x W x 4 elements of dtype), e.g. a torch tensor
            self._ck(lib().nmr_render_views_format(self._h, nerf._id, cams.shape[0], _ptr(cams), int(width), int(height), int(bool(linear)), fmt, C.c_void_p(int(out_device_ptr))))
            return None
        out = _pinned_array((cams.shape[0], height, width, 4), dtype=np.dtype(dtype)) if to_host else None
        self._ck(lib().nmr_render_views_format(self._h, nerf._id, cams.shape[0], _ptr(cams), int(width), int(height), int(bool(linear)), fmt, _ptr(out) if to_host else None))
        return out

    def set_shard(self, rank: int, world: int, band: int = 8):
        self._ck(lib().nmr_set_shard(self._h, int(rank), int(world), int(band)))

    SURFACE_AUTO, SURFACE_EXACT, SURFACE_BATCH8 = 0, 1, 2

    def set_surface_insertion(self, mode: int):
        """Where a partially covering mesh surface enters the compositing order (include/nmr.h: nmr_surface_mode)."""
        self._ck(lib().nmr_set_surface_insertion(self._h, int(mode)))

    def set_overlap(self, enabled: bool = True):
        """March kernel consuming the ray queue while the set-up kernel fills it (include/nmr.h: nmr_set_overlap); pixels do not change."""
        self._ck(lib().nmr_set_overlap(self._h, int(bool(enabled))))

    def set_lens(self, enabled: bool = True, ior: float = -1.0, transmission: float = -1.0, tint=None):
        """Lens surfaces and their secondary rays (include/nmr.h: nmr_set_lens).  Negative / None keeps a parameter."""
        self._ck(lib().nmr_set_lens(self._h, int(bool(enabled)), float(ior), float(transmission), _f3(tint) if tint is not None else None))

    LENS_THIN, LENS_PLATE = 0, 1

    def set_lens_model(self, model: int = 0, thickness: float = 0.0):
        """Thin sheet (0, no bending) or a pane of `thickness` with parallel faces (1: two Snell interfaces) - include/nmr.h."""
        self._ck(lib().nmr_set_lens_model(self._h, int(model), float(thickness)))

    def stream_ptr(self) -> int:
        """cudaStream_t of this renderer as an integer (e.g. for torch.cuda.ExternalStream)."""
        p = C.c_void_p()
        self._ck(lib().nmr_get_stream(self._h, C.byref(p)))
        return int(p.value or 0)

    def gather_create(self):
        """Destination rank of a tile-sharded frame: -> (64-byte IPC handle as bytes, device pointer of the shared image)."""
        h = (C.c_uint8 * 64)(); p = C.c_void_p()
        self._ck(lib().nmr_gather_create(self._h, h, C.byref(p)))
        return bytes(h), p.value

    def gather_attach(self, handle: bytes):
        """Other ranks: frame() writes this rank's rows into the destination's shared image from now on."""
        assert len(handle) == 64
        self._ck(lib().nmr_gather_attach(self._h, (C.c_uint8 * 64)(*handle)))

    def gather_detach(self):
        self._ck(lib().nmr_gather_detach(self._h))

    def device_image(self):
        """(device pointer, width, height) of the float4 image of the last render."""
        p = C.c_void_p(); w = C.c_int(); h = C.c_int()
        self._ck(lib().nmr_get_device_image(self._h, C.byref(p), C.byref(w), C.byref(h)))
        return p.value, w.value, h.value

    def copy_device_image(self, dst_device_ptr: int, dst_bytes: int):
        """Last image (its own resolution / format) -> caller-owned device memory of dst_bytes bytes."""
        self._ck(lib().nmr_copy_device_image(self._h, C.c_void_p(dst_device_ptr), int(dst_bytes)))

    def flush_l2(self):
        self._ck(lib().nmr_flush_l2(self._h))

    def measure_l2(self, nbytes: int = 32 << 20, gather: bool = False) -> float:
        """GB/s the L2 delivers to the SMs from a resident buffer (include/nmr.h: nmr_measure_l2)."""
        v = C.c_float(0)
        self._ck(lib().nmr_measure_l2(self._h, int(nbytes), 1 if gather else 0, C.byref(v)))
        return float(v.value)

    def stats(self) -> dict:
        s = Stats()
        self._ck(lib().nmr_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}


def free_temporary_memory():
    """tcnn::free_all_gpu_memory_arenas in the reference (S/python_api.cu:159); here: the pinned host buffer pool."""
    free_pinned_pool()
