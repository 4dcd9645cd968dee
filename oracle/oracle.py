"""ctypes front-end of the CPU oracle (oracle/nmr_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (nerf-glasses_b200/) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "nmr_oracle.c")
LIB = os.path.join(HERE, "liboracle.so")

CFLAGS = ["-O3", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fvisibility=hidden",
          "-shared", "-fPIC", "-std=c11"]


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, OpenMP).  Rebuilds when the source is newer than the library."""
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", *CFLAGS, "-o", LIB, SRC, "-lm"])
    return LIB


class RenderParams(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("camera", C.c_float * 12),
        ("aabb_min", C.c_float * 3), ("aabb_max", C.c_float * 3),
        ("render_aabb_to_local", C.c_float * 9),
        ("train_aabb_min", C.c_float * 3), ("train_aabb_max", C.c_float * 3),
        ("cone_angle", C.c_float),
        ("spp_index", C.c_uint32),
        ("min_transmittance", C.c_float),
        ("rgb_activation", C.c_int32), ("density_activation", C.c_int32),
        ("n_steps_mode", C.c_int32),
        ("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32),
        ("model_rot", C.c_float * 9), ("model_trans", C.c_float * 3),
    ]


class Camera(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("eye", C.c_float * 3), ("look", C.c_float * 3),
                ("pivot", C.c_float * 3), ("up", C.c_float * 3)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp, f32p, u16p, u8p, u32p, i32p, u64p, i64p = (C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint16),
                                                   C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_int32),
                                                   C.POINTER(C.c_uint64), C.POINTER(C.c_int64))
    sig = {
        "orc_model_create": (vp, [vp, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]),
        "orc_model_destroy": (None, [vp]),
        "orc_model_level_table": (None, [vp, vp, vp, vp, vp]),
        "orc_model_per_level_scale": (C.c_float, [vp]),
        "orc_model_cone_angle": (C.c_float, [vp]),
        "orc_model_bitfield": (vp, [vp]),
        "orc_model_density_mean": (C.c_float, [vp]),
        "orc_model_set_density_grid": (C.c_int, [vp, vp, C.c_uint64]),
        "orc_model_set_bitfield": (None, [vp, vp]),
        "orc_encode": (None, [vp, vp, C.c_int64, vp]),
        "orc_sh4": (None, [vp, C.c_int64, vp]),
        "orc_mlp": (None, [vp, C.c_int, vp, C.c_int64, vp]),
        "orc_network": (None, [vp, vp, vp, C.c_int64, vp]),
        "orc_render": (C.c_int, [vp, C.POINTER(RenderParams), vp, vp, vp, vp, vp, vp]),
        "orc_accumulate_tonemap": (None, [vp, vp, C.c_int64, C.c_uint32, vp, C.c_int, vp]),
        "orc_accumulate_tonemap_curve": (None, [vp, vp, C.c_int64, C.c_uint32, vp, C.c_int, C.c_int, vp]),
        "orc_probe_points": (None, [vp, C.POINTER(RenderParams), vp, vp, C.c_int64, vp]),
        "orc_probe_rays": (None, [vp, C.POINTER(RenderParams), vp, vp, C.c_int64, vp]),
        "orc_trace_samples": (None, [vp, C.POINTER(RenderParams), vp, C.c_int64, C.c_uint32, vp, vp, vp, vp, vp, vp]),
        "orc_mesh_create": (vp, [vp, vp, vp, C.c_uint32, vp, C.c_uint32, vp, vp, vp, vp, C.c_float, C.c_float, vp, vp, C.c_int, C.c_int]),
        "orc_mesh_destroy": (None, [vp]),
        "orc_mesh_set_texture": (None, [vp, C.c_int, vp, C.c_int, C.c_int, C.c_float]),
        "orc_mesh_set_tangents": (None, [vp, vp, vp, vp, vp]),
        "orc_mesh_world_positions": (None, [vp, vp]),
        "orc_mesh_render": (None, [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp]),
        "orc_mesh_resolve": (None, [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]),
        "orc_bitfield_to_cells": (None, [vp, vp]),
        "orc_cells_to_bitfield": (None, [vp, vp]),
        "orc_remove_floaties": (C.c_int, [vp, vp, vp]),
        "orc_camera_init": (None, [C.POINTER(Camera)]),
        "orc_camera_orbit": (None, [C.POINTER(Camera), C.c_float, C.c_float, C.c_float]),
        "orc_camera_matrix": (None, [C.POINTER(Camera), C.c_int, C.c_int, vp]),
        "orc_camera_trajectory_pose": (None, [C.POINTER(Camera), C.c_float, C.c_float, C.c_float, vp]),
        "orc_f2h": (None, [vp, vp, C.c_int64]),
        "orc_h2f": (None, [vp, vp, C.c_int64]),
        "orc_morton3D": (C.c_uint32, [C.c_uint32, C.c_uint32, C.c_uint32]),
        "orc_morton3D_invert": (C.c_uint32, [C.c_uint32]),
        "orc_ld_random_val": (C.c_float, [C.c_uint32, C.c_uint32]),
        "orc_linear_to_srgb": (C.c_float, [C.c_float]),
        "orc_srgb_to_linear": (C.c_float, [C.c_float]),
        "orc_aabb_ray_intersect": (None, [vp, vp, vp, vp, vp]),
        "orc_cascaded_grid_idx_at": (C.c_uint32, [vp, C.c_uint32]),
        "orc_mip_from_pos": (C.c_int, [vp]),
        "orc_calc_dt": (C.c_float, [C.c_float, C.c_float]),
        "orc_num_threads": (C.c_int, []),
        "orc_set_num_threads": (None, [C.c_int]),
        "orc_render_lens": (C.c_int, [vp, C.POINTER(RenderParams), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "orc_mesh_set_lens": (None, [vp, vp]),
        "orc_mesh_render_layers": (None, [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]),
        "orc_lens_resolve": (None, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


HASH_TYPES = {"Prime": 0, "CoherentPrime": 1, "ReversedPrime": 2}
BITFIELD_BYTES = 8 * 128 ** 3 // 8


class OrbitCamera:
    """The NerfMeshRenderer camera state (cam_pos/cam_look/pivot/viewMat) and its 3x4 matrix."""

    def __init__(self, width: int, height: int):
        self.w, self.h = width, height
        self.c = Camera()
        lib().orc_camera_init(C.byref(self.c))

    def orbit(self, delta_azimuth: float, delta_polar: float, delta_zoom: float):
        lib().orc_camera_orbit(C.byref(self.c), delta_azimuth, delta_polar, delta_zoom)

    def trajectory_pose(self, angle: float, distance: float = 1.1, height: float = 0.1, lookat=(0.0, 0.0, 0.0)):
        """One pose of the GUI's trajectory tool (S/nerf_mesh_renderer.cu:649-658)."""
        la = np.ascontiguousarray(lookat, dtype=np.float32)
        lib().orc_camera_trajectory_pose(C.byref(self.c), C.c_float(angle), C.c_float(distance), C.c_float(height), _p(la))

    @property
    def view(self):
        return np.array(list(self.c.view), dtype=np.float32)

    def matrix(self) -> np.ndarray:
        """float32[12], column-major 3x4 (col0=right*uLen, col1=up*vLen, col2=fwd, col3=eye)."""
        out = np.empty(12, dtype=np.float32)
        lib().orc_camera_matrix(C.byref(self.c), self.w, self.h, _p(out))
        return out

    @property
    def eye(self):
        return np.array(list(self.c.eye), dtype=np.float32)


class Model:
    def __init__(self, params: np.ndarray, density_grid: np.ndarray | None, n_levels=16, log2_hashmap_size=19,
                 base_resolution=16, per_level_scale=0.0, aabb_scale=1, hash="CoherentPrime", density_hidden=1, rgb_hidden=2):
        params = np.ascontiguousarray(params.view(np.uint16) if params.dtype == np.float16 else params, dtype=np.uint16)
        self.h = lib().orc_model_create(_p(params), params.size, n_levels, log2_hashmap_size, base_resolution,
                                        float(per_level_scale), aabb_scale, HASH_TYPES[hash], density_hidden, rgb_hidden)
        if not self.h:
            raise ValueError("orc_model_create: parameter count / configuration mismatch")
        self.n_levels, self.aabb_scale = n_levels, aabb_scale
        half = 0.5 * min(128, aabb_scale)
        self.aabb_min = np.full(3, 0.5 - half, dtype=np.float32)
        self.aabb_max = np.full(3, 0.5 + half, dtype=np.float32)
        if density_grid is not None:
            g = np.ascontiguousarray(density_grid.view(np.uint16) if density_grid.dtype == np.float16 else density_grid, dtype=np.uint16)
            if lib().orc_model_set_density_grid(self.h, _p(g), g.size) != 0:
                raise ValueError("incompatible number of grid cascades")

    @classmethod
    def from_snapshot(cls, snap: dict) -> "Model":
        """snap: tools.synth.read_snapshot() output."""
        return cls(snap["params"], snap["density_grid"], snap["n_levels"], snap["log2_hashmap_size"], snap["base_resolution"],
                   snap.get("per_level_scale", 0.0), snap.get("aabb_scale", 1), snap.get("hash", "CoherentPrime"))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_model_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def level_table(self):
        offs = np.zeros(17, dtype=np.uint32); sc = np.zeros(16, dtype=np.float32)
        res = np.zeros(16, dtype=np.uint32); dense = np.zeros(16, dtype=np.int32)
        lib().orc_model_level_table(self.h, _p(offs), _p(sc), _p(res), _p(dense))
        n = self.n_levels
        return offs[:n + 1], sc[:n], res[:n], dense[:n]

    @property
    def cone_angle(self) -> float:
        return float(lib().orc_model_cone_angle(self.h))

    @property
    def density_mean(self) -> float:
        return float(lib().orc_model_density_mean(self.h))

    def bitfield(self) -> np.ndarray:
        ptr = lib().orc_model_bitfield(self.h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(BITFIELD_BYTES,)).copy()

    def set_bitfield(self, bits: np.ndarray):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        assert bits.size == BITFIELD_BYTES
        lib().orc_model_set_bitfield(self.h, _p(bits))

    def encode(self, pos: np.ndarray) -> np.ndarray:
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        out = np.empty((pos.shape[0], self.n_levels * 2), dtype=np.uint16)
        lib().orc_encode(self.h, _p(pos), pos.shape[0], _p(out))
        return out.view(np.float16)

    def mlp(self, which: int, x16: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x16.view(np.uint16))
        out = np.empty((x.shape[0], 16), dtype=np.uint16)
        lib().orc_mlp(self.h, which, _p(x), x.shape[0], _p(out))
        return out.view(np.float16)

    def network(self, pos: np.ndarray, dir01: np.ndarray) -> np.ndarray:
        pos = np.ascontiguousarray(pos, dtype=np.float32); d = np.ascontiguousarray(dir01, dtype=np.float32)
        out = np.empty((pos.shape[0], 4), dtype=np.uint16)
        lib().orc_network(self.h, _p(pos), _p(d), pos.shape[0], _p(out))
        return out.view(np.float16)

    def params_struct(self, width, height, camera12, aabb_min=None, aabb_max=None, spp_index=0, n_steps_mode=0,
                      min_transmittance=0.01, rgb_activation=2, density_activation=3, window=None, model_rot=None, model_trans=None) -> RenderParams:
        P = RenderParams()
        P.width, P.height = width, height
        P.camera[:] = [float(x) for x in camera12]
        amin = self.aabb_min if aabb_min is None else np.asarray(aabb_min, dtype=np.float32)
        amax = self.aabb_max if aabb_max is None else np.asarray(aabb_max, dtype=np.float32)
        P.aabb_min[:] = [float(x) for x in amin]; P.aabb_max[:] = [float(x) for x in amax]
        P.render_aabb_to_local[:] = [1, 0, 0, 0, 1, 0, 0, 0, 1]
        P.train_aabb_min[:] = [float(x) for x in self.aabb_min]; P.train_aabb_max[:] = [float(x) for x in self.aabb_max]
        P.cone_angle = self.cone_angle
        P.spp_index = spp_index
        P.min_transmittance = min_transmittance
        P.rgb_activation, P.density_activation = rgb_activation, density_activation
        P.n_steps_mode = n_steps_mode
        if window is not None:
            P.x0, P.y0, P.x1, P.y1 = window
        P.model_rot[:] = [1, 0, 0, 0, 1, 0, 0, 0, 1] if model_rot is None else [float(x) for x in np.asarray(model_rot, dtype=np.float32).reshape(9)]
        P.model_trans[:] = [0, 0, 0] if model_trans is None else [float(x) for x in np.asarray(model_trans, dtype=np.float32).reshape(3)]
        return P

    def render_frame(self, P: RenderParams, surf_rgba=None, t_surface=None, lens=None):
        """-> (frame f32[H,W,4] linear premultiplied, depth f32[H,W], n_samples u32[H,W], stats dict)
        lens (new functionality, no reference equivalent) = dict(w, t, n, f0, k (3), background (4)): per-pixel lens hand-off."""
        W, H = P.width, P.height
        frame = np.zeros((H, W, 4), dtype=np.float32); depth = np.zeros((H, W), dtype=np.float32)
        ns = np.zeros((H, W), dtype=np.uint32); stats = np.zeros(4, dtype=np.uint64)
        sp = tp = None
        if t_surface is not None:
            surf_rgba = np.ascontiguousarray(surf_rgba, dtype=np.float32); t_surface = np.ascontiguousarray(t_surface, dtype=np.float32)
            sp, tp = _p(surf_rgba), _p(t_surface)
        if lens is not None:
            lw = np.ascontiguousarray(lens["w"], dtype=np.float32); lt = np.ascontiguousarray(lens["t"], dtype=np.float32)
            ln = np.ascontiguousarray(lens["n"], dtype=np.float32)
            k = np.asarray(lens["k"], dtype=np.float32)
            l9 = np.array([lens["f0"], k[0], k[1], k[2], float((k[0] + k[1] + k[2]) / np.float32(3.0)), *lens["background"],
                           float(lens.get("model", 0)), float(lens.get("thickness", 0.0)), float(lens.get("ior", 1.5))], dtype=np.float32)
            rc = lib().orc_render_lens(self.h, C.byref(P), sp, tp, _p(lw), _p(lt), _p(ln), _p(l9), _p(frame), _p(depth), _p(ns), _p(stats))
        else:
            rc = lib().orc_render(self.h, C.byref(P), sp, tp, _p(frame), _p(depth), _p(ns), _p(stats))
        if rc != 0:
            raise RuntimeError("orc_render failed")
        return frame, depth, ns, {"alive_after_first_hit": int(stats[0]), "samples": int(stats[1]),
                                 "iterations": int(stats[2]), "rays_hit": int(stats[3])}

    def probe_points(self, P: RenderParams, points_world: np.ndarray, direction) -> np.ndarray:
        """NerfTracer::intersects: alpha of one minimum step at each point, 0 where the cell is not occupied."""
        pts = np.ascontiguousarray(points_world, dtype=np.float32).reshape(-1, 3); d = np.ascontiguousarray(direction, dtype=np.float32)
        out = np.zeros(len(pts), dtype=np.float32)
        lib().orc_probe_points(self.h, C.byref(P), _p(pts), _p(d), len(pts), _p(out))
        return out

    def probe_rays(self, P: RenderParams, origins_world: np.ndarray, direction) -> np.ndarray:
        """NerfTracer::collide + check_collision: distance to the first sample with positive alpha, 0 when there is none."""
        pts = np.ascontiguousarray(origins_world, dtype=np.float32).reshape(-1, 3); d = np.ascontiguousarray(direction, dtype=np.float32)
        out = np.zeros(len(pts), dtype=np.float32)
        lib().orc_probe_rays(self.h, C.byref(P), _p(pts), _p(d), len(pts), _p(out))
        return out

    def trace_samples(self, P: RenderParams, pixels: np.ndarray, max_samples: int):
        pixels = np.ascontiguousarray(pixels, dtype=np.uint32); n = pixels.size
        t = np.zeros((n, max_samples), dtype=np.float32); cell = np.zeros((n, max_samples), dtype=np.uint32)
        mip = np.zeros((n, max_samples), dtype=np.uint32); pos = np.zeros((n, max_samples, 3), dtype=np.float32)
        cnt = np.zeros(n, dtype=np.uint32); ray = np.zeros((n, 8), dtype=np.float32)
        lib().orc_trace_samples(self.h, C.byref(P), _p(pixels), n, max_samples, _p(t), _p(cell), _p(mip), _p(pos), _p(cnt), _p(ray))
        return {"t": t, "cell": cell, "mip": mip, "pos": pos, "count": cnt, "ray": ray}


def accumulate_tonemap(frame: np.ndarray, accum: np.ndarray | None, spp_index: int, background=(1.0, 1.0, 1.0, 1.0), to_srgb=True, curve: int = 0):
    """-> (image f32[H,W,4], accum).  accum None starts a new accumulation."""
    frame = np.ascontiguousarray(frame, dtype=np.float32)
    if accum is None:
        accum = np.zeros_like(frame)
    out = np.empty_like(frame)
    bg = np.asarray(background, dtype=np.float32)
    lib().orc_accumulate_tonemap_curve(_p(frame), _p(accum), frame.shape[0] * frame.shape[1], spp_index, _p(bg), int(to_srgb), int(curve), _p(out))
    return out, accum


class Mesh:
    def __init__(self, positions, normals, texcoords, indices, t=(0, 0, 0), s=(1, 1, 1), r_wxyz=(0, 0, 0, 1),
                 base_color=(1, 1, 1, 1), metallic=1.0, roughness=1.0, emissive=(0, 0, 0), texture_rgba8: np.ndarray | None = None):
        pos = np.ascontiguousarray(positions, dtype=np.float32); nrm = np.ascontiguousarray(normals, dtype=np.float32)
        uv = np.ascontiguousarray(texcoords, dtype=np.float32); idx = np.ascontiguousarray(indices, dtype=np.uint16)
        tt = np.asarray(t, dtype=np.float32); ss = np.asarray(s, dtype=np.float32); rr = np.asarray(r_wxyz, dtype=np.float32)
        bc = np.asarray(base_color, dtype=np.float32); em = np.asarray(emissive, dtype=np.float32)
        tex, tw, th = None, 0, 0
        if texture_rgba8 is not None:
            tex = np.ascontiguousarray(texture_rgba8, dtype=np.uint8); th, tw = tex.shape[:2]
        self.n_verts = pos.shape[0]
        self.h = lib().orc_mesh_create(_p(pos), _p(nrm), _p(uv), pos.shape[0], _p(idx), idx.size, _p(tt), _p(ss), _p(rr), _p(bc),
                                       float(metallic), float(roughness), _p(em), _p(tex) if tex is not None else None, tw, th)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_mesh_destroy(self.h)
                self.h = None
        except Exception:
            pass

    TEXTURES = {"emissive": 0, "metallic_roughness": 1, "normal": 2, "occlusion": 3}

    def set_texture(self, which: str, rgba8, factor: float = 1.0):
        """The other textures of the reference's closest-hit program; factor = normalTexture.scale / occlusionTexture.strength."""
        if rgba8 is None:
            lib().orc_mesh_set_texture(self.h, self.TEXTURES[which], None, 0, 0, float(factor))
        else:
            t = np.ascontiguousarray(rgba8, dtype=np.uint8)
            lib().orc_mesh_set_texture(self.h, self.TEXTURES[which], _p(t), t.shape[1], t.shape[0], float(factor))

    def set_tangents(self, normals_obj, tangents4_obj, s, r_wxyz):
        """Object-space normals / tangents (xyz + handedness) and the mesh's scale / rotation, for the normal map's TBN matrix."""
        n = np.ascontiguousarray(normals_obj, dtype=np.float32); t = np.ascontiguousarray(tangents4_obj, dtype=np.float32)
        ss = np.asarray(s, dtype=np.float32); rr = np.asarray(r_wxyz, dtype=np.float32)
        lib().orc_mesh_set_tangents(self.h, _p(n), _p(t), _p(ss), _p(rr))

    def set_lens(self, tri_lens):
        """Per-triangle lens flags (uint8) or None."""
        if tri_lens is None:
            lib().orc_mesh_set_lens(self.h, None)
        else:
            f = np.ascontiguousarray(tri_lens, dtype=np.uint8)
            lib().orc_mesh_set_lens(self.h, _p(f))

    def render_layers(self, camera12, W2: int, H2: int, light=(1.0, 1.0, 1.0), window=None):
        """-> (opaque rgba, opaque hitT (NaN on miss), lens hitT (0 on miss), lens unit normal); window = (x0, y0, x1, y1) in
        supersampled pixels renders only that sub-rectangle (the rest reads as a miss in both layers)."""
        cam = np.asarray(camera12, dtype=np.float32); lp = np.asarray(light, dtype=np.float32)
        rgba = np.zeros((H2, W2, 4), dtype=np.float32); depth = np.full((H2, W2), np.nan, dtype=np.float32)
        ld = np.zeros((H2, W2), dtype=np.float32); ln = np.zeros((H2, W2, 3), dtype=np.float32); ln[..., 2] = 1.0
        win = None if window is None else np.asarray(window, dtype=np.int32)
        lib().orc_mesh_render_layers(self.h, _p(cam), _p(lp), W2, H2, _p(rgba), _p(depth), _p(ld), _p(ln), _p(win) if win is not None else None)
        return rgba, depth, ld, ln

    def world_positions(self) -> np.ndarray:
        out = np.empty((self.n_verts, 3), dtype=np.float32)
        lib().orc_mesh_world_positions(self.h, _p(out))
        return out

    def render(self, camera12, W2: int, H2: int, light=(1.0, 1.0, 1.0), window=None):
        """-> (rgba f32[H2,W2,4], hitT f32[H2,W2] (NaN on miss), tri i32[H2,W2]); window = (x0, y0, x1, y1) in
        supersampled pixels renders only that sub-rectangle (the rest reads as a miss)."""
        cam = np.asarray(camera12, dtype=np.float32); lp = np.asarray(light, dtype=np.float32)
        rgba = np.zeros((H2, W2, 4), dtype=np.float32); depth = np.full((H2, W2), np.nan, dtype=np.float32); tri = np.full((H2, W2), -1, dtype=np.int32)
        win = None if window is None else np.asarray(window, dtype=np.int32)
        lib().orc_mesh_render(self.h, _p(cam), _p(lp), W2, H2, _p(rgba), _p(depth), _p(tri), _p(win) if win is not None else None)
        return rgba, depth, tri


def mesh_resolve(rgba2: np.ndarray, depth2: np.ndarray, W: int, H: int, mesh_scale: int = 2):
    """-> (surface_color f32[H,W,4], t_surface f32[H,W])"""
    rgba2 = np.ascontiguousarray(rgba2, dtype=np.float32); depth2 = np.ascontiguousarray(depth2, dtype=np.float32)
    surf = np.empty((H, W, 4), dtype=np.float32); ts = np.empty((H, W), dtype=np.float32)
    lib().orc_mesh_resolve(_p(rgba2), _p(depth2), W, H, mesh_scale, _p(surf), _p(ts))
    return surf, ts


def lens_resolve(depth2, lens_depth2, lens_normal2, t_surface, W: int, H: int, mesh_scale: int = 2):
    """-> (coverage w f32[H,W], t_lens f32[H,W], normal f32[H,W,3])"""
    d2 = np.ascontiguousarray(depth2, dtype=np.float32); l2 = np.ascontiguousarray(lens_depth2, dtype=np.float32)
    n2 = np.ascontiguousarray(lens_normal2, dtype=np.float32); ts = np.ascontiguousarray(t_surface, dtype=np.float32)
    w = np.empty((H, W), dtype=np.float32); t = np.empty((H, W), dtype=np.float32); n = np.empty((H, W, 3), dtype=np.float32)
    lib().orc_lens_resolve(_p(d2), _p(l2), _p(n2), _p(ts), W, H, mesh_scale, _p(w), _p(t), _p(n))
    return w, t, n


def lens_f0(ior: float) -> np.float32:
    r0 = (np.float32(ior) - np.float32(1.0)) / (np.float32(ior) + np.float32(1.0))
    return np.float32(r0 * r0)


def bitfield_to_cells(bitfield: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(bitfield, dtype=np.uint8)
    cells = np.empty(8 * 128 ** 3, dtype=np.uint8)
    lib().orc_bitfield_to_cells(_p(b), _p(cells))
    return cells


def cells_to_bitfield(cells: np.ndarray) -> np.ndarray:
    c = np.ascontiguousarray(cells, dtype=np.uint8)
    b = np.empty(BITFIELD_BYTES, dtype=np.uint8)
    lib().orc_cells_to_bitfield(_p(c), _p(b))
    return b


def remove_floaties_cells(cells: np.ndarray):
    """In-place NgpGrid pipeline on the dumped byte grid -> (n_clusters, best_size, best_score)."""
    assert cells.dtype == np.uint8 and cells.size == 8 * 128 ** 3 and cells.flags.c_contiguous
    size = C.c_int64(0); score = C.c_int64(0)
    n = lib().orc_remove_floaties(_p(cells), C.byref(size), C.byref(score))
    return n, size.value, score.value


def remove_floaties_bitfield(bitfield: np.ndarray):
    """NerfMeshRenderer::removeFloaties on a bitfield -> (new bitfield, n_clusters, best_size)."""
    cells = bitfield_to_cells(bitfield)
    n, size, _ = remove_floaties_cells(cells)
    return cells_to_bitfield(cells), n, size


def f2h(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32); out = np.empty(x.shape, dtype=np.uint16)
    lib().orc_f2h(_p(x), _p(out), x.size)
    return out
