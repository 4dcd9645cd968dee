"""Synthetic stand-ins for the reference's git-LFS inputs.

The reference's bundled model (volume/datasets/alice/base.msgpack) and the
glasses texture are LFS pointers, so every test / bench input is generated here:

* ``write_snapshot``  - a schema-valid Instant-NGP snapshot (.msgpack) with the
  stock network (16-level hash grid, 64-wide density/rgb MLPs, SH degree 4),
  laid out exactly as the reference loader expects it
  (S/ngp/testbed.cu:939-1002 for the snapshot keys, S/ngp/nerf_network.cuh:359-392
  for the parameter order  density net | rgb net | hash grid,
  T/include/tiny-cuda-nn/encodings/grid.h:985-1016 for the per-level offsets).
* ``write_glasses_gltf`` - re-materialises the glasses mesh fixture
  (tests/golden/glasses_mesh.npz) as .gltf + .bin + a small PNG texture.
* a vectorised numpy restatement of the encoder and the MLPs, used to calibrate
  the synthetic density and as a second implementation to check the C oracle.

Nothing here is on the product path: the product reads the files these
functions write through its own C++ parsers.
"""
from __future__ import annotations

import json
import math
import os
import struct
import zlib

import numpy as np

NERF_GRIDSIZE = 128
STOCK_CONFIG = {
    "loss": {"otype": "Huber"},
    "optimizer": {"otype": "Ema", "decay": 0.95, "nested": {"otype": "Adam"}},
    "encoding": {
        "otype": "HashGrid",
        "n_levels": 16,
        "n_features_per_level": 2,
        "log2_hashmap_size": 19,
        "base_resolution": 16,
    },
    "network": {
        "otype": "FullyFusedMLP",
        "activation": "ReLU",
        "output_activation": "None",
        "n_neurons": 64,
        "n_hidden_layers": 1,
    },
    "dir_encoding": {
        "otype": "Composite",
        "nested": [
            {"n_dims_to_encode": 3, "otype": "SphericalHarmonics", "degree": 4},
            {"otype": "Identity", "n_bins": 4, "degree": 4},
        ],
    },
    "rgb_network": {
        "otype": "FullyFusedMLP",
        "activation": "ReLU",
        "output_activation": "None",
        "n_neurons": 64,
        "n_hidden_layers": 2,
    },
}


# ----------------------------------------------------------------------------
# hash-grid geometry (T/.../grid.h:196-203, 985-1016; S/ngp/testbed.cu:1197-1204)
# ----------------------------------------------------------------------------
def per_level_scale(n_levels: int, base_resolution: int, aabb_scale: int = 1) -> float:
    """std::exp(std::log(2048*aabb_scale/base)/(L-1)) evaluated in float32."""
    v = np.float32(2048.0) * np.float32(aabb_scale) / np.float32(base_resolution)
    return float(np.exp(np.log(v, dtype=np.float32) / np.float32(n_levels - 1), dtype=np.float32))


def grid_scale(level: int, log2_pls: np.float32, base_resolution: int) -> np.float32:
    return np.float32(np.exp2(np.float32(level) * log2_pls, dtype=np.float32) * np.float32(base_resolution) - np.float32(1.0))


def grid_resolution(scale: np.float32) -> int:
    return int(np.ceil(scale)) + 1


def level_table(n_levels: int, base_resolution: int, pls: float, log2_hashmap_size: int):
    """-> (offsets[n_levels+1] in entries, scales f32[n_levels], resolutions[n_levels])."""
    log2_pls = np.float32(np.log2(np.float32(pls), dtype=np.float32))
    offsets = [0]
    scales, ress = [], []
    for lvl in range(n_levels):
        s = grid_scale(lvl, log2_pls, base_resolution)
        r = grid_resolution(s)
        max_params = (2**32 - 1) // 2
        dense = r**3
        n = max_params if float(np.float32(r) ** 3) > float(max_params) else dense
        n = (n + 7) // 8 * 8
        n = min(n, 1 << log2_hashmap_size)
        offsets.append(offsets[-1] + n)
        scales.append(s)
        ress.append(r)
    return np.array(offsets, dtype=np.uint32), np.array(scales, dtype=np.float32), np.array(ress, dtype=np.uint32)


def level_indexing(res: int, size: int):
    """grid_index()'s dense/hash decision with its uint32 stride arithmetic (T/.../grid.h:164-186).

    -> (dense?, stride_y, stride_z).  The stride loop stops adding dimensions once stride > size;
    the level is hashed iff size < final stride.  The final stride wraps in uint32 (res=2048 with
    log2_hashmap_size >= 22 wraps to 0 and therefore indexes *densely*, modulo size) - kept as is."""
    stride, strides = 1, []
    for _ in range(3):
        if stride > size:
            break
        strides.append(stride)
        stride = (stride * res) & 0xFFFFFFFF
    dense = not (size < stride)
    while len(strides) < 3:
        strides.append(0)
    return dense, strides[1], strides[2]


def morton3d(x, y, z):
    def expand(v):
        v = np.asarray(v, dtype=np.uint64)
        v = (v * np.uint64(0x00010001)) & np.uint64(0xFF0000FF)
        v = (v * np.uint64(0x00000101)) & np.uint64(0x0F00F00F)
        v = (v * np.uint64(0x00000011)) & np.uint64(0xC30C30C3)
        v = (v * np.uint64(0x00000005)) & np.uint64(0x49249249)
        return v
    return (expand(x) | (expand(y) << np.uint64(1)) | (expand(z) << np.uint64(2))).astype(np.uint32)


# ----------------------------------------------------------------------------
# numpy restatement of the network (fp32 accumulate, fp16 storage points)
# ----------------------------------------------------------------------------
class NetParams:
    """Views into one flat fp16 parameter vector, in params_binary order."""

    def __init__(self, params: np.ndarray, n_levels=16, log2_hashmap_size=19, base_resolution=16,
                 pls: float | None = None, density_hidden=1, rgb_hidden=2, width=64, hash_type="CoherentPrime"):
        assert params.dtype == np.float16
        self.n_levels, self.log2T, self.base = n_levels, log2_hashmap_size, base_resolution
        self.pls = per_level_scale(n_levels, base_resolution) if pls is None else pls
        self.hash_type = hash_type
        self.offsets, self.scales, self.ress = level_table(n_levels, base_resolution, self.pls, log2_hashmap_size)
        enc_w = n_levels * 2
        o = 0
        self.density = []
        shapes = [(width, enc_w)] + [(width, width)] * (density_hidden - 1) + [(16, width)]
        for s in shapes:
            self.density.append(params[o:o + s[0] * s[1]].reshape(s)); o += s[0] * s[1]
        self.rgb = []
        shapes = [(width, 32)] + [(width, width)] * (rgb_hidden - 1) + [(16, width)]
        for s in shapes:
            self.rgb.append(params[o:o + s[0] * s[1]].reshape(s)); o += s[0] * s[1]
        self.mlp_params = o
        n_grid = int(self.offsets[-1]) * 2
        self.grid = params[o:o + n_grid].reshape(-1, 2)
        self.n_params = o + n_grid
        assert params.size == self.n_params, (params.size, self.n_params)


def n_params_for(n_levels=16, log2_hashmap_size=19, base_resolution=16, density_hidden=1, rgb_hidden=2, width=64, aabb_scale=1):
    offs, _, _ = level_table(n_levels, base_resolution, per_level_scale(n_levels, base_resolution, aabb_scale), log2_hashmap_size)
    mlp = width * n_levels * 2 + (density_hidden - 1) * width * width + 16 * width
    mlp += width * 32 + (rgb_hidden - 1) * width * width + 16 * width
    return mlp + int(offs[-1]) * 2


def np_encode(net: NetParams, pos: np.ndarray) -> np.ndarray:
    """pos f32[n,3] in [0,1] -> fp16[n, 2L]; fp16 accumulation in corner order 0..7
    exactly like T/.../grid.h:317-343 (result += (half)(weight*data))."""
    pos = np.asarray(pos, dtype=np.float32)
    n = pos.shape[0]
    out = np.zeros((n, net.n_levels * 2), dtype=np.float16)
    primes = (np.uint32(1 if net.hash_type == "CoherentPrime" else 1958374283), np.uint32(2654435761), np.uint32(805459861))
    for lvl in range(net.n_levels):
        scale = net.scales[lvl]
        res = np.uint32(net.ress[lvl])
        size = np.uint32(net.offsets[lvl + 1] - net.offsets[lvl])
        p = pos * scale + np.float32(0.5)
        fl = np.floor(p)
        frac = (p - fl).astype(np.float32)
        g = fl.astype(np.int64).astype(np.uint32)
        acc = np.zeros((n, 2), dtype=np.float16)
        dense, st1, st2 = level_indexing(int(res), int(size))
        for corner in range(8):
            w = np.ones(n, dtype=np.float32)
            c = np.empty((n, 3), dtype=np.uint32)
            for d in range(3):
                if corner & (1 << d):
                    w = w * frac[:, d]
                    c[:, d] = g[:, d] + np.uint32(1)
                else:
                    w = w * (np.float32(1) - frac[:, d])
                    c[:, d] = g[:, d]
            if dense:
                idx = c[:, 0] + c[:, 1] * np.uint32(st1) + c[:, 2] * np.uint32(st2)
            else:
                idx = (c[:, 0] * primes[0]) ^ (c[:, 1] * primes[1]) ^ (c[:, 2] * primes[2])
            idx = idx % size
            val = net.grid[int(net.offsets[lvl]) + idx.astype(np.int64)].astype(np.float32)
            acc = (acc + (w[:, None] * val).astype(np.float16)).astype(np.float16)
        out[:, 2 * lvl:2 * lvl + 2] = acc
    return out


def np_sh4(d01: np.ndarray) -> np.ndarray:
    """dir in [0,1]^3 -> 16 SH coefficients as fp16 (T/.../spherical_harmonics.h:65-98)."""
    d = np.asarray(d01, dtype=np.float32) * np.float32(2) - np.float32(1)
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    f = np.float32
    o = np.stack([
        np.full_like(x, f(0.28209479177387814)),
        f(-0.48860251190291987) * y, f(0.48860251190291987) * z, f(-0.48860251190291987) * x,
        f(1.0925484305920792) * xy, f(-1.0925484305920792) * yz,
        f(0.94617469575755997) * z2 - f(0.31539156525251999),
        f(-1.0925484305920792) * xz,
        f(0.54627421529603959) * x2 - f(0.54627421529603959) * y2,
        f(0.59004358992664352) * y * (f(-3.0) * x2 + y2),
        f(2.8906114426405538) * xy * z,
        f(0.45704579946446572) * y * (f(1.0) - f(5.0) * z2),
        f(0.3731763325901154) * z * (f(5.0) * z2 - f(3.0)),
        f(0.45704579946446572) * x * (f(1.0) - f(5.0) * z2),
        f(1.4453057213202769) * z * (x2 - y2),
        f(0.59004358992664352) * x * (-x2 + f(3.0) * y2),
    ], axis=1)
    return o.astype(np.float16)


def np_mlp(layers, x16: np.ndarray) -> np.ndarray:
    """y = W x per layer, ReLU on hidden layers, fp16 storage between layers, fp32 accumulate."""
    h = x16.astype(np.float32)
    for i, w in enumerate(layers):
        h = h @ w.astype(np.float32).T
        if i + 1 < len(layers):
            h = np.maximum(h, 0)
        h = h.astype(np.float16).astype(np.float32)
    return h.astype(np.float16)


def np_network(net: NetParams, pos: np.ndarray, dir01: np.ndarray) -> np.ndarray:
    """-> fp16[n,4] = (r,g,b raw, sigma raw) as NerfNetwork::inference_mixed_precision_impl."""
    enc = np_encode(net, pos)
    dens = np_mlp(net.density, enc)
    rgb_in = np.concatenate([dens, np_sh4(dir01)], axis=1)
    rgb = np_mlp(net.rgb, rgb_in)
    return np.concatenate([rgb[:, :3], dens[:, :1]], axis=1)


# ----------------------------------------------------------------------------
# snapshot writer
# ----------------------------------------------------------------------------
HEAD_CENTER = (0.5, 0.5, 0.5)
HEAD_AXES = (0.18, 0.24, 0.20)
CROP_MIN = (-0.2, 0.15, -0.2)   # V/render.py:234-235
CROP_MAX = (1.0, 1.0, 1.0)


def make_density_grid(rng: np.random.Generator, n_floaters: int = 64) -> np.ndarray:
    """fp16[128^3] in Morton order: 1.0 inside a head-like ellipsoid plus seeded floaters."""
    g = np.arange(NERF_GRIDSIZE, dtype=np.float32)
    c = (g + 0.5) / NERF_GRIDSIZE
    X, Y, Z = np.meshgrid(c, c, c, indexing="ij")
    inside = (((X - HEAD_CENTER[0]) / HEAD_AXES[0]) ** 2 + ((Y - HEAD_CENTER[1]) / HEAD_AXES[1]) ** 2
              + ((Z - HEAD_CENTER[2]) / HEAD_AXES[2]) ** 2) <= 1.0
    occ = inside.copy()
    for _ in range(n_floaters):
        p = rng.integers(8, NERF_GRIDSIZE - 8, size=3)
        s = rng.integers(1, 4, size=3)
        occ[p[0]:p[0] + s[0], p[1]:p[1] + s[1], p[2]:p[2] + s[2]] = True
    xs, ys, zs = np.nonzero(occ)
    flat = np.zeros(NERF_GRIDSIZE ** 3, dtype=np.float16)
    flat[morton3d(xs, ys, zs)] = np.float16(1.0)
    return flat


def make_params(rng: np.random.Generator, n_levels=16, log2_hashmap_size=19, base_resolution=16,
                regime: str = "opaque", calibrate: bool = True, aabb_scale: int = 1) -> np.ndarray:
    """Random-init fp16 parameter vector in params_binary order.

    regime "opaque": density output row made non-negative and scaled so the median raw sigma
    inside the head ellipsoid is ~ +6 (rays saturate in ~5-10 samples);
    regime "translucent": median raw sigma ~ 0 (full-length marches, worst case)."""
    n = n_params_for(n_levels, log2_hashmap_size, base_resolution, aabb_scale=aabb_scale)
    calibrate = calibrate and aabb_scale == 1      # (the calibration helper assumes the unit-cube level table)
    p = np.empty(n, dtype=np.float16)
    o = 0
    shapes = [(64, n_levels * 2), (16, 64), (64, 32), (64, 64), (16, 64)]
    views = []
    for (r, c) in shapes:
        w = rng.normal(0.0, math.sqrt(2.0 / c), size=(r, c)).astype(np.float32)
        views.append((o, r, c))
        p[o:o + r * c] = w.astype(np.float16).ravel(); o += r * c
    n_grid = n - o
    # chunked to bound peak memory for log2T = 24
    step = 1 << 24
    for s in range(0, n_grid, step):
        e = min(n_grid, s + step)
        p[o + s:o + e] = (rng.uniform(-1.0, 1.0, size=e - s).astype(np.float32) * np.float32(0.1)).astype(np.float16)
    if calibrate:
        net = NetParams(p, n_levels, log2_hashmap_size, base_resolution)
        do, dr, dc = views[1]
        w_out = p[do:do + dr * dc].reshape(dr, dc)
        if regime == "opaque":
            w_out[0] = np.abs(w_out[0])
        pts = rng.uniform(-1, 1, size=(4096, 3))
        pts = pts[(pts ** 2).sum(1) <= 1.0][:1024].astype(np.float32)
        pts = pts * np.array(HEAD_AXES, dtype=np.float32) + np.array(HEAD_CENTER, dtype=np.float32)
        enc = np_encode(net, pts)
        hid = np.maximum(enc.astype(np.float32) @ net.density[0].astype(np.float32).T, 0).astype(np.float16).astype(np.float32)
        raw = hid @ w_out[0].astype(np.float32)
        if regime == "opaque":
            med = float(np.median(raw))
            w_out[0] = (w_out[0].astype(np.float32) * (6.0 / max(med, 1e-6))).astype(np.float16)
        else:
            w_out[0] = (w_out[0].astype(np.float32) * (1.0 / max(float(np.std(raw)), 1e-6))).astype(np.float16)
    return p


def snapshot_dict(params: np.ndarray, density_grid: np.ndarray, n_levels=16, log2_hashmap_size=19,
                  base_resolution=16, render_aabb=(CROP_MIN, CROP_MAX), aabb_scale=1) -> dict:
    cfg = json.loads(json.dumps(STOCK_CONFIG))
    cfg["encoding"].update(n_levels=n_levels, log2_hashmap_size=log2_hashmap_size, base_resolution=base_resolution)
    half = 0.5 * aabb_scale
    ra_min = [max(float(a), 0.5 - half) for a in render_aabb[0]]
    ra_max = [min(float(a), 0.5 + half) for a in render_aabb[1]]
    ident = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]
    xf = [[1.0, 0.0, 0.0, 0.5], [0.0, 1.0, 0.0, 0.5], [0.0, 0.0, 1.0, 2.5]]
    cfg["snapshot"] = {
        "version": 1,
        "aabb": {"min": [0.5 - half] * 3, "max": [0.5 + half] * 3},
        "bounding_radius": 1.0,
        "density_grid_size": NERF_GRIDSIZE,
        "density_grid_binary": density_grid.astype(np.float16).tobytes(),
        "nerf": {
            "rgb": {"rays_per_batch": 4096, "measured_batch_size": 262144, "measured_batch_size_before_compaction": 1048576},
            "dataset": {
                "n_images": 1,
                "paths": [""],
                "xforms": [{"start": xf, "end": xf}],
                "metadata": [{"resolution": [800, 800], "focal_length": [1118.484365425203, 1118.484365425203],
                              "principal_point": [0.5, 0.5], "rolling_shutter": [0.0, 0.0, 0.0, 0.0], "lens": {}}],
                "render_aabb": {"min": ra_min, "max": ra_max},
                "render_aabb_to_local": ident,
                "up": [0.0, 1.0, 0.0],
                "offset": [0.5, 0.5, 0.5],
                "envmap_resolution": [0, 0],
                "scale": 0.33,
                "aabb_scale": aabb_scale,
                "from_mitsuba": False,
                "is_hdr": False,
                "wants_importance_sampling": True,
            },
        },
        "render_aabb": {"min": ra_min, "max": ra_max},
        "render_aabb_to_local": ident,
        "training_step": 10000,
        "loss": 0.00175,
        "n_params": int(params.size),
        "params_type": "__half",
        "params_binary": params.astype(np.float16).tobytes(),
    }
    return cfg


def write_snapshot(path: str, seed: int = 1337, n_levels=16, log2_hashmap_size=19, base_resolution=16,
                   regime: str = "opaque", n_floaters: int = 64, calibrate: bool = True, aabb_scale: int = 1) -> dict:
    """Write a synthetic snapshot; returns {"params": fp16[], "density_grid": fp16[], "config": dict-without-binaries}.
    aabb_scale > 1: log2(aabb_scale) + 1 occupancy cascades (a few blobs outside the unit cube in the coarser ones), a render
    box of that size and the cone-angle step growth of the reference (S/ngp/testbed.cu:1098-1115)."""
    import msgpack
    rng = np.random.default_rng(seed)
    grid = make_density_grid(rng, n_floaters)
    if aabb_scale > 1:
        n_casc = int(np.log2(aabb_scale)) + 1
        full = np.zeros(NERF_GRIDSIZE ** 3 * n_casc, dtype=np.float16)
        full[:NERF_GRIDSIZE ** 3] = grid
        for c in range(1, n_casc):       # blobs outside the centre octant (which the loader fills by pooling the finer cascade)
            occ = np.zeros((NERF_GRIDSIZE,) * 3, dtype=bool)
            for _ in range(6):
                p = rng.integers(4, NERF_GRIDSIZE - 12, size=3)
                if np.all((p > 24) & (p < 100)):
                    p[int(rng.integers(0, 3))] = int(rng.choice([6, 108]))
                s3 = rng.integers(3, 9, size=3)
                occ[p[0]:p[0] + s3[0], p[1]:p[1] + s3[1], p[2]:p[2] + s3[2]] = True
            xs, ys, zs = np.nonzero(occ)
            full[c * NERF_GRIDSIZE ** 3 + morton3d(xs, ys, zs)] = np.float16(1.0)
        grid = full
    params = make_params(rng, n_levels, log2_hashmap_size, base_resolution, regime, calibrate, aabb_scale)
    half = 0.5 * aabb_scale
    d = snapshot_dict(params, grid, n_levels, log2_hashmap_size, base_resolution,
                      render_aabb=(CROP_MIN, CROP_MAX) if aabb_scale == 1 else ([0.5 - half] * 3, [0.5 + half] * 3), aabb_scale=aabb_scale)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(msgpack.packb(d, use_bin_type=True))
    return {"params": params, "density_grid": grid, "n_levels": n_levels, "log2_hashmap_size": log2_hashmap_size,
            "base_resolution": base_resolution}


def read_snapshot(path: str) -> dict:
    """Parse a snapshot with the python msgpack package (test-side loader, independent of the C++ reader)."""
    import msgpack
    with open(path, "rb") as f:
        d = msgpack.unpackb(f.read(), raw=False, strict_map_key=False)
    s = d["snapshot"]
    enc = d["encoding"]
    dt = np.float16 if s.get("params_type", "__half") == "__half" else np.float32
    params = np.frombuffer(s["params_binary"], dtype=dt).astype(np.float16)
    grid = np.frombuffer(s["density_grid_binary"], dtype=np.float16)
    ds = s["nerf"].get("dataset", {})
    aabb_scale = int(ds.get("aabb_scale", s["nerf"].get("aabb_scale", 1)))
    ra = s.get("render_aabb", ds.get("render_aabb"))
    return {
        "config": d, "params": params, "density_grid": grid, "aabb_scale": aabb_scale,
        "n_levels": int(enc.get("n_levels", 16)), "log2_hashmap_size": int(enc.get("log2_hashmap_size", 19)),
        "base_resolution": int(enc.get("base_resolution", 16)),
        "per_level_scale": float(enc.get("per_level_scale", 0.0)),
        "hash": enc.get("hash", "CoherentPrime"),
        "render_aabb_min": np.array(ra["min"], dtype=np.float32), "render_aabb_max": np.array(ra["max"], dtype=np.float32),
        "is_hdr": bool(ds.get("is_hdr", False)),
    }


# ----------------------------------------------------------------------------
# glasses mesh fixture -> .gltf/.bin/.png on disk
# ----------------------------------------------------------------------------
def _png_bytes(rgba: np.ndarray) -> bytes:
    h, w, _ = rgba.shape
    raw = b"".join(b"\x00" + rgba[y].astype(np.uint8).tobytes() for y in range(h))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0))
            + chunk(b"IDAT", zlib.compress(raw, 9)) + chunk(b"IEND", b""))


GLASSES_NPZ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "glasses_mesh.npz")
# fixed placement replacing the MediaPipe-derived transform
GLASSES_T = (0.0, 0.03, 0.215)   # in front of the synthetic head ellipsoid (semi-axis z = 0.20)
GLASSES_S = (0.17, 0.17, 0.17)
GLASSES_R_WXYZ = (0.7071068, 0.7071067, 0.0, 0.0)


def write_glasses_gltf(out_dir: str, texture_rgba=(128, 128, 128, 255), npz_path: str = GLASSES_NPZ) -> str:
    """Write glasses.gltf + glasses.bin + glasses.png from the committed geometry fixture.

    The fixture holds the reference asset's accessors verbatim (positions/normals/uv f32, indices u16;
    R/assets/meshes/glasses/glasses.gltf); the texture is a constant stand-in because the
    reference's glasses.png is a git-LFS pointer."""
    m = np.load(npz_path)
    pos, nrm, uv, idx = m["positions"], m["normals"], m["texcoords"], m["indices"]
    os.makedirs(out_dir, exist_ok=True)
    blob = pos.astype("<f4").tobytes() + nrm.astype("<f4").tobytes() + uv.astype("<f4").tobytes() + idx.astype("<u2").tobytes()
    with open(os.path.join(out_dir, "glasses.bin"), "wb") as f:
        f.write(blob)
    tex = np.tile(np.array(texture_rgba, dtype=np.uint8), (4, 4, 1))
    with open(os.path.join(out_dir, "glasses.png"), "wb") as f:
        f.write(_png_bytes(tex))
    o1, o2, o3 = pos.nbytes, pos.nbytes + nrm.nbytes, pos.nbytes + nrm.nbytes + uv.nbytes
    doc = {
        "asset": {"generator": "nmr-b200 fixture writer", "version": "2.0"},
        "scene": 0,
        "scenes": [{"name": "Scene", "nodes": [0]}],
        "nodes": [{"mesh": 0, "name": "Glasses.001",
                   "rotation": [float(x) for x in m["node_rotation_xyzw"]],
                   "translation": [float(x) for x in m["node_translation"]]}],
        "materials": [{"doubleSided": True, "name": "oculos",
                       "pbrMetallicRoughness": {"baseColorTexture": {"index": 0}, "metallicFactor": 0,
                                                "roughnessFactor": float(m["roughness"])}}],
        "meshes": [{"name": "Glasses.001", "primitives": [
            {"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}, "indices": 3, "material": 0}]}],
        "textures": [{"sampler": 0, "source": 0}],
        "images": [{"mimeType": "image/png", "name": "glasses", "uri": "glasses.png"}],
        "accessors": [
            {"bufferView": 0, "componentType": 5126, "count": int(pos.shape[0]), "type": "VEC3",
             "max": [float(x) for x in pos.max(0)], "min": [float(x) for x in pos.min(0)]},
            {"bufferView": 1, "componentType": 5126, "count": int(nrm.shape[0]), "type": "VEC3"},
            {"bufferView": 2, "componentType": 5126, "count": int(uv.shape[0]), "type": "VEC2"},
            {"bufferView": 3, "componentType": 5123, "count": int(idx.shape[0]), "type": "SCALAR"},
        ],
        "bufferViews": [
            {"buffer": 0, "byteLength": pos.nbytes, "byteOffset": 0, "target": 34962},
            {"buffer": 0, "byteLength": nrm.nbytes, "byteOffset": o1, "target": 34962},
            {"buffer": 0, "byteLength": uv.nbytes, "byteOffset": o2, "target": 34962},
            {"buffer": 0, "byteLength": idx.nbytes, "byteOffset": o3, "target": 34963},
        ],
        "samplers": [{"magFilter": 9729, "minFilter": 9987}],
        "buffers": [{"byteLength": len(blob), "uri": "glasses.bin"}],
    }
    path = os.path.join(out_dir, "glasses.gltf")
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
    return path


LENS_IOR, LENS_TRANSMISSION, LENS_TINT = 1.5, 0.85, (0.75, 0.9, 1.0)


def write_lens_glasses_gltf(out_dir: str, texture_rgba=(128, 128, 128, 255), npz_path: str = GLASSES_NPZ,
                            ior: float = LENS_IOR, transmission: float = LENS_TRANSMISSION, tint=LENS_TINT) -> str:
    """glasses.gltf with a second primitive: two lens panes inside the rims (two-sided quads in the plane of the front frame),
    whose material is transmissive (KHR_materials_transmission + KHR_materials_ior).  The reference asset carries no such
    material (one opaque primitive); this is the fixture of the lens / secondary-ray path."""
    m = np.load(npz_path)
    pos, nrm, uv, idx = m["positions"], m["normals"], m["texcoords"], m["indices"]
    lp, ln, li = [], [], []
    y0 = -0.03
    for (xa, xb) in ((-0.68, -0.12), (0.12, 0.68)):
        za, zb = -0.12, 0.28
        quad = [(xa, y0, za), (xb, y0, za), (xb, y0, zb), (xa, y0, zb)]
        b = len(lp)
        lp += quad; ln += [(0.0, 1.0, 0.0)] * 4; li += [b, b + 2, b + 1, b, b + 3, b + 2]          # front face (+y)
        b = len(lp)
        lp += quad; ln += [(0.0, -1.0, 0.0)] * 4; li += [b, b + 1, b + 2, b, b + 2, b + 3]         # back face (-y)
    lp = np.array(lp, dtype=np.float32); ln = np.array(ln, dtype=np.float32); luv = np.zeros((len(lp), 2), dtype=np.float32)
    li = np.array(li, dtype=np.uint16)
    os.makedirs(out_dir, exist_ok=True)
    parts = [pos.astype("<f4").tobytes(), nrm.astype("<f4").tobytes(), uv.astype("<f4").tobytes(), idx.astype("<u2").tobytes()]
    if len(parts[3]) % 4:
        parts[3] += b"\0" * (4 - len(parts[3]) % 4)
    parts += [lp.astype("<f4").tobytes(), ln.astype("<f4").tobytes(), luv.astype("<f4").tobytes(), li.astype("<u2").tobytes()]
    offs = np.cumsum([0] + [len(p) for p in parts])
    with open(os.path.join(out_dir, "glasses.bin"), "wb") as f:
        f.write(b"".join(parts))
    tex = np.tile(np.array(texture_rgba, dtype=np.uint8), (4, 4, 1))
    with open(os.path.join(out_dir, "glasses.png"), "wb") as f:
        f.write(_png_bytes(tex))
    counts = [pos.shape[0], nrm.shape[0], uv.shape[0], idx.shape[0], lp.shape[0], ln.shape[0], luv.shape[0], li.shape[0]]
    types = ["VEC3", "VEC3", "VEC2", "SCALAR"] * 2
    comp = [5126, 5126, 5126, 5123] * 2
    doc = {
        "asset": {"generator": "nmr-b200 fixture writer", "version": "2.0"},
        "extensionsUsed": ["KHR_materials_transmission", "KHR_materials_ior"],
        "scene": 0,
        "scenes": [{"name": "Scene", "nodes": [0]}],
        "nodes": [{"mesh": 0, "name": "Glasses.001", "rotation": [float(x) for x in m["node_rotation_xyzw"]],
                   "translation": [float(x) for x in m["node_translation"]]}],
        "materials": [{"doubleSided": True, "name": "oculos",
                       "pbrMetallicRoughness": {"baseColorTexture": {"index": 0}, "metallicFactor": 0, "roughnessFactor": float(m["roughness"])}},
                      {"doubleSided": True, "name": "lens", "alphaMode": "BLEND",
                       "pbrMetallicRoughness": {"baseColorFactor": [float(tint[0]), float(tint[1]), float(tint[2]), 1.0], "metallicFactor": 0, "roughnessFactor": 0.05},
                       "extensions": {"KHR_materials_transmission": {"transmissionFactor": float(transmission)}, "KHR_materials_ior": {"ior": float(ior)}}}],
        "meshes": [{"name": "Glasses.001", "primitives": [
            {"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}, "indices": 3, "material": 0},
            {"attributes": {"POSITION": 4, "NORMAL": 5, "TEXCOORD_0": 6}, "indices": 7, "material": 1}]}],
        "textures": [{"sampler": 0, "source": 0}],
        "images": [{"mimeType": "image/png", "name": "glasses", "uri": "glasses.png"}],
        "accessors": [dict({"bufferView": i, "componentType": comp[i], "count": int(counts[i]), "type": types[i]},
                           **({"max": [float(x) for x in (pos if i == 0 else lp).max(0)], "min": [float(x) for x in (pos if i == 0 else lp).min(0)]} if i in (0, 4) else {}))
                      for i in range(8)],
        "bufferViews": [{"buffer": 0, "byteLength": len(parts[i]) if i != 3 else idx.nbytes, "byteOffset": int(offs[i]), "target": 34963 if i % 4 == 3 else 34962} for i in range(8)],
        "samplers": [{"magFilter": 9729, "minFilter": 9987}],
        "buffers": [{"byteLength": int(offs[-1]), "uri": "glasses.bin"}],
    }
    path = os.path.join(out_dir, "glasses.gltf")
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
    return path


def np_tangents(pos: np.ndarray, nrm: np.ndarray, uv: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """Per-vertex tangents (xyz + handedness) from the UV derivatives: per-triangle tangent / bitangent accumulated per vertex, then
    Gram-Schmidt against the normal.  What the textured fixture writes into its TANGENT attribute (any sane tangent field would do:
    a file's tangents are taken as they are).  Not what the loader generates for a file WITHOUT the attribute - that is the reference's
    MikkTSpace generator, csrc/mikk.cpp."""
    pos = pos.astype(np.float32); nrm = nrm.astype(np.float32); uv = uv.astype(np.float32)
    tri = idx.reshape(-1, 3).astype(np.int64)
    tacc = np.zeros_like(pos); bacc = np.zeros_like(pos)
    for i0, i1, i2 in tri:
        e1, e2 = pos[i1] - pos[i0], pos[i2] - pos[i0]
        du1, dv1 = uv[i1] - uv[i0]; du2, dv2 = uv[i2] - uv[i0]
        det = np.float32(du1 * dv2 - du2 * dv1)
        if not abs(det) > 1e-20:
            continue
        r = np.float32(1.0) / det
        t = (e1 * dv2 - e2 * dv1) * r; b = (e2 * du1 - e1 * du2) * r
        for v in (i0, i1, i2):
            tacc[v] += t; bacc[v] += b
    out = np.zeros((pos.shape[0], 4), dtype=np.float32)
    for v in range(pos.shape[0]):
        n = nrm[v]; nl = np.sqrt(np.float32(n @ n))
        nn = n / nl if nl > 0 else np.array([0, 0, 1], np.float32)
        tv = tacc[v] - nn * np.float32(nn @ tacc[v]); tl = np.sqrt(np.float32(tv @ tv))
        if not tl > 1e-20:
            ax = np.array([1, 0, 0], np.float32) if abs(nn[0]) < 0.9 else np.array([0, 1, 0], np.float32)
            tv = ax - nn * np.float32(nn @ ax); tl = np.sqrt(np.float32(tv @ tv))
        t = tv / tl
        out[v, :3] = t
        out[v, 3] = -1.0 if np.cross(nn, t) @ bacc[v] < 0 else 1.0
    return out


def _test_texture(kind: str, n: int = 16) -> np.ndarray:
    """Small procedural RGBA8 textures of the textured fixture (deterministic)."""
    y, x = np.mgrid[0:n, 0:n].astype(np.float32) / (n - 1)
    if kind == "base":
        rgb = np.stack([0.35 + 0.6 * x, 0.3 + 0.5 * y, 0.8 - 0.5 * x * y], -1)
    elif kind == "emissive":
        rgb = np.stack([0.4 * (1 - x), 0.15 + 0.2 * y, 0.5 * x * (1 - y)], -1)
    elif kind == "metallic_roughness":
        rgb = np.stack([np.zeros_like(x), 0.25 + 0.7 * y, 0.2 + 0.75 * x], -1)          # G roughness, B metallic
    elif kind == "normal":
        nx, ny = 0.45 * np.sin(6.3 * x), 0.45 * np.cos(5.1 * y)
        nz = np.sqrt(np.maximum(0.0, 1 - nx * nx - ny * ny))
        rgb = np.stack([nx, ny, nz], -1) * 0.5 + 0.5
    else:                                                                                # occlusion
        rgb = np.stack([0.35 + 0.65 * (0.5 + 0.5 * np.sin(9 * x + 4 * y))] * 3, -1)
    out = np.empty((n, n, 4), dtype=np.uint8)
    out[..., :3] = np.clip(np.round(rgb * 255), 0, 255).astype(np.uint8); out[..., 3] = 255
    return out


def write_textured_glasses_gltf(out_dir: str, with_tangents: bool = True, npz_path: str = GLASSES_NPZ) -> str:
    """The glasses geometry with a material that uses every texture slot of the reference's closest-hit program (base colour,
    emissive, metallic-roughness, normal with scale, occlusion with strength; S/optix/optix_scene.cu:221-258).  The reference asset
    carries only a base colour texture; this fixture exercises the rest of the shader.  with_tangents: write a TANGENT attribute
    (np_tangents) so that every reader works from the same tangents; False leaves their generation to the loader."""
    m = np.load(npz_path)
    pos, nrm, uv, idx = m["positions"], m["normals"], m["texcoords"], m["indices"]
    os.makedirs(out_dir, exist_ok=True)
    parts = [pos.astype("<f4").tobytes(), nrm.astype("<f4").tobytes(), uv.astype("<f4").tobytes(), idx.astype("<u2").tobytes()]
    if len(parts[3]) % 4:
        parts[3] += b"\0" * (4 - len(parts[3]) % 4)
    if with_tangents:
        parts.append(np_tangents(pos, nrm, uv, idx).astype("<f4").tobytes())
    offs = np.cumsum([0] + [len(p) for p in parts])
    with open(os.path.join(out_dir, "glasses.bin"), "wb") as f:
        f.write(b"".join(parts))
    kinds = ["base", "emissive", "metallic_roughness", "normal", "occlusion"]
    for k in kinds:
        with open(os.path.join(out_dir, f"{k}.png"), "wb") as f:
            f.write(_png_bytes(_test_texture(k)))
    attrs = {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}
    accessors = [
        {"bufferView": 0, "componentType": 5126, "count": int(pos.shape[0]), "type": "VEC3", "max": [float(x) for x in pos.max(0)], "min": [float(x) for x in pos.min(0)]},
        {"bufferView": 1, "componentType": 5126, "count": int(nrm.shape[0]), "type": "VEC3"},
        {"bufferView": 2, "componentType": 5126, "count": int(uv.shape[0]), "type": "VEC2"},
        {"bufferView": 3, "componentType": 5123, "count": int(idx.shape[0]), "type": "SCALAR"}]
    views = [{"buffer": 0, "byteLength": len(parts[i]) if i != 3 else idx.nbytes, "byteOffset": int(offs[i]), "target": 34963 if i == 3 else 34962} for i in range(len(parts))]
    if with_tangents:
        attrs["TANGENT"] = 4
        accessors.append({"bufferView": 4, "componentType": 5126, "count": int(pos.shape[0]), "type": "VEC4"})
    doc = {
        "asset": {"generator": "nmr-b200 fixture writer", "version": "2.0"}, "scene": 0, "scenes": [{"name": "Scene", "nodes": [0]}],
        "nodes": [{"mesh": 0, "name": "Glasses.001", "rotation": [float(x) for x in m["node_rotation_xyzw"]], "translation": [float(x) for x in m["node_translation"]]}],
        "materials": [{"doubleSided": True, "name": "textured",
                       "emissiveFactor": [0.8, 0.9, 1.0], "emissiveTexture": {"index": 1},
                       "normalTexture": {"index": 3, "scale": 0.8}, "occlusionTexture": {"index": 4, "strength": 0.7},
                       "pbrMetallicRoughness": {"baseColorFactor": [0.9, 0.95, 1.0, 1.0], "baseColorTexture": {"index": 0}, "metallicFactor": 0.8,
                                                "roughnessFactor": 0.9, "metallicRoughnessTexture": {"index": 2}}}],
        "meshes": [{"name": "Glasses.001", "primitives": [{"attributes": attrs, "indices": 3, "material": 0}]}],
        "textures": [{"sampler": 0, "source": i} for i in range(5)],
        "images": [{"mimeType": "image/png", "name": k, "uri": f"{k}.png"} for k in kinds],
        "accessors": accessors, "bufferViews": views,
        "samplers": [{"magFilter": 9729, "minFilter": 9987}],
        "buffers": [{"byteLength": int(offs[-1]), "uri": "glasses.bin"}],
    }
    path = os.path.join(out_dir, "glasses.gltf")
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
    return path


def _read_png_rgba8(path: str) -> np.ndarray:
    """Reader for the PNGs _png_bytes writes (8-bit RGBA, filter type 0 on every row)."""
    data = open(path, "rb").read()
    p, idat, w, h = 8, b"", 0, 0
    while p < len(data):
        n = struct.unpack(">I", data[p:p + 4])[0]; tag = data[p + 4:p + 8]; body = data[p + 8:p + 8 + n]
        if tag == b"IHDR":
            w, h = struct.unpack(">II", body[:8])
        elif tag == b"IDAT":
            idat += body
        p += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 4 * w)
    assert (raw[:, 0] == 0).all()
    return raw[:, 1:].reshape(h, w, 4).copy()


def read_gltf(path: str) -> dict:
    """Test-side glTF reader (json + numpy), independent of the C++ loader."""
    with open(path) as f:
        doc = json.load(f)
    base = os.path.dirname(os.path.abspath(path))
    bufs = [open(os.path.join(base, b["uri"]), "rb").read() for b in doc["buffers"]]

    def acc(i):
        a = doc["accessors"][i]
        bv = doc["bufferViews"][a["bufferView"]]
        ncomp = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4}[a["type"]]
        dt = {5126: "<f4", 5123: "<u2", 5125: "<u4", 5121: "u1"}[a["componentType"]]
        off = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
        arr = np.frombuffer(bufs[bv["buffer"]], dtype=dt, count=a["count"] * ncomp, offset=off)
        return arr.reshape(a["count"], ncomp) if ncomp > 1 else arr
    node = doc["nodes"][doc["scenes"][doc.get("scene", 0)]["nodes"][0]]
    prims = doc["meshes"][node["mesh"]]["primitives"]
    mat0 = doc["materials"][prims[0]["material"]]
    mat = mat0["pbrMetallicRoughness"]
    P, N, T, I, L = [], [], [], [], []
    TG = []
    lens = None
    base = 0
    for prim in prims:      # concatenated like the C++ loader; material of the first primitive; lens flags per triangle
        p = acc(prim["attributes"]["POSITION"]).astype(np.float32)
        P.append(p); N.append(acc(prim["attributes"]["NORMAL"]).astype(np.float32)); T.append(acc(prim["attributes"]["TEXCOORD_0"]).astype(np.float32))
        i = acc(prim["indices"]).astype(np.int64) + base
        I.append(i); base += p.shape[0]
        TG.append(acc(prim["attributes"]["TANGENT"]).astype(np.float32) if "TANGENT" in prim["attributes"] else None)
        md = doc["materials"][prim["material"]]
        ext = md.get("extensions", {})
        tf = float(ext.get("KHR_materials_transmission", {}).get("transmissionFactor", 0.0))
        bc = md.get("pbrMetallicRoughness", {}).get("baseColorFactor", [1, 1, 1, 1])
        is_lens = tf > 0.0 or (md.get("alphaMode") == "BLEND" and bc[3] < 1.0)
        if is_lens and tf <= 0.0:
            tf = 1.0 - bc[3]
        L.append(np.full(i.size // 3, 1 if is_lens else 0, dtype=np.uint8))
        if is_lens and lens is None:
            lens = {"ior": float(ext.get("KHR_materials_ior", {}).get("ior", 1.5)), "transmission": tf, "tint": np.array(bc[:3], dtype=np.float32)}
    def tex(ref):
        if ref is None:
            return None
        img = doc["images"][doc["textures"][ref["index"]]["source"]]
        return _read_png_rgba8(os.path.join(os.path.dirname(os.path.abspath(path)), img["uri"]))
    textures = {"emissive": tex(mat0.get("emissiveTexture")), "metallic_roughness": tex(mat.get("metallicRoughnessTexture")),
                "normal": tex(mat0.get("normalTexture")), "occlusion": tex(mat0.get("occlusionTexture"))}
    return {
        "tangents": np.concatenate(TG) if all(t is not None for t in TG) else None,
        "textures": textures, "base_texture_ref": mat.get("baseColorTexture"),
        "normal_scale": float(mat0.get("normalTexture", {}).get("scale", 1.0)), "occlusion_strength": float(mat0.get("occlusionTexture", {}).get("strength", 1.0)),
        "emissive": np.array(mat0.get("emissiveFactor", [0, 0, 0]), dtype=np.float32),
        "positions": np.concatenate(P), "normals": np.concatenate(N), "texcoords": np.concatenate(T),
        "indices": np.concatenate(I).astype(np.uint16), "tri_lens": np.concatenate(L), "lens": lens,
        "base_color": np.array(mat.get("baseColorFactor", [1, 1, 1, 1]), dtype=np.float32),
        "metallic": float(mat.get("metallicFactor", 1.0)), "roughness": float(mat.get("roughnessFactor", 1.0)),
    }


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="write a synthetic iNGP snapshot")
    ap.add_argument("out")
    ap.add_argument("--log2-hashmap-size", type=int, default=19)
    ap.add_argument("--regime", default="opaque")
    ap.add_argument("--seed", type=int, default=1337)
    a = ap.parse_args()
    r = write_snapshot(a.out, a.seed, log2_hashmap_size=a.log2_hashmap_size, regime=a.regime)
    print(a.out, r["params"].size, "params")


# ---- meshes for the tangent generator (nmr_mikk_tangents vs the reference's mikktspace.c) ---------------------------------------

def tangent_test_grid(nu: int, nv: int, rng, mirror: bool = False, jitter: float = 0.0, explode: bool = False):
    """A band of a sphere as an indexed grid (shared vertices, seam columns equal in position but not in uv).  mirror folds the u
    coordinate (both UV orientations in one mesh), jitter roughens it, explode gives every face its own three vertices (the welding
    step has to join them again).  Faces are shuffled.  -> positions, normals, texcoords, indices (n, 3)."""
    u, v = np.meshgrid(np.linspace(0, 1, nu), np.linspace(0, 1, nv), indexing="ij")
    th = u * 2 * np.pi; ph = (v * 0.8 + 0.1) * np.pi
    p = np.stack([np.sin(ph) * np.cos(th), np.cos(ph), np.sin(ph) * np.sin(th)], -1).reshape(-1, 3)
    p = p * (1 + jitter * rng.standard_normal(p.shape))
    n = p / np.linalg.norm(p, axis=1, keepdims=True)
    uv = np.stack([u, v], -1).reshape(-1, 2).copy()
    if mirror:
        uv[:, 0] = np.abs(uv[:, 0] - 0.5) * 2
    a, b = np.meshgrid(np.arange(nu - 1), np.arange(nv - 1), indexing="ij")
    i0 = (a * nv + b).reshape(-1); i1 = ((a + 1) * nv + b).reshape(-1); i2 = i1 + 1; i3 = i0 + 1
    f = np.concatenate([np.stack([i0, i1, i2], 1), np.stack([i0, i2, i3], 1)]).astype(np.uint32)
    f = f[rng.permutation(len(f))]
    p, n, uv = p.astype(np.float32), n.astype(np.float32), uv.astype(np.float32)
    if explode:
        p, n, uv = p[f.reshape(-1)], n[f.reshape(-1)], uv[f.reshape(-1)]
        f = np.arange(len(p), dtype=np.uint32).reshape(-1, 3)
    return p, n, uv, f


def tangent_test_soup(rng, nv: int, nf: int, dup: float = 0.3, degen: float = 0.05, flat: float = 0.05):
    """Random triangles over few vertices: edges shared by many triangles, whole-vertex duplicates (welded), position-only duplicates
    (not welded), triangles with two equal corners (set aside) and triangles without UV area (group with anything)."""
    p = rng.integers(-3, 4, (nv, 3)).astype(np.float32) * 0.25 + (rng.standard_normal((nv, 3)) * 0.01).astype(np.float32)
    n = rng.standard_normal((nv, 3)).astype(np.float32); n /= np.linalg.norm(n, axis=1, keepdims=True)
    uv = rng.random((nv, 2)).astype(np.float32)
    k = int(nv * dup); src = rng.integers(0, nv, k); dst = rng.integers(0, nv, k)
    p[dst] = p[src]; n[dst] = n[src]; uv[dst] = uv[src]
    k = int(nv * 0.1); src = rng.integers(0, nv, k); dst = rng.integers(0, nv, k)
    p[dst] = p[src]
    f = rng.integers(0, nv, (nf, 3)).astype(np.uint32)
    d = rng.random(nf) < degen; f[d, 1] = f[d, 0]
    for t in np.nonzero(rng.random(nf) < flat)[0]:
        uv[f[t, 2]] = uv[f[t, 1]]
    return p, n, uv, f


def tangent_test_cases(seed: int = 20260102, n_soup: int = 24, n_grid: int = 8):
    """The fixed list of meshes behind tests/golden/ref_mikk.npz."""
    rng = np.random.default_rng(seed)
    cases = []
    for s in range(n_soup):
        cases.append(tangent_test_soup(rng, int(rng.integers(4, 60)), int(rng.integers(1, 400)), dup=float(rng.random() * 0.5),
                                       degen=float(rng.random() * 0.2), flat=float(rng.random() * 0.2)))
    for s in range(n_grid):
        cases.append(tangent_test_grid(int(rng.integers(3, 28)), int(rng.integers(3, 28)), rng, mirror=bool(s & 1), jitter=0.02 * (s % 3), explode=(s % 4 == 0)))
    g = np.load(GLASSES_NPZ)
    keys = set(g.files)
    if {"positions", "normals", "texcoords", "indices"} <= keys:
        cases.append((g["positions"].astype(np.float32), g["normals"].astype(np.float32), g["texcoords"].astype(np.float32), g["indices"].astype(np.uint32).reshape(-1, 3)))
    return cases
