"""Generates tests/golden/ref_vectors.npz by calling the REFERENCE's own host code through
oracle/_ref/libnmr_ref.so (built by oracle/build_ref.py from the headers under /root/reference).

Authoring-container only.  The .npz travels to the GPU box, where /root/reference does not exist;
tests/test_oracle_golden.py replays these inputs through the C oracle and demands identical outputs.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle import build_ref  # noqa: E402


class RefCamera(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("eye", C.c_float * 3), ("look", C.c_float * 3), ("pivot", C.c_float * 3), ("up", C.c_float * 3)]


def load_ref():
    so = build_ref.build()
    if so is None:
        raise SystemExit("reference tree / nvcc not available")
    R = C.CDLL(so)
    vp = C.c_void_p
    R.ref_ld_random_val.restype = C.c_float; R.ref_ld_random_val.argtypes = [C.c_uint32, C.c_uint32]
    R.ref_morton3D.restype = C.c_uint32; R.ref_morton3D.argtypes = [C.c_uint32] * 3
    R.ref_morton3D_invert.restype = C.c_uint32; R.ref_morton3D_invert.argtypes = [C.c_uint32]
    R.ref_linear_to_srgb.restype = C.c_float; R.ref_linear_to_srgb.argtypes = [C.c_float]
    R.ref_srgb_to_linear.restype = C.c_float; R.ref_srgb_to_linear.argtypes = [C.c_float]
    R.ref_grid_scale.restype = C.c_float; R.ref_grid_scale.argtypes = [C.c_uint32, C.c_float, C.c_uint32]
    R.ref_grid_resolution.restype = C.c_uint32; R.ref_grid_resolution.argtypes = [C.c_float]
    R.ref_aabb_ray_intersect.argtypes = [vp] * 5
    R.ref_aabb_contains.restype = C.c_int; R.ref_aabb_contains.argtypes = [vp] * 3
    R.ref_pixel_to_ray.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    R.ref_remove_floaties.restype = C.c_int; R.ref_remove_floaties.argtypes = [vp, vp, vp]
    R.ref_camera_init.argtypes = [C.POINTER(RefCamera)]
    R.ref_camera_orbit.argtypes = [C.POINTER(RefCamera), C.c_float, C.c_float, C.c_float]
    return R


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def floaty_cases(rng):
    """-> list of (name, sparse set-cell indices into the [8][128][128][128] byte grid)."""
    import synth
    cases = []
    # (a) the synthetic snapshot's occupancy (cascade 0 head ellipsoid + 64 floaters), upper cascades = max-pool
    from oracle import oracle as O
    grid = synth.make_density_grid(np.random.default_rng(1337), 64)
    m = O.Model(synth.make_params(np.random.default_rng(1), log2_hashmap_size=12, calibrate=False), grid, log2_hashmap_size=12)
    cells = O.bitfield_to_cells(m.bitfield())
    cases.append(("synthetic_head", np.flatnonzero(cells).astype(np.uint32)))
    # (b) two blobs + isolated cells, cascade 0 only
    c = np.zeros((8, 128, 128, 128), dtype=np.uint8)  # [lvl][z][y][x]
    c[0, 10:20, 10:20, 10:20] = 1
    c[0, 60:90, 50:70, 40:100] = 1
    for q in rng.integers(0, 128, size=(40, 3)):
        c[0, q[0], q[1], q[2]] = 1
    cases.append(("two_blobs", np.flatnonzero(c.ravel()).astype(np.uint32)))
    # (c) a bar crossing the cascade 0 -> 1 -> 2 boundary along +x (each level stores its own shell), plus a far blob in cascade 1
    c = np.zeros((8, 128, 128, 128), dtype=np.uint8)
    c[0, 64:66, 64:66, 100:128] = 1            # reaches x = 127 of cascade 0
    c[1, 64, 64, 96:128] = 1                   # continues in cascade 1 (x 96.. is outside the interior)
    c[2, 64, 64, 96:110] = 1                   # and in cascade 2
    c[1, 5:9, 5:9, 5:9] = 1                    # separate blob
    c[0, 3, 3, 3] = 1                          # isolated
    cases.append(("cross_cascade_bar", np.flatnonzero(c.ravel()).astype(np.uint32)))
    return cases


def _floaty_worker(sparse):
    R = load_ref()
    cells = np.zeros(8 * 128 ** 3, dtype=np.uint8); cells[sparse] = 1
    size = C.c_int64(0); score = C.c_int64(0)
    ncl = R.ref_remove_floaties(p(cells), C.byref(size), C.byref(score))
    return np.flatnonzero(cells).astype(np.uint32), ncl, size.value, score.value


def main():
    R = load_ref()
    rng = np.random.default_rng(20240610)
    out = {}
    # --- ld_random_val
    idx = rng.integers(0, 2 ** 32, size=4000, dtype=np.uint64).astype(np.uint32)
    idx[::3] = rng.integers(0, 16, size=idx[::3].size)
    seed = (rng.integers(0, 2 ** 22, size=4000, dtype=np.uint64) * np.uint64(786433)).astype(np.uint32)
    out["ld_index"], out["ld_seed"] = idx, seed
    out["ld_out"] = np.array([R.ref_ld_random_val(int(a), int(b)) for a, b in zip(idx, seed)], dtype=np.float32)
    # --- morton
    xyz = rng.integers(0, 1024, size=(2000, 3)).astype(np.uint32)
    out["morton_xyz"] = xyz
    out["morton_out"] = np.array([R.ref_morton3D(int(a), int(b), int(c)) for a, b, c in xyz], dtype=np.uint32)
    codes = rng.integers(0, 2 ** 30, size=2000).astype(np.uint32)
    out["morton_inv_in"] = codes
    out["morton_inv_out"] = np.array([R.ref_morton3D_invert(int(a)) for a in codes], dtype=np.uint32)
    # --- colour transfer
    v = np.concatenate([rng.uniform(0, 1, 2000), rng.uniform(0, 0.01, 500), [0.0, 1.0, 0.0031308, 0.04045]]).astype(np.float32)
    out["color_in"] = v
    out["lin2srgb"] = np.array([R.ref_linear_to_srgb(float(a)) for a in v], dtype=np.float32)
    out["srgb2lin"] = np.array([R.ref_srgb_to_linear(float(a)) for a in v], dtype=np.float32)
    # --- grid scale table for the stock network
    import synth
    pls = np.float32(synth.per_level_scale(16, 16))
    l2 = np.float32(np.log2(pls))
    out["grid_log2_pls"] = np.array([l2], dtype=np.float32)
    out["grid_scale"] = np.array([R.ref_grid_scale(l, float(l2), 16) for l in range(16)], dtype=np.float32)
    out["grid_res"] = np.array([R.ref_grid_resolution(float(s)) for s in out["grid_scale"]], dtype=np.uint32)
    # --- AABB slab test
    n = 3000
    bmin = rng.uniform(-0.3, 0.4, size=(n, 3)).astype(np.float32); bmax = (bmin + rng.uniform(0.2, 1.0, size=(n, 3))).astype(np.float32)
    pos = rng.uniform(-2, 3, size=(n, 3)).astype(np.float32); d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    d[::50, 0] = 0.0
    d[25::50, 2] = 0.0
    tt = np.zeros((n, 2), dtype=np.float32); cont = np.zeros(n, dtype=np.int32)
    for i in range(n):
        R.ref_aabb_ray_intersect(p(bmin[i]), p(bmax[i]), p(pos[i]), p(d[i]), p(tt[i]))
        cont[i] = R.ref_aabb_contains(p(bmin[i]), p(bmax[i]), p(pos[i]))
    out.update(aabb_min=bmin, aabb_max=bmax, aabb_pos=pos, aabb_dir=d, aabb_t=tt, aabb_contains=cont)
    # --- orbit camera sequence (render.py's loop plus zooms and pole clamps)
    cam = RefCamera(); R.ref_camera_init(C.byref(cam))
    deltas = [(0.0, 0.0, 0.0), (0.1, 0.0, float(np.sin(1.0))), (np.deg2rad(60), np.deg2rad(-15), 0.0), (0.0, 0.0, 2.0), (-np.pi / 2, 0.0, 0.0)]
    a = 0.0
    for _ in range(60):
        a += 0.03
        deltas.append((-(np.sin(a * 1.733)) / 100, np.cos(a * 1.733) / 200, 0.0))
    deltas += [(0.3, 2.0, -3.0), (-7.0, -4.0, 0.5), (0.0, 0.0, 20.0)]
    deltas = np.array(deltas, dtype=np.float32)
    states = np.zeros((len(deltas) + 1, 16 + 3 + 3), dtype=np.float32)
    states[0] = list(cam.view) + list(cam.eye) + list(cam.look)
    for i, (da, dp, dz) in enumerate(deltas):
        R.ref_camera_orbit(C.byref(cam), float(da), float(dp), float(dz))
        states[i + 1] = list(cam.view) + list(cam.eye) + list(cam.look)
    out["orbit_deltas"], out["orbit_states"] = deltas, states
    # --- pixel_to_ray
    from oracle import oracle as O
    oc = O.OrbitCamera(1920, 1080); oc.orbit(0.4, -0.2, 1.0)
    cam12 = oc.matrix()
    pix = np.stack([rng.integers(0, 1920, 500), rng.integers(0, 1080, 500)], axis=1).astype(np.int32)
    rays = np.zeros((500, 9), dtype=np.float32)
    for i in range(500):
        R.ref_pixel_to_ray(0, int(pix[i, 0]), int(pix[i, 1]), 1920, 1080, p(cam12), p(rays[i]))
    out.update(p2r_cam=cam12, p2r_pix=pix, p2r_out=rays)
    # --- floaties
    # one fresh process per case: NgpGrid::point_set_importance caches scores in a function-local static map keyed
    # by the cluster's ADDRESS (S/floatyremover.h:254-265), so a second call in one process can read stale scores.
    import multiprocessing as mp
    names = []
    ctx = mp.get_context("spawn")
    for name, sparse in floaty_cases(rng):
        with ctx.Pool(1) as pool:
            cells_out, ncl, size, score = pool.apply(_floaty_worker, (sparse,))
        out[f"floaty_{name}_in"] = sparse
        out[f"floaty_{name}_out"] = cells_out
        out[f"floaty_{name}_meta"] = np.array([ncl, size, score], dtype=np.int64)
        names.append(name)
        print(name, "clusters", ncl, "best size", size, "score", score, "cells out", int(cells_out.size))
    out["floaty_names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    print("wrote ref_vectors.npz")


if __name__ == "__main__":
    main()
