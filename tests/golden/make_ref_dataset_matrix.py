"""Golden vectors of NerfDataset::nerf_matrix_to_ngp / ngp_matrix_to_nerf (S/ngp/nerf_loader.cuh:115-153) from the reference's OWN header -
run in the authoring container, where oracle/_ref/libnmr_ref.so exists (oracle/build_ref.py):
    python tests/golden/make_ref_dataset_matrix.py        ->  tests/golden/ref_dataset_matrix.npz
These conversions sit behind Testbed.crop_box / set_crop_box / crop_box_corners with nerf_space=True (S/ngp/testbed.cu:1421-1477)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
R = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "libnmr_ref.so"))
R.ref_dataset_matrix.argtypes = [C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]

rng = np.random.default_rng(20260103)
n = 240
to_ngp = rng.integers(0, 2, n).astype(np.int32); scale_cols = rng.integers(0, 2, n).astype(np.int32); mitsuba = (rng.random(n) < 0.25).astype(np.int32)
scale = rng.choice(np.array([0.33, 1.0, 0.25, 2.5, 0.0123], np.float32), n).astype(np.float32)
offset = rng.uniform(-1, 1, (n, 3)).astype(np.float32); offset[::3] = 0.5
mats = rng.normal(size=(n, 3, 4)).astype(np.float32)
out = np.zeros_like(mats)
for i in range(n):
    m = np.ascontiguousarray(mats[i]); o = np.zeros((3, 4), np.float32); off = np.ascontiguousarray(offset[i])
    R.ref_dataset_matrix(int(to_ngp[i]), int(scale_cols[i]), float(scale[i]), off.ctypes.data, int(mitsuba[i]), m.ctypes.data, o.ctypes.data)
    out[i] = o
np.savez_compressed(os.path.join(HERE, "ref_dataset_matrix.npz"), to_ngp=to_ngp, scale_columns=scale_cols, from_mitsuba=mitsuba, scale=scale, offset=offset, mats=mats, out=out)
print("wrote ref_dataset_matrix.npz:", n, "matrices")
