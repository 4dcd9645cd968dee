"""Live comparison of the C oracle with the reference's own host code (oracle/_ref/libnmr_ref.so).
Skipped where the reference library was not built (it needs /root/reference at build time)."""
import ctypes as C

import numpy as np
import pytest

from oracle import build_ref
from oracle import oracle as O

pytestmark = pytest.mark.skipif(not build_ref.available(), reason="oracle/_ref not built (no /root/reference here)")


def _ref():
    R = C.CDLL(build_ref.OUT)
    R.ref_ld_random_val.restype = C.c_float; R.ref_ld_random_val.argtypes = [C.c_uint32, C.c_uint32]
    R.ref_remove_floaties.restype = C.c_int; R.ref_remove_floaties.argtypes = [C.c_void_p] * 3
    return R


def test_jitter_render_loop_seeds():
    """The seeds advance_pos_nerf actually uses: index = spp, seed = pixel * 786433 (S/ngp/testbed.cu:503)."""
    R, L = _ref(), O.lib()
    for spp in (0, 1, 2, 7):
        for pix in list(range(0, 4000, 7)) + [1920 * 1080 - 1, 3840 * 2160 - 1]:
            seed = (pix * 786433) & 0xFFFFFFFF
            assert R.ref_ld_random_val(spp, seed) == L.orc_ld_random_val(spp, seed)


def test_floaties_random_cascade0_grid():
    R = _ref()
    rng = np.random.default_rng(11)
    c = np.zeros((8, 128, 128, 128), dtype=np.uint8)
    for _ in range(30):
        q = rng.integers(4, 100, size=3); s = rng.integers(1, 12, size=3)
        c[0, q[0]:q[0] + s[0], q[1]:q[1] + s[1], q[2]:q[2] + s[2]] = 1
    a = c.ravel().copy(); b = c.ravel().copy()
    n1 = R.ref_remove_floaties(a.ctypes.data_as(C.c_void_p), None, None)
    n2, _, _ = O.remove_floaties_cells(b)
    assert n1 == n2 and np.array_equal(a, b)
