"""Golden vectors of the GUI's trajectory tool (S/nerf_mesh_renderer.cu:604-659) from the reference's OWN code - run in the
authoring container, where oracle/_ref/libnmr_ref.so (oracle/build_ref.py: the reference's flythrough_camera.h, glm and Eigen
compiled around oracle/ref_harness.cu) exists:
    python tests/golden/make_ref_trajectory.py        ->  tests/golden/ref_trajectory.npz
 * poses: (angle, distance, height, lookat) -> cam_pos, cam_look, viewMat of ref_trajectory_camera;
 * text: 3 x 4 matrices -> the string `viewProjectionMat.format(Eigen::IOFormat(FullPrecision, 0, ", ", ",\\n", "[", "]", "[", "]"))`."""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
R = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "libnmr_ref.so"))
R.ref_trajectory_camera.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
R.ref_format_transform.argtypes = [C.c_void_p, C.c_char_p, C.c_int]; R.ref_format_transform.restype = C.c_int
rng = np.random.default_rng(20260101)
n = 200
inp = np.zeros((n, 6), np.float32)
inp[:, 0] = rng.uniform(-7, 7, n); inp[:, 1] = rng.uniform(0.3, 4.0, n); inp[:, 2] = rng.uniform(-1.5, 1.5, n); inp[:, 3:] = rng.uniform(-0.4, 0.4, (n, 3))
inp[0] = (0.5, 1.1, 0.1, 0, 0, 0)                      # the sliders' initial values
inp[1:11, 0] = 0.5 + 0.2 * np.arange(1, 11, dtype=np.float32); inp[1:11, 1:] = inp[0, 1:]
out = np.zeros((n, 3 + 3 + 16), np.float32)
for i in range(n):
    la = np.ascontiguousarray(inp[i, 3:]); eye = np.zeros(3, np.float32); look = np.zeros(3, np.float32); view = np.zeros(16, np.float32)
    R.ref_trajectory_camera(float(inp[i, 0]), float(inp[i, 1]), float(inp[i, 2]), la.ctypes.data, eye.ctypes.data, look.ctypes.data, view.ctypes.data)
    out[i] = np.concatenate([eye, look, view])
mats, texts = [], []
for k in range(120):
    kind = k % 4
    m = rng.normal(size=12)
    if kind == 1: m = m * 10.0 ** rng.integers(-8, 9, 12)
    if kind == 2: m = np.round(m * 3)
    if kind == 3: m[rng.integers(0, 12)] = -0.0; m[rng.integers(0, 12)] = 1e-7
    m = m.astype(np.float32)
    buf = C.create_string_buffer(2048)
    assert R.ref_format_transform(m.ctypes.data, buf, 2048) > 0
    mats.append(m); texts.append(buf.value.decode())
np.savez_compressed(os.path.join(HERE, "ref_trajectory.npz"), pose_in=inp, pose_out=out, text_mats=np.stack(mats), texts=np.array(texts))
print("wrote ref_trajectory.npz:", n, "poses,", len(texts), "matrices")
