"""GPU parity of SURVEY.md 8f.4 (run with -m gpu on the B200 box): the model transform of a NeRF (Testbed::m_model_translation /
m_model_rotation, S/ngp/testbed.cu:1537-1542, 442-446), several NeRFs in one frame with the depth merge of combineBuffersKernel
(S/nerf_mesh_renderer.cu:34-48, 582-597), and the density-grid side format of dumpDensityGrid / loadDensityGrid (:239-358).
Each against the CPU oracle, and against the reference's own kernels (oracle/_ref/libnmr_refgpu.so) where that library exists."""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

W, HH = 192, 108
PIX_TOL = 2.0 / 255.0


def cam12(r):
    return np.ascontiguousarray(r.view_projection_mat.T.reshape(-1))


@pytest.fixture(scope="module")
def second_snapshot(tmp_path_factory):
    import synth
    path = str(tmp_path_factory.mktemp("snap2") / "second.msgpack")
    synth.write_snapshot(path, seed=4242, log2_hashmap_size=15)
    return path, synth.read_snapshot(path)


def oracle_nerf(snap, w, h, c12, model_rot=None, model_trans=None, surf=None, ts=None, n_steps_mode=1):
    from oracle import oracle as O
    m = O.Model.from_snapshot(snap)
    P = m.params_struct(w, h, c12, aabb_min=snap["render_aabb_min"], aabb_max=snap["render_aabb_max"], n_steps_mode=n_steps_mode,
                        model_rot=model_rot, model_trans=model_trans)
    frame, depth, ns, stats = m.render_frame(P, surf, ts)
    return m, P, frame, depth, stats


# ---------------------------------------------------------------------------------------------------------------------------
def test_model_transform_matches_oracle_and_reference(small_snapshot):
    import pynmr
    from oracle import oracle as O
    from oracle import refgpu
    path, snap = small_snapshot
    r = pynmr.NerfMeshRenderer(W, HH)
    nerf = r.load_nerf(path)
    r.orbit(0.35, -0.2, 4.0)
    H.set_flags(r, 0)
    plain = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
    assert np.array_equal(nerf.model_matrix, np.eye(3, dtype=np.float32))
    trans, rot_pi = (0.06, -0.04, 0.03), (0.05, -0.12, 0.08)
    nerf.model_translation = trans
    nerf.model_rotation = rot_pi
    R = nerf.model_matrix
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-6) and not np.allclose(R, np.eye(3), atol=1e-2)
    assert np.allclose(nerf.model_translation, trans) and np.allclose(nerf.model_rotation, rot_pi)
    # traversal: ray origin / direction, t, cell and mip of the samples bit-exact against the oracle given the same 3x3 matrix
    m = O.Model.from_snapshot(snap)
    P = m.params_struct(W, HH, cam12(r), aabb_min=snap["render_aabb_min"], aabb_max=snap["render_aabb_max"], model_rot=R, model_trans=trans)
    pixels = np.arange(0, W * HH, 5, dtype=np.uint32)
    want = m.trace_samples(P, pixels, 32)
    got = H.debug_trace(r, nerf, W, HH, pixels, 32)
    assert want["count"].sum() > 1000
    for k in ("count", "cell", "mip"):
        assert np.array_equal(got[k], want[k]), k
    for k in ("t", "pos"):
        assert np.array_equal(got[k].view(np.uint32), want[k].view(np.uint32)), k
    gr, wr = got["ray"].view(np.uint32), want["ray"].view(np.uint32)
    assert np.array_equal(gr[:, :6], wr[:, :6]) and np.array_equal(gr[:, 7], wr[:, 7])
    # pixels
    img = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
    assert float(np.abs(img - plain).max()) > 0.2                       # the head has visibly moved
    _, _, frame, _, stats = oracle_nerf(snap, W, HH, cam12(r), R, trans)
    want_img, _ = O.accumulate_tonemap(frame, None, 0, to_srgb=True)
    assert r.stats()["rays_alive"] == stats["alive_after_first_hit"]
    assert float(np.abs(img - want_img).max()) <= PIX_TOL and H.psnr(img, want_img) >= 45.0
    if refgpu.available() and hasattr(refgpu.lib(), "refgpu_set_model_transform"):
        ref = refgpu.ReferenceRenderer(path)
        try:
            ref.set_model_transform(trans, rot_pi)
            theirs, _ = ref.render(cam12(r), W, HH, 1, False)
        finally:
            ref.close()
        d = np.abs(img - theirs)
        assert H.psnr(img, theirs) >= 45.0 and float(np.mean(d.max(axis=2) > PIX_TOL)) <= 0.004
    # back to the identity: the very same bits as before
    nerf.model_translation = (0, 0, 0)
    nerf.model_rotation = (0, 0, 0)
    again = np.asarray(nerf.render(W, HH, 1, linear=False))
    assert np.array_equal(again.view(np.uint32), plain.view(np.uint32))


# ---------------------------------------------------------------------------------------------------------------------------
def test_density_grid_dump_and_load(small_snapshot, tmp_path):
    import pynmr
    from oracle import oracle as O
    path, snap = small_snapshot
    r = pynmr.NerfMeshRenderer(64, 64)
    nerf = r.load_nerf(path)
    bits = H.get_bitfield(r, nerf)
    f = str(tmp_path / "density_grid.bin")
    cells = nerf.dump_density_grid(f)
    assert os.path.getsize(f) == 8 * 128 ** 3
    assert cells.shape == (8, 128, 128, 128) and set(np.unique(cells)) <= {0, 1}
    assert np.array_equal(cells.reshape(-1), O.bitfield_to_cells(bits).reshape(-1))          # the oracle's (reference-pinned) layout
    assert np.array_equal(np.fromfile(f, dtype=np.uint8), cells.reshape(-1))
    # the reference's workflow: dump, prune the floaters on the host, load the result back (S/nerf_mesh_renderer.cu:901-917)
    pruned_bits, n_clusters, _ = O.remove_floaties_bitfield(bits)
    pruned_cells = O.bitfield_to_cells(pruned_bits)
    assert int(pruned_cells.sum()) < int(cells.sum())
    nerf.load_density_grid(pruned_cells)
    assert np.array_equal(H.get_bitfield(r, nerf), pruned_bits)
    r2 = pynmr.NerfMeshRenderer(64, 64)
    nerf2 = r2.load_nerf(path)
    r2.remove_floaties()
    assert np.array_equal(H.get_bitfield(r2, nerf2), pruned_bits)                            # same result as the device-side pruning
    # and through the file
    nerf.load_density_grid(f)
    assert np.array_equal(H.get_bitfield(r, nerf), bits)
    bad = tmp_path / "short.bin"
    bad.write_bytes(b"\x01" * 100)
    with pytest.raises(RuntimeError):
        nerf.load_density_grid(str(bad))


# ---------------------------------------------------------------------------------------------------------------------------
def test_two_nerfs_merge_by_depth(small_snapshot, second_snapshot, glasses_gltf):
    """frame() with two NeRFs: per-NeRF frame / depth buffers merged like NerfMeshRenderer::render_frame does, the mesh hand-off
    going to the first NeRF only."""
    import pynmr
    import synth
    from oracle import oracle as O
    from oracle import refgpu
    path_a, snap_a = small_snapshot
    path_b, snap_b = second_snapshot
    r = pynmr.NerfMeshRenderer(W, HH)
    a = r.load_nerf(path_a)
    b = r.load_nerf(path_b)
    assert a is not None and b is not None
    trans_b = (0.22, 0.02, -0.05)
    b.model_translation = trans_b                                     # the second head stands next to the first one
    g = {"path": glasses_gltf, "t": synth.GLASSES_T, "s": synth.GLASSES_S, "r": synth.GLASSES_R_WXYZ,
         "texture": np.tile(np.array([128, 128, 128, 255], dtype=np.uint8), (4, 4, 1))}
    assert r.load_mesh(g["path"], t=g["t"], s=g["s"], r=g["r"]) is not None
    r.orbit(0.35, -0.2, 3.0)
    c12 = cam12(r)
    assert r.frame()
    img = np.asarray(r.read_frame()).copy()
    frame, depth = r.read_combined()
    # oracle: NeRF a with the mesh hand-off, NeRF b without, then the merge
    _, _, _, _, (surf, ts) = H.oracle_scene(snap_a, W, HH, c12, glasses=g, n_steps_mode=1)
    _, _, fa, da, sa = oracle_nerf(snap_a, W, HH, c12, surf=surf, ts=ts)
    _, _, fb, db, sb = oracle_nerf(snap_b, W, HH, c12, model_trans=trans_b)
    take_b = db < da
    want_frame = np.where(take_b[..., None], fb, fa); want_depth = np.where(take_b, db, da)
    assert take_b.mean() > 0.02 and (da < 1e9).mean() > 0.02 and ((da < 1e9) & ~take_b).mean() > 0.02      # both heads own pixels
    # a pixel whose two depths are within rounding of each other may go either way
    close = (np.abs(da - db) <= 1e-3 * np.minimum(da, db)) & ((da < 1e9) | (db < 1e9))
    ok = ~close
    assert close.mean() < 0.01
    assert float(np.abs(frame - want_frame)[ok].max()) <= 6e-3
    assert np.array_equal((depth < 1e9)[ok], (want_depth < 1e9)[ok])
    hit = ok & (want_depth < 1e9)
    assert float(np.abs(depth - want_depth)[hit].max()) <= 2e-2      # depth = distance of the max-weight sample: a sample apart at most
    want_img, _ = O.accumulate_tonemap(want_frame, None, 0, to_srgb=True)
    assert float(np.abs(img - want_img)[ok].max()) <= PIX_TOL
    # one NeRF again: the plain path, bit-identical to a renderer that never saw a second NeRF
    r1 = pynmr.NerfMeshRenderer(W, HH)
    r1.load_nerf(path_a)
    assert r1.load_mesh(g["path"], t=g["t"], s=g["s"], r=g["r"]) is not None
    r1.view_projection_mat = r.view_projection_mat
    assert r1.frame()
    single = np.asarray(r1.read_frame()).copy()
    only_a = ~take_b
    assert float(np.abs(single - img)[only_a & ok].max()) <= 1e-6
    with pytest.raises(RuntimeError):
        r1.read_combined()                                            # one NeRF, probes off: no merged buffers are kept
    if refgpu.available() and hasattr(refgpu.lib(), "refgpu_get_buffers"):
        ra, rb = refgpu.ReferenceRenderer(path_a), refgpu.ReferenceRenderer(path_b)
        try:
            rb.set_model_transform(trans_b, (0, 0, 0))
            ra.render(c12, W, HH, 1, False, surf=surf, ts=ts)
            rfa, rda = ra.buffers(W, HH)
            rb.render(c12, W, HH, 1, False)
            rfb, rdb = rb.buffers(W, HH)
        finally:
            ra.close(); rb.close()
        tb = rdb < rda                                                 # combineBuffersKernel
        ref_frame = np.where(tb[..., None], rfb, rfa)
        same_owner = tb == take_b
        assert same_owner.mean() > 0.995
        d = np.abs(frame - ref_frame)[same_owner]
        assert float(np.mean(d.max(axis=-1) > 6e-3)) <= 0.004


# ---------------------------------------------------------------------------------------------------------------------------
def _read_png_rgb8(path):
    """8-bit RGB, filter type 0 on every row (what pynmr writes)"""
    import struct, zlib
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    p, idat, w, h = 8, b"", 0, 0
    while p < len(data):
        n = struct.unpack(">I", data[p:p + 4])[0]; tag = data[p + 4:p + 8]; body = data[p + 8:p + 8 + n]
        assert struct.unpack(">I", data[p + 8 + n:p + 12 + n])[0] == (zlib.crc32(tag + body) & 0xFFFFFFFF)
        if tag == b"IHDR":
            w, h, depth, colour = struct.unpack(">IIBB", body[:10]); assert (depth, colour) == (8, 2)
        elif tag == b"IDAT":
            idat += body
        p += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 3 * w)
    assert (raw[:, 0] == 0).all()
    return raw[:, 1:].reshape(h, w, 3)


def test_trajectory_export(small_snapshot, glasses_gltf, tmp_path):
    """The GUI's trajectory tool without the GUI (SURVEY 8f.4 side formats): poses bit-exact vs the oracle (which is pinned on the
    reference's own camera code, tests/golden/ref_trajectory.npz), transform_N in Eigen's text layout, one PNG per step."""
    import pynmr
    import synth
    from oracle import oracle as O
    path, snap = small_snapshot
    r = pynmr.NerfMeshRenderer(W, HH)
    nerf = r.load_nerf(path)
    assert r.load_mesh(glasses_gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is not None
    cam = O.OrbitCamera(W, HH)
    rng = np.random.default_rng(3)
    for _ in range(50):
        a, d, h = float(rng.uniform(-7, 7)), float(rng.uniform(0.3, 4.0)), float(rng.uniform(-1.5, 1.5))
        la = rng.uniform(-0.4, 0.4, 3).astype(np.float32)
        r.trajectory_pose(a, d, h, la); cam.trajectory_pose(a, d, h, la)
        assert np.array_equal(cam12(r).view(np.uint32), cam.matrix().view(np.uint32))
    out = str(tmp_path / "traj")
    n = r.export_trajectory(out, num_images=5)                  # the sliders' other defaults: arc 0.5 .. 2.5 at distance 1.1, height 0.1
    assert n == 4 and sorted(os.listdir(out)) == sorted([f"trajectory_{k}.png" for k in range(1, 5)] + [f"transform_{k}" for k in range(1, 5)])
    step = np.float32(np.float32(2.5) - np.float32(0.5)) / np.float32(5)
    angle = np.float32(0.5)
    for k in range(1, 5):
        angle = np.float32(angle + step)
        cam.trajectory_pose(float(angle), 1.1, 0.1)
        want = cam.matrix().reshape(4, 3).T                     # 3 x 4, like view_projection_mat
        assert open(os.path.join(out, f"transform_{k}")).read() == pynmr.format_transform(want)
        png = _read_png_rgb8(os.path.join(out, f"trajectory_{k}.png"))
        r.trajectory_pose(float(angle), 1.1, 0.1)
        assert r.frame()
        img = np.asarray(r.read_frame())
        assert png.shape[:2] == (HH, W)
        assert np.array_equal(png[..., :3], np.uint8(np.clip(img[::-1, :, :3], 0.0, 1.0) * np.float32(255.0)))
        assert png[..., :3].std() > 5                           # a picture, not a constant


def test_rotated_crop_box_and_secondary_testbed_properties(small_snapshot):
    """Testbed.set_crop_box / crop_box / crop_box_corners with a ROTATED box (S/ngp/testbed.cu:1421-1477; render_aabb_to_local is
    writable in the reference, S/python_api.cu:411): the walks test samples against the box in its own frame (S/ngp/testbed.cu:505, 596).
    Traversal stays bit-exact against the oracle given the same matrix, pixels within tolerance; plus the camera helpers and the
    plain members of the reference's Testbed binding (S/python_api.cu:408-447)."""
    import pynmr
    from oracle import oracle as O
    path, snap = small_snapshot
    W, HH = 160, 90
    r = pynmr.NerfMeshRenderer(W, HH)
    nerf = r.load_nerf(path)
    r.orbit(0.3, -0.15, 3.0)
    c12 = np.ascontiguousarray(r.view_projection_mat.T.reshape(-1))
    plain = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
    # a box turned by 30 degrees about y and 20 about x, centred on the head, in NeRF coordinates (nerf_space=False)
    a, b = np.deg2rad(30.0), np.deg2rad(20.0)
    Ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32)
    Rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]], np.float32)
    axes = (Ry @ Rx).astype(np.float32)
    radius = np.array([0.12, 0.2, 0.16], np.float32)
    m = np.concatenate([axes * radius[None, :], np.array([[0.5], [0.52], [0.5]], np.float32)], axis=1)
    nerf.set_crop_box(m, nerf_space=False)
    r2l = nerf.render_aabb_to_local
    assert float(np.abs(r2l - axes.T).max()) <= 1e-6
    assert float(np.abs(nerf.crop_box(nerf_space=False) - m).max()) <= 1e-6                      # round trip
    rt = nerf.crop_box(nerf_space=True)
    corners = nerf.crop_box_corners(nerf_space=False)
    assert len(corners) == 8 and float(np.abs(np.mean(corners, axis=0) - m[:, 3]).max()) <= 1e-6
    mn, mx = np.asarray(nerf.render_aabb.min), np.asarray(nerf.render_aabb.max)
    assert float(np.abs((mx - mn) * 0.5 - radius).max()) <= 1e-6
    img = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
    assert float(np.abs(img - plain).max()) > 0.05                                               # the crop cuts into the head
    model = O.Model.from_snapshot(snap)
    P = model.params_struct(W, HH, c12, aabb_min=mn, aabb_max=mx, n_steps_mode=1)
    P.render_aabb_to_local[:] = [float(x) for x in r2l.reshape(-1)]
    frame, _, ns, _ = model.render_frame(P)
    want, _ = O.accumulate_tonemap(frame, None, 0, to_srgb=True)
    assert float(np.abs(img - want).max()) <= 2.0 / 255.0 and H.psnr(img, want) >= 45.0
    # the same box handed over in dataset coordinates gives the same state back
    nerf.set_crop_box(rt, nerf_space=True)
    assert float(np.abs(nerf.crop_box(nerf_space=False) - m).max()) <= 2e-6
    assert np.array_equal(np.asarray(nerf.render(W, HH, 1, linear=False)), img) or float(np.abs(np.asarray(nerf.render(W, HH, 1, linear=False)) - img).max()) <= 2.0 / 255.0
    # camera helpers (S/ngp/testbed.cu:1319-1349) act on camera_matrix
    cam = nerf.camera_matrix
    assert np.allclose(nerf.look_at, cam[:, 3] + cam[:, 2] * nerf.scale)
    la = nerf.look_at.copy()
    assert nerf.scale == 1.5                                      # Testbed::reset_camera (S/ngp/testbed.cu:1388)
    nerf.scale = 2.0
    assert np.allclose(nerf.look_at, la, atol=1e-6) and np.allclose(nerf.camera_matrix[:, 3], (cam[:, 3] - la) * np.float32(2.0 / 1.5) + la, atol=1e-6)
    nerf.look_at = [0.1, 0.2, 0.3]
    assert np.allclose(nerf.look_at, [0.1, 0.2, 0.3], atol=1e-6)
    nerf.view_dir = [0.0, 0.0, 1.0]
    cm = nerf.camera_matrix
    assert np.allclose(cm[:, 2], [0, 0, 1], atol=1e-6) and np.allclose(cm[:, 0], np.cross([0, 0, 1], nerf.up_dir) / np.linalg.norm(np.cross([0, 0, 1], nerf.up_dir)), atol=1e-6)
    assert np.allclose(nerf.look_at, [0.1, 0.2, 0.3], atol=1e-5)
    assert nerf.bounding_radius > 0 and np.allclose(nerf.raw_aabb.min, nerf.aabb.min)
    # plain members keep what they are given; members that would change the picture refuse anything but the default
    nerf.zoom = 2.0; nerf.screen_center = [0.4, 0.6]; nerf.snap_to_pixel_centers = True; nerf.camera_smoothing = True; nerf.sun_dir = [0, 1, 0]
    assert nerf.zoom == 2.0 and np.allclose(nerf.screen_center, [0.4, 0.6]) and np.allclose(nerf.sun_dir, [0, 1, 0])
    nerf.color_space = pynmr.ColorSpace.Linear
    with pytest.raises(NotImplementedError):
        nerf.color_space = pynmr.ColorSpace.SRGB
    nerf.parallax_shift = [0, 0, 1]
    with pytest.raises(NotImplementedError):
        nerf.parallax_shift = [0.1, 0, 1]
    with pytest.raises(RuntimeError):
        nerf.render_aabb_to_local = np.full((3, 3), np.nan, np.float32)


def test_load_snapshot_into_an_existing_testbed(small_snapshot, second_snapshot, tmp_path):
    """Testbed.load_snapshot(path) (S/python_api.cu:319): the second model replaces the first in the same Testbed - frames equal those of
    a renderer that loaded the second model directly; the background colour set through the property survives; a file that does
    not load raises and leaves the model in place."""
    import pynmr
    path_a, _ = small_snapshot
    path_b, _ = second_snapshot
    r = pynmr.NerfMeshRenderer(W, HH)
    nerf = r.load_nerf(path_a)
    nerf.background_color = [0.2, 0.4, 0.6, 1.0]
    r.orbit(0.3, -0.2, 3.0)
    img_a = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
    nerf.load_snapshot(path_b)
    img_b = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
    fresh = pynmr.NerfMeshRenderer(W, HH)
    nb = fresh.load_nerf(path_b)
    nb.background_color = [0.2, 0.4, 0.6, 1.0]
    fresh.view_projection_mat = r.view_projection_mat
    want = np.asarray(nb.render(W, HH, 1, linear=False))
    assert np.array_equal(img_b, want)
    assert float(np.abs(img_b - img_a).max()) > 0.05
    assert np.allclose(nerf.background_color, [0.2, 0.4, 0.6, 1.0])
    bad = tmp_path / "broken.msgpack"
    bad.write_bytes(open(path_a, "rb").read()[:1000])
    with pytest.raises(RuntimeError):
        nerf.load_snapshot(str(bad))
    assert np.array_equal(np.asarray(nerf.render(W, HH, 1, linear=False)), img_b)
    assert r.frame()
