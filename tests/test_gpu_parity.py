"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path behind the C ABI vs the CPU oracle on the same
seeded inputs.  Bars (BASELINE.json north_star): ray set-up, occupancy traversal (t, Morton cell, mip) and the fp16
hash-grid features bit-exact; network outputs within fp16 rounding of the fp32-accumulating oracle; pixels within
2/255 max-abs and >= 45 dB PSNR."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

W, HH = 192, 108
PIX_TOL = 2.0 / 255.0


@pytest.fixture(scope="module")
def scene(small_snapshot, glasses_gltf):
    import pynmr
    import synth
    path, snap = small_snapshot
    r = pynmr.NerfMeshRenderer(W, HH)
    nerf = r.load_nerf(path)
    assert nerf is not None
    r.orbit(0.35, -0.2, 4.0)     # zoom in so the head fills a good part of the frame
    H.set_flags(r, 0)            # keep the probe surfaces (debug_last_frame)
    return {"r": r, "nerf": nerf, "snap": snap, "path": path, "gltf": glasses_gltf,
            "glasses": {"path": glasses_gltf, "t": synth.GLASSES_T, "s": synth.GLASSES_S, "r": synth.GLASSES_R_WXYZ,
                        "texture": np.tile(np.array([128, 128, 128, 255], dtype=np.uint8), (4, 4, 1))}}


def cam12(r):
    return np.ascontiguousarray(r.view_projection_mat.T.reshape(-1))


def test_camera_matches_oracle(scene):
    from oracle import oracle as O
    oc = O.OrbitCamera(W, HH)
    oc.orbit(0.35, -0.2, 4.0)
    assert np.array_equal(oc.matrix().view(np.uint32), cam12(scene["r"]).view(np.uint32))


def test_occupancy_bitfield_bit_exact(scene):
    from oracle import oracle as O
    m = O.Model.from_snapshot(scene["snap"])
    assert np.array_equal(H.get_bitfield(scene["r"], scene["nerf"]), m.bitfield())


def test_encoding_bit_exact(scene):
    from oracle import oracle as O
    m = O.Model.from_snapshot(scene["snap"])
    rng = np.random.default_rng(5)
    pos = rng.uniform(0, 1, size=(20000, 3)).astype(np.float32)
    pos[:8] = [[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0.999999, 0.5, 0.25], [1e-7, 1e-7, 1e-7]]
    got = H.debug_encode(scene["r"], scene["nerf"], pos)
    want = m.encode(pos).view(np.uint16)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("flags", [1, 0], ids=["cuda_core_mlp", "tcgen05_mlp"])
def test_network_outputs(scene, flags):
    from oracle import oracle as O
    m = O.Model.from_snapshot(scene["snap"])
    rng = np.random.default_rng(6)
    n = 128 * 37 + 5
    pos = rng.uniform(0.3, 0.7, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d01 = ((d + 1) * 0.5).astype(np.float32)
    H.set_flags(scene["r"], flags)
    try:
        got = H.debug_network(scene["r"], scene["nerf"], pos, d01).astype(np.float32)
    finally:
        H.set_flags(scene["r"], 0)
    want = m.network(pos, d01).astype(np.float32)
    if flags == 1:   # same k-ordered fp32 accumulation as the oracle: bit-exact
        assert np.array_equal(got, want)
    else:            # tensor cores sum in a different order: a few fp16 ulps on values of magnitude ~10
        err = np.abs(got - want)
        tol = 2.0 ** -8 * np.maximum(np.abs(want), 1.0) * 4
        assert np.all(err <= tol), float(err.max())


def test_traversal_bit_exact(scene):
    from oracle import oracle as O
    r, nerf, snap = scene["r"], scene["nerf"], scene["snap"]
    m = O.Model.from_snapshot(snap)
    P = m.params_struct(W, HH, cam12(r), aabb_min=snap["render_aabb_min"], aabb_max=snap["render_aabb_max"])
    pixels = np.arange(0, W * HH, 7, dtype=np.uint32)
    want = m.trace_samples(P, pixels, 48)
    got = H.debug_trace(r, nerf, W, HH, pixels, 48)
    assert want["count"].sum() > 1000
    for k in ("count", "cell", "mip"):
        assert np.array_equal(got[k], want[k]), k
    for k in ("t", "pos"):
        assert np.array_equal(got[k].view(np.uint32), want[k].view(np.uint32)), k
    # ray = origin3, dir3, t of the first occupied sample, alive.  t is compared on live rays only: a ray that dies keeps
    # whatever t its empty-space walk stopped at, which nothing reads (the product stops that walk at the far side of the
    # box around the occupied cells, the oracle - like the reference - at the far side of the render box).
    gr, wr = got["ray"].view(np.uint32), want["ray"].view(np.uint32)
    assert np.array_equal(gr[:, :6], wr[:, :6]) and np.array_equal(gr[:, 7], wr[:, 7])
    live = want["ray"][:, 7] > 0
    assert live.sum() > 100 and (~live).sum() > 100
    assert np.array_equal(gr[live, 6], wr[live, 6])


def test_render_no_mesh_pixels(scene):
    r, nerf, snap = scene["r"], scene["nerf"], scene["snap"]
    img = nerf.render(W, HH, 1, linear=False)
    want, frame, ns, stats, _ = H.oracle_scene(snap, W, HH, cam12(r))
    fr, dp, gns = H.debug_last_frame(r, W, HH)
    st = r.stats()
    assert st["rays_alive"] == stats["alive_after_first_hit"]
    assert stats["samples"] > 5000
    # sample counts can differ only where a termination threshold is crossed within rounding of the network outputs
    assert np.mean(gns == ns) > 0.98
    assert abs(int(gns.sum()) - stats["samples"]) <= 0.01 * stats["samples"]
    # the kernel evaluates rays in batches of 8 samples and discards what lies past a ray's termination
    assert int(gns.sum()) <= int(st["samples"]) <= int(gns.sum()) + 8 * int(st["rays_alive"])
    assert np.max(np.abs(np.asarray(img) - want)) <= PIX_TOL
    assert H.psnr(np.asarray(img), want) >= 45.0


def test_render_hybrid_pixels(scene):
    import pynmr
    r, nerf, snap, g = scene["r"], scene["nerf"], scene["snap"], scene["glasses"]
    mesh = r.load_mesh(g["path"], t=g["t"], s=g["s"], r=g["r"])
    assert mesh is not None
    rgba2, d2, tri2, surf, ts = H.debug_mesh(r, W, HH)
    # n_steps_mode 1: the oracle replays the reference's wavefront loop, n_steps = clamp(pixels / live rays, 1, 8), which
    # decides in front of which sample a mesh surface is blended (S/ngp/testbed.cu:843, 1996) - pinned against the
    # reference's own renderer in test_gpu_vs_reference.py.  n_steps_mode 0: surface at its exact position.
    want, frame, ns, stats, (osurf, ots) = H.oracle_scene(snap, W, HH, cam12(r), glasses=g, n_steps_mode=1)
    want_exact = H.oracle_scene(snap, W, HH, cam12(r), glasses=g, n_steps_mode=0)[0]
    covered = float((ots > 0).mean())
    assert covered > 0.003, "glasses should cover part of the frame"
    assert stats["alive_after_first_hit"] * 8 <= W * HH          # the framing where the reference runs 8-sample batches
    assert np.max(np.abs(want - want_exact)) > 10 * PIX_TOL      # ... and where that choice is visible
    # mesh stage: same triangle test arithmetic -> identical hit depths; colours within libm rounding
    assert np.array_equal(ts.view(np.uint32), ots.view(np.uint32))
    assert np.max(np.abs(surf - osurf)) <= 1e-5
    img = nerf.render(W, HH, 1, linear=False)                    # NMR_SURFACE_AUTO -> the reference's batches here
    assert np.max(np.abs(np.asarray(img) - want)) <= PIX_TOL
    assert H.psnr(np.asarray(img), want) >= 45.0
    # frame() renders the same hybrid image at the constructor resolution
    assert r.frame()
    assert np.max(np.abs(np.asarray(r.read_frame()) - want)) <= PIX_TOL
    r.set_surface_insertion(pynmr.NerfMeshRenderer.SURFACE_EXACT)
    try:
        img = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
    finally:
        r.set_surface_insertion(pynmr.NerfMeshRenderer.SURFACE_AUTO)
    assert np.max(np.abs(img - want_exact)) <= PIX_TOL


def test_render_hybrid_pixels_close_up(scene):
    """More than 1/8 of the pixels live: the reference's n_steps = clamp(pixels / live rays, 1, 8) varies from wavefront
    iteration to iteration.  NMR_SURFACE_AUTO replays that schedule (death histogram of a first pass -> schedule_kernel ->
    second pass over the rays that carry a mesh surface); the oracle's n_steps_mode 1 is the reference's loop itself."""
    import pynmr
    snap, g = scene["snap"], scene["glasses"]
    w, h = 160, 90
    r = pynmr.NerfMeshRenderer(w, h)
    nerf = r.load_nerf(scene["path"])
    assert r.load_mesh(g["path"], t=g["t"], s=g["s"], r=g["r"]) is not None
    r.orbit(0.2, -0.1, 9.0)                     # orbit radius is clamped to >= 1: walk the eye half-way in along the view axis
    m = r.view_projection_mat
    m[:, 3] += 0.5 * m[:, 2]
    r.view_projection_mat = m
    want, _, _, stats, _ = H.oracle_scene(snap, w, h, cam12(r), glasses=g, n_steps_mode=1)
    want_exact = H.oracle_scene(snap, w, h, cam12(r), glasses=g, n_steps_mode=0)[0]
    assert stats["alive_after_first_hit"] * 8 > w * h
    assert np.max(np.abs(want - want_exact)) > 10 * PIX_TOL          # the schedule is visible in this frame
    img = np.asarray(nerf.render(w, h, 1, linear=False)).copy()
    assert np.max(np.abs(img - want)) <= PIX_TOL
    assert H.psnr(img, want) >= 45.0
    r.set_surface_insertion(pynmr.NerfMeshRenderer.SURFACE_EXACT)
    img = np.asarray(nerf.render(w, h, 1, linear=False)).copy()
    assert np.max(np.abs(img - want_exact)) <= PIX_TOL
    # same frame through the copy-overlapped band path of render() (height >= 256) and through frame()
    r2 = pynmr.NerfMeshRenderer(480, 270)
    nerf2 = r2.load_nerf(scene["path"])
    assert r2.load_mesh(g["path"], t=g["t"], s=g["s"], r=g["r"]) is not None
    r2.view_projection_mat = m
    a = np.asarray(nerf2.render(480, 270, 1, linear=False)).copy()
    assert r2.frame()
    b = np.asarray(r2.read_frame())
    assert r2.stats()["rays_alive"] * 8 > 480 * 270
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_accumulation_and_linear_output(scene):
    r, nerf = scene["r"], scene["nerf"]
    a = np.asarray(nerf.render(W, HH, 1, linear=True))
    b = np.asarray(nerf.render(W, HH, 2, linear=True))
    assert a.shape == (HH, W, 4) and np.isfinite(a).all() and np.isfinite(b).all()
    d = np.abs(a - b)
    assert 0 < float(d.mean()) < 0.01 and float(d.max()) <= 0.5 + 1e-6   # the second sample uses another start jitter; mean of two samples


def test_sharded_render_equals_full(scene):
    import pynmr
    r, nerf = scene["r"], scene["nerf"]
    full = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
    merged = np.zeros_like(full)
    for rank in range(3):
        r.set_shard(rank, 3, 8)
        part = np.asarray(nerf.render(W, HH, 1, linear=False))
        rows = [y for y in range(HH) if (y // 8) % 3 == rank]
        merged[rows] = part[rows]
    r.set_shard(0, 1, 8)
    assert np.array_equal(merged.view(np.uint32), full.view(np.uint32))


def test_remove_floaties_matches_oracle(scene, small_snapshot):
    import pynmr
    from oracle import oracle as O
    path, snap = small_snapshot
    r2 = pynmr.NerfMeshRenderer(64, 64)
    nerf2 = r2.load_nerf(path)
    before = H.get_bitfield(r2, nerf2)
    n, kept = r2.remove_floaties()
    after = H.get_bitfield(r2, nerf2)
    want, n_ref, size_ref = O.remove_floaties_bitfield(before)
    assert (n, kept) == (n_ref, size_ref)
    assert np.array_equal(after, want)
    assert after.sum() < before.sum()


def test_render_views_batched(scene):
    r, nerf = scene["r"], scene["nerf"]
    cams = []
    for k in range(3):
        r.orbit(0.05, 0.01, 0)
        cams.append(r.view_projection_mat)
    out = np.asarray(r.render_views(nerf, np.stack(cams), 96, 54, linear=False))
    assert out.shape == (3, 54, 96, 4)
    r.view_projection_mat = cams[1]
    single = np.asarray(nerf.render(96, 54, 1, linear=False))
    assert np.array_equal(out[1].view(np.uint32), single.view(np.uint32))


def test_errors_are_reported_not_raised(scene, tmp_path):
    r = scene["r"]
    assert r.load_nerf(str(tmp_path / "missing.msgpack")) is None
    bad = tmp_path / "bad.msgpack"
    bad.write_bytes(b"\x81\xa1x\x01")
    assert r.load_nerf(str(bad)) is None
    assert r.load_mesh(str(tmp_path / "missing.gltf")) is None


def test_render_py_call_sequence(small_snapshot, glasses_gltf):
    """The calls volume/render.py makes, in its order and with its argument types (numpy arrays, keyword names), minus
    MediaPipe / OpenCV / numpy-quaternion which are not in this image (V/render.py:62-66, 70-95, 122-186, 196-258)."""
    import pynmr as nmr
    import synth
    path, _ = small_snapshot
    Wr, Hr = 256, 144
    renderer = nmr.NerfMeshRenderer(Wr, Hr)
    renderer.envmap("sunflowers_puresky_1k.png")                       # render.py:228 (absent from the reference module)
    nerf = renderer.load_nerf(path)
    assert nerf is not None
    before = np.asarray(nerf.render(Wr, Hr, linear=False)).copy()
    nerf.render_aabb.min = np.array([-0.2, 0.15, -0.2])                # render.py:234-235
    nerf.render_aabb.max = np.array([1, 1, 1])
    assert np.allclose(np.asarray(nerf.render_aabb.min), [-0.2, 0.15, -0.2])
    # rotate_camera_to_face_face / find_3d_landmarks: frame(), render_image(), orbit(), view_projection_mat
    views = []
    for i in range(3):
        assert renderer.frame() is True
        im = np.uint8(np.asarray(nerf.render(Wr, Hr, linear=False)) * 255)[::-1, :]       # render_image()
        assert im.shape == (Hr, Wr, 4) and im.dtype == np.uint8
        cam = renderer.view_projection_mat
        assert cam.shape == (3, 4)
        origin = np.squeeze(np.array(np.transpose(cam[:, 3])))         # class Ray
        direction = np.squeeze(np.array(cam[0:3, 0:3].dot(np.array([0.1, -0.2, 1]))))
        assert origin.shape == (3,) and direction.shape == (3,)
        views.append(im)
        renderer.orbit(0.1, 0, np.sin(i))
    assert not np.array_equal(views[0], views[2])
    # place_glasses: load_mesh with numpy t / s / r (w, x, y, z)
    mesh = renderer.load_mesh(glasses_gltf, t=np.array(synth.GLASSES_T), s=np.array([0.17, 0.17, 0.17]), r=np.array(synth.GLASSES_R_WXYZ))
    assert mesh is not None
    renderer.remove_floaties()                                         # render.py:238 (commented out there, bound in the module)
    a, shown = 0.0, []
    for _ in range(4):                                                 # the orbit loop, render.py:252-254
        assert renderer.frame()
        a += 0.03
        renderer.orbit(-(np.sin(a * 1.733)) / 100, np.cos(a * 1.733) / 200, 0)
        shown.append(np.asarray(renderer.read_frame()).copy())
    assert all(np.isfinite(s).all() for s in shown) and not np.array_equal(shown[0], shown[-1])
    nmr.free_temporary_memory()
    # secondary Testbed properties of the reference module
    assert nerf.training_step == 35000 or nerf.training_step >= 0
    assert nerf.loss >= 0.0 and nerf.n_params > 10240
    assert np.array_equal(nerf.render_aabb_to_local, np.eye(3, dtype=np.float32))
    assert nerf.nerf.rgb_activation == "Logistic" and nerf.nerf.density_activation == "Exponential" and nerf.nerf.cone_angle_constant == 0.0
    assert np.array_equal(nerf.camera_matrix, renderer.view_projection_mat)
    after = np.asarray(nerf.render(Wr, Hr, linear=False))
    assert after.shape == before.shape == (Hr, Wr, 4)


def test_banded_render_equals_frame_under_camera_jumps(small_snapshot, glasses_gltf):
    """Testbed.render() at >= 256 rows copies bands of rows out while later bands are still marched, and skips the march of
    bands that were empty in the previous frame; when that guess is wrong the frame is rendered again.  Whatever the camera
    does between calls, the image must equal frame()'s bit for bit."""
    import pynmr
    import synth
    path, _ = small_snapshot
    w, h = 384, 288
    r = pynmr.NerfMeshRenderer(w, h)
    nerf = r.load_nerf(path)
    assert r.load_mesh(glasses_gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is not None
    r.orbit(0.0, 0.0, 3.0)
    base = r.view_projection_mat
    poses = []
    for dy in (0.0, 0.0, 0.45, -0.45, 0.0, 0.45):          # head in the middle, twice, then pushed to the top / bottom bands and back
        m = base.copy()
        m[:, 3] += dy * m[:, 1] / np.linalg.norm(m[:, 1])
        poses.append(m)
    for m in poses:
        r.view_projection_mat = m
        a = np.asarray(nerf.render(w, h, 1, linear=False)).copy()
        r.view_projection_mat = m
        assert r.frame()
        b = np.asarray(r.read_frame())
        assert r.stats()["rays_alive"] > 500
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_aabb_scale_4_model(tmp_path):
    """A snapshot with aabb_scale 4: three occupancy cascades (+ pooled parents), a render box of size 4, the cone-angle step
    growth (cone_angle_constant 1/256, dt = clamp(t / 256)) and mip selection from both position and step size
    (S/ngp/testbed.cu:188-202, 1098-1115).  Same bars as the unit-cube model; also checked against the reference's own
    renderer where its library is present."""
    import pynmr
    import synth
    from oracle import oracle as O
    from oracle import refgpu
    path = str(tmp_path / "s4.msgpack")
    synth.write_snapshot(path, seed=7, log2_hashmap_size=15, aabb_scale=4)
    snap = synth.read_snapshot(path)
    w, h = 160, 90
    r = pynmr.NerfMeshRenderer(w, h)
    nerf = r.load_nerf(path)
    assert nerf is not None
    assert nerf.nerf.cone_angle_constant == 1.0 / 256.0
    H.set_flags(r, 0)
    r.orbit(0.4, -0.25, -1.5)
    m = O.Model.from_snapshot(snap)
    assert np.array_equal(H.get_bitfield(r, nerf), m.bitfield())
    c12 = cam12(r)
    P = m.params_struct(w, h, c12, aabb_min=snap["render_aabb_min"], aabb_max=snap["render_aabb_max"])
    pixels = np.arange(0, w * h, 3, dtype=np.uint32)
    want = m.trace_samples(P, pixels, 64)
    got = H.debug_trace(r, nerf, w, h, pixels, 64)
    assert want["count"].sum() > 2000 and int(want["mip"].max()) >= 1
    for k in ("count", "cell", "mip"):
        assert np.array_equal(got[k], want[k]), k
    for k in ("t", "pos"):
        assert np.array_equal(got[k].view(np.uint32), want[k].view(np.uint32)), k
    img = np.asarray(nerf.render(w, h, 1, linear=False)).copy()
    ref_img, _, ns, stats, _ = H.oracle_scene(snap, w, h, c12)
    assert r.stats()["rays_alive"] == stats["alive_after_first_hit"] and stats["samples"] > 5000
    assert np.max(np.abs(img - ref_img)) <= PIX_TOL and H.psnr(img, ref_img) >= 45.0
    if refgpu.available():
        ref = refgpu.ReferenceRenderer(path)
        try:
            assert np.array_equal(ref.bitfield(), m.bitfield())
            theirs, _ = ref.render(c12, w, h, 1, False)
        finally:
            ref.close()
        d = np.abs(img - theirs)
        assert H.psnr(img, theirs) >= 45.0
        assert float(np.mean(d.max(axis=2) > PIX_TOL)) <= 0.004


def test_edge_cases(small_snapshot, glasses_gltf, tmp_path):
    """Odd and tiny frames, an empty occupancy grid, a mesh out of view / behind the eye, the eye inside the head."""
    import msgpack
    import pynmr
    import synth
    path, snap = small_snapshot
    g = {"path": glasses_gltf, "t": synth.GLASSES_T, "s": synth.GLASSES_S, "r": synth.GLASSES_R_WXYZ,
         "texture": np.tile(np.array([128, 128, 128, 255], dtype=np.uint8), (4, 4, 1))}
    # odd sizes, hybrid, against the oracle (the camera's aspect stays that of the constructor resolution, like the reference)
    for (w, h) in ((37, 23), (1, 1), (2, 300)):
        r = pynmr.NerfMeshRenderer(w, h)
        nerf = r.load_nerf(path)
        assert r.load_mesh(g["path"], t=g["t"], s=g["s"], r=g["r"]) is not None
        r.orbit(0.3, -0.1, 5.0)
        img = np.asarray(nerf.render(w, h, 1, linear=False)).copy()
        want = H.oracle_scene(snap, w, h, cam12(r), glasses=g, n_steps_mode=1)[0]
        assert img.shape == (h, w, 4)
        assert np.max(np.abs(img - want)) <= PIX_TOL, (w, h)
    # nothing occupied: every pixel is the background, no ray is queued
    d = msgpack.unpackb(open(path, "rb").read(), raw=False, strict_map_key=False)
    d["snapshot"]["density_grid_binary"] = np.zeros(128 ** 3, dtype=np.float16).tobytes()
    empty = str(tmp_path / "empty.msgpack")
    open(empty, "wb").write(msgpack.packb(d, use_bin_type=True))
    r = pynmr.NerfMeshRenderer(64, 48)
    nerf = r.load_nerf(empty)
    img = np.asarray(nerf.render(64, 48, 1, linear=False))
    assert r.stats()["rays_alive"] == 0 and r.stats()["samples"] == 0
    assert np.allclose(img, 1.0, atol=1e-6)
    assert r.remove_floaties() is not None          # no clusters: nothing to keep, nothing breaks
    # mesh out of view and behind the eye: the frame equals the NeRF-only frame
    r = pynmr.NerfMeshRenderer(96, 54)
    nerf = r.load_nerf(path)
    r.orbit(0.3, -0.1, 4.0)
    plain = np.asarray(nerf.render(96, 54, 1, linear=False)).copy()
    assert r.load_mesh(g["path"], t=(0.0, 30.0, 0.0), s=g["s"], r=g["r"]) is not None      # far above
    assert np.array_equal(np.asarray(nerf.render(96, 54, 1, linear=False)).view(np.uint32), plain.view(np.uint32))
    eye = r.view_projection_mat[:, 3]; fwd = r.view_projection_mat[:, 2]
    r2 = pynmr.NerfMeshRenderer(96, 54)
    nerf2 = r2.load_nerf(path)
    r2.view_projection_mat = r.view_projection_mat
    assert r2.load_mesh(g["path"], t=tuple(eye - 0.8 * fwd), s=g["s"], r=g["r"]) is not None    # behind the eye
    assert np.array_equal(np.asarray(nerf2.render(96, 54, 1, linear=False)).view(np.uint32), plain.view(np.uint32))
    # the eye inside the head: rays start in occupied cells
    r3 = pynmr.NerfMeshRenderer(96, 54)
    nerf3 = r3.load_nerf(path)
    m = r3.view_projection_mat
    m[:, 3] = (0.0, 0.0, 0.05)
    r3.view_projection_mat = m
    img = np.asarray(nerf3.render(96, 54, 1, linear=False)).copy()
    want = H.oracle_scene(snap, 96, 54, cam12(r3))[0]
    assert np.max(np.abs(img - want)) <= PIX_TOL


@pytest.mark.parametrize("curve", [1, 2, 3], ids=["ACES", "Hable", "Reinhard"])
def test_tonemap_curves_match_oracle(scene, curve):
    """Testbed.tonemap_curve: the displayed image equals the oracle's accumulate + tonemap of the same linear frame."""
    from oracle import oracle as O
    r, nerf = scene["r"], scene["nerf"]
    try:
        nerf.tonemap_curve = curve
        img = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
        fr, _, _ = H.debug_last_frame(r, W, HH)
        want, _ = O.accumulate_tonemap(fr, None, 0, to_srgb=True, curve=curve)
        assert np.max(np.abs(img - want)) <= 1e-5
        lin = np.asarray(nerf.render(W, HH, 1, linear=True)).copy()
        fr, _, _ = H.debug_last_frame(r, W, HH)
        want_lin, _ = O.accumulate_tonemap(fr, None, 0, to_srgb=False, curve=curve)
        assert np.max(np.abs(lin - want_lin)) <= 1e-5
        with pytest.raises(RuntimeError):
            nerf.tonemap_curve = 7
    finally:
        nerf.tonemap_curve = 0


def test_full_size_frame_properties(tmp_path, glasses_gltf):
    """BASELINE configs[1] at its full size (1920x1080 hybrid frame, log2_hashmap_size 19 model, floatie removal on), through
    size-independent properties: the same camera renders the same bits twice; Testbed.render() (rows copied out under the
    rendering) equals frame(); row shards of three ranks reassemble the full frame bit for bit; a crop around the glasses
    agrees with the oracle rendering that window of the full-resolution frame."""
    import pynmr
    import synth
    from oracle import oracle as O
    FW, FH = 1920, 1080
    path = str(tmp_path / "full.msgpack")
    synth.write_snapshot(path, seed=1337, log2_hashmap_size=19)
    snap = synth.read_snapshot(path)
    r = pynmr.NerfMeshRenderer(FW, FH)
    nerf = r.load_nerf(path)
    assert nerf is not None and r.load_mesh(glasses_gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is not None
    r.remove_floaties()
    r.set_surface_insertion(pynmr.NerfMeshRenderer.SURFACE_BATCH8)       # ray-local rule: shards decide like the full frame
    r.orbit(-0.01, 0.004, 0.0)
    cam = r.view_projection_mat
    assert r.frame()
    a = np.asarray(r.read_frame()).copy()
    st = r.stats()
    assert st["rays"] == FW * FH and st["rays_alive"] > 20000 and st["samples"] > 200000
    r.view_projection_mat = cam                                           # restarts the accumulation
    assert r.frame()
    b = np.asarray(r.read_frame()).copy()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    c = np.asarray(nerf.render(FW, FH, 1, linear=False)).copy()
    assert np.array_equal(a.view(np.uint32), c.view(np.uint32))
    merged = np.zeros_like(a)
    for rank in range(3):
        r.set_shard(rank, 3, 16)
        r.view_projection_mat = cam
        assert r.frame()
        part = np.asarray(r.read_frame())
        rows = [y for y in range(FH) if (y // 16) % 3 == rank]
        merged[rows] = part[rows]
    r.set_shard(0, 1, 16)
    assert np.array_equal(merged.view(np.uint32), a.view(np.uint32))
    # oracle on a 96 x 54 window around the glasses of the same full-resolution frame
    x0, y0, cw, ch = 912, 540, 96, 54
    m = O.Model.from_snapshot(snap)
    m.set_bitfield(O.remove_floaties_bitfield(m.bitfield())[0])
    g = synth.read_gltf(glasses_gltf)
    mesh = O.Mesh(g["positions"], g["normals"], g["texcoords"], g["indices"], synth.GLASSES_T, synth.GLASSES_S, synth.GLASSES_R_WXYZ,
                  g["base_color"], g["metallic"], g["roughness"], (0, 0, 0), np.tile(np.array([128, 128, 128, 255], dtype=np.uint8), (4, 4, 1)))
    c12 = np.ascontiguousarray(cam.T.reshape(-1))
    rgba2, d2, _ = mesh.render(c12, 2 * FW, 2 * FH, window=(2 * x0, 2 * y0, 2 * (x0 + cw), 2 * (y0 + ch)))
    surf, ts = O.mesh_resolve(rgba2, d2, FW, FH, 2)
    P = m.params_struct(FW, FH, c12, aabb_min=snap["render_aabb_min"], aabb_max=snap["render_aabb_max"], n_steps_mode=2, window=(x0, y0, x0 + cw, y0 + ch))
    frame, _, _, _ = m.render_frame(P, surf, ts)
    want, _ = O.accumulate_tonemap(frame[y0:y0 + ch, x0:x0 + cw].copy(), None, 0, to_srgb=True)
    got = a[y0:y0 + ch, x0:x0 + cw]
    assert (ts[y0:y0 + ch, x0:x0 + cw] > 0).mean() > 0.02                 # the window does see the glasses
    assert np.max(np.abs(got - want)) <= PIX_TOL and H.psnr(got, want) >= 45.0


def test_render_views_lanes_equal_single_renders(small_snapshot, glasses_gltf):
    """nmr_render_views keeps up to 8 views in flight (helper contexts on their own streams).  19 hybrid views - more than there are
    lanes, so lanes are reused behind their copies - must equal 19 separate Testbed.render() calls bit for bit; with
    to_host=False the last view is what stays readable on the device."""
    import pynmr
    import synth
    path, _ = small_snapshot
    w, h = 160, 90
    r = pynmr.NerfMeshRenderer(w, h)
    nerf = r.load_nerf(path)
    assert nerf is not None and r.load_mesh(glasses_gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is not None
    r.remove_floaties()
    r.orbit(0.2, -0.1, 3.0)
    cams = []
    for k in range(19):
        r.orbit(0.11, 0.013 * ((k % 5) - 2), 0.1 * ((k % 3) - 1))
        cams.append(r.view_projection_mat)
    cams = np.stack(cams)
    out = np.asarray(r.render_views(nerf, cams, w, h, linear=False)).copy()
    assert out.shape == (19, h, w, 4)
    for k in range(19):
        r.view_projection_mat = cams[k]
        single = np.asarray(nerf.render(w, h, 1, linear=False))
        assert np.array_equal(out[k].view(np.uint32), single.view(np.uint32)), k
    assert r.render_views(nerf, cams, w, h, linear=False, to_host=False) is None
    ptr, dw, dh = r.device_image()
    assert (dw, dh) == (w, h) and ptr
    import torch
    last = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
    r.copy_device_image(last.data_ptr(), last.numel() * 4)
    assert np.array_equal(last.cpu().numpy().view(np.uint32), out[18].view(np.uint32))
    assert r.stats()["rays"] == w * h
