"""GPU parity at the sizes BASELINE.json's configs name (run with -m gpu on the B200 box).

  configs[2]  render.py's landmark pass: the 64 camera poses of tests/golden/alice_views64.npy at 512 x 512 through
              nmr_render_views == 64 separate Testbed.render() calls bit for bit; two of them against the oracle.
  configs[3]  3840 x 2160 hybrid frame with lens secondary rays: a window around the lens panes against the oracle, three row
              shards reassemble the full frame bit for bit.
  configs[4]  log2_hashmap_size 22 and 24 (hash tables of 150 / 507 MiB: 64-bit level bases, power-of-two masks): the encoder
              bit-exact against the oracle, and a window of a 1080p hybrid frame of the 2^22 model against the oracle.
  plus        spp = 2 and 4 against the oracle's accumulation, and the first-hit walk (lattice jumps + coarse empty-space
              skips) bit-exact against the oracle's plain walk over 40 random poses x 512 x 512 rays (> 10^7 rays).
Bars: integer / traversal results bit-exact, pixels <= 2/255 max-abs and >= 45 dB PSNR (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PIX_TOL = 2.0 / 255.0
GREY = np.tile(np.array([128, 128, 128, 255], dtype=np.uint8), (4, 4, 1))


def cam12_of(mat34):
    return np.ascontiguousarray(np.asarray(mat34, dtype=np.float32).T.reshape(-1))


@pytest.fixture(scope="module")
def model19(tmp_path_factory):
    """The bench model: synthetic snapshot seed 1337, log2_hashmap_size 19."""
    import synth
    path = str(tmp_path_factory.mktemp("m19") / "s19.msgpack")
    synth.write_snapshot(path, seed=1337, log2_hashmap_size=19)
    return path, synth.read_snapshot(path)


def glasses_dict(gltf):
    import synth
    return {"path": gltf, "t": synth.GLASSES_T, "s": synth.GLASSES_S, "r": synth.GLASSES_R_WXYZ, "texture": GREY}


# ---------------------------------------------------------------------------------------------------------------------------
# configs[2]: 64 views at 512 x 512
# ---------------------------------------------------------------------------------------------------------------------------
def test_config3_views64_equal_single_renders_and_oracle(model19, glasses_gltf):
    import pynmr
    import synth
    from oracle import oracle as O
    path, snap = model19
    w = h = 512
    cams = np.load(os.path.join(ROOT, "tests", "golden", "alice_views64.npy"))
    assert cams.shape == (64, 3, 4)
    r = pynmr.NerfMeshRenderer(w, h)
    nerf = r.load_nerf(path)
    assert nerf is not None and r.load_mesh(glasses_gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is not None
    r.remove_floaties()
    out = np.asarray(r.render_views(nerf, cams, w, h, linear=False))
    assert out.shape == (64, h, w, 4)
    live_views = 0
    for k in range(64):
        r.view_projection_mat = cams[k]
        single = np.asarray(nerf.render(w, h, 1, linear=False))
        assert np.array_equal(out[k].view(np.uint32), single.view(np.uint32)), k
        live_views += int(r.stats()["rays_alive"] > 2000)
    assert live_views >= 32                                    # the dataset's cameras look at the head
    # two views against the oracle (floaties removed there as well; the reference's wavefront batching for the mesh surface)
    m_bits = O.remove_floaties_bitfield(O.Model.from_snapshot(snap).bitfield())[0]
    for k in (0, 37):
        want = oracle_hybrid(snap, m_bits, glasses_dict(glasses_gltf), w, h, cam12_of(cams[k]), n_steps_mode=1)
        d = np.abs(out[k] - want)
        assert float(d.max()) <= PIX_TOL and H.psnr(out[k], want) >= 45.0, (k, float(d.max()))


def oracle_hybrid(snap, bitfield, g, w, h, c12, n_steps_mode, window=None, lens_gltf=None, spp_index=0):
    """Oracle image of the hybrid scene with a given occupancy bitfield; window = (x0, y0, x1, y1) renders that part only."""
    import synth
    from oracle import oracle as O
    m = O.Model.from_snapshot(snap)
    m.set_bitfield(bitfield)
    gl = synth.read_gltf(lens_gltf if lens_gltf else g["path"])
    mesh = O.Mesh(gl["positions"], gl["normals"], gl["texcoords"], gl["indices"], g["t"], g["s"], g["r"],
                  gl["base_color"], gl["metallic"], gl["roughness"], (0, 0, 0), g["texture"])
    win2 = None if window is None else (2 * window[0], 2 * window[1], 2 * window[2], 2 * window[3])
    lens = None
    if lens_gltf:
        mesh.set_lens(gl["tri_lens"])
        rgba2, d2, ld2, ln2 = mesh.render_layers(c12, 2 * w, 2 * h, window=win2)
        surf, ts = O.mesh_resolve(rgba2, d2, w, h, 2)
        lw, lt, lnn = O.lens_resolve(d2, ld2, ln2, ts, w, h, 2)
        lp = gl["lens"]
        lens = {"w": lw, "t": lt, "n": lnn, "f0": O.lens_f0(lp["ior"]), "k": np.float32(lp["transmission"]) * lp["tint"].astype(np.float32),
                "background": (1.0, 1.0, 1.0, 1.0)}
    else:
        rgba2, d2, _ = mesh.render(c12, 2 * w, 2 * h, window=win2)
        surf, ts = O.mesh_resolve(rgba2, d2, w, h, 2)
    P = m.params_struct(w, h, c12, aabb_min=snap["render_aabb_min"], aabb_max=snap["render_aabb_max"], n_steps_mode=n_steps_mode,
                        window=window, spp_index=spp_index)
    frame, _, _, _ = m.render_frame(P, surf, ts, lens=lens)
    if window is not None:
        x0, y0, x1, y1 = window
        frame = frame[y0:y1, x0:x1].copy()
    return O.accumulate_tonemap(frame, None, 0, to_srgb=True)[0]


# ---------------------------------------------------------------------------------------------------------------------------
# configs[3]: 4K hybrid frame with lens secondary rays
# ---------------------------------------------------------------------------------------------------------------------------
def test_config4_4k_lens_frame_window_vs_oracle_and_row_shards(model19, tmp_path):
    import pynmr
    import synth
    from oracle import oracle as O
    path, snap = model19
    FW, FH = 3840, 2160
    lens_gltf = synth.write_lens_glasses_gltf(str(tmp_path / "lensmesh"))
    r = pynmr.NerfMeshRenderer(FW, FH)
    nerf = r.load_nerf(path)
    assert nerf is not None and r.load_mesh(lens_gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is not None
    r.remove_floaties()
    r.set_surface_insertion(pynmr.NerfMeshRenderer.SURFACE_BATCH8)        # ray-local rule: shards decide like the full frame
    r.orbit(-0.02, 0.01, 0.0)
    cam = r.view_projection_mat
    assert r.frame()
    full = np.asarray(r.read_frame()).copy()
    st = r.stats()
    assert st["rays"] == FW * FH and st["rays_alive"] > 80000
    lw, lt, _ = H.debug_lens(r, FW, FH)
    ys, xs = np.nonzero(lw > 0)
    assert ys.size > 5000, "the lens panes should cover part of the frame"
    # three row shards of the same frame reassemble it bit for bit (lens rays included)
    merged = np.zeros_like(full)
    for rank in range(3):
        r.set_shard(rank, 3, 16)
        r.view_projection_mat = cam
        assert r.frame()
        part = np.asarray(r.read_frame())
        rows = [y for y in range(FH) if (y // 16) % 3 == rank]
        merged[rows] = part[rows]
    r.set_shard(0, 1, 16)
    assert np.array_equal(merged.view(np.uint32), full.view(np.uint32))
    # a 96 x 54 window centred on the lens pixels against the oracle rendering that window of the 4K frame
    left = xs < np.median(xs)                                   # (two panes: centre the window on one of them, not on the bridge)
    cx, cy = int(np.median(xs[left])), int(np.median(ys[left]))
    x0, y0 = max(0, min(FW - 96, cx - 48)), max(0, min(FH - 54, cy - 27))
    win = (x0, y0, x0 + 96, y0 + 54)
    assert float((lw[y0:y0 + 54, x0:x0 + 96] > 0).mean()) > 0.2
    bits = O.remove_floaties_bitfield(O.Model.from_snapshot(snap).bitfield())[0]
    g = {"path": lens_gltf, "t": synth.GLASSES_T, "s": synth.GLASSES_S, "r": synth.GLASSES_R_WXYZ, "texture": GREY}
    want = oracle_hybrid(snap, bits, g, FW, FH, cam12_of(cam), n_steps_mode=2, window=win, lens_gltf=lens_gltf)
    got = full[y0:y0 + 54, x0:x0 + 96]
    assert float(np.abs(got - want).max()) <= PIX_TOL and H.psnr(got, want) >= 45.0


# ---------------------------------------------------------------------------------------------------------------------------
# configs[4]: tables that leave the L2
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("log2T", [22, 24])
def test_config5_encoding_bit_exact_on_large_tables(tmp_path, log2T):
    import pynmr
    import synth
    from oracle import oracle as O
    path = str(tmp_path / f"s{log2T}.msgpack")
    synth.write_snapshot(path, seed=1337, log2_hashmap_size=log2T)
    snap = synth.read_snapshot(path)
    r = pynmr.NerfMeshRenderer(64, 64)
    nerf = r.load_nerf(path)
    os.remove(path)
    assert nerf is not None
    m = O.Model.from_snapshot(snap)
    offs = m.level_table()[0]
    assert int(offs[-1]) * 4 > (140 << 20)                      # the table is larger than the L2
    rng = np.random.default_rng(50 + log2T)
    pos = rng.uniform(0, 1, size=(30000, 3)).astype(np.float32)
    pos[:6] = [[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5], [1, 0, 0], [0.999999, 0.5, 0.25], [1e-7, 1e-7, 1e-7]]
    got = H.debug_encode(r, nerf, pos)
    want = m.encode(pos).view(np.uint16)
    assert np.array_equal(got, want)


def test_config5_1080p_window_vs_oracle_log2T22(tmp_path, glasses_gltf):
    import pynmr
    import synth
    from oracle import oracle as O
    FW, FH = 1920, 1080
    path = str(tmp_path / "s22.msgpack")
    synth.write_snapshot(path, seed=1337, log2_hashmap_size=22)
    snap = synth.read_snapshot(path)
    r = pynmr.NerfMeshRenderer(FW, FH)
    nerf = r.load_nerf(path)
    os.remove(path)
    assert nerf is not None and r.load_mesh(glasses_gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is not None
    r.remove_floaties()
    r.set_surface_insertion(pynmr.NerfMeshRenderer.SURFACE_BATCH8)
    r.orbit(0.0, 0.0, 4.0)                                      # the encoding-bound framing of tools/run_configs.py
    cam = r.view_projection_mat
    assert r.frame()
    full = np.asarray(r.read_frame()).copy()
    assert r.stats()["samples"] > 500000
    bits = O.remove_floaties_bitfield(O.Model.from_snapshot(snap).bitfield())[0]
    win = (912, 560, 912 + 96, 560 + 54)
    want = oracle_hybrid(snap, bits, glasses_dict(glasses_gltf), FW, FH, cam12_of(cam), n_steps_mode=2, window=win)
    got = full[win[1]:win[3], win[0]:win[2]]
    assert float((np.abs(got[..., :3] - 1.0).max(axis=2) > 0.01).mean()) > 0.5       # the window is on the head
    assert float(np.abs(got - want).max()) <= PIX_TOL and H.psnr(got, want) >= 45.0


# ---------------------------------------------------------------------------------------------------------------------------
# spp > 1 (accumulation, S/ngp/render_buffer.cu:232-267) against the oracle
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("spp", [2, 4])
def test_spp_accumulation_matches_oracle(small_snapshot, glasses_gltf, spp):
    import pynmr
    import synth
    from oracle import oracle as O
    path, snap = small_snapshot
    w, h = 192, 108
    r = pynmr.NerfMeshRenderer(w, h)
    nerf = r.load_nerf(path)
    assert r.load_mesh(glasses_gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is not None
    r.orbit(0.35, -0.2, 4.0)
    c12 = cam12_of(r.view_projection_mat)
    g = glasses_dict(glasses_gltf)
    for linear in (False, True):
        img = np.asarray(nerf.render(w, h, spp, linear=linear)).copy()
        accum, want = None, None
        for k in range(spp):
            frame = H.oracle_scene(snap, w, h, c12, glasses=g, spp_index=k, n_steps_mode=1)[1]
            want, accum = O.accumulate_tonemap(frame, accum, k, to_srgb=not linear)
        assert float(np.abs(img - want).max()) <= PIX_TOL and H.psnr(img, want) >= 45.0, (spp, linear)
    one = np.asarray(nerf.render(w, h, 1, linear=True))
    assert float(np.abs(one - img).max()) > 0                    # further samples use other start jitters


# ---------------------------------------------------------------------------------------------------------------------------
# first-hit walk over random poses
# ---------------------------------------------------------------------------------------------------------------------------
def _random_pose(rng):
    """A camera somewhere around (and sometimes inside) the unit cube, looking roughly at the head."""
    eye = rng.normal(size=3); eye /= np.linalg.norm(eye)
    eye *= rng.choice([0.05, 0.3, 0.6, 1.0, 1.7, 2.5, 4.0]) * rng.uniform(0.8, 1.25)
    target = rng.uniform(-0.15, 0.15, size=3)
    fwd = target - eye; fwd /= np.linalg.norm(fwd)
    up0 = rng.normal(size=3)
    right = np.cross(fwd, up0); right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    fov = rng.uniform(0.25, 1.7)
    m = np.zeros((3, 4), dtype=np.float32)
    m[:, 0] = right * fov; m[:, 1] = up * fov; m[:, 2] = fwd; m[:, 3] = eye
    return m


def test_first_hit_bit_exact_over_random_poses(small_snapshot):
    """coarse_skip / lattice_advance are exact by an argument about roundings (device_common.cuh); this pins the argument on
    40 random poses x 512 x 512 = 10.5 M rays: alive masks, t, Morton cell and mip of the first sample equal the oracle's
    plain walk bit for bit."""
    import pynmr
    from oracle import oracle as O
    path, snap = small_snapshot
    w = h = 512
    r = pynmr.NerfMeshRenderer(w, h)
    nerf = r.load_nerf(path)
    m = O.Model.from_snapshot(snap)
    rng = np.random.default_rng(20261018)
    pixels = np.arange(w * h, dtype=np.uint32)
    total, total_live = 0, 0
    for _ in range(40):
        pose = _random_pose(rng)
        r.view_projection_mat = pose
        P = m.params_struct(w, h, cam12_of(pose), aabb_min=snap["render_aabb_min"], aabb_max=snap["render_aabb_max"])
        want = m.trace_samples(P, pixels, 1)
        got = H.debug_trace(r, nerf, w, h, pixels, 1)
        gr, wr = got["ray"].view(np.uint32), want["ray"].view(np.uint32)
        assert np.array_equal(gr[:, 7], wr[:, 7])
        live = want["ray"][:, 7] > 0
        assert np.array_equal(gr[live, 6], wr[live, 6])
        assert np.array_equal(got["count"], want["count"])
        assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
        assert np.array_equal(got["cell"], want["cell"]) and np.array_equal(got["mip"], want["mip"])
        total += w * h; total_live += int(live.sum())
    assert total >= 10_000_000 and total_live > 500_000


# ---------------------------------------------------------------------------------------------------------------------------
# compact image outputs: converted on the device, equal to what render.py computes on the host
# ---------------------------------------------------------------------------------------------------------------------------
def test_u8_and_f16_outputs_equal_host_conversion(small_snapshot, glasses_gltf):
    import pynmr
    import synth
    path, _ = small_snapshot
    for (w, h) in ((192, 108), (384, 288)):                    # below / above the row-banded copy path of render()
        r = pynmr.NerfMeshRenderer(w, h)
        nerf = r.load_nerf(path)
        assert r.load_mesh(glasses_gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is not None
        r.orbit(0.35, -0.2, 4.0)
        for linear in (False, True):
            f32 = np.asarray(nerf.render(w, h, 1, linear=linear)).copy()
            u8 = np.asarray(nerf.render(w, h, 1, linear=linear, dtype=np.uint8)).copy()
            f16 = np.asarray(nerf.render(w, h, 1, linear=linear, dtype=np.float16)).copy()
            assert u8.dtype == np.uint8 and u8.shape == (h, w, 4) and f16.dtype == np.float16
            assert np.array_equal(u8, np.uint8(np.clip(f32, 0.0, 1.0) * 255))            # render.py: np.uint8(img * 255)
            assert np.array_equal(f16.view(np.uint16), f32.astype(np.float16).view(np.uint16))
            assert len(np.unique(u8[..., 0])) > 20
        # render.py's frame loop: frame() then render(); read_frame() refuses to hand out a render() image as the frame
        assert r.frame()
        fr = np.asarray(r.read_frame()).copy()
        assert np.array_equal(np.uint8(fr * 255), np.asarray(nerf.render(w, h, 1, linear=False, dtype=np.uint8)))
        with pytest.raises(RuntimeError):
            r.read_frame()
        # batched views in u8
        cams = []
        for k in range(5):
            r.orbit(0.07, 0.01, 0)
            cams.append(r.view_projection_mat)
        v32 = np.asarray(r.render_views(nerf, np.stack(cams), w, h)).copy()
        v8 = np.asarray(r.render_views(nerf, np.stack(cams), w, h, dtype=np.uint8)).copy()
        assert np.array_equal(v8, np.uint8(v32 * 255))


def test_read_frame_and_copy_device_image_check_sizes(small_snapshot):
    """a render() at another resolution must not let read_frame() overrun the caller's buffer."""
    import pynmr
    path, _ = small_snapshot
    r = pynmr.NerfMeshRenderer(128, 72)
    nerf = r.load_nerf(path)
    assert r.frame()
    a = np.asarray(r.read_frame()).copy()
    nerf.render(320, 180, 1, linear=False)                     # resizes the surfaces
    with pytest.raises(RuntimeError):
        r.read_frame()
    import torch
    small = torch.empty(16, dtype=torch.float32, device="cuda")
    with pytest.raises(RuntimeError):
        r.copy_device_image(small.data_ptr(), small.numel() * 4)
    r.view_projection_mat = r.view_projection_mat
    assert r.frame()
    assert np.array_equal(np.asarray(r.read_frame()).view(np.uint32), a.view(np.uint32))
