"""Lens surfaces and their secondary rays (BASELINE config 4).  NEW functionality: the published
reference traces primary rays only, so there is no reference output to compare with ("parity unpinned"); the executable
specification is the CPU oracle (oracle/nmr_oracle.c: march_lens_ray, orc_mesh_render_layers, orc_lens_resolve), and the
CUDA path must match it: lens hand-off bit-exact, pixels within 2/255 and >= 45 dB."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

W, HH = 192, 108
PIX_TOL = 2.0 / 255.0


@pytest.fixture(scope="module")
def lens_scene(small_snapshot, tmp_path_factory):
    import pynmr
    import synth
    path, snap = small_snapshot
    gltf = synth.write_lens_glasses_gltf(str(tmp_path_factory.mktemp("lensmesh")))
    r = pynmr.NerfMeshRenderer(W, HH)
    nerf = r.load_nerf(path)
    assert nerf is not None
    assert r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is not None
    r.orbit(0.35, -0.2, 4.0)
    g = {"path": gltf, "t": synth.GLASSES_T, "s": synth.GLASSES_S, "r": synth.GLASSES_R_WXYZ,
         "texture": np.tile(np.array([128, 128, 128, 255], dtype=np.uint8), (4, 4, 1))}
    return {"r": r, "nerf": nerf, "snap": snap, "path": path, "glasses": g}


def cam12(r):
    return np.ascontiguousarray(r.view_projection_mat.T.reshape(-1))


def test_lens_handoff_matches_oracle(lens_scene):
    r, snap, g = lens_scene["r"], lens_scene["snap"], lens_scene["glasses"]
    w, t, n = H.debug_lens(r, W, HH)
    _, _, _, stats, _ = H.oracle_scene(snap, W, HH, cam12(r), glasses=g, n_steps_mode=1)
    L = stats["lens"]
    assert (L["w"] > 0).mean() > 0.004, "the lens panes should cover part of the frame"
    assert np.array_equal(w, L["w"])
    assert np.array_equal(t.view(np.uint32), L["t"].view(np.uint32))          # hit distances: same triangle test arithmetic
    on = L["w"] > 0
    assert np.max(np.abs(n[on] - L["n"][on])) <= 1e-6


def test_lens_pixels_match_oracle(lens_scene):
    r, nerf, snap, g = lens_scene["r"], lens_scene["nerf"], lens_scene["snap"], lens_scene["glasses"]
    want, _, ns, stats, _ = H.oracle_scene(snap, W, HH, cam12(r), glasses=g, n_steps_mode=1)
    assert stats["alive_after_first_hit"] * 8 <= W * HH
    img = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
    assert np.max(np.abs(img - want)) <= PIX_TOL
    assert H.psnr(img, want) >= 45.0
    # the lenses are visible: with lens surfaces switched off their triangles are ordinary opaque surfaces
    r.set_lens(False)
    try:
        opaque = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
    finally:
        r.set_lens(True)
    assert np.max(np.abs(opaque - img)) > 0.1
    want_off = H.oracle_scene(snap, W, HH, cam12(r), glasses=dict(g, lens=False), n_steps_mode=1)[0]
    assert np.max(np.abs(opaque - want_off)) <= PIX_TOL


def test_lens_parameters_and_sharding(lens_scene):
    import pynmr
    r, nerf = lens_scene["r"], lens_scene["nerf"]
    base = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
    r.set_lens(True, ior=2.4, transmission=0.3, tint=(1.0, 0.2, 0.2))
    try:
        tinted = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
    finally:
        r.set_lens(True, ior=1.5, transmission=0.85, tint=(0.75, 0.9, 1.0))
    assert np.max(np.abs(tinted - base)) > 0.05
    again = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
    assert np.array_equal(again.view(np.uint32), base.view(np.uint32))
    # row-band shards reproduce the full frame bit for bit (surface rule pinned: shards see other live-pixel ratios)
    r.set_surface_insertion(pynmr.NerfMeshRenderer.SURFACE_BATCH8)
    try:
        full = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
        merged = np.zeros_like(full)
        for rank in range(2):
            r.set_shard(rank, 2, 8)
            part = np.asarray(nerf.render(W, HH, 1, linear=False))
            rows = [y for y in range(HH) if (y // 8) % 2 == rank]
            merged[rows] = part[rows]
    finally:
        r.set_shard(0, 1, 8)
        r.set_surface_insertion(pynmr.NerfMeshRenderer.SURFACE_AUTO)
    assert np.array_equal(merged.view(np.uint32), full.view(np.uint32))


def test_lens_plate_model_matches_oracle(lens_scene):
    """Lens model 1: a pane of glass with two parallel Snell interfaces - the transmitted segment is shifted sideways and resumes
    behind the pane (include/nmr.h: nmr_set_lens_model; oracle: lens_plate_shift).  Thickness 0 is the thin sheet, bit for bit."""
    import pynmr
    r, nerf, snap, g = lens_scene["r"], lens_scene["nerf"], lens_scene["snap"], lens_scene["glasses"]
    thin = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
    try:
        r.set_lens_model(pynmr.NerfMeshRenderer.LENS_PLATE, 0.0)
        assert np.array_equal(np.asarray(nerf.render(W, HH, 1, linear=False)).view(np.uint32), thin.view(np.uint32))
        thickness = 0.03
        r.set_lens_model(pynmr.NerfMeshRenderer.LENS_PLATE, thickness)
        plate = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
        want = H.oracle_scene(snap, W, HH, cam12(r), glasses=dict(g, lens_model=1, lens_thickness=thickness), n_steps_mode=1)[0]
        assert np.max(np.abs(plate - want)) <= PIX_TOL and H.psnr(plate, want) >= 45.0
        assert np.max(np.abs(plate - thin)) > 0.02                       # the shift is visible through the lenses
        w, _, _ = H.debug_lens(r, W, HH)
        assert np.array_equal((np.abs(plate - thin).max(axis=2) > 0) & (w == 0), np.zeros_like(w, dtype=bool))      # ... and only there
        with pytest.raises(RuntimeError):
            r.set_lens_model(7, 0.01)
    finally:
        r.set_lens_model(pynmr.NerfMeshRenderer.LENS_THIN, 0.0)
    assert np.array_equal(np.asarray(nerf.render(W, HH, 1, linear=False)).view(np.uint32), thin.view(np.uint32))
