"""The mesh shader with every texture slot of the reference's closest-hit program in use (S/optix/optix_scene.cu:221-258: base
colour, emissive, metallic-roughness, normal map through a Gram-Schmidt TBN matrix, occlusion), against the oracle's restatement.
Like the rest of the mesh stage this is pinned on the oracle only - OptiX itself is not available - but a glTF that
uses these slots is no longer shaded as if it did not."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

W, HH = 192, 108
PIX_TOL = 2.0 / 255.0


def oracle_mesh(gltf, t, s, r):
    import synth
    from oracle import oracle as O
    g = synth.read_gltf(gltf)
    base = synth._read_png_rgba8(gltf.replace("glasses.gltf", "base.png")) if g["base_texture_ref"] is not None else None
    mesh = O.Mesh(g["positions"], g["normals"], g["texcoords"], g["indices"], t, s, r, g["base_color"], g["metallic"], g["roughness"], g["emissive"], base)
    for name, tex in g["textures"].items():
        if tex is not None:
            mesh.set_texture(name, tex, {"normal": g["normal_scale"], "occlusion": g["occlusion_strength"]}.get(name, 1.0))
    mesh.set_tangents(g["normals"], g["tangents"], s, r)
    return mesh, g


def test_textured_mesh_stage_matches_oracle(small_snapshot, glasses_gltf, tmp_path):
    import pynmr
    import synth
    from oracle import oracle as O
    path, snap = small_snapshot
    gltf = synth.write_textured_glasses_gltf(str(tmp_path / "tex"))
    t, s, rq = synth.GLASSES_T, (0.17, 0.2, 0.15), (0.6830127, 0.6830127, 0.1830127, -0.1830127)       # non-uniform scale, a tilted rotation
    r = pynmr.NerfMeshRenderer(W, HH)
    nerf = r.load_nerf(path)
    assert r.load_mesh(gltf, t=t, s=s, r=rq) is not None
    r.orbit(0.35, -0.2, 4.0)
    c12 = np.ascontiguousarray(r.view_projection_mat.T.reshape(-1))
    rgba2, d2, tri2, surf, ts = H.debug_mesh(r, W, HH)
    mesh, g = oracle_mesh(gltf, t, s, rq)
    want_rgba2, want_d2, want_tri = mesh.render(c12, 2 * W, 2 * HH)
    hit = want_tri >= 0
    assert hit.mean() > 0.002
    assert np.array_equal(tri2, want_tri)
    assert np.array_equal(d2[hit].view(np.uint32), want_d2[hit].view(np.uint32))
    assert float(np.abs(rgba2 - want_rgba2).max()) <= 2e-4
    # the textures matter: the same geometry with the plain fixture's material looks different
    plain = pynmr.NerfMeshRenderer(W, HH)
    plain.load_nerf(path)
    assert plain.load_mesh(glasses_gltf, t=t, s=s, r=rq) is not None
    plain.view_projection_mat = r.view_projection_mat
    p_rgba2 = H.debug_mesh(plain, W, HH)[0]
    assert float(np.abs(p_rgba2 - rgba2)[hit].mean()) > 0.02
    # each slot on its own moves the picture (a slot silently ignored would not)
    full = want_rgba2[hit]
    for name in ("emissive", "metallic_roughness", "normal", "occlusion"):
        m2, _ = oracle_mesh(gltf, t, s, rq)
        m2.set_texture(name, None, 1.0)
        assert float(np.abs(m2.render(c12, 2 * W, 2 * HH)[0][hit] - full).max()) > 1e-3, name
    # and the hybrid frame
    want_surf, want_ts = O.mesh_resolve(want_rgba2, want_d2, W, HH, 2)
    m = O.Model.from_snapshot(snap)
    P = m.params_struct(W, HH, c12, aabb_min=snap["render_aabb_min"], aabb_max=snap["render_aabb_max"], n_steps_mode=1)
    frame, _, _, _ = m.render_frame(P, want_surf, want_ts)
    want_img, _ = O.accumulate_tonemap(frame, None, 0, to_srgb=True)
    img = np.asarray(nerf.render(W, HH, 1, linear=False))
    assert float(np.abs(img - want_img).max()) <= PIX_TOL and H.psnr(img, want_img) >= 45.0


def test_normal_mapped_mesh_without_tangent_attribute(small_snapshot, tmp_path):
    """A glTF whose primitive has a normal map but no TANGENT attribute: load_mesh generates the tangents the way the reference does
    (S/gltf_scene.cpp:150-155 -> mikktspace.c; the generator itself is pinned bit for bit on the reference's file in
    tests/test_host_cpu.py), and the shader's TBN matrix uses them - the mesh stage equals the oracle fed with those tangents, and
    differs from the same file shaded with the fixture's own (unweighted) tangents."""
    import pynmr
    import synth
    path, _ = small_snapshot
    without = synth.write_textured_glasses_gltf(str(tmp_path / "no_tangent"), with_tangents=False)
    with_t = synth.write_textured_glasses_gltf(str(tmp_path / "tangent"), with_tangents=True)
    t, s, rq = synth.GLASSES_T, (0.17, 0.2, 0.15), (0.6830127, 0.6830127, 0.1830127, -0.1830127)
    r = pynmr.NerfMeshRenderer(W, HH)
    r.load_nerf(path)
    assert r.load_mesh(without, t=t, s=s, r=rq) is not None
    r.orbit(0.35, -0.2, 4.0)
    c12 = np.ascontiguousarray(r.view_projection_mat.T.reshape(-1))
    rgba2, d2, tri2, _, _ = H.debug_mesh(r, W, HH)
    mesh, g = oracle_mesh(with_t, t, s, rq)                              # same geometry, material and textures
    gen = pynmr.mikk_tangents(g["positions"], g["normals"], g["texcoords"], g["indices"])
    assert np.array_equal(pynmr.parse_gltf(without, tangents=True)["tangents"].view(np.uint32), gen.view(np.uint32))
    file_tangents_picture = mesh.render(c12, 2 * W, 2 * HH)[0]
    mesh.set_tangents(g["normals"], gen, s, rq)
    want_rgba2, want_d2, want_tri = mesh.render(c12, 2 * W, 2 * HH)
    hit = want_tri >= 0
    assert hit.mean() > 0.002 and np.array_equal(tri2, want_tri)
    assert np.array_equal(d2[hit].view(np.uint32), want_d2[hit].view(np.uint32))
    assert float(np.abs(rgba2 - want_rgba2).max()) <= 2e-4
    assert float(np.abs(file_tangents_picture - want_rgba2)[hit].max()) > 1e-3      # the tangents reach the picture
