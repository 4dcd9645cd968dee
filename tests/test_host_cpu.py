"""CPU-side tests: the C ABI library loads and exports everything include/nmr.h declares, the oracle agrees with the
numpy restatement, fixtures are well-formed, and sharding arithmetic is consistent.  No GPU compute here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol():
    import pynmr
    header = open(os.path.join(ROOT, "include", "nmr.h")).read()
    declared = set(re.findall(r"NMR_API\s+[\w\s\*]+?\b(nmr_\w+)\s*\(", header))
    assert len(declared) >= 35
    assert declared == set(pynmr.EXPORTED_SYMBOLS)
    lib = C.CDLL(pynmr.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_create_fails_loudly_without_gpu():
    """No CPU fallback: without a usable sm_100 device construction raises (skipped where a GPU is present)."""
    import pynmr
    import helpers
    if helpers.gpu_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        pynmr.NerfMeshRenderer(64, 64)
    assert b"CUDA" in pynmr.lib().nmr_last_error(None) or b"device" in pynmr.lib().nmr_last_error(None)


def test_product_does_not_reference_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "nerf-glasses_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"oracle[/\.]|nmr_oracle|liboracle", txt):
                    bad.append(f)
    assert not bad, bad


def test_oracle_encoding_matches_numpy(small_snapshot):
    import synth
    from oracle import oracle as O
    _, snap = small_snapshot
    m = O.Model.from_snapshot(snap)
    net = synth.NetParams(snap["params"], 16, snap["log2_hashmap_size"], 16)
    net.scales = m.level_table()[1]      # libm exp2f vs numpy exp2 differ by 1 ulp on two levels
    rng = np.random.default_rng(0)
    pos = rng.uniform(0, 1, (3000, 3)).astype(np.float32)
    assert np.array_equal(m.encode(pos).view(np.uint16), synth.np_encode(net, pos).view(np.uint16))
    d = rng.uniform(0, 1, (3000, 3)).astype(np.float32)
    a = m.network(pos, d).astype(np.float32); b = synth.np_network(net, pos, d).astype(np.float32)
    assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)) < 2e-3      # numpy matmul sums in another order


def test_oracle_level_table_matches_survey(small_snapshot):
    import synth
    offs, _, res = synth.level_table(16, 16, synth.per_level_scale(16, 16), 19)
    sizes = np.diff(offs.astype(np.int64))
    assert list(sizes[:5]) == [4096, 12168, 29792, 79512, 205384] and all(s == 524288 for s in sizes[5:])
    assert synth.n_params_for() == 12206480 and int(res[-1]) == 2048


def test_oracle_render_modes_agree_without_mesh(small_snapshot):
    """Without a mesh, compositing does not depend on the wavefront batching rule (n = 1 vs the reference's 1..8)."""
    from oracle import oracle as O
    _, snap = small_snapshot
    m = O.Model.from_snapshot(snap)
    cam = O.OrbitCamera(96, 54); cam.orbit(0.3, -0.1, 4.0)
    kw = dict(aabb_min=snap["render_aabb_min"], aabb_max=snap["render_aabb_max"])
    f0, _, n0, s0 = m.render_frame(m.params_struct(96, 54, cam.matrix(), n_steps_mode=0, **kw))
    f1, _, n1, s1 = m.render_frame(m.params_struct(96, 54, cam.matrix(), n_steps_mode=1, **kw))
    assert s0["alive_after_first_hit"] > 200 and np.array_equal(f0, f1)
    assert s1["samples"] >= s0["samples"] and s1["iterations"] < s0["iterations"]


def test_oracle_hybrid_modes_differ_only_on_mesh_pixels(small_snapshot, glasses_gltf):
    """n_steps_mode 1 replays the reference's wavefront loop (the surface enters in front of the n_steps batch whose end passed
    it), n_steps_mode 0 inserts it at the exact sample: the two may only differ where a mesh surface is in play."""
    import helpers
    import synth
    from oracle import oracle as O
    _, snap = small_snapshot
    cam = O.OrbitCamera(96, 54); cam.orbit(0.3, -0.1, 4.0)
    g = {"path": glasses_gltf, "t": synth.GLASSES_T, "s": synth.GLASSES_S, "r": synth.GLASSES_R_WXYZ}
    a = helpers.oracle_scene(snap, 96, 54, cam.matrix(), glasses=g, n_steps_mode=0)
    b = helpers.oracle_scene(snap, 96, 54, cam.matrix(), glasses=g, n_steps_mode=1)
    mesh_px = a[4][1] > 0
    assert float(mesh_px.mean()) > 0.003
    assert np.array_equal(a[0][~mesh_px], b[0][~mesh_px])
    assert helpers.psnr(a[0], b[0]) > 30.0


def test_glasses_fixture_roundtrip(glasses_gltf):
    import synth
    m = synth.read_gltf(glasses_gltf)
    assert m["positions"].shape == (1864, 3) and m["indices"].shape == (8856,) and int(m["indices"].max()) == 1863
    assert abs(m["roughness"] - 0.525658369064331) < 1e-12 and m["metallic"] == 0.0


def test_shard_rows_partition():
    for H in (54, 108, 1080, 2160):
        for world in (1, 2, 3, 4, 8):
            owned = [[y for y in range(H) if (y // 8) % world == r] for r in range(world)]
            assert sorted(sum(owned, [])) == list(range(H))


def _lattice_advance_np(t, K):
    """numpy restatement of device_common.cuh: lattice_advance (same integer arithmetic, same fp32 operations)."""
    dt0 = np.float32(np.float32(1.73205080757) / np.float32(1024.0))
    t = np.float32(t)
    while K > 0:
        b = int(t.view(np.uint32))
        e = b & 0xFF800000
        inc = int((np.uint32(e).view(np.float32) + dt0).view(np.uint32)) - e
        room = (e | 0x007FFFFF) - b
        fit = int(np.float32(room) / np.float32(inc)) - 2
        if K <= fit:
            return np.uint32(b + K * inc).view(np.float32)
        if fit > 0:
            t = np.uint32(b + fit * inc).view(np.float32)
            K -= fit
        while True:
            t = np.float32(t + dt0); K -= 1
            if not (K > 0 and (int(t.view(np.uint32)) & 0xFF800000) == e):
                break
    return t


def _uniform_steps_past_np(t, t_target):
    """numpy restatement of device_common.cuh: uniform_steps_past (same integer arithmetic; the rounded-down quotient is taken a
    little low, which the two corrections absorb exactly as on the device).  Returns (landing, took_the_direct_path)."""
    f32 = np.float32
    dt0 = f32(np.sqrt(f32(3.0)) / f32(1024.0))
    b = int(np.array(t, f32).view(np.uint32)); e = b & 0xFF800000; bt = int(np.array(t_target, f32).view(np.uint32))
    inc = (int(f32(np.array(e, np.uint32).view(f32) + dt0).view(np.uint32)) - e) & 0xFFFFFFFF
    if (bt & 0xFF800000) == e and 0 < inc <= 0x400000:
        k = 1
        if bt > b:
            diff = bt - b
            k = int(np.float64(diff) * np.float64(np.nextafter(f32(1.0) / f32(inc), f32(0))))
            for _ in range(2):
                if k * inc < diff:
                    k += 1
        d = (bt - b) & 0xFFFFFFFF
        r = b + k * inc
        if k * inc >= d and (k == 1 or (k - 1) * inc < d) and r <= (e | 0x7FFFFF):
            return np.array(r, np.uint32).view(f32)[()], True
    while True:
        t = f32(t + dt0)
        if not (t < t_target):
            return t, False


def test_direct_landing_equals_the_step_loop():
    """advance_to_next_voxel with a zero cone angle lands on `do { t += dt0; } while (t < t_target);` without running it
    (device_common.cuh: uniform_steps_past).  Against the loop itself in fp32, bit for bit: random t in every binade a walk visits,
    voxel distances of every cascade, t exactly on and just above powers of two, and zero distances."""
    f32 = np.float32
    dt0 = f32(np.sqrt(f32(3.0)) / f32(1024.0))
    rng = np.random.default_rng(5)
    direct = 0
    for i in range(60000):
        t = f32(rng.uniform(0.001, 8.0)) if i % 3 else f32(2.0 ** int(rng.integers(-3, 3))) * f32(1 + rng.random() * 1e-3 * (i % 2))
        dist = f32(rng.random() * 0.0136 * 2 ** int(rng.integers(0, 8))) if i % 7 else f32(0)
        tt = f32(t + dist)
        want = t
        while True:
            want = f32(want + dt0)
            if not (want < tt):
                break
        got, fast = _uniform_steps_past_np(t, tt)
        direct += fast
        assert np.array(got, f32).view(np.uint32) == np.array(want, f32).view(np.uint32), (t, tt, got, want)
    assert direct > 40000          # the direct path is the common one


def test_uniform_step_lattice_is_exact():
    """The first-hit walk's empty-space jumps (device_common.cuh: lattice_advance) rest on `t += dt0` advancing the bit pattern
    of t by a constant inside a binade: no rounding tie for dt0 = sqrt(3)/1024 in any binade a walk can reach, and the
    closed form equals the K-fold fp32 loop bit for bit, across binade boundaries."""
    dt0 = np.float32(np.float32(1.73205080757) / np.float32(1024.0))
    for e in range(-8, 9):
        frac = (np.float64(dt0) / 2.0 ** (e - 23)) % 1.0
        assert abs(frac - 0.5) > 1e-6, (e, frac)
    rng = np.random.default_rng(7)
    for _ in range(300):
        t0 = np.float32(rng.uniform(0.01, 6.0))
        K = int(rng.integers(1, 1200))
        t = t0
        for _ in range(K):
            t = np.float32(t + dt0)
        got = _lattice_advance_np(t0, K)
        assert got.view(np.uint32) == t.view(np.uint32), (t0, K, got, t)


def test_oracle_probes_are_consistent(small_snapshot):
    """Oracle restatement of the collision tool's probes (NerfTracer::intersects / collide): a collision lies in an occupied
    cell of the density grid, a ray outside the render box reports 0, and a point probe is positive only on occupied cells."""
    from oracle import oracle as O
    _, snap = small_snapshot
    m = O.Model.from_snapshot(snap)
    P = m.params_struct(64, 64, O.OrbitCamera(64, 64).matrix(), aabb_min=snap["render_aabb_min"], aabb_max=snap["render_aabb_max"])
    rng = np.random.default_rng(3)
    n = 512
    org = np.stack([rng.uniform(-0.3, 0.3, n), np.full(n, 0.45), rng.uniform(-0.3, 0.3, n)], axis=1).astype(np.float32)
    d = np.array([0.0, -1.0, 0.0], dtype=np.float32)
    dist = m.probe_rays(P, org, d)
    hit = dist > 0
    assert 50 < hit.sum() < n - 50
    bits = m.bitfield()
    pos = org[hit] + 0.5 + d[None, :] * dist[hit, None]
    cells = np.array([O.lib().orc_cascaded_grid_idx_at(np.ascontiguousarray(q, dtype=np.float32).ctypes.data_as(C.c_void_p), 0) for q in pos], dtype=np.uint32)
    assert np.all((bits[cells // 8] >> (cells % 8)) & 1)
    assert np.array_equal(m.probe_rays(P, np.array([[5, 5, 5], [-3, 0, 0]], np.float32), d), np.zeros(2, np.float32))
    pts = rng.uniform(-0.35, 0.35, size=(800, 3)).astype(np.float32)
    alpha = m.probe_points(P, pts, d)
    pc = np.array([O.lib().orc_cascaded_grid_idx_at(np.ascontiguousarray(q + 0.5, dtype=np.float32).ctypes.data_as(C.c_void_p), 0) for q in pts], dtype=np.uint32)
    occ = ((bits[pc // 8] >> (pc % 8)) & 1).astype(bool)
    assert np.all(alpha[~occ] == 0) and (alpha[occ] > 0).mean() > 0.95


# ---- the glTF / PNG loader on malformed input (host-only entry point nmr_debug_parse_gltf) ---------------------
def _tiny_gltf(tmp_path, mutate=None, name="m.gltf"):
    """A one-triangle glTF with embedded buffers; `mutate(doc)` breaks it."""
    import base64
    import json
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype="<f4").tobytes()
    nrm = np.array([[0, 0, 1]] * 3, dtype="<f4").tobytes()
    uv = np.zeros((3, 2), dtype="<f4").tobytes()
    idx = np.array([0, 1, 2, 0], dtype="<u2").tobytes()
    blob = pos + nrm + uv + idx
    doc = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
           "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}, "indices": 3}]}],
           "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"},
                         {"bufferView": 1, "componentType": 5126, "count": 3, "type": "VEC3"},
                         {"bufferView": 2, "componentType": 5126, "count": 3, "type": "VEC2"},
                         {"bufferView": 3, "componentType": 5123, "count": 3, "type": "SCALAR"}],
           "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 36},
                           {"buffer": 0, "byteOffset": 72, "byteLength": 24}, {"buffer": 0, "byteOffset": 96, "byteLength": 8}],
           "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}]}
    if mutate:
        mutate(doc)
    p = tmp_path / name
    p.write_text(json.dumps(doc))
    return str(p)


def test_gltf_loader_accepts_the_fixtures(tmp_path, glasses_gltf):
    import pynmr
    import synth
    assert pynmr.parse_gltf(_tiny_gltf(tmp_path))["triangles"] == 1
    g = pynmr.parse_gltf(glasses_gltf)
    assert (g["vertices"], g["triangles"], g["lens_triangles"], g["texture"]) == (1864, 2952, 0, (4, 4))
    lens = pynmr.parse_gltf(synth.write_lens_glasses_gltf(str(tmp_path / "lens")))
    assert lens["triangles"] == 2952 + 8 and lens["lens_triangles"] == 8


@pytest.mark.parametrize("case", ["negative_view_offset", "huge_view_length", "negative_accessor_offset", "negative_count", "count_overflow",
                                  "short_normals", "short_texcoords", "node_cycle", "node_self_child", "bad_buffer_index", "bad_accessor_index",
                                  "index_out_of_range", "negative_scene", "huge_stride"])
def test_gltf_loader_rejects_malformed_files(tmp_path, case):
    """Untrusted offsets / lengths / counts / indices: every one of these used to crash, hang or read out of bounds
    (negative byteOffset passing an unsigned bounds check, NORMAL shorter than POSITION, a child list leading back to its parent)."""
    import pynmr

    def mutate(d):
        if case == "negative_view_offset": d["bufferViews"][0]["byteOffset"] = -64
        elif case == "huge_view_length": d["bufferViews"][1]["byteLength"] = 2 ** 62
        elif case == "negative_accessor_offset": d["accessors"][0]["byteOffset"] = -12
        elif case == "negative_count": d["accessors"][3]["count"] = -3
        elif case == "count_overflow": d["accessors"][0]["count"] = 2 ** 61
        elif case == "short_normals": d["accessors"][1]["count"] = 1
        elif case == "short_texcoords": d["accessors"][2]["count"] = 2
        elif case == "node_cycle": d["nodes"] = [{"children": [1]}, {"children": [0], "mesh": 0}]
        elif case == "node_self_child": d["nodes"] = [{"children": [0], "mesh": 0}]
        elif case == "bad_buffer_index": d["bufferViews"][0]["buffer"] = 7
        elif case == "bad_accessor_index": d["meshes"][0]["primitives"][0]["attributes"]["POSITION"] = -1
        elif case == "index_out_of_range": d["accessors"][3]["byteOffset"] = 2      # reads indices 1, 2, 0 -> fine; then break one
        elif case == "negative_scene": d["scene"] = -1
        elif case == "huge_stride": d["bufferViews"][0]["byteStride"] = 2 ** 40

    if case == "index_out_of_range":
        import base64
        import json
        path = _tiny_gltf(tmp_path)
        d = json.loads(open(path).read())
        raw = bytearray(base64.b64decode(d["buffers"][0]["uri"].split(",", 1)[1]))
        raw[96:98] = np.array([9], dtype="<u2").tobytes()
        d["buffers"][0]["uri"] = "data:application/octet-stream;base64," + base64.b64encode(bytes(raw)).decode()
        open(path, "w").write(json.dumps(d))
    else:
        path = _tiny_gltf(tmp_path, mutate)
    with pytest.raises(RuntimeError, match="libnmr error -3"):
        pynmr.parse_gltf(path)


def test_png_palette_index_is_bounds_checked(tmp_path):
    """decode_png: palette index k with k*3+2 == palette.size() reads one byte past the palette - such a texture is
    now reported as unusable (the loader falls back to baseColorFactor and says so) instead of being read out of bounds."""
    import base64
    import struct
    import zlib
    import pynmr

    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xFFFFFFFF)

    def png(palette_bytes, index):
        ihdr = struct.pack(">IIBBBBB", 1, 1, 8, 3, 0, 0, 0)
        return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"PLTE", palette_bytes) + chunk(b"IDAT", zlib.compress(bytes([0, index]))) + chunk(b"IEND", b"")

    def with_texture(data):
        def mutate(d):
            d["materials"] = [{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}}}]
            d["meshes"][0]["primitives"][0]["material"] = 0
            d["textures"] = [{"source": 0}]
            d["images"] = [{"uri": "data:image/png;base64," + base64.b64encode(data).decode()}]
        return mutate

    ok = pynmr.parse_gltf(_tiny_gltf(tmp_path, with_texture(png(bytes([10, 20, 30, 40, 50, 60]), 1)), "ok.gltf"))
    assert ok["texture"] == (1, 1) and ok["warning"] == ""
    for pal, idx in ((bytes([10, 20, 30, 40, 50]), 1), (bytes([10, 20, 30]), 1), (bytes([10, 20]), 0)):       # k*3+2 == size, > size
        bad = pynmr.parse_gltf(_tiny_gltf(tmp_path, with_texture(png(pal, idx)), "bad.gltf"))
        assert bad["texture"] == (0, 0) and "palette" in bad["warning"]


def test_gltf_loader_tangents_and_texture_slots(tmp_path):
    """Every texture slot of the reference's closest-hit program is read (S/optix/optix_scene.cu:221-258), a TANGENT attribute is
    taken as is, and a file without one gets the tangents of the reference's generator for such primitives
    (S/gltf_scene.cpp:150-155 -> mikktspace.c), i.e. what nmr_mikk_tangents returns for the primitive's arrays."""
    import pynmr
    import synth
    with_t = synth.write_textured_glasses_gltf(str(tmp_path / "a"), with_tangents=True)
    without = synth.write_textured_glasses_gltf(str(tmp_path / "b"), with_tangents=False)
    g = synth.read_gltf(with_t)
    a = pynmr.parse_gltf(with_t, tangents=True)
    b = pynmr.parse_gltf(without, tangents=True)
    assert a["warning"] == "" and b["warning"] == "" and a["texture"] == (16, 16)
    assert np.array_equal(a["tangents"], g["tangents"])                    # read verbatim
    want = pynmr.mikk_tangents(g["positions"], g["normals"], g["texcoords"], g["indices"])
    assert np.array_equal(b["tangents"].view(np.uint32), want.view(np.uint32))
    rough = synth.np_tangents(g["positions"], g["normals"], g["texcoords"], g["indices"])  # unweighted UV-derivative average
    used = np.zeros(len(want), bool); used[np.asarray(g["indices"]).reshape(-1)] = True
    cos = np.einsum("ij,ij->i", want[used, :3], rough[used, :3])
    assert float(np.median(cos)) > 0.99 and float(np.mean(want[used, 3] == rough[used, 3])) > 0.95   # the same direction field
    n = g["normals"] / np.linalg.norm(g["normals"], axis=1, keepdims=True)
    assert float(np.abs(np.linalg.norm(want[used, :3], axis=1) - 1).max()) <= 1e-5
    # built in the plane of the vertex normal - except where the method falls back to its default frame (1, 0, 0): UV-flat triangles
    assert float(np.mean(np.abs(np.einsum("ij,ij->i", want[used, :3], n[used])) <= 2e-3)) > 0.9


def test_mikk_tangents_equal_the_references_golden_vectors():
    """nmr_mikk_tangents against tests/golden/ref_mikk.npz: the answers of the reference's own dependencies/MikkTSpace/mikktspace.c
    (tests/golden/make_ref_mikk.py) for 33 seeded meshes - triangle soups with edges shared by many triangles, welded and unwelded
    duplicates, degenerate and UV-flat triangles; sphere bands with mirrored UVs and exploded faces; the glasses mesh.  Bit for bit,
    including the method's quirks (the weld names a vertex after whichever corner its median splits leave in front; the last run of
    every edge-sort pass stays unsorted, which drops some neighbour links)."""
    import pynmr
    import synth
    gold = np.load(os.path.join(GOLDEN, "ref_mikk.npz"))
    cases = synth.tangent_test_cases()
    assert len(cases) == len(gold.files)
    for k, (p, n, uv, f) in enumerate(cases):
        got = pynmr.mikk_tangents(p, n, uv, f)
        want = gold[f"t{k}"]
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"mesh {k}: {int((got != want).any(1).sum())} of {len(p)} vertices differ"


def test_mikk_tangents_vs_reference_library_on_fresh_meshes():
    """Where oracle/_ref/libmikk_ref.so exists (the authoring container; it travels to the GPU box): 200 more random meshes, not in
    the golden file, bit for bit against the reference's mikktspace.c."""
    import ctypes as C
    import pynmr
    import synth
    lib_path = os.path.join(ROOT, "oracle", "_ref", "libmikk_ref.so")
    if not os.path.exists(lib_path):
        pytest.skip("oracle/_ref/libmikk_ref.so not built (needs /root/reference)")
    R = C.CDLL(lib_path)
    R.ref_mikk_tangents.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    rng = np.random.default_rng(99)
    for k in range(200):
        if k % 5 == 4:
            p, n, uv, f = synth.tangent_test_grid(int(rng.integers(3, 20)), int(rng.integers(3, 20)), rng, mirror=bool(k & 1), jitter=0.03, explode=bool(k & 2))
        else:
            p, n, uv, f = synth.tangent_test_soup(rng, int(rng.integers(4, 40)), int(rng.integers(1, 300)), dup=float(rng.random() * 0.5),
                                                  degen=float(rng.random() * 0.2), flat=float(rng.random() * 0.2))
        i = np.ascontiguousarray(f, np.uint32).reshape(-1)
        want = np.zeros((len(p), 4), np.float32); want[:] = (1, 0, 0, -1)
        assert R.ref_mikk_tangents(p.ctypes.data, n.ctypes.data, uv.ctypes.data, i.ctypes.data, i.size, want.ctypes.data) == 0
        got = pynmr.mikk_tangents(p, n, uv, f)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"mesh {k}"


def test_mikk_tangents_rejects_bad_arguments():
    import pynmr
    p = np.zeros((3, 3), np.float32); p[1, 0] = 1; p[2, 1] = 1
    n = np.tile(np.array([0, 0, 1], np.float32), (3, 1)); uv = p[:, :2].copy()
    with pytest.raises(RuntimeError):
        pynmr.mikk_tangents(p, n, uv, np.array([0, 1, 3], np.uint32))          # index out of range
    with pytest.raises(ValueError):
        pynmr.mikk_tangents(p, n[:2], uv, np.array([0, 1, 2], np.uint32))
    t = pynmr.mikk_tangents(p, n, uv, np.array([0, 1, 2], np.uint32))
    assert np.array_equal(t, np.tile(np.array([1, 0, 0, 1], np.float32), (3, 1)))                # ds = +x, UV orientation kept
    t = pynmr.mikk_tangents(p, n, uv, np.zeros(0, np.uint32))
    assert np.array_equal(t, np.tile(np.array([1, 0, 0, -1], np.float32), (3, 1)))               # no triangle: the default frame


def test_integration_pybind11_stub_compiles_and_binds(tmp_path):
    """INTEGRATION.md section B shows the pybind11 module a reference maintainer would compile instead of src/python_api.cu.
    The test compiles exactly that text against include/nmr.h and libnmr.so and imports the result: the stub cannot rot."""
    import subprocess
    import sys
    import sysconfig
    import pybind11
    import pynmr
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```cpp\n(// pynmr_libnmr\.cpp.*?)```", text, re.S).group(1)
    src = tmp_path / "pynmr_libnmr.cpp"
    src.write_text(code)
    libdir = os.path.dirname(pynmr.LIB_PATH)
    out = tmp_path / ("pynmr_stub" + sysconfig.get_config_var("EXT_SUFFIX"))
    code_named = code.replace("PYBIND11_MODULE(pynmr, m)", "PYBIND11_MODULE(pynmr_stub, m)")     # (do not shadow the ctypes shim)
    src.write_text(code_named)
    cmd = ["g++", "-O1", "-std=c++17", "-shared", "-fPIC", f"-I{pybind11.get_include()}", f"-I{sysconfig.get_paths()['include']}",
           f"-I{os.path.join(ROOT, 'include')}", str(src), f"-L{libdir}", "-l:libnmr.so", f"-Wl,-rpath,{libdir}", "-o", str(out)]
    subprocess.check_call(cmd)
    probe = ("import sys; sys.path.insert(0, %r); import pynmr_stub as m; "
             "assert hasattr(m, 'NerfMeshRenderer') and hasattr(m, 'Testbed') and hasattr(m, 'free_temporary_memory'); "
             "names = dir(m.NerfMeshRenderer); "
             "assert all(n in names for n in ('frame', 'load_nerf', 'load_mesh', 'remove_floaties', 'orbit', 'envmap', 'view_projection_mat')); "
             "print('ok')") % str(tmp_path)
    res = subprocess.run([sys.executable, "-c", probe], capture_output=True, text=True)
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr[-2000:]


def test_pinned_pool_is_bounded(monkeypatch):
    """The shim's pool of page-locked blocks gives the least recently used sizes back to the driver beyond its cap - also when
    sizes whose blocks are all in use sit at the front of the pool (no GPU: the library is replaced by a recorder)."""
    import pynmr
    freed = []

    class Fake:
        def nmr_host_free(self, p):
            freed.append(p)
    monkeypatch.setattr(pynmr, "lib", lambda: Fake())
    monkeypatch.setattr(pynmr, "_PINNED_POOL_CAP", 1000)
    monkeypatch.setattr(pynmr, "_pinned_pool", {300: []})
    for n, p in [(400, 1), (400, 2), (300, 3), (200, 4), (500, 5)]:
        pynmr._pinned_release(n, p)
    assert freed == [2, 1]
    assert pynmr._pinned_pool == {300: [3], 200: [4], 500: [5]}
    assert sum(n * len(b) for n, b in pynmr._pinned_pool.items()) <= 1000


def test_bounding_box_ray_intersect_equals_the_oracle():
    """pynmr.BoundingBox.ray_intersect (S/python_api.cu:256 -> S/ngp/bounding_box.cuh:106-147) against the oracle's slab test, which is
    pinned bit for bit on the reference's own BoundingBox (tests/test_oracle_golden.py, tests/test_oracle_vs_ref.py)."""
    import pynmr
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    box = pynmr.BoundingBox([-0.2, 0.15, -0.2], [1.0, 1.0, 1.0])
    bmin = np.array([-0.2, 0.15, -0.2], np.float32); bmax = np.array([1, 1, 1], np.float32)
    for i in range(2000):
        pos = rng.uniform(-2, 3, 3).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32)
        if i % 10 == 0:
            d[int(rng.integers(0, 3))] = 0.0                      # a ray parallel to a slab
        want = np.zeros(2, np.float32)
        O.lib().orc_aabb_ray_intersect(bmin.ctypes.data, bmax.ctypes.data, pos.ctypes.data, d.ctypes.data, want.ctypes.data)
        got = np.array(box.ray_intersect(pos, d), np.float32)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)) or (np.isnan(got).any() and np.isnan(want).any()), (pos, d, got, want)


def test_testbed_crop_box_and_camera_helpers_host_arithmetic():
    """The host arithmetic behind Testbed.crop_box / set_crop_box / crop_box_corners (S/ngp/testbed.cu:1421-1477 with
    NerfDataset::ngp_matrix_to_nerf / nerf_matrix_to_ngp, S/ngp/nerf_loader.cuh:115-153) and scale / look_at / view_dir
    (S/ngp/testbed.cu:1328-1349), on a Testbed whose device calls are replaced by plain members (no GPU needed)."""
    import pynmr

    class HostOnly(pynmr.Testbed):
        def __init__(self):
            self._scale = 1.5; self._up_dir = None
            self._mn = np.array([0.3, 0.15, 0.3], np.float32); self._mx = np.array([1, 1, 1], np.float32)
            self._r2l = np.eye(3, dtype=np.float32)
            self._cam = np.array([[1, 0, 0, 0.1], [0, 1, 0, 0.2], [0, 0, 1, 2.0]], np.float32)

        def _dataset(self):
            d = pynmr.NerfDataset(); d.scale = 0.33; d.offset[:] = [0.5, 0.5, 0.5]; d.up[:] = [0, 1, 0]; d.from_mitsuba = 0; d.bounding_radius = 2.0
            d.raw_aabb_min[:] = [0, 0, 0]; d.raw_aabb_max[:] = [1, 1, 1]
            return d

        def _get_render_aabb(self):
            return self._mn.copy(), self._mx.copy()

        def _set_render_aabb(self, mn, mx):
            self._mn = np.asarray(mn, np.float32); self._mx = np.asarray(mx, np.float32)

        render_aabb_to_local = property(lambda s: s._r2l.copy(), lambda s, m: setattr(s, "_r2l", np.asarray(m, np.float32).reshape(3, 3)))
        camera_matrix = property(lambda s: s._cam.copy(), lambda s, m: setattr(s, "_cam", np.asarray(m, np.float32).reshape(3, 4)))

    t = HostOnly()
    ngp = t.crop_box(nerf_space=False)
    assert np.allclose(ngp, [[0.35, 0, 0, 0.65], [0, 0.425, 0, 0.575], [0, 0, 0.35, 0.65]])
    nerf = t.crop_box(nerf_space=True)
    c = ngp[:, 3]
    assert np.allclose(nerf[:, 3], (np.array([c[2], c[0], c[1]]) - 0.5) / 0.33, atol=1e-6)          # ngp_position_to_nerf of the centre
    assert np.allclose(np.abs(nerf[:, :3]).sum(axis=0), np.array([0.35, 0.425, 0.35]) / 0.33, atol=1e-5)   # half axes scaled, axes cycled
    # a rotated box survives the round trip through dataset coordinates
    a = np.deg2rad(25.0)
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float32)
    m = np.concatenate([R * np.array([0.1, 0.2, 0.3], np.float32)[None, :], np.array([[0.5], [0.5], [0.6]], np.float32)], axis=1)
    t.set_crop_box(m, nerf_space=False)
    assert np.allclose(t.render_aabb_to_local, R.T, atol=1e-6) and np.allclose(t.crop_box(False), m, atol=1e-6)
    back = t.crop_box(True)
    t.set_crop_box(np.eye(3, 4, dtype=np.float32), nerf_space=False)
    t.set_crop_box(back, nerf_space=True)
    assert np.allclose(t.crop_box(False), m, atol=2e-6)
    corners = np.array(t.crop_box_corners(False))
    assert corners.shape == (8, 3) and np.allclose(corners.mean(axis=0), m[:, 3], atol=1e-6)
    assert np.allclose(np.linalg.norm(corners[1] - corners[0]), 0.2, atol=1e-6)                     # corner 1 - corner 0 = 2 x the first half axis
    # camera helpers
    assert np.allclose(t.look_at, [0.1, 0.2, 3.5]) and np.allclose(t.view_dir, [0, 0, 1])
    t.scale = 3.0
    assert np.allclose(t.look_at, [0.1, 0.2, 3.5]) and np.allclose(t.camera_matrix[:, 3], [0.1, 0.2, 0.5])
    t.view_dir = [1, 0, 0]
    cam = t.camera_matrix
    assert np.allclose(cam[:, 2], [1, 0, 0]) and np.allclose(cam[:, 0], [0, 0, 1]) and np.allclose(cam[:, 1], [0, -1, 0])
    assert np.allclose(t.look_at, [0.1, 0.2, 3.5], atol=1e-6)
    t.translate_camera([0, 0, 1])
    assert np.allclose(t.camera_matrix[:, 3], cam[:, 3] + 2.0 * cam[:, 2])                           # bounding_radius x the view direction
    assert t.bounding_radius == 2.0 and np.allclose(t.up_dir, [0, 1, 0])


def test_mikk_tangents_on_hostile_meshes():
    """Inputs a file can contain and a modeller never produces: every vertex identical, NaN / infinite coordinates, a fan of thousands
    of triangles around one vertex (the method compares all of a vertex's triangles pairwise; the reference recurses once per
    triangle of the fan).  The generator answers or refuses; it does not crash, hang or overflow the stack."""
    import pynmr
    import synth
    rng = np.random.default_rng(11)
    p = np.zeros((30, 3), np.float32); n = np.tile(np.array([0, 0, 1], np.float32), (30, 1)); t = np.zeros((30, 2), np.float32)
    out = pynmr.mikk_tangents(p, n, t, rng.integers(0, 30, (40, 3)).astype(np.uint32))
    assert np.array_equal(out, np.tile(np.array([1, 0, 0, -1], np.float32), (30, 1)))           # every triangle degenerate: default frames

    def fan(k):
        ang = np.linspace(0, 2 * np.pi, k + 1)[:-1]
        p = np.concatenate([[[0, 0, 0]], np.stack([np.cos(ang), np.sin(ang), 0 * ang], 1)]).astype(np.float32)
        i = np.stack([np.zeros(k, np.uint32), np.arange(1, k + 1, dtype=np.uint32), np.roll(np.arange(1, k + 1, dtype=np.uint32), -1)], 1)
        return p, np.tile(np.array([0, 0, 1], np.float32), (k + 1, 1)), (p[:, :2] * 0.5 + 0.5).astype(np.float32), i
    p, n, t, i = fan(1500)
    out = pynmr.mikk_tangents(p, n, t, i)
    assert float(np.abs(out[:, :3] - np.array([1, 0, 0], np.float32)).max()) <= 1e-3 and np.all(out[:, 3] == 1)   # a flat disc mapped by its own x, y
    with pytest.raises(RuntimeError):
        pynmr.mikk_tangents(*fan(20000))                                                         # refused, not ground through
    p, n, t, i = synth.tangent_test_soup(rng, 40, 200)
    p[3] = np.nan; p[7] = np.inf; t[5] = np.inf; n[::3] = 0
    out = pynmr.mikk_tangents(p, n, t, i)
    assert out.shape == (40, 4) and set(np.unique(out[:, 3])) <= {-1.0, 1.0}


def test_gltf_with_a_fan_the_tangent_generator_refuses_still_loads(tmp_path):
    """A primitive without TANGENT whose fan exceeds what the generator accepts: the file loads with a warning and default frames."""
    import base64
    import json
    import pynmr
    k = 9000
    ang = np.linspace(0, 2 * np.pi, k + 1)[:-1]
    pos = np.concatenate([[[0, 0, 0]], np.stack([np.cos(ang), np.sin(ang), 0 * ang], 1)]).astype("<f4")
    nrm = np.tile(np.array([0, 0, 1], "<f4"), (k + 1, 1)); uv = (pos[:, :2] * 0.5 + 0.5).astype("<f4")
    idx = np.stack([np.zeros(k, "<u4"), np.arange(1, k + 1, dtype="<u4"), np.roll(np.arange(1, k + 1, dtype="<u4"), -1)], 1)
    parts = [pos.tobytes(), nrm.tobytes(), uv.tobytes(), idx.tobytes()]
    offs = np.cumsum([0] + [len(b) for b in parts])
    doc = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
           "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}, "indices": 3}]}],
           "buffers": [{"byteLength": int(offs[-1]), "uri": "data:application/octet-stream;base64," + base64.b64encode(b"".join(parts)).decode()}],
           "bufferViews": [{"buffer": 0, "byteOffset": int(offs[j]), "byteLength": len(parts[j])} for j in range(4)],
           "accessors": [{"bufferView": 0, "componentType": 5126, "count": k + 1, "type": "VEC3", "min": [-1, -1, 0], "max": [1, 1, 0]},
                         {"bufferView": 1, "componentType": 5126, "count": k + 1, "type": "VEC3"},
                         {"bufferView": 2, "componentType": 5126, "count": k + 1, "type": "VEC2"},
                         {"bufferView": 3, "componentType": 5125, "count": 3 * k, "type": "SCALAR"}]}
    path = tmp_path / "fan.gltf"
    path.write_text(json.dumps(doc))
    g = pynmr.parse_gltf(str(path), tangents=True)
    assert g["triangles"] == k and "tangents not generated" in g["warning"]
    assert np.array_equal(g["tangents"], np.tile(np.array([1, 0, 0, -1], np.float32), (k + 1, 1)))


def test_dataset_matrix_conversions_equal_the_references_golden_vectors():
    """The shim's NeRF <-> dataset coordinate conversions behind crop_box / set_crop_box (nerf_space=True) against
    tests/golden/ref_dataset_matrix.npz: 240 matrices through the reference's own NerfDataset::nerf_matrix_to_ngp /
    ngp_matrix_to_nerf (S/ngp/nerf_loader.cuh:115-153; tests/golden/make_ref_dataset_matrix.py), with and without column scaling,
    both axis conventions - bit for bit."""
    import pynmr
    g = np.load(os.path.join(GOLDEN, "ref_dataset_matrix.npz"))

    class T(pynmr.Testbed):
        def __init__(self):
            pass

        def _dataset(self):
            return self._d
    t = T()
    for i in range(len(g["mats"])):
        d = pynmr.NerfDataset(); d.scale = float(g["scale"][i]); d.offset[:] = [float(x) for x in g["offset"][i]]; d.from_mitsuba = int(g["from_mitsuba"][i])
        t._d = d
        f = t._nerf_matrix_to_ngp if g["to_ngp"][i] else t._ngp_matrix_to_nerf
        got = f(g["mats"][i], bool(g["scale_columns"][i]))
        assert np.array_equal(got.view(np.uint32), g["out"][i].view(np.uint32)), (i, got, g["out"][i])
