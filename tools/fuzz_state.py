"""State-machine fuzz (GPU box): a renderer is driven through a random sequence of API calls - camera, crop box, background,
tonemap curve, minimum transmittance, model transform, mesh translation / scale, lenses on / off / plate, surface rule, overlap,
shards, renders at other sizes and formats, render_views, render_update, accumulated frames, probes - and then has to produce,
for its final state, exactly the picture a fresh renderer produces when it is put into that state directly.  Finds state that
sticks where it should not (caches keyed on too little, counters not reset, surfaces not resized).  A failing sequence is
reduced to a minimal sub-sequence (greedy removal of calls) before it is printed.
    python tools/fuzz_state.py [n_sequences] [seed]"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "nerf-glasses_b200")):
    sys.path.insert(0, p)
import numpy as np
import pynmr, synth

W, HH = 256, 256
DT = {"f32": np.float32, "u8": np.uint8, "f16": np.float16}


def make_ops(rng, box0):
    """a random call sequence: [(name, args)] - all arguments drawn here, so that any sub-sequence can be replayed"""
    ops = []
    for _ in range(int(rng.integers(8, 30))):
        k = int(rng.integers(0, 22))
        if k == 0: ops.append(("orbit", (float(rng.uniform(-1, 1)), float(rng.uniform(-0.5, 0.5)), float(rng.uniform(-1, 4)))))
        elif k == 1: ops.append(("dolly", (float(rng.uniform(0, 0.8)),)))
        elif k == 2:
            c = rng.uniform(0.35, 0.65, 3); h = rng.uniform(0.1, 0.45, 3)
            ops.append(("aabb", (np.maximum(box0[0], c - h).astype(np.float32), np.minimum(box0[1], c + h).astype(np.float32))))
        elif k == 3: ops.append(("aabb", box0))
        elif k == 4: ops.append(("background", ([float(v) for v in rng.uniform(0, 1, 3)] + [1.0],)))
        elif k == 5: ops.append(("curve", (int(rng.integers(0, 4)),)))
        elif k == 6: ops.append(("min_t", (float(rng.choice([0.01, 0.05, 0.2])),)))
        elif k == 7: ops.append(("model", (rng.uniform(-0.1, 0.1, 3).astype(np.float32), rng.uniform(-0.1, 0.1, 3).astype(np.float32))))
        elif k == 8: ops.append(("mesh_t", ([float(v) for v in np.array(synth.GLASSES_T) + rng.uniform(-0.05, 0.05, 3)],)))
        elif k == 9: ops.append(("mesh_s", ([float(v) for v in np.array(synth.GLASSES_S) * rng.uniform(0.8, 1.3, 3)],)))
        elif k == 10: ops.append(("lens_on", (bool(rng.integers(0, 2)),)))
        elif k == 11: ops.append(("lens_model", (int(rng.integers(0, 2)), float(rng.uniform(0.005, 0.05)))))
        elif k == 12: ops.append(("surface", (int(rng.integers(0, 3)),)))
        elif k == 13: ops.append(("overlap", (bool(rng.integers(0, 2)),)))
        elif k == 14: ops.append(("frames", (int(rng.integers(1, 4)),)))
        elif k == 15: ops.append(("render", (int(rng.integers(17, 400)), int(rng.integers(9, 400)), int(rng.integers(1, 3)), bool(rng.integers(0, 2)), str(rng.choice(["f32", "u8", "f16"])))))
        elif k == 16: ops.append(("views", (int(rng.integers(1, 5)), int(rng.integers(32, 200)), int(rng.integers(32, 200)), str(rng.choice(["f32", "u8"])))))
        elif k == 17: ops.append(("update", ()))
        elif k == 18: ops.append(("shard", (int(rng.integers(0, 3)),)))
        elif k == 19: ops.append(("probes", (rng.uniform(-0.3, 0.3, (64, 3)).astype(np.float32),)))
        elif k == 20: ops.append(("floaties", ()))
        else: ops.append(("render", (W, HH, 1, False, "u8")))
    return ops


def apply_state(r, nerf, mesh, st):
    """puts a renderer into state `st` with the public setters, in a fixed order"""
    nerf.render_aabb = pynmr.BoundingBox(st["aabb"][0], st["aabb"][1])
    nerf.background_color = st["background"]
    nerf.tonemap_curve = st["curve"]
    nerf.nerf.render_min_transmittance = st["min_t"]
    nerf.model_translation = st["model_t"]; nerf.model_rotation = st["model_r"]
    mesh.nodes[0].translation = st["mesh_t"]; mesh.nodes[0].scale = st["mesh_s"]
    if st["floaties"]: r.remove_floaties()
    r.set_lens(st["lens_on"]); r.set_lens_model(st["lens_model"], st["lens_thickness"])
    r.set_surface_insertion(st["surface"])
    r.set_overlap(st["overlap"])
    r.view_projection_mat = st["cam"]


def run(n_seq: int = 20, seed: int = 0, verbose: bool = True):
    say = print if verbose else (lambda *a, **k: None)
    failures = []
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "s.msgpack"); synth.write_snapshot(path, seed=1337, log2_hashmap_size=15)
        gltf = synth.write_lens_glasses_gltf(os.path.join(d, "lens"))

        def fresh():
            r = pynmr.NerfMeshRenderer(W, HH, 0)
            nerf = r.load_nerf(path)
            mesh = r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ)
            return r, nerf, mesh
        r0, n0, _ = fresh()
        box0 = (n0.render_aabb.min.copy(), n0.render_aabb.max.copy())
        cam0 = r0.view_projection_mat.copy()
        bg0 = [float(v) for v in n0.background_color]
        del r0, n0

        def execute(ops, report=False):
            r, nerf, mesh = fresh()
            st = {"aabb": box0, "background": bg0, "curve": 0, "min_t": 0.01, "model_t": np.zeros(3, np.float32), "model_r": np.zeros(3, np.float32),
                  "mesh_t": list(synth.GLASSES_T), "mesh_s": list(synth.GLASSES_S), "lens_on": True, "lens_model": 0, "lens_thickness": 0.0,
                  "surface": 0, "overlap": True, "cam": cam0, "floaties": False}
            kept = None
            for name, a in ops:
                if name == "orbit": r.orbit(*a); st["cam"] = r.view_projection_mat.copy()
                elif name == "dolly":
                    m = r.view_projection_mat; m[:, 3] += a[0] * m[:, 2] * float(np.linalg.norm(m[:, 3])); r.view_projection_mat = m; st["cam"] = m.copy()
                elif name == "aabb": st["aabb"] = a; nerf.render_aabb = pynmr.BoundingBox(*a)
                elif name == "background": st["background"] = a[0]; nerf.background_color = a[0]
                elif name == "curve": st["curve"] = a[0]; nerf.tonemap_curve = a[0]
                elif name == "min_t": st["min_t"] = a[0]; nerf.nerf.render_min_transmittance = a[0]
                elif name == "model": st["model_t"], st["model_r"] = a; nerf.model_translation = a[0]; nerf.model_rotation = a[1]
                elif name == "mesh_t": st["mesh_t"] = a[0]; mesh.nodes[0].translation = a[0]
                elif name == "mesh_s": st["mesh_s"] = a[0]; mesh.nodes[0].scale = a[0]
                elif name == "lens_on": st["lens_on"] = a[0]; r.set_lens(a[0])
                elif name == "lens_model": st["lens_model"], st["lens_thickness"] = a; r.set_lens_model(*a)
                elif name == "surface": st["surface"] = a[0]; r.set_surface_insertion(a[0])
                elif name == "overlap": st["overlap"] = a[0]; r.set_overlap(a[0])
                elif name == "frames":
                    for _ in range(a[0]): assert r.frame()            # accumulates on a still camera
                elif name == "render": nerf.render(a[0], a[1], a[2], linear=a[3], dtype=DT[a[4]])
                elif name == "views": r.render_views(nerf, np.stack([st["cam"]] * a[0]), a[1], a[2], dtype=DT[a[3]])
                elif name == "update": kept = nerf.render_update(kept, W, HH, linear=False)
                elif name == "shard": r.set_shard(a[0], 3, 16); r.frame(); r.set_shard(0, 1, 16)
                elif name == "floaties": st["floaties"] = True; r.remove_floaties()
                elif name == "probes": nerf.probe_points(a[0], [0, -1, 0]); nerf.probe_rays(a[0], [0, -1, 0])
            # the picture of the final state, by the driven renderer and by a fresh one
            r.view_projection_mat = st["cam"]
            ok = r.frame(); got_frame = np.asarray(r.read_frame()).copy()
            got = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
            got_u8 = np.asarray(nerf.render(W, HH, 1, linear=False, dtype=np.uint8)).copy()
            upd = np.asarray(nerf.render_update(kept, W, HH, linear=False)).copy() if kept is not None else got
            r2, n2, m2 = fresh()
            apply_state(r2, n2, m2, st)
            assert r2.frame()
            want_frame = np.asarray(r2.read_frame()).copy()
            want = np.asarray(n2.render(W, HH, 1, linear=False)).copy()
            bad = []
            if not (ok and np.array_equal(got_frame.view(np.uint32), want_frame.view(np.uint32))): bad.append("frame()")
            if not np.array_equal(got.view(np.uint32), want.view(np.uint32)): bad.append("render()")
            if not np.array_equal(got_u8, np.uint8(np.clip(want, 0.0, 1.0) * np.float32(255.0))): bad.append("render(uint8)")
            if not np.array_equal(upd.view(np.uint32), want.view(np.uint32)): bad.append("render_update()")
            if bad and report:
                for name, a, b in (("render", got, want), ("frame", got_frame, want_frame)):
                    m = (a.view(np.uint32) != b.view(np.uint32)).any(axis=2)
                    if m.any():
                        ys, xs = np.nonzero(m)
                        say(f"  {name}: {int(m.sum())} pixels differ, rows {ys.min()}..{ys.max()} columns {xs.min()}..{xs.max()}, max |d| {float(np.abs(a - b).max()):.5f}")
            return bad

        for s in range(n_seq):
            ops = make_ops(np.random.default_rng([seed, s]), box0)
            bad = execute(ops)
            if bad:
                # greedy reduction: drop every call whose removal keeps the sequence failing
                i = 0
                while i < len(ops):
                    trial = ops[:i] + ops[i + 1:]
                    if execute(trial): ops = trial
                    else: i += 1
                say(f"sequence {s}: {bad} differ from a fresh renderer; minimal failing calls:", flush=True)
                for name, a in ops:
                    say("   ", name, [np.round(v, 4).tolist() if isinstance(v, np.ndarray) and v.size <= 6 else (v if not isinstance(v, np.ndarray) else f"array{v.shape}") for v in a])
                execute(ops, report=True)
                failures.append((s, bad, [n for n, _ in ops]))
    say(f"{n_seq} sequences, seed {seed}: {len(failures)} failures")
    return failures


if __name__ == "__main__":
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 20, int(sys.argv[2]) if len(sys.argv) > 2 else 0) else 0)
