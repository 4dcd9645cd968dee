"""Where do lens frames differ from the oracle?  (development aid)"""
import os, sys, tempfile, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + "/tools", ROOT + "/nerf-glasses_b200", ROOT + "/tests"): sys.path.insert(0, p)
import helpers as H, pynmr, synth
W, HH = 192, 108
with tempfile.TemporaryDirectory() as d:
    sp = os.path.join(d, "s.msgpack"); synth.write_snapshot(sp, seed=1337, log2_hashmap_size=15)
    snap = synth.read_snapshot(sp)
    gltf = synth.write_lens_glasses_gltf(os.path.join(d, "m"))
    r = pynmr.NerfMeshRenderer(W, HH); nerf = r.load_nerf(sp)
    r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ)
    r.orbit(0.35, -0.2, 4.0)
    H.set_flags(r, 0)
    cam12 = np.ascontiguousarray(r.view_projection_mat.T.reshape(-1))
    g = {"path": gltf, "t": synth.GLASSES_T, "s": synth.GLASSES_S, "r": synth.GLASSES_R_WXYZ, "texture": np.tile(np.array([128, 128, 128, 255], dtype=np.uint8), (4, 4, 1))}
    for mode_name, smode, omode in (("auto/batch8", 0, 1), ("exact", 1, 0)):
        r.set_surface_insertion(smode)
        want, fr, ns, st, (surf, ts) = H.oracle_scene(snap, W, HH, cam12, glasses=g, n_steps_mode=omode)
        img = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
        gfr, gdp, gns = H.debug_last_frame(r, W, HH)
        L = st["lens"]
        dd = np.abs(img - want).max(axis=2)
        bad = dd > 2 / 255
        lw = L["w"]
        print(f"== {mode_name}: max {dd.max():.4f} bad {int(bad.sum())}; lens px {int((lw > 0).sum())}; bad by class: lens {int((bad & (lw > 0)).sum())} nonlens {int((bad & (lw == 0)).sum())}")
        print("   n_samples equal on lens px:", float(np.mean(gns[lw > 0] == ns[lw > 0])), " on others:", float(np.mean(gns[lw == 0] == ns[lw == 0])))
        ys, xs = np.nonzero(bad)
        for y, x in list(zip(ys, xs))[:8]:
            print(f"   px ({x},{y}) w={lw[y, x]:.2f} t_lens={L['t'][y, x]:.4f} surf_w={surf[y, x, 3]:.2f} ts={ts[y, x]:.4f} ns gpu/orc {gns[y, x]}/{ns[y, x]} frame gpu {np.round(gfr[y, x], 4)} orc {np.round(fr[y, x], 4)}")
