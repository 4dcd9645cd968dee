"""Random camera poses, GPU against the CPU oracle (GPU box; pairs well with the checked build, NMR_LIB=.../libnmr_checked.so):
    python tools/fuzz_poses.py [n_poses] [seed]
Every pose renders the hybrid lens scene at a small size through frame() and compares live-ray count, samples and pixels with
the oracle's render of the same camera.  Poses: orbit steps of any size, zooms from far outside to inside the head, sideways
shifts that push the head partly or wholly out of the picture, model transforms.  Prints the worst case and exits 1 on a violation."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "nerf-glasses_b200")):
    sys.path.insert(0, p)
import numpy as np
import pynmr, synth
import helpers as H

W, HH = 160, 96
TOL = 2.0 / 255.0


def run(n_poses: int = 100, seed: int = 0, with_lens: bool = True, only=None, verbose: bool = True, regime: str = "opaque", aabb_scale: int = 1, snap_seed: int = 1337, plate: bool = False, size=None):
    W, HH = size if size else (globals()["W"], globals()["HH"])
    """-> (violations, worst pixel difference, pixels over tolerance in total)"""
    rng = np.random.default_rng(seed)
    say = print if verbose else (lambda *a, **k: None)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "s.msgpack"); synth.write_snapshot(path, seed=snap_seed, log2_hashmap_size=15, regime=regime, aabb_scale=aabb_scale)
        snap = synth.read_snapshot(path)
        gltf = synth.write_lens_glasses_gltf(os.path.join(d, "lens")) if with_lens else synth.write_glasses_gltf(os.path.join(d, "mesh"))
        g = {"path": gltf, "t": synth.GLASSES_T, "s": synth.GLASSES_S, "r": synth.GLASSES_R_WXYZ,
             "texture": np.tile(np.array([128, 128, 128, 255], dtype=np.uint8), (4, 4, 1))}
        r = pynmr.NerfMeshRenderer(W, HH, 0)
        nerf = r.load_nerf(path)
        r.load_mesh(gltf, t=g["t"], s=g["s"], r=g["r"])
        H.set_flags(r, 0)
        base = r.view_projection_mat.copy()
        worst = {"pix": 0.0, "psnr": 999.0}
        over_total = 0
        alive_all, lens_all = [], []
        bad = 0
        for k in range(n_poses):
            r.view_projection_mat = base
            r.orbit(float(rng.uniform(-3, 3)), float(rng.uniform(-1.2, 1.2)), float(rng.uniform(-2, 5.3)))
            m = r.view_projection_mat
            if rng.random() < 0.6:                      # dolly towards / into / through the head
                m[:, 3] += float(rng.uniform(0.0, 1.1)) * m[:, 2] * float(np.linalg.norm(m[:, 3]))
                r.view_projection_mat = m
            if rng.random() < 0.4:
                m[:, 3] += float(rng.uniform(-1.0, 1.0)) * m[:, 0] + float(rng.uniform(-0.6, 0.6)) * m[:, 1]
                r.view_projection_mat = m
            if plate and with_lens:                     # the plate model of the lens panes (two Snell interfaces), a random thickness per pose
                th = float(rng.uniform(0.004, 0.06))
                r.set_lens_model(1, th); g["lens_model"] = 1; g["lens_thickness"] = th
            c12 = np.ascontiguousarray(r.view_projection_mat.T.reshape(-1))
            if only is not None and k != only:
                continue
            assert r.frame()
            img = np.asarray(r.read_frame()).copy()
            st = r.stats()
            want, _, ns, ost, (osurf, ots) = H.oracle_scene(snap, W, HH, c12, glasses=g, n_steps_mode=1)
            _, _, gns = H.debug_last_frame(r, W, HH)
            diff = np.abs(img - want)
            px = float(diff.max()); ps = H.psnr(img, want)
            frac = float(np.mean(diff.max(axis=2) > TOL))
            alive_ok = st["rays_alive"] == ost["alive_after_first_hit"]
            alive_all.append(st["rays_alive"]); lens_all.append(int((ost["lens"]["w"] > 0).sum()) if "lens" in ost else 0)
            worst["pix"] = max(worst["pix"], px); worst["psnr"] = min(worst["psnr"], ps)
            # (lens frames: a pixel whose reflected segment starts a sample apart may differ more; the tests allow 0.4 % of them)
            if px > TOL:
                over = diff.max(axis=2) > TOL
                over_total += int(over.sum())
                lw = ost["lens"]["w"] if "lens" in ost else np.zeros((HH, W), np.float32)
                ys, xs = np.nonzero(over)
                say(f"pose {k}: alive {st['rays_alive']} ({st['rays_alive'] / (W * HH):.3f} of the pixels), {int(over.sum())} pixels over tolerance, {int((over & (lw > 0)).sum())} of them lens pixels, "
                      f"{int((over & (osurf[..., 3] > 0)).sum())} mesh pixels, {int((over & (gns != ns)).sum())} with another sample count; first: " +
                      "; ".join(f"({x},{y}) gpu {np.round(img[y, x, :3], 3).tolist()} oracle {np.round(want[y, x, :3], 3).tolist()} lens w {lw[y, x]:.2f} mesh w {osurf[y, x, 3]:.2f} t_surface {ots[y, x]:.4f} ns {int(gns[y, x])}/{int(ns[y, x])}" for y, x in list(zip(ys, xs))[:3]), flush=True)
            if only is not None and px > TOL:
                gw, gt, gn = H.debug_lens(r, W, HH)
                _, _, _, gsurf, gts = H.debug_mesh(r, W, HH)
                L = ost.get("lens")
                for y, x in list(zip(ys, xs))[:6]:
                    say(f"  ({x},{y}): lens w gpu {gw[y, x]:.3f} oracle {L['w'][y, x]:.3f} | lens t gpu {gt[y, x]:.6f} oracle {L['t'][y, x]:.6f} | normal gpu {np.round(gn[y, x], 4).tolist()} oracle {np.round(L['n'][y, x], 4).tolist()}"
                          f" | surf gpu {np.round(gsurf[y, x], 4).tolist()} oracle {np.round(osurf[y, x], 4).tolist()} | t_surface gpu {gts[y, x]:.6f} oracle {ots[y, x]:.6f}")
            if not alive_ok or ps < 45.0 or frac > 0.004:
                bad += 1
                say(f"pose {k}: alive {st['rays_alive']} vs {ost['alive_after_first_hit']}, max |d| {px:.4f}, psnr {ps:.1f} dB, pixels over tolerance {frac:.4%}", flush=True)
        a = np.array(alive_all); l = np.array(lens_all)
        say(f"{n_poses} poses, seed {seed}, scene {regime} / aabb_scale {aabb_scale} / snapshot seed {snap_seed}{" / plate lenses" if plate else ""} / {W}x{HH}: violations {bad}, worst pixel difference {worst['pix']:.4f}, worst psnr {worst['psnr']:.1f} dB; live rays per pose: "
              f"min {a.min()} median {int(np.median(a))} max {a.max()} of {W * HH} ({int((a == 0).sum())} poses see nothing), lens pixels median {int(np.median(l))} max {l.max()}")
        return bad, worst["pix"], over_total


if __name__ == "__main__":
    n_poses = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    with_lens = (sys.argv[3] != "nolens") if len(sys.argv) > 3 else True
    only = int(os.environ["FUZZ_ONLY"]) if "FUZZ_ONLY" in os.environ else None      # one pose of the sequence, with the hand-offs of the offending pixels
    kw = dict(a.split("=") for a in sys.argv[4:])        # regime=translucent aabb_scale=4 snap_seed=7
    sys.exit(1 if run(n_poses, seed, with_lens, only, regime=kw.get("regime", "opaque"), aabb_scale=int(kw.get("aabb_scale", 1)), snap_seed=int(kw.get("snap_seed", 1337)), plate=kw.get("plate", "0") == "1", size=tuple(int(v) for v in kw["size"].split("x")) if "size" in kw else None)[0] else 0)
