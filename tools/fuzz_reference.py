"""Random camera poses (and crop boxes, and model transforms), our frame against the REFERENCE'S OWN renderer on the same GPU
(oracle/_ref/libnmr_refgpu.so, oracle/refgpu.py):
    python tools/fuzz_reference.py [n_poses] [seed]
Hybrid frames: the reference receives the mesh hand-off buffers our mesh stage produced (its OptiX stage cannot be built) and
consumes them with its own hand-off + compositing code.  The reference binary contracts to FMA, ours does not, so silhouette
pixels may take one sample more or less: per pose >= 45 dB and at most 0.4 % of the pixels over 2/255, like the fixed-pose tests."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "nerf-glasses_b200")):
    sys.path.insert(0, p)
import numpy as np
import pynmr, synth
import helpers as H

W, HH = 192, 108
TOL = 2.0 / 255.0


def run(n_poses: int = 100, seed: int = 0, verbose: bool = True, regime: str = "opaque", aabb_scale: int = 1, snap_seed: int = 1337):
    """-> (violations, worst psnr, worst fraction of pixels over tolerance)"""
    from oracle import refgpu
    rng = np.random.default_rng(seed)
    say = print if verbose else (lambda *a, **k: None)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "s.msgpack"); synth.write_snapshot(path, seed=snap_seed, log2_hashmap_size=15, regime=regime, aabb_scale=aabb_scale)
        gltf = synth.write_glasses_gltf(os.path.join(d, "mesh"))
        ref = refgpu.ReferenceRenderer(path)
        r = pynmr.NerfMeshRenderer(W, HH, 0)
        nerf = r.load_nerf(path)
        r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ)
        base = r.view_projection_mat.copy()
        box0 = nerf.render_aabb.min.copy(), nerf.render_aabb.max.copy()
        bad, worst_ps, worst_frac, alive_all = 0, 999.0, 0.0, []
        try:
            for k in range(n_poses):
                r.view_projection_mat = base
                r.orbit(float(rng.uniform(-3, 3)), float(rng.uniform(-1.2, 1.2)), float(rng.uniform(-2, 5.3)))
                m = r.view_projection_mat
                if rng.random() < 0.6:
                    m[:, 3] += float(rng.uniform(0.0, 1.1)) * m[:, 2] * float(np.linalg.norm(m[:, 3])); r.view_projection_mat = m
                if rng.random() < 0.4:
                    m[:, 3] += float(rng.uniform(-1.0, 1.0)) * m[:, 0] + float(rng.uniform(-0.6, 0.6)) * m[:, 1]; r.view_projection_mat = m
                what = []
                # a crop box through the head (Testbed.render_aabb), one pose in four
                if rng.random() < 0.25:
                    c = rng.uniform(0.35, 0.65, 3); hsz = rng.uniform(0.08, 0.45, 3)
                    mn = np.maximum(box0[0], c - hsz).astype(np.float32); mx = np.minimum(box0[1], c + hsz).astype(np.float32)
                    what.append("crop")
                else:
                    mn, mx = box0
                nerf.render_aabb = pynmr.BoundingBox(mn, mx); ref.set_render_aabb(mn, mx)
                # the model transform of the GUI sliders, one pose in four
                if rng.random() < 0.25:
                    tr = rng.uniform(-0.15, 0.15, 3).astype(np.float32); ro = rng.uniform(-0.2, 0.2, 3).astype(np.float32)
                    what.append("model")
                else:
                    tr = np.zeros(3, np.float32); ro = np.zeros(3, np.float32)
                nerf.model_translation = tr; nerf.model_rotation = ro; ref.set_model_transform(tr, ro)
                c12 = np.ascontiguousarray(r.view_projection_mat.T.reshape(-1))
                _, _, _, surf, ts = H.debug_mesh(r, W, HH)
                want, _ = ref.render(c12, W, HH, 1, False, surf=surf, ts=ts)
                got = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
                alive_all.append(r.stats()["rays_alive"])
                d = np.abs(got - want)
                ps = H.psnr(got, want); frac = float(np.mean(d.max(axis=2) > TOL))
                worst_ps = min(worst_ps, ps); worst_frac = max(worst_frac, frac)
                if ps < 45.0 or frac > 0.004:
                    bad += 1
                    say(f"pose {k} {what}: alive {alive_all[-1]} mesh pixels {int((ts > 0).sum())} max |d| {float(d.max()):.4f} psnr {ps:.1f} dB, pixels over tolerance {frac:.4%}", flush=True)
        finally:
            ref.close()
        a = np.array(alive_all)
        say(f"{n_poses} poses, seed {seed}, scene {regime} / aabb_scale {aabb_scale} / snapshot seed {snap_seed}: violations {bad}, worst psnr {worst_ps:.1f} dB, worst fraction over tolerance {worst_frac:.4%}; live rays per pose min {a.min()} median {int(np.median(a))} max {a.max()} of {W * HH}")
        return bad, worst_ps, worst_frac


if __name__ == "__main__":
    kw = dict(a.split("=") for a in sys.argv[3:])        # regime=translucent aabb_scale=4 snap_seed=7
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 0,
                      regime=kw.get("regime", "opaque"), aabb_scale=int(kw.get("aabb_scale", 1)), snap_seed=int(kw.get("snap_seed", 1337)))[0] else 0)
