"""Shared helpers of the GPU parity tests: thin ctypes calls into the parity probes of libnmr.so."""
import ctypes as C

import numpy as np


def gpu_available() -> bool:
    try:
        import pynmr
        L = pynmr.lib()
    except Exception:
        return False
    h = C.c_void_p()
    if L.nmr_create(64, 64, -1, C.byref(h)) != 0:
        return False
    L.nmr_destroy(h)
    return True


def psnr(a: np.ndarray, b: np.ndarray) -> float:
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 99.0 if mse == 0 else 10.0 * np.log10(1.0 / mse)


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def debug_encode(r, nerf, pos):
    import pynmr
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    out = np.zeros((pos.shape[0], 32), dtype=np.uint16)
    r._ck(pynmr.lib().nmr_debug_encode(r._h, nerf._id, p(pos), pos.shape[0], p(out)))
    return out


def debug_network(r, nerf, pos, dir01):
    import pynmr
    pos = np.ascontiguousarray(pos, dtype=np.float32); d = np.ascontiguousarray(dir01, dtype=np.float32)
    out = np.zeros((pos.shape[0], 4), dtype=np.uint16)
    r._ck(pynmr.lib().nmr_debug_network(r._h, nerf._id, p(pos), p(d), pos.shape[0], p(out)))
    return out.view(np.float16)


def debug_trace(r, nerf, W, H, pixels, max_samples):
    import pynmr
    pixels = np.ascontiguousarray(pixels, dtype=np.uint32); n = pixels.size
    t = np.zeros((n, max_samples), dtype=np.float32); cell = np.zeros((n, max_samples), dtype=np.uint32)
    mip = np.zeros((n, max_samples), dtype=np.uint32); pos = np.zeros((n, max_samples, 3), dtype=np.float32)
    cnt = np.zeros(n, dtype=np.uint32); ray = np.zeros((n, 8), dtype=np.float32)
    r._ck(pynmr.lib().nmr_debug_trace(r._h, nerf._id, W, H, p(pixels), n, max_samples, p(t), p(cell), p(mip), p(pos), p(cnt), p(ray)))
    return {"t": t, "cell": cell, "mip": mip, "pos": pos, "count": cnt, "ray": ray}


def debug_mesh(r, W, H, ms=2):
    import pynmr
    rgba2 = np.zeros((H * ms, W * ms, 4), dtype=np.float32); d2 = np.zeros((H * ms, W * ms), dtype=np.float32)
    tri2 = np.zeros((H * ms, W * ms), dtype=np.int32); surf = np.zeros((H, W, 4), dtype=np.float32); ts = np.zeros((H, W), dtype=np.float32)
    r._ck(pynmr.lib().nmr_debug_mesh(r._h, W, H, p(rgba2), p(d2), p(tri2), p(surf), p(ts)))
    return rgba2, d2, tri2, surf, ts


def debug_lens(r, W, H):
    """Lens hand-off per pixel: coverage, hit distance, unit normal."""
    import pynmr
    w = np.zeros((H, W), dtype=np.float32); t = np.zeros((H, W), dtype=np.float32); n = np.zeros((H, W, 3), dtype=np.float32)
    r._ck(pynmr.lib().nmr_debug_lens(r._h, W, H, p(w), p(t), p(n)))
    return w, t, n


def debug_last_frame(r, W, H):
    import pynmr
    fr = np.zeros((H, W, 4), dtype=np.float32); dp = np.zeros((H, W), dtype=np.float32); ns = np.zeros((H, W), dtype=np.uint32)
    r._ck(pynmr.lib().nmr_debug_last_frame(r._h, p(fr), p(dp), p(ns)))
    return fr, dp, ns


def get_bitfield(r, nerf):
    import pynmr
    b = np.zeros(2 * 1024 * 1024, dtype=np.uint8)
    r._ck(pynmr.lib().nmr_get_density_bitfield(r._h, nerf._id, p(b)))
    return b


KEEP_PROBES = 4   # debug flag bit 2: renders also write the probe surfaces read by debug_last_frame


def set_flags(r, flags):
    import pynmr
    r._ck(pynmr.lib().nmr_debug_set_flags(r._h, flags | KEEP_PROBES))


def oracle_scene(snap, width, height, cam12, glasses=None, spp_index=0, n_steps_mode=0, aabb=None):
    """Oracle render of the same scene -> (image srgb f32[H,W,4], frame linear, n_samples, stats, (surf, ts))."""
    import synth
    from oracle import oracle as O
    m = O.Model.from_snapshot(snap)
    amin, amax = (snap["render_aabb_min"], snap["render_aabb_max"]) if aabb is None else aabb
    P = m.params_struct(width, height, cam12, aabb_min=amin, aabb_max=amax, spp_index=spp_index, n_steps_mode=n_steps_mode)
    surf = ts = lens = None
    if glasses is not None:
        g = synth.read_gltf(glasses["path"])
        mesh = O.Mesh(g["positions"], g["normals"], g["texcoords"], g["indices"], glasses["t"], glasses["s"], glasses["r"],
                      g["base_color"], g["metallic"], g["roughness"], (0, 0, 0), glasses.get("texture"))
        if g["lens"] is not None and glasses.get("lens", True):
            # lens surfaces (new functionality): two visibility layers, lens hand-off, secondary rays in the oracle
            mesh.set_lens(g["tri_lens"])
            rgba2, d2, ld2, ln2 = mesh.render_layers(cam12, 2 * width, 2 * height)
            surf, ts = O.mesh_resolve(rgba2, d2, width, height, 2)
            lw, lt, lnn = O.lens_resolve(d2, ld2, ln2, ts, width, height, 2)
            lp = g["lens"]
            lens = {"w": lw, "t": lt, "n": lnn, "f0": O.lens_f0(lp["ior"]), "k": np.float32(lp["transmission"]) * lp["tint"].astype(np.float32),
                    "background": (1.0, 1.0, 1.0, 1.0),
                    "model": glasses.get("lens_model", 0), "thickness": glasses.get("lens_thickness", 0.0), "ior": lp["ior"]}
        else:
            rgba2, d2, _ = mesh.render(cam12, 2 * width, 2 * height)
            surf, ts = O.mesh_resolve(rgba2, d2, width, height, 2)
    frame, depth, ns, stats = m.render_frame(P, surf, ts, lens=lens)
    if lens is not None:
        stats = dict(stats, lens=lens)
    img, _ = O.accumulate_tonemap(frame, None, 0, to_srgb=True)
    return img, frame, ns, stats, (surf, ts)
