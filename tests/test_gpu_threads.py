"""Two host threads, each with its own context on the same GPU, rendering at the same time (ctypes drops the GIL inside libnmr),
plus two threads sharing ONE context (its calls are serialised by the context's mutex): every image equals the one a single
thread renders for that camera.  (include/nmr.h: 'one context = one device + one stream; calls on a context are serialised'.)"""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, HH = 256, 144


def _scene(path, gltf):
    import pynmr
    import synth
    r = pynmr.NerfMeshRenderer(W, HH)
    nerf = r.load_nerf(path)
    assert r.load_mesh(gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is not None
    return r, nerf


def test_contexts_in_parallel_threads(small_snapshot, glasses_gltf):
    path, _ = small_snapshot
    r0, n0 = _scene(path, glasses_gltf)
    cams, want = [], []
    for k in range(24):
        r0.orbit(0.11, 0.02 * ((k % 5) - 2), 0.5 if k % 7 == 0 else 0.0)
        cams.append(r0.view_projection_mat.copy())
        want.append(np.asarray(n0.render(W, HH, 1, linear=False)).copy())
    errors = []

    def worker(tid, r, nerf, order):
        try:
            for k in order:
                r.view_projection_mat = cams[k]
                if (k + tid) % 3 == 0:
                    assert r.frame()
                    img = np.asarray(r.read_frame())
                else:
                    img = np.asarray(nerf.render(W, HH, 1, linear=False))
                if not np.array_equal(img.view(np.uint32), want[k].view(np.uint32)):
                    errors.append((tid, k))
        except Exception as e:      # noqa: BLE001
            errors.append((tid, repr(e)))

    # own contexts
    ctxs = [_scene(path, glasses_gltf) for _ in range(3)]
    threads = [threading.Thread(target=worker, args=(t, ctxs[t][0], ctxs[t][1], list(range(t, 24)) + list(range(t)))) for t in range(3)]
    for t in threads: t.start()
    for t in threads: t.join()
    assert errors == []
    # one shared context: set-camera + render must not interleave between threads, so each thread takes the context for a whole
    # step under its own lock; libnmr's mutex keeps the individual calls whole
    r, nerf = ctxs[0]
    step = threading.Lock()

    def shared(tid):
        try:
            for k in range(tid, 24, 2):
                with step:
                    r.view_projection_mat = cams[k]
                    img = np.asarray(nerf.render(W, HH, 1, linear=False)).copy()
                if not np.array_equal(img.view(np.uint32), want[k].view(np.uint32)):
                    errors.append(("shared", tid, k))
        except Exception as e:      # noqa: BLE001
            errors.append(("shared", tid, repr(e)))
    threads = [threading.Thread(target=shared, args=(t,)) for t in range(2)]
    for t in threads: t.start()
    for t in threads: t.join()
    assert errors == []
