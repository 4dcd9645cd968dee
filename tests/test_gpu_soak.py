"""Short soak (tools/soak.py is the long one): hundreds of mixed calls with a different image size every few calls, then contexts
created and destroyed - device memory, the page-locked pool and host RSS come to rest, and a fixed camera renders the same bits at
the end as at the start.  Round 2's long run found the shim's page-locked pool growing by one block per new size (now capped,
least recently used sizes go back to the driver) and renderers kept alive until the garbage collector ran by a reference cycle
Testbed <-> Testbed.nerf (now created per access)."""
import gc
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_memory_comes_to_rest(small_snapshot, glasses_gltf):
    import psutil
    import torch
    import pynmr
    import synth
    proc = psutil.Process()

    def used():
        free, total = torch.cuda.mem_get_info(0)
        return (total - free) / 2 ** 20, proc.memory_info().rss / 2 ** 20

    path, _ = small_snapshot
    W, H = 1280, 720
    r = pynmr.NerfMeshRenderer(W, H, 0)
    nerf = r.load_nerf(path)
    assert r.load_mesh(glasses_gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is not None
    cam0 = r.view_projection_mat.copy()
    first = np.asarray(nerf.render(W, H, 1, linear=False)).copy()
    kept, a, marks = None, 0.0, []
    for k in range(1200):
        a += 0.03; r.orbit(-math.sin(a * 1.733) / 100.0, math.cos(a * 1.733) / 200.0, 0.0)
        m = k % 6
        if m == 0: r.frame()
        elif m == 1: nerf.render(W, H, 1, linear=False)
        elif m == 2: nerf.render(W, H, 1, linear=False, dtype=np.uint8)
        elif m == 3: kept = nerf.render_update(kept, W, H, linear=False)
        elif m == 4: r.render_views(nerf, np.stack([r.view_projection_mat] * 3), 256, 256)
        else: nerf.render(int(200 + (k * 37) % 900), int(100 + (k * 53) % 700), 1, linear=False)
        if k % 150 == 149:
            marks.append(used() + (sum(n * len(b) for n, b in pynmr._pinned_pool.items()) / 2 ** 20, len(pynmr._pinned_pool)))
    r.view_projection_mat = cam0
    assert np.array_equal(np.asarray(nerf.render(W, H, 1, linear=False)).view(np.uint32), first.view(np.uint32))
    assert marks[-1][0] - marks[1][0] < 64, marks               # device MiB
    assert marks[-1][1] - marks[-2][1] < 64, marks              # host RSS MiB: flat once the page-locked pool has reached its cap (512 MiB)
    assert marks[-1][1] - marks[0][1] < 700, marks
    assert sum(n * len(b) for n, b in pynmr._pinned_pool.items()) <= pynmr._PINNED_POOL_CAP
    base = used()
    for _ in range(12):
        r2 = pynmr.NerfMeshRenderer(640, 360, 0); n2 = r2.load_nerf(path)
        assert r2.load_mesh(glasses_gltf, t=synth.GLASSES_T, s=synth.GLASSES_S, r=synth.GLASSES_R_WXYZ) is not None
        r2.frame(); n2.render(640, 360, 1, linear=False)
        del n2, r2                                              # no gc.collect(): reference counting alone must release the context
    end = used()
    assert end[0] - base[0] < 64, (base, end)
    gc.collect()
