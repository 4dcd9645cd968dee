"""Random camera poses, GPU frame against the CPU oracle (tools/fuzz_poses.py): orbit steps of any size, dollies from far
outside to inside the head, sideways shifts - live rays from none to every pixel, the batch-8 regime and the replayed
n_steps schedule of close-ups, lens pixels from none to the whole frame.  Round 2's runs of this (900 poses) found the two
cases that the hand-picked poses of the other tests do not reach: lens rays inside the schedule replay of a close-up, and a
transmitted lens segment that starts behind an opaque surface with a short batch (march_kernel: `pre_blend`)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.parametrize("with_lens,seed,n", [(True, 4, 110), (True, 6, 60), (False, 11, 60)])
def test_random_poses_match_the_oracle(with_lens, seed, n):
    import fuzz_poses
    violations, worst, over = fuzz_poses.run(n, seed, with_lens, verbose=False)
    assert violations == 0
    assert over == 0 and worst <= 2.0 / 255.0          # not one pixel over the tolerance (seed 4 holds both cases named above)


def test_random_poses_match_the_reference_renderer():
    """The same kind of poses, plus random crop boxes (Testbed.render_aabb) and model transforms, against the reference's own
    kernels on this GPU (tools/fuzz_reference.py); per pose the tolerances of tests/test_gpu_vs_reference.py."""
    from oracle import refgpu
    if not refgpu.available():
        pytest.skip("oracle/_ref/libnmr_refgpu.so not built")
    import fuzz_reference
    violations, worst_psnr, worst_frac = fuzz_reference.run(80, 1, verbose=False)
    assert violations == 0 and worst_psnr >= 45.0 and worst_frac <= 0.004


def test_random_poses_every_path_gives_the_same_bits():
    """frame() with / without overlap, render() in three formats, render_update(), render_views(), three row shards: one picture
    (tools/fuzz_paths.py)."""
    import fuzz_paths
    assert fuzz_paths.run(50, 3, verbose=False) == []


def test_random_mesh_transforms_match_the_oracle():
    """Random rotations / scales / offsets of the textured glTF (all texture slots) and random cameras: visibility bit for bit,
    shading and the hybrid frame within tolerance (tools/fuzz_mesh.py)."""
    import fuzz_mesh
    assert fuzz_mesh.run(25, 2, verbose=False) == []


def test_random_poses_plate_lenses_match_the_oracle():
    import fuzz_poses
    violations, worst, over = fuzz_poses.run(60, 8, True, verbose=False, plate=True)
    assert violations == 0 and over == 0 and worst <= 2.0 / 255.0


def test_random_poses_accumulation_and_two_nerfs_match_the_oracle():
    """frame() repeated 1 - 4 times on a still camera (the reference's per-sample jitter) and two NeRFs merged by depth, at random
    poses (tools/fuzz_scene.py)."""
    import fuzz_scene
    assert fuzz_scene.run(20, 5, verbose=False) == []


def test_random_call_sequences_leave_no_state_behind():
    """A renderer driven through random API calls ends up rendering its final state exactly like a fresh renderer put into that
    state (tools/fuzz_state.py).  Round 2's runs found a lens ray's shifted origin (plate model) leaking into the next ray of its
    ray group - visible only in frames large enough for a group to march more than one ray."""
    import fuzz_state
    assert fuzz_state.run(25, 1, verbose=False) == []


def test_plate_lenses_in_frames_where_ray_groups_march_several_rays():
    import fuzz_poses
    violations, worst, over = fuzz_poses.run(8, 31, True, verbose=False, plate=True, size=(512, 384))
    assert violations == 0 and over == 0 and worst <= 2.0 / 255.0
