import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "nerf-glasses_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def small_snapshot(tmp_path_factory):
    """Synthetic snapshot with a 2^15-entry hash table (fast on CPU); -> (path, parsed dict)."""
    import synth
    path = str(tmp_path_factory.mktemp("snap") / "small.msgpack")
    synth.write_snapshot(path, seed=1337, log2_hashmap_size=15)
    return path, synth.read_snapshot(path)


@pytest.fixture(scope="session")
def glasses_gltf(tmp_path_factory):
    import synth
    return synth.write_glasses_gltf(str(tmp_path_factory.mktemp("mesh")))
