"""The host-side file readers (msgpack snapshot, glTF with its tangent generator, PNG: csrc/host.cpp, csrc/value.cpp, csrc/mikk.cpp) under AddressSanitizer + UBSan, fed
with valid files and with hundreds of mutated ones (byte flips, truncations, numbers replaced by extremes inside the JSON).
A reader may reject a file; it must not touch memory it does not own, overflow, or hang.  CPU only (g++ -fsanitize)."""
import json
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "nerf-glasses_b200", "csrc")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    out = str(tmp_path_factory.mktemp("asan") / "host_fuzz")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
           "-I", CSRC, os.path.join(ROOT, "tests", "native", "host_fuzz.cpp"), os.path.join(CSRC, "host.cpp"), os.path.join(CSRC, "value.cpp"), os.path.join(CSRC, "mikk.cpp"), "-lz", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if "asan" in r.stderr.lower() or "sanitize" in r.stderr.lower():
            pytest.skip("sanitizer runtime not installed: " + r.stderr[-200:])
        raise AssertionError(r.stderr[-2000:])
    return out


def run(harness, kind, files):
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0:allocator_may_return_null=1:max_allocation_size_mb=2048", UBSAN_OPTIONS="print_stacktrace=1")
    for i in range(0, len(files), 200):
        r = subprocess.run([harness, kind] + files[i:i + 200], capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, (r.stdout[-300:], r.stderr[-3000:])
    return r.stdout


def mutate_bytes(data: bytes, rng, n_variants: int):
    out = []
    for _ in range(n_variants):
        b = bytearray(data)
        kind = int(rng.integers(0, 4))
        if kind == 0:                                   # a few random byte flips
            for _ in range(int(rng.integers(1, 8))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        elif kind == 1:                                 # truncation
            b = b[: int(rng.integers(1, len(b)))]
        elif kind == 2:                                 # a run overwritten with 0xFF / 0x00
            p = int(rng.integers(0, len(b))); n = int(rng.integers(1, 64))
            b[p:p + n] = bytes([int(rng.choice([0, 255]))]) * len(b[p:p + n])
        else:                                           # a chunk duplicated in place (lengths no longer match)
            p = int(rng.integers(0, len(b))); n = int(rng.integers(1, 128))
            b[p:p] = b[p:p + n]
        out.append(bytes(b))
    return out


def mutate_json_numbers(text: str, rng, n_variants: int):
    nums = [m for m in re.finditer(r"(?<![\w.\"])-?\d+(?:\.\d+)?(?![\w.\"])", text)]
    out = []
    for _ in range(n_variants):
        t = text
        for m in sorted(rng.choice(len(nums), size=min(len(nums), int(rng.integers(1, 4))), replace=False), reverse=True):
            m = nums[int(m)]
            t = t[:m.start()] + str(rng.choice(["-1", "0", "4294967295", "4294967296", "-2147483649", "9223372036854775807", "1e30", "-7", "65536", "3.5"])) + t[m.end():]
        out.append(t)
    return out


def test_gltf_and_png_readers_survive_mutated_files(harness, tmp_path, glasses_gltf):
    import synth
    rng = np.random.default_rng(7)
    files = []
    # valid inputs first: the fixtures (external .bin + .png) and the textured / lens variants
    tex_gltf = synth.write_textured_glasses_gltf(str(tmp_path / "tex"))
    lens_gltf = synth.write_lens_glasses_gltf(str(tmp_path / "lens"))
    good = [glasses_gltf, tex_gltf, lens_gltf]
    out = run(harness, "gltf", good)
    assert "accepted 3 rejected 0" in out
    # mutated JSON (numbers replaced by extremes; raw byte damage) next to the original side files
    for src in good:
        base = os.path.dirname(src)
        text = open(src).read()
        for k, t in enumerate(mutate_json_numbers(text, rng, 60)):
            p = os.path.join(base, f"num_{k}.gltf"); open(p, "w").write(t); files.append(p)
        for k, b in enumerate(mutate_bytes(text.encode(), rng, 40)):
            p = os.path.join(base, f"bytes_{k}.gltf"); open(p, "wb").write(b); files.append(p)
    # damaged side files: the .bin and every .png of the textured scene, one at a time, under the untouched .gltf
    base = os.path.dirname(tex_gltf)
    doc = json.load(open(tex_gltf))
    side = [b["uri"] for b in doc["buffers"]] + [i["uri"] for i in doc.get("images", [])]
    for uri in side:
        if uri.startswith("data:"):
            continue
        data = open(os.path.join(base, uri), "rb").read()
        for k, b in enumerate(mutate_bytes(data, rng, 25)):
            d = tmp_path / f"side_{uri.replace('.', '_')}_{k}"
            shutil.copytree(base, d, ignore=shutil.ignore_patterns("num_*", "bytes_*"))
            open(d / uri, "wb").write(b)
            files.append(str(d / os.path.basename(tex_gltf)))
    out = run(harness, "gltf", files)
    assert "rejected" in out


def test_snapshot_reader_survives_mutated_files(harness, tmp_path):
    import synth
    rng = np.random.default_rng(8)
    good = str(tmp_path / "s.msgpack")
    synth.write_snapshot(good, seed=3, log2_hashmap_size=12, n_floaters=4)
    assert "accepted 1 rejected 0" in run(harness, "snapshot", [good])
    data = open(good, "rb").read()
    files = []
    # the msgpack header (keys, lengths, type tags) lives in the first kilobytes and between the two big binaries: damage both
    # the whole file and, more densely, its structural parts
    for k, b in enumerate(mutate_bytes(data, rng, 60)):
        p = str(tmp_path / f"all_{k}.msgpack"); open(p, "wb").write(b); files.append(p)
    head = 4096
    for k in range(120):
        b = bytearray(data)
        for _ in range(int(rng.integers(1, 6))):
            b[int(rng.integers(0, min(head, len(b))))] = int(rng.integers(0, 256))
        p = str(tmp_path / f"head_{k}.msgpack"); open(p, "wb").write(bytes(b)); files.append(p)
    out = run(harness, "snapshot", files)
    assert "rejected" in out


def test_readers_reject_pathological_nesting_and_lengths(harness, tmp_path):
    """millions of nested containers, declared lengths of 4 G entries, empty files: rejected, no stack overflow, no huge allocation"""
    snaps = {"nest_arr": b"\x91" * 2000000 + b"\x00", "nest_map": b"\x81\xa1k" * 500000 + b"\x00", "huge_arr": b"\xdd\xff\xff\xff\xff" + b"\x00" * 10,
             "huge_map": b"\xdf\xff\xff\xff\xff" + b"\xa1k\x00" * 3, "huge_bin": b"\x81\xa1k\xc6\xff\xff\xff\xff" + b"\x00" * 10,
             "huge_str": b"\xdb\xff\xff\xff\xff" + b"a" * 10, "empty": b""}
    files = []
    for name, data in snaps.items():
        p = tmp_path / f"{name}.msgpack"; p.write_bytes(data); files.append(str(p))
    assert f"accepted 0 rejected {len(files)}" in run(harness, "snapshot", files)
    docs = {"deep_arr": "[" * 1000000, "deep_obj": '{"a":' * 500000 + "1" + "}" * 500000, "empty": "", "just_null": "null"}
    files = []
    for name, text in docs.items():
        p = tmp_path / f"{name}.gltf"; p.write_text(text); files.append(str(p))
    assert f"accepted 0 rejected {len(files)}" in run(harness, "gltf", files)
