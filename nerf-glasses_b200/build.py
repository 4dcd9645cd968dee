"""Builds nerf-glasses_b200/libnmr.so (C ABI of include/nmr.h) with nvcc for sm_100a, in-tree.

    python nerf-glasses_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with gpurun.
--fmad=false / -ffp-contract=off: see DESIGN.md "Numerics" (bit-exact traversal needs unfused fp32 arithmetic).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libnmr.so")
SOURCES = ["api.cu", "host.cpp", "value.cpp", "mikk.cpp", "kernels.cu", "floaties.cu"]
HEADERS = ["host.h", "value.h", "kernels.cuh", "device_common.cuh", os.path.join("..", "..", "include", "nmr.h")]

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--fmad=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-Wall", "-shared", "-cudart", "static",
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, out: str = OUT, defines=()) -> str:
    """defines: extra -D macros for tuning builds (e.g. NMR_ENCODE_UNROLL=4) written to another `out`."""
    if not force and out == OUT and not needs_build():
        return OUT
    cmd = ["nvcc", *NVCC_FLAGS] + [f"-D{d}" for d in defines]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", out] + [os.path.join(CSRC, f) for f in SOURCES] + ["-lz"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv or bool(defs), verbose="--verbose" in sys.argv, out=outs[0] if outs else OUT, defines=defs))
