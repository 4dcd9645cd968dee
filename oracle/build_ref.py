"""Builds oracle/_ref/libnmr_ref.so from the reference's OWN headers (TEST INFRASTRUCTURE ONLY).

The sources stay where they lie under /root/reference; only oracle/ref_harness.cu (ours) is compiled,
with include paths pointing into the reference tree.  Output goes to oracle/_ref/ (git-ignored, shipped
to the GPU box by gpurun).  No-op when /root/reference is absent (GPU box) or nvcc is missing.
What it pins: floatie removal (S/floatyremover.h), the orbit camera (S/orbit_camera.h,
flythrough_camera.h), and the HOST_DEVICE helpers of the ray set-up path.  The reference's kernels
(__global__/__device__ only, OptiX programs) cannot be run on a CPU and stay unpinned.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
R = "/root/reference/nerf_mesh_renderer"
T = R + "/dependencies/tiny-cuda-nn"
OUT = os.path.join(HERE, "_ref", "libnmr_ref.so")
SRC = os.path.join(HERE, "ref_harness.cu")


def available() -> bool:
    return os.path.exists(OUT)


MIKK_OUT = os.path.join(HERE, "_ref", "libmikk_ref.so")
MIKK_SRC = os.path.join(HERE, "ref_mikk_harness.c")


def build_mikk(force: bool = False) -> str | None:
    """oracle/_ref/libmikk_ref.so: the reference's vendored mikktspace.c (compiled where it lies) + oracle/ref_mikk_harness.c (ours).
    Plain gcc -O2 for x86-64 (no FMA contraction possible without -mfma), like the reference's CMake default for this C file."""
    mk = R + "/dependencies/MikkTSpace"
    if not os.path.isdir(mk) or shutil.which("gcc") is None:
        return MIKK_OUT if os.path.exists(MIKK_OUT) else None
    if not force and os.path.exists(MIKK_OUT) and os.path.getmtime(MIKK_OUT) >= os.path.getmtime(MIKK_SRC):
        return MIKK_OUT
    os.makedirs(os.path.dirname(MIKK_OUT), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-w", "-I", mk, "-o", MIKK_OUT,
                           MIKK_SRC, mk + "/mikktspace.c", "-lm"])
    return MIKK_OUT


def build(force: bool = False) -> str | None:
    build_mikk(force)
    if not os.path.isdir(R) or shutil.which("nvcc") is None:
        return OUT if os.path.exists(OUT) else None
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    inc = [R + "/src", R + "/dependencies", R + "/dependencies/eigen", R + "/dependencies/glm",
           R + "/dependencies/spdlog/include", R + "/dependencies/json", T + "/include", T + "/dependencies",
           T + "/dependencies/fmt/include", R + "/dependencies/filesystem"]
    cmd = ["nvcc", "-std=c++17", "-O2", "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off", "--fmad=false", "-shared",
           "-arch=sm_100", "-DTCNN_MIN_GPU_ARCH=100", "-DNDEBUG", "--extended-lambda", "--expt-relaxed-constexpr", "-w"]
    for i in inc:
        cmd += ["-I", i]
    cmd += ["-o", OUT, SRC]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True), MIKK_OUT)
