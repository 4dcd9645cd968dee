"""Generates tests/golden/alice_views64.npy: the first 64 camera poses of the reference's bundled dataset
(/root/reference/volume/datasets/alice/transforms.json) as 3x4 view matrices of this renderer (BASELINE config 3,
SURVEY.md 8d): nerf_matrix_to_ngp (S/ngp/nerf_loader.cuh:115-134, scale 0.33, offset 0.5), then columns
(right * uLen, up * vLen, forward, eye - 0.5) with the renderer's fixed vLen = tanf(0.5f * 45) and a square aspect.
Run in the authoring container only (the reference tree is not on the GPU box); the .npy is committed."""
import json
import math
import os

import numpy as np

SRC = "/root/reference/volume/datasets/alice/transforms.json"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "alice_views64.npy")


def nerf_matrix_to_ngp(m, scale=0.33, offset=0.5):
    r = np.array(m, dtype=np.float64)[:3, :4].copy()
    r[:, 1] *= -1; r[:, 2] *= -1
    r[:, 3] = r[:, 3] * scale + offset
    return r[[1, 2, 0], :]                      # cycle axes xyz <- yzx


def main():
    frames = json.load(open(SRC))["frames"][:64]
    v_len = abs(math.tan(np.float32(0.5) * np.float32(45.0)))     # the reference passes degrees to tanf (radians)
    out = np.zeros((len(frames), 3, 4), dtype=np.float32)
    for i, f in enumerate(frames):
        m = nerf_matrix_to_ngp(f["transform_matrix"])
        right, down, fwd, pos = m[:, 0], m[:, 1], m[:, 2], m[:, 3]
        out[i, :, 0] = right / np.linalg.norm(right) * v_len      # square views: uLen = vLen
        out[i, :, 1] = -down / np.linalg.norm(down) * v_len
        out[i, :, 2] = fwd / np.linalg.norm(fwd)
        out[i, :, 3] = pos - 0.5
    np.save(OUT, out)
    print(OUT, out.shape, "eye distance to centre: min %.2f max %.2f" % (np.linalg.norm(out[:, :, 3], axis=1).min(), np.linalg.norm(out[:, :, 3], axis=1).max()))


if __name__ == "__main__":
    main()
