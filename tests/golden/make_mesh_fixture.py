"""Regenerates tests/golden/glasses_mesh.npz from the reference's bundled glasses asset.

Run in the authoring container only (/root/reference is absent on the GPU box):
    python tests/golden/make_mesh_fixture.py
The fixture carries the accessor payloads of
/root/reference/nerf_mesh_renderer/assets/meshes/glasses/glasses.{gltf,bin}
(1864 vertices, 8856 u16 indices) plus the node TRS and material scalars, as data;
tools/synth.write_glasses_gltf() turns it back into a .gltf/.bin pair at test time.
"""
import json
import os
import sys

import numpy as np

REF = "/root/reference/nerf_mesh_renderer/assets/meshes/glasses"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    doc = json.load(open(os.path.join(REF, "glasses.gltf")))
    blob = open(os.path.join(REF, doc["buffers"][0]["uri"]), "rb").read()

    def acc(i, dt, ncomp):
        a = doc["accessors"][i]
        bv = doc["bufferViews"][a["bufferView"]]
        off = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
        arr = np.frombuffer(blob, dtype=dt, count=a["count"] * ncomp, offset=off)
        return arr.reshape(a["count"], ncomp) if ncomp > 1 else arr.copy()
    prim = doc["meshes"][0]["primitives"][0]
    node = doc["nodes"][0]
    pbr = doc["materials"][0]["pbrMetallicRoughness"]
    np.savez_compressed(
        os.path.join(HERE, "glasses_mesh.npz"),
        positions=acc(prim["attributes"]["POSITION"], "<f4", 3),
        normals=acc(prim["attributes"]["NORMAL"], "<f4", 3),
        texcoords=acc(prim["attributes"]["TEXCOORD_0"], "<f4", 2),
        indices=acc(prim["indices"], "<u2", 1),
        node_rotation_xyzw=np.array(node["rotation"], dtype=np.float64),
        node_translation=np.array(node["translation"], dtype=np.float64),
        metallic=np.float64(pbr["metallicFactor"]),
        roughness=np.float64(pbr["roughnessFactor"]),
    )
    print("wrote glasses_mesh.npz")


if __name__ == "__main__":
    sys.exit(main())
