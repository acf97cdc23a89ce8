"""Generates tests/golden/mesh_golden.npz from the REFERENCE's own cr::Mesh (oracle/_ref/libref_mesh.so = mesh.cpp +
triangle.cpp of /root/reference compiled against the tinyobjloader 1.2.0 of its tree; `make -C oracle ref`).

    python tests/golden/make_mesh_golden.py

Holds, for every tests/golden/obj/*.obj, the TriangleGPU records Mesh::load returns (as raw uint32 words), the three
primitives, MeshModelGPU after fixed sequences of setPosition/setScale/setRotation/setMaterial calls, and centroids.
Only needed when the fixtures change; the tests read the .npz and never the reference."""
import ctypes as C
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from mesh_cases import TRANSFORM_CASES, centroid_inputs, ref_mesh_lib, ref_load, ref_primitive, ref_transform, ref_centroid  # noqa: E402


def main():
    L = ref_mesh_lib()
    if L is None:
        raise SystemExit("oracle/_ref/libref_mesh.so is missing: run `make -C oracle ref` where /root/reference exists")
    out = {}
    for path in sorted(glob.glob(os.path.join(HERE, "obj", "*.obj"))):
        tris, _ = ref_load(L, path)
        out["obj_" + os.path.basename(path)[:-4]] = tris
    for which in range(4):
        out["primitive_%d" % which] = ref_primitive(L, which)[0]
    for i, ops in enumerate(TRANSFORM_CASES):
        out["transform_%d" % i] = ref_transform(L, ops)
    tris, models = centroid_inputs()
    out["centroids"] = np.stack([ref_centroid(L, tris[i], models[i]) for i in range(len(tris))]).view(np.uint32)
    np.savez_compressed(os.path.join(HERE, "mesh_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
