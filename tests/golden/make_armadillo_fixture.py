"""Generates tests/golden/armadillo_mesh.npz: the triangles the REFERENCE's own Mesh::load (mesh.cpp against the
tinyobjloader 1.2.0 of its tree, oracle/_ref/libref_mesh.so) returns for resources/models/armadillo.obj -- the
~100 k-triangle real mesh of BASELINE config 2.  Run in the build container only:

    python tests/golden/make_armadillo_fixture.py

The GPU box has no /root/reference, so the 99 976 TriangleGPU records travel as an indexed mesh: unique vertex
positions (float32 words, exactly the loader's bits) + three indices per triangle; tests/raytracer_cases.armadillo()
expands it again.  The reference's default scene transform is identity (application.cpp:182-185)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from mesh_cases import ref_load, ref_mesh_lib  # noqa: E402


def main():
    L = ref_mesh_lib()
    if L is None:
        raise SystemExit("oracle/_ref/libref_mesh.so is missing: run `make -C oracle ref` where /root/reference exists")
    words, _ = ref_load(L, "/root/reference/resources/models/armadillo.obj")
    n = words.shape[0]
    pos = np.stack([words[:, 0:3], words[:, 4:7], words[:, 8:11]], axis=1).reshape(-1, 3)      # [3n, 3] uint32 words
    assert (words[:, [3, 7, 11]] == np.float32(1.0).view(np.uint32)).all()                      # w = 1, triangle.cpp:22-24
    verts, inverse = np.unique(pos, axis=0, return_inverse=True)
    faces = inverse.reshape(n, 3).astype(np.uint32)
    assert np.array_equal(verts[faces].reshape(-1, 3), pos)
    np.savez_compressed(os.path.join(HERE, "armadillo_mesh.npz"), vertex_words=verts, faces=faces)
    print(n, "triangles,", verts.shape[0], "vertices")


if __name__ == "__main__":
    main()
