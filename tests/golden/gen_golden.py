"""Generates tests/golden/bvh_golden.npz from the REFERENCE's own bvh.cpp (oracle/_ref, built by
oracle/Makefile from /root/reference).  Run in the build container only:

    OMP_NUM_THREADS=1 python tests/golden/gen_golden.py

The scenes are regenerated from seeds by tests/scenes.py, so the fixture holds reference OUTPUTS only:
unsorted Morton codes, _TriangleIndices, and the FNV hashes of the cluster arrays / flattened array
for every scene, plus the complete arrays for one small scene.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
os.environ["OMP_NUM_THREADS"] = "1"

import scenes  # noqa: E402
from oracle import Oracle, Reference  # noqa: E402
from realtimeraytracing_b200 import synth  # noqa: E402
from realtimeraytracing_b200.layouts import node_words  # noqa: E402


def scene_table():
    t = {}
    t["survey20k"] = synth.survey_known_answer_scene() + (None,)
    tris, meshes, _ = scenes.soup(512)
    t["soup512"] = (tris, meshes, None)
    tris, meshes, _ = scenes.soup(65536)
    t["soup65536"] = (tris, meshes, None)
    t["two_mesh"] = scenes.two_mesh_scene() + (None,)
    tris, meshes, n = scenes.padded_scene()
    t["padded"] = (tris, meshes, n)
    t["dupcodes"] = scenes.duplicate_codes_scene() + (None,)
    t["grid"] = synth.grid_mesh(40, 30) + (None,)
    return t


def main():
    o = Oracle()
    ref = Reference(65536)
    out = {}
    for name, (tris, meshes, n) in scene_table().items():
        rb = ref.bvh_build(tris, meshes, n=n)
        flat = o.flatten(rb.clusters, rb.left, rb.right)  # scene.cpp:189-208 cannot be compiled (GL); restated
        out[name + "/morton_hash"] = np.uint64(o.hash_words(rb.morton_unsorted))
        out[name + "/tri_idx_hash"] = np.uint64(o.hash_words(rb.triangle_indices))
        out[name + "/cluster_hash"] = np.uint64(o.hash_words(np.concatenate(
            [node_words(rb.clusters).ravel(), rb.left, rb.right, rb.parent])))
        out[name + "/flat_hash"] = np.uint64(o.hash_flat_nodes(flat))
        out[name + "/n"] = np.uint32(rb.n)
        if name == "soup512":
            out[name + "/morton"] = rb.morton_unsorted
            out[name + "/tri_idx"] = rb.triangle_indices
            out[name + "/flat_words"] = node_words(flat)
            out[name + "/left"] = rb.left
            out[name + "/right"] = rb.right
            out[name + "/parent"] = rb.parent
        print(name, rb.n, "%016x" % int(out[name + "/flat_hash"]))
    np.savez_compressed(os.path.join(HERE, "bvh_golden.npz"), **out)


if __name__ == "__main__":
    main()
