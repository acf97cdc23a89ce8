"""OBJ ingestion feeding the device path: meshes made by scene.Mesh (rtr_obj_load, rtr_mesh_primitive, rtr_mesh_set_*)
go through Scene -> BVH build -> primary rays, and every stage equals the oracle on the same records.  The records
themselves are pinned against the reference's mesh.cpp in tests/test_mesh_cpu.py."""
import os

import numpy as np
import pytest

from realtimeraytracing_b200 import capi, scene as rscene, synth
from realtimeraytracing_b200.layouts import node_words

pytestmark = pytest.mark.gpu
OBJ_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "obj")


def _scene(ctx):
    rscene.Mesh._IdGenerator = 0  # ids = mesh slots, like a fresh reference process (mesh.cpp:9)
    a = rscene.Mesh.load(os.path.join(OBJ_DIR, "random.obj"))
    a.setScale(2e-5)
    a.setRotation(0.3, -0.2, 0.1)
    b = rscene.Mesh.primitiveCube()
    b.setRotation(0.5, 0.7, -0.4)
    b.setScale(0.75)
    b.setPosition((1.5, -0.5, 2.0))
    b.setMaterial(1)
    c = rscene.Mesh.load(os.path.join(OBJ_DIR, "polygons.obj"))
    c.setPosition((-3.0, -1.0, 1.0))
    s = rscene.Scene(ctx)
    for m in (a, b, c):
        s.addMesh(m)
    return s


def test_obj_scene_builds_and_traces_like_the_oracle(ctx, oracle):
    s = _scene(ctx)
    tris, meshes = s.getTriangleToGPUData(), s.getMeshModelToGPUData()
    assert tris.size == 473 + 12 + 38 and meshes.size == 3
    assert sorted(set(tris["model_id"].tolist())) == [0, 1, 2]
    flat = s.sendDataToGpu()
    ob = oracle.bvh_build(tris, meshes)
    assert np.array_equal(node_words(flat), node_words(oracle.flatten(ob.clusters, ob.left, ob.right)))
    W, H = 160, 96
    cam = synth.soup_camera(6.0, W, H)
    got = s._BVH.handle.trace_primary(cam, W, H, W, H)
    exp = oracle.trace_primary(flat, tris, meshes, cam, W, H, W, H)
    assert np.array_equal(got.view(np.uint8), exp.view(np.uint8))
    assert got["did_hit"].sum() > 100
    hit_models = set(tris["model_id"][got["tri"][got["did_hit"] != 0]].tolist())
    assert hit_models == {0, 1, 2}
