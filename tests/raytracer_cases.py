"""Seeded scenes / cameras / ray batches shared by the traversal pin: the generator of
tests/golden/raytracer_golden.npz, the CPU tests of the oracle against the compiled shader and the
-m gpu tests of the CUDA path against both."""
import numpy as np

import scenes
from realtimeraytracing_b200 import synth
from realtimeraytracing_b200.layouts import RAY

# name -> (width, height): deliberately not multiples of 16 where the Q5 rule matters
FRAMES = {"soup3000": (96, 64), "grid": (100, 70), "two_mesh": (96, 64), "survey20k": (64, 48)}
CASE_NAMES = list(FRAMES)


def case(name):
    """(triangles, meshes, materials, camera, width, height)"""
    w, h = FRAMES[name]
    if name == "soup3000":
        tris, meshes, L = scenes.soup(3000)
        cam = synth.soup_camera(L, w, h)
    elif name == "grid":
        # connected mesh with vertices on the planes x = 0 and y = 0 under an axis-aligned camera at x = y = 0:
        # exact-t ties on shared edges and the NaN slab case of Q11 on the centre row
        tris, meshes = synth.grid_mesh(24, 24)
        cam = synth.reference_camera(aspect=w / h)
    elif name == "two_mesh":
        tris, meshes = scenes.two_mesh_scene()
        cam = synth.reference_camera(eye=(0.0, 0.0, -9.0), aspect=w / h)
    elif name == "survey20k":
        tris, meshes = synth.survey_known_answer_scene()
        cam = synth.reference_camera(eye=(0.0, 0.0, -4.0), aspect=w / h)
    else:
        raise KeyError(name)
    nb_mat = int(meshes["material_id"].max()) + 1
    rng = np.random.RandomState(17)
    materials = rng.uniform(0.0, 1.0, size=(nb_mat, 4)).astype(np.float32)
    return tris, meshes, materials, cam, w, h


def random_rays(n, extent, seed):
    """Rays from a shell around the scene towards points inside it (finite, no zero components)."""
    rng = np.random.RandomState(seed)
    o = rng.normal(size=(n, 3))
    o = o / np.linalg.norm(o, axis=1, keepdims=True) * extent * rng.uniform(0.2, 1.6, size=(n, 1))
    tgt = rng.uniform(-0.5, 0.5, size=(n, 3)) * extent
    d = tgt - o
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros(n, dtype=RAY)
    rays["o"][:, :3] = o.astype(np.float32)
    rays["o"][:, 3] = 1.0
    rays["d"][:, :3] = d.astype(np.float32)
    return rays


def _unit(v):
    return v / np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-30)


def rays_at_triangles(tris, meshes, tri_index, seed):
    """One ray per chosen triangle, aimed at a point of its plane with barycentrics in [-0.25, 1.25] (inside,
    on-edge and outside), from the front, from the back and -- a quarter of them -- at grazing incidence, where
    |a| falls through the shader's 1e-4 rejection (raytracer.glsl:117-119)."""
    rng = np.random.RandomState(seed)
    n = tri_index.size
    t = tris[tri_index]
    M = meshes["m"][t["model_id"]].reshape(n, 4, 4).transpose(0, 2, 1).astype(np.float64)  # [row][col]
    P = [np.einsum("nij,nj->ni", M, t[k].astype(np.float64))[:, :3] for k in ("p0", "p1", "p2")]
    b = rng.uniform(-0.25, 1.25, size=(n, 2))
    tgt = P[0] + b[:, :1] * (P[1] - P[0]) + b[:, 1:] * (P[2] - P[0])
    nrm = _unit(np.cross(P[1] - P[0], P[2] - P[0]))
    inplane = _unit(P[1] - P[0])
    side = np.where(rng.uniform(size=(n, 1)) < 0.5, 1.0, -1.0)
    graze = rng.uniform(size=(n, 1)) < 0.25
    tilt = np.where(graze, 10.0 ** rng.uniform(-5.0, -1.5, size=(n, 1)), rng.uniform(0.2, 1.0, size=(n, 1)))
    d = _unit(-(side * nrm * tilt + inplane * (1.0 - tilt) + 0.05 * rng.normal(size=(n, 3)) * (~graze)))
    dist = rng.uniform(0.5, 8.0, size=(n, 1))
    o = tgt - d * dist
    rays = np.zeros(n, dtype=RAY)
    rays["o"][:, :3] = o.astype(np.float32)
    rays["o"][:, 3] = 1.0
    rays["d"][:, :3] = d.astype(np.float32)
    return rays


def rays_at_boxes(flat, node_index, seed):
    """One ray per chosen node aimed at a point within 1.3x of its box (hits, edge entries and near misses)."""
    rng = np.random.RandomState(seed)
    n = node_index.size
    lo = flat["bmin"][node_index].astype(np.float64)
    hi = flat["bmax"][node_index].astype(np.float64)
    c, e = 0.5 * (lo + hi), 0.5 * (hi - lo) + 1e-3
    tgt = c + e * rng.uniform(-1.3, 1.3, size=(n, 3))
    d = _unit(rng.normal(size=(n, 3)))
    o = tgt - d * rng.uniform(0.1, 6.0, size=(n, 1)) * np.where(rng.uniform(size=(n, 1)) < 0.15, -1.0, 1.0)
    rays = np.zeros(n, dtype=RAY)
    rays["o"][:, :3] = o.astype(np.float32)
    rays["o"][:, 3] = 1.0
    rays["d"][:, :3] = d.astype(np.float32)
    return rays


def hit_points(tris, meshes, hits):
    """World-space point named by the barycentrics of hit records (all must be hits).  Q8: the shader's struct
    swaps the names of the second and third vertex, so b0 weighs host P2, b1 host P1 and b2 host P0."""
    t = tris[hits["tri"]]
    n = t.size
    M = meshes["m"][t["model_id"]].reshape(n, 4, 4).transpose(0, 2, 1).astype(np.float64)
    P = [np.einsum("nij,nj->ni", M, t[k].astype(np.float64))[:, :3] for k in ("p0", "p1", "p2")]
    return (hits["b0"].astype(np.float64)[:, None] * P[2] + hits["b1"].astype(np.float64)[:, None] * P[1]
            + hits["b2"].astype(np.float64)[:, None] * P[0])


def incidence_cos(tris, meshes, hits, rays):
    """|cos| of the angle between the ray and the normal of the triangle each hit names (float64).  The distance
    of a hit is ill-conditioned at grazing incidence: a relative change eps of the direction moves t by about
    eps * t / |cos|, and the shader accepts |cos| down to ~1e-4 / (2 * area) (its |a| >= 1e-4 test)."""
    t = tris[hits["tri"]]
    n = t.size
    M = meshes["m"][t["model_id"]].reshape(n, 4, 4).transpose(0, 2, 1).astype(np.float64)
    P = [np.einsum("nij,nj->ni", M, t[k].astype(np.float64))[:, :3] for k in ("p0", "p1", "p2")]
    nrm = _unit(np.cross(P[1] - P[0], P[2] - P[0]))
    d = _unit(rays["d"][:, :3].astype(np.float64))
    return np.abs((nrm * d).sum(axis=1))


def has_zero_component(rays):
    """Q11: a direction component of exactly 0 makes the slab test divide by zero; where the origin also lies on
    a slab plane the result is NaN and GLSL min/max leave the outcome to the implementation."""
    return (rays["d"][:, :3] == 0).any(axis=1)


def armadillo():
    """(triangles, meshes): resources/models/armadillo.obj as the reference's Mesh::load returns it (99 976 triangles),
    from tests/golden/armadillo_mesh.npz (make_armadillo_fixture.py); identity model matrix, ModelId 0."""
    import os
    from realtimeraytracing_b200.layouts import TRIANGLE
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "armadillo_mesh.npz"))
    v = z["vertex_words"].view(np.float32)
    f = z["faces"]
    tris = np.zeros(f.shape[0], dtype=TRIANGLE)
    for k, name in enumerate(("p0", "p1", "p2")):
        tris[name][:, :3] = v[f[:, k]]
        tris[name][:, 3] = 1.0
    return tris, synth.identity_meshes(1)
