"""Shared seeded inputs of the parity tests (same scenes for the oracle, the reference build and the GPU)."""
import numpy as np

from realtimeraytracing_b200 import synth
from realtimeraytracing_b200.layouts import MESH, TRIANGLE


def soup(n, seed=None):
    tris, meshes, L = synth.triangle_soup(n, seed=seed)
    return tris, meshes, L


def two_mesh_scene(n_each=700, seed=11):
    """Two meshes with non-trivial model matrices (rotation+scale+translation) -- exercises M*P order."""
    rng = np.random.RandomState(seed)
    tris = np.zeros(2 * n_each, dtype=TRIANGLE)
    c = rng.uniform(-2, 2, size=(2 * n_each, 1, 3)).astype(np.float32)
    v = c + rng.uniform(-0.3, 0.3, size=(2 * n_each, 3, 3)).astype(np.float32)
    for k, name in enumerate(("p0", "p1", "p2")):
        tris[name][:, :3] = v[:, k]
        tris[name][:, 3] = 1.0
    tris["model_id"][n_each:] = 1
    meshes = np.zeros(2, dtype=MESH)

    def model(angle, scale, t):
        ca, sa = np.cos(angle), np.sin(angle)
        m = np.array([[ca * scale, -sa * scale, 0, t[0]], [sa * scale, ca * scale, 0, t[1]],
                      [0, 0, scale, t[2]], [0, 0, 0, 1]], dtype=np.float32)
        return m.T.reshape(16)  # column-major storage

    meshes["m"][0] = model(0.3, 1.7, (0.5, -1.0, 2.0))
    meshes["m"][1] = model(-1.1, 0.6, (-3.0, 0.25, -1.5))
    meshes["material_id"] = [1, 2]
    return tris, meshes


def padded_scene(n=3000, pad_to=4096, seed=5):
    """Q2: the scene box is taken over the whole zero-padded vector."""
    tris, meshes, _ = synth.triangle_soup(n, seed=seed, extent=6.0)
    tris["p0"][:, 0] += 10.0  # keep the origin outside the real geometry so the padding matters
    tris["p1"][:, 0] += 10.0
    tris["p2"][:, 0] += 10.0
    out = np.zeros(pad_to, dtype=TRIANGLE)
    out[:n] = tris
    return out, meshes, n


def duplicate_codes_scene(n=2048, seed=3):
    """Many triangles share a centroid cell => equal Morton codes => tie order of the stable sort matters."""
    rng = np.random.RandomState(seed)
    tris = np.zeros(n, dtype=TRIANGLE)
    cells = rng.randint(0, 8, size=(n, 3)).astype(np.float32) * 4.0
    for k, name in enumerate(("p0", "p1", "p2")):
        off = rng.uniform(-0.001, 0.001, size=(n, 3)).astype(np.float32) if k else 0.0
        tris[name][:, :3] = cells + off
        tris[name][:, 3] = 1.0
    tris["p1"][:, 0] += 0.002
    tris["p2"][:, 1] += 0.002
    return tris, synth.identity_meshes(1)
