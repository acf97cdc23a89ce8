"""Shared by tests/test_mesh_cpu.py and tests/golden/make_mesh_golden.py: the transform / centroid cases and the
ctypes view of oracle/_ref/libref_mesh.so (the reference's own mesh.cpp; TEST INFRASTRUCTURE)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MESH = os.path.join(ROOT, "oracle", "_ref", "libref_mesh.so")

# op = (kind, a, b, c): 0 setPosition(a,b,c), 1 setScale(a), 2 setRotation(a,b,c), 3 setMaterial(int(a))
TRANSFORM_CASES = [
    [],
    [(1, 2.0, 0, 0)],
    [(0, 1.5, -2.25, 3.125)],
    [(2, 0.3, 0.0, 0.0)],
    [(2, 0.0, -1.1, 0.0)],
    [(2, 0.0, 0.0, 2.7)],
    [(2, 0.1, 0.2, 0.3), (1, 0.37, 0, 0), (0, 4.0, 5.0, 6.0), (3, 7, 0, 0)],
    [(0, 1.0, 2.0, 3.0), (2, 3.14159274, 1.57079637, -0.785398185), (1, 1.0 / 3.0, 0, 0), (2, -0.5, 0.25, 1e-3)],
    [(1, 1e-3, 0, 0), (1, 1e3, 0, 0), (2, 100.0, -200.0, 300.0), (0, -1e6, 1e-6, 0.0), (1, 0.7, 0, 0), (3, 63, 0, 0)],
]


def _random_cases():
    rng = np.random.default_rng(99)
    cases = []
    for _ in range(24):
        ops = []
        for _ in range(int(rng.integers(1, 8))):
            k = int(rng.integers(0, 4))
            a, b, c = (float(np.float32(x)) for x in rng.uniform(-7, 7, 3))
            ops.append((k, abs(a) + 0.05 if k == 1 else (float(int(abs(a) * 9)) if k == 3 else a), b, c))
        cases.append(ops)
    return cases


TRANSFORM_CASES = TRANSFORM_CASES + _random_cases()


def centroid_inputs():
    rng = np.random.default_rng(7)
    n = 64
    tris = np.zeros((n, 16), dtype=np.float32)
    tris[:, :12] = rng.uniform(-50, 50, (n, 12)).astype(np.float32)
    tris[:, 3] = tris[:, 7] = tris[:, 11] = 1.0
    models = rng.uniform(-2, 2, (n, 16)).astype(np.float32)
    models[: n // 2] = np.eye(4, dtype=np.float32).reshape(16)
    models[: n // 2, 12:15] = rng.uniform(-5, 5, (n // 2, 3)).astype(np.float32)
    return tris, models


def ref_mesh_lib():
    if not os.path.exists(REF_MESH):
        return None
    L = C.CDLL(REF_MESH)
    L.ref_mesh_load.restype = C.c_int64
    L.ref_mesh_load.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.POINTER(C.c_uint32)]
    L.ref_mesh_primitive.restype = C.c_int64
    L.ref_mesh_primitive.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_uint32)]
    L.ref_mesh_transform.restype = None
    L.ref_mesh_transform.argtypes = [C.c_int] + [C.c_void_p] * 6
    L.ref_triangle_centroid.restype = None
    L.ref_triangle_centroid.argtypes = [C.c_void_p] * 3
    return L


def _strip_id(words: np.ndarray) -> np.ndarray:
    """TriangleGPU words with _ModelId zeroed (Mesh::_Id counts every Mesh the process ever made)."""
    w = words.copy()
    w[:, 12] = 0
    return w


def ref_load(L, path, cap=1 << 19):
    buf = np.zeros((cap, 16), dtype=np.uint32)
    mid = C.c_uint32()
    n = L.ref_mesh_load(os.fsencode(path), buf.ctypes.data, cap, C.byref(mid))
    assert n <= cap
    assert n == 0 or (buf[:n, 12] == mid.value).all()
    return _strip_id(buf[:n]), mid.value


def ref_primitive(L, which):
    buf = np.zeros((16, 16), dtype=np.uint32)
    mid = C.c_uint32()
    n = L.ref_mesh_primitive(which, buf.ctypes.data, 16, C.byref(mid))
    return _strip_id(buf[:n]), mid.value


def ref_transform(L, ops):
    kind = np.array([o[0] for o in ops], dtype=np.int32)
    a, b, c = (np.array([o[i] for o in ops], dtype=np.float32) for i in (1, 2, 3))
    out = np.zeros(17, dtype=np.uint32)
    L.ref_mesh_transform(len(ops), kind.ctypes.data, a.ctypes.data, b.ctypes.data, c.ctypes.data, None, out.ctypes.data)
    return out


def ref_centroid(L, tri16, model16):
    out = np.zeros(3, dtype=np.float32)
    t = np.ascontiguousarray(tri16)
    m = np.ascontiguousarray(model16)
    L.ref_triangle_centroid(t.ctypes.data, m.ctypes.data, out.ctypes.data)
    return out
