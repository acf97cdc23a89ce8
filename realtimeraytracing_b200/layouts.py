"""Binary layouts of the reference's "GPU structs", as numpy dtypes.

Sizes/offsets follow SURVEY.md App. A rule 10 (probed from the reference headers):

* ``TRIANGLE``  <- cr::TriangleGPU   srcCommon/scene/geometry/triangle.hpp:9-14   64 B
* ``MESH``      <- cr::MeshModelGPU  srcCommon/scene/geometry/mesh.hpp:12-15      68 B
* ``NODE``      <- cr::BVH_NodeGPU   srcCommon/scene/geometry/bvh.hpp:22-42       48 B
* ``CAMERA``    <- cr::CameraGPU     srcCommon/scene/camera.hpp:21-30            284 B
* ``RAY``       <- Ray               srcCommon/shaders/raytracer.glsl:14-17       32 B
* ``HIT``       <- Hit               srcCommon/shaders/raytracer.glsl:36-40       24 B
"""
import numpy as np

TRIANGLE = np.dtype([("p0", "<f4", 4), ("p1", "<f4", 4), ("p2", "<f4", 4),
                     ("model_id", "<u4"), ("pad", "<u4", 3)])
MESH = np.dtype([("m", "<f4", 16), ("material_id", "<u4")])
NODE = np.dtype([("bmin", "<f4", 3), ("pad0", "<u4"), ("bmax", "<f4", 3), ("pad1", "<u4"),
                 ("tri", "<u4"), ("left", "<u4"), ("right", "<u4"), ("pad2", "<u4")])
CAMERA = np.dtype([("view", "<f4", 16), ("proj", "<f4", 16), ("inv_view", "<f4", 16),
                   ("inv_proj", "<f4", 16), ("eye", "<f4", 4),
                   ("plane_width", "<f4"), ("plane_height", "<f4"), ("plane_near", "<f4")])
RAY = np.dtype([("o", "<f4", 4), ("d", "<f4", 4)])
HIT = np.dtype([("b0", "<f4"), ("b1", "<f4"), ("b2", "<f4"), ("t", "<f4"),
                ("did_hit", "<u4"), ("tri", "<u4")])

assert TRIANGLE.itemsize == 64 and TRIANGLE.fields["model_id"][1] == 48
assert MESH.itemsize == 68
assert NODE.itemsize == 48 and NODE.fields["bmax"][1] == 16 and NODE.fields["tri"][1] == 32
assert CAMERA.itemsize == 284
assert RAY.itemsize == 32 and HIT.itemsize == 24

NONE = 0xFFFFFFFF


def node_words(nodes: np.ndarray) -> np.ndarray:
    """The 9 meaningful 32-bit words of each node (padding excluded), shape [n, 9]."""
    raw = np.ascontiguousarray(nodes).view(np.uint32).reshape(-1, 12)
    return raw[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]]


def hash_words(words: np.ndarray) -> int:
    """FNV-1a-64 variant over 32-bit words."""
    h = 1469598103934665603
    mask = (1 << 64) - 1
    for w in np.ascontiguousarray(words, dtype=np.uint32).ravel().tolist():
        h = ((h ^ w) * 1099511628211) & mask
    return h
