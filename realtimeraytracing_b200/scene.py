"""Host-side mirror of the reference's srcCommon scene/BVH interface for the accelerated path.

Same names and argument meaning as the reference so parity tests read like its own code:

* ``BVH(nbTriangles, unsortedTriangles, meshesInTheScene)``  <- cr::BVH::BVH, bvh.hpp:89-91.
  The constructor runs the whole build synchronously (bvh.cpp:11-24) -- on the GPU, through
  rtr_bvh_build -- and the result is read from the public ``_InternalStruct`` (bvh.hpp:86)
  whose members carry the reference's names (BVH_Params, bvh.hpp:44-63).
* ``Scene`` with ``addMesh`` / ``getBVH_NodesToGPUData`` / ``sendDataToGpu``  <- glr::Scene,
  srcOpenGL/scene/scene.hpp:50-54, scene.cpp:110-220: pads nothing (no MAX_NB_TRIANGLES cap)
  unless ``reference_padding=True`` reproduces the 65 536 / 64 entry vectors of scene.cpp:18,27 (Q2).

Error behaviour: the reference prints and exit()s on fatal errors (errorHandler.cpp:13-29);
here every failure raises ``RtrError`` carrying the C ABI's code and message.
"""
import numpy as np

from . import capi
from .layouts import MESH, NONE, TRIANGLE

MAX_NB_TRIANGLES = 2 << 15  # triangle.hpp:21
MAX_NB_MESHES = 2 << 5      # mesh.hpp:26
MAX_NB_MATERIALS = 2 << 15  # material.hpp:19


class BVH_Params:
    """BVH_Params (bvh.hpp:44-63).  Optionals become arrays with RTR_NONE (0xFFFFFFFF) for nullopt."""

    def __init__(self):
        self._NbTriangles = 0
        self._Clusters = None         # NODE[2n-1] by cluster id (internal links are 0, Q10)
        self._IsLeaf = None           # uint8[2n-1]; 1 <=> has_value() in the reference (Q9)
        self._Parent = None           # uint32[2n-1]
        self._LeftChild = None        # uint32[2n-1]
        self._RightChild = None       # uint32[2n-1]
        self._TriangleIndices = None  # uint32[n]
        self._MortonCodes = None      # uint32[n] sorted (PlocParams::_MortonCodes)


class BVH:
    def __init__(self, nbTriangles, unsortedTriangles, meshesInTheScene, ctx: capi.Context = None,
                 search_radius: int = 16):
        self._ctx = ctx if ctx is not None else capi.Context(0)
        tris = np.ascontiguousarray(unsortedTriangles, dtype=TRIANGLE)
        meshes = np.ascontiguousarray(meshesInTheScene, dtype=MESH)
        self._handle = capi.Bvh(self._ctx).build(tris, meshes, n=int(nbTriangles), search_radius=search_radius)
        p = BVH_Params()
        p._NbTriangles = int(nbTriangles)
        p._Clusters, p._Parent, p._LeftChild, p._RightChild, p._IsLeaf = self._handle.clusters()
        p._TriangleIndices = self._handle.triangle_indices()
        p._MortonCodes = self._handle.morton_codes()
        self._InternalStruct = p

    @property
    def handle(self) -> capi.Bvh:
        return self._handle


def getBVH_NodesToGPUData(bvh: BVH) -> np.ndarray:
    """glr::Scene::getBVH_NodesToGPUData (scene.cpp:203-208): DFS pre-order BVH_NodeGPU array."""
    return bvh.handle.flat_nodes()


FORWARD, BACKWARD, LEFT, RIGHT, UP, DOWN = range(6)  # cr::CameraMovement, camera.hpp:10-17


class Camera:
    """cr::Camera (srcCommon/scene/camera.hpp:35-106): keeps the construction arguments and the input events and lets
    librtr_b200 replay them through the same code as the C++ shim's cr::Camera, so getGpuData() is the reference's
    CameraGPU bit for bit (tests/test_camera_cpu.py)."""

    def __init__(self, position, aspectRatio: float, fov: float = 45.0, near: float = 0.1, far: float = 200.0):
        self._position = tuple(float(v) for v in position)
        self._AspectRatio, self._Fov, self._Near, self._Far = float(aspectRatio), float(fov), float(near), float(far)
        self._Accelerate = False
        self._events = []

    def processKeyboard(self, direction: int, deltaTime: float):
        self._events.append((int(direction) + 1, float(deltaTime), 1.0 if self._Accelerate else 0.0))

    def ProcessMouseMovement(self, xoffset: float, yoffset: float):
        self._events.append((0, float(xoffset), float(yoffset)))

    def getGpuData(self) -> np.ndarray:
        return capi.camera_gpu_data(self._position, self._AspectRatio, self._Fov, self._Near, self._Far, self._events)


class Mesh:
    """cr::Mesh (srcCommon/scene/geometry/mesh.hpp:17-44): a triangle list, a MeshModelGPU and the id every triangle
    of the mesh carries in _ModelId.  All arithmetic is librtr_b200's (rtr_obj_load / rtr_mesh_*), so the records are
    the bits the reference produces."""

    _IdGenerator = 0  # mesh.cpp:9
    MAX_NB_MESHES = MAX_NB_MESHES

    def __init__(self, triangles=None):
        self._Id = Mesh._IdGenerator
        Mesh._IdGenerator += 1
        self._InternalStruct = np.zeros(1, dtype=MESH)
        capi.load_library().rtr_mesh_init(capi._ptr(self._InternalStruct))
        self._Triangles = np.zeros(0, dtype=TRIANGLE) if triangles is None else np.array(triangles, dtype=TRIANGLE)
        self._Triangles["model_id"] = self._Id

    def setModel(self, model):
        """mesh.cpp:24-26; `model` is 16 floats, column-major like glm::mat4."""
        m = np.ascontiguousarray(model, dtype=np.float32).reshape(16)
        capi.load_library().rtr_mesh_set_model(capi._ptr(self._InternalStruct), capi._ptr(m))

    def setPosition(self, position):
        x, y, z = (float(v) for v in position)
        capi.load_library().rtr_mesh_set_position(capi._ptr(self._InternalStruct), x, y, z)

    def setScale(self, scale: float):
        capi.load_library().rtr_mesh_set_scale(capi._ptr(self._InternalStruct), float(scale))

    def setRotation(self, thetaX: float, thetaY: float, thetaZ: float):
        capi.load_library().rtr_mesh_set_rotation(capi._ptr(self._InternalStruct), float(thetaX), float(thetaY), float(thetaZ))

    def setMaterial(self, materialId: int):
        capi.load_library().rtr_mesh_set_material(capi._ptr(self._InternalStruct), int(materialId))

    @staticmethod
    def _primitive(which):
        m = Mesh()
        m._Triangles = capi.mesh_primitive(which, m._Id)
        return m

    @staticmethod
    def primitiveTriangle():
        return Mesh._primitive(capi.PRIMITIVE_TRIANGLE)

    @staticmethod
    def primitiveSquare():
        return Mesh._primitive(capi.PRIMITIVE_SQUARE)

    @staticmethod
    def primitiveCube():
        return Mesh._primitive(capi.PRIMITIVE_CUBE)

    @staticmethod
    def primitiveSphere():
        return Mesh._primitive(capi.PRIMITIVE_SPHERE)

    @staticmethod
    def load(path: str):
        """mesh.cpp:186-263"""
        m = Mesh()
        m._Triangles = capi.load_obj(path, m._Id)
        return m


class Scene:
    """The slice of glr::Scene that feeds the accelerated path."""

    def __init__(self, ctx: capi.Context = None, reference_padding: bool = False):
        self._ctx = ctx
        self._Meshes = []  # list of (triangles TRIANGLE[], mesh MESH[1])
        self._NbTriangles = 0
        self._NbMeshes = 0
        self._BVH = None
        self._reference_padding = reference_padding
        self._Materials = []  # [r, g, b, a] float32 each
        self._NbMaterials = 0

    def addMesh(self, triangles, model=None, material_id: int = 0):
        """scene.cpp:42-47.  A `Mesh` is taken as it is (its triangles carry Mesh::_Id, which is the mesh slot only
        when meshes are added in creation order -- as in the reference); a bare TRIANGLE array gets a mesh record
        made here and _ModelId = the mesh slot."""
        if self._reference_padding and len(self._Meshes) == MAX_NB_MESHES:
            return
        if isinstance(triangles, Mesh):
            tris, mesh = triangles._Triangles, triangles._InternalStruct
            self._Meshes.append((tris, mesh))
            self._NbMeshes += 1
            self._NbTriangles += tris.size if not self._reference_padding else min(tris.size, MAX_NB_TRIANGLES)
            return
        mesh = np.zeros(1, dtype=MESH)
        mesh["m"][0] = (np.eye(4, dtype=np.float32) if model is None else np.asarray(model, dtype=np.float32)).T.reshape(16)
        mesh["material_id"][0] = material_id
        tris = np.array(triangles, dtype=TRIANGLE)
        tris["model_id"] = len(self._Meshes)
        self._Meshes.append((tris, mesh))
        self._NbMeshes += 1
        self._NbTriangles += tris.size if not self._reference_padding else min(tris.size, MAX_NB_TRIANGLES)

    def addMaterial(self, color):
        """scene.cpp:57-61"""
        if len(self._Materials) == MAX_NB_MATERIALS:
            return
        self._Materials.append(np.asarray(color, dtype=np.float32).reshape(4).copy())
        self._NbMaterials += 1

    def addRandomMaterial(self):
        """scene.cpp:63-67 + material.cpp:17-27: three draws of the C library's rand()."""
        import ctypes
        libc = ctypes.CDLL(None)
        libc.rand.restype = ctypes.c_int
        rand_max = np.float32(2147483647)  # RAND_MAX of glibc
        rgb = [np.float32(libc.rand()) / rand_max for _ in range(3)]
        self.addMaterial(rgb + [1.0])

    def getMaterialToGPUData(self) -> np.ndarray:
        """scene.cpp:42-48: [n, 4]; with reference_padding MAX_NB_MATERIALS entries, the unused ones a default
        MaterialGPU = (1, 1, 1, 1) (material.hpp:10)."""
        mats = np.stack(self._Materials) if self._Materials else np.zeros((0, 4), dtype=np.float32)
        if not self._reference_padding:
            return mats
        out = np.ones((MAX_NB_MATERIALS, 4), dtype=np.float32)
        out[:mats.shape[0]] = mats
        return out

    def getTriangleToGPUData(self) -> np.ndarray:
        """scene.cpp:26-40 (zero padded to MAX_NB_TRIANGLES only with reference_padding)."""
        allt = np.concatenate([t for t, _ in self._Meshes]) if self._Meshes else np.zeros(0, dtype=TRIANGLE)
        if not self._reference_padding:
            return allt
        out = np.zeros(MAX_NB_TRIANGLES, dtype=TRIANGLE)
        k = min(allt.size, MAX_NB_TRIANGLES)
        out[:k] = allt[:k]
        return out

    def getMeshModelToGPUData(self) -> np.ndarray:
        """scene.cpp:17-23"""
        allm = np.concatenate([m for _, m in self._Meshes]) if self._Meshes else np.zeros(0, dtype=MESH)
        if not self._reference_padding:
            return allm
        out = np.zeros(MAX_NB_MESHES, dtype=MESH)
        out["m"][:] = np.eye(4, dtype=np.float32).reshape(16)  # MeshModelGPU default, mesh.hpp:13
        out[:allm.size] = allm
        return out

    def sendDataToGpu(self):
        """scene.cpp:210-220 -> bindSSBO :110-187: build the BVH (scene.cpp:148) and return the flat
        node array that the reference uploads to SSBO binding 5."""
        nb = min(self._NbTriangles, MAX_NB_TRIANGLES) if self._reference_padding else self._NbTriangles
        self._BVH = BVH(nb, self.getTriangleToGPUData(), self.getMeshModelToGPUData(), ctx=self._ctx)
        return getBVH_NodesToGPUData(self._BVH)


__all__ = ["BVH", "BVH_Params", "Scene", "Mesh", "Camera", "FORWARD", "BACKWARD", "LEFT", "RIGHT", "UP", "DOWN", "getBVH_NodesToGPUData", "MAX_NB_TRIANGLES", "MAX_NB_MESHES", "MAX_NB_MATERIALS", "NONE"]
