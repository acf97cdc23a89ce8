"""ctypes binding of librtr_b200.so (include/rtr.h).  No fallback: if the CUDA library is
missing or no B200 is present the constructors raise -- nothing here computes on the CPU."""
import ctypes as C
import os

import numpy as np

from . import build as _build
from .layouts import CAMERA, HIT, MESH, NODE, RAY, TRIANGLE

RTR_OK = 0
ERROR_NAMES = {0: "RTR_OK", -1: "RTR_E_INVALID", -2: "RTR_E_CUDA", -3: "RTR_E_NOMEM", -4: "RTR_E_NODEVICE",
               -5: "RTR_E_UNSUPPORTED", -6: "RTR_E_STATE", -7: "RTR_E_COMM"}
TRACE_DEFAULT = 0
SHADE_WIREFRAME = 1  # RTR_SHADE_WIREFRAME
SHADE_BVH = 2        # RTR_SHADE_BVH
TRACE_REFERENCE_ORDER = 1
TRACE_DEEP_STACK = 2   # the shader's loop with the shader's own 1024-entry stack (raytracer.glsl:251)
NCCL_UNIQUE_ID_BYTES = 128

# every symbol include/rtr.h declares (tests/test_abi.py checks the header against this list and the .so)
SYMBOLS = [
    "rtr_version", "rtr_ctx_create", "rtr_ctx_destroy", "rtr_last_error", "rtr_ctx_sync", "rtr_ctx_stream",
    "rtr_ctx_set_stream", "rtr_ctx_device", "rtr_ctx_sm_count", "rtr_ctx_launch_count",
    "rtr_host_alloc", "rtr_host_free", "rtr_host_register", "rtr_host_unregister", "rtr_dev_alloc", "rtr_dev_free", "rtr_dev_upload", "rtr_dev_download",
    "rtr_dev_zero",
    "rtr_bit_histogram32", "rtr_bit_histogram32_dev", "rtr_digitplace_exclusive_scan",
    "rtr_digitplace_exclusive_scan_dev",
    "rtr_sort_keys_u32", "rtr_sort_pairs_u32", "rtr_sort_keys_u64", "rtr_sort_pairs_u64",
    "rtr_sort_pairs_u32_dev", "rtr_sort_pairs_u64_dev",
    "rtr_morton_codes", "rtr_morton_codes_dev", "rtr_morton_codes64", "rtr_scene_bounds",
    "rtr_bvh_build", "rtr_bvh_build_dev", "rtr_bvh_destroy", "rtr_bvh_nb_triangles", "rtr_bvh_nb_nodes",
    "rtr_bvh_iteration_trace", "rtr_bvh_iteration_times", "rtr_bvh_enable_stage_timing", "rtr_bvh_stage_ms", "rtr_bvh_morton_codes",
    "rtr_bvh_triangle_indices", "rtr_bvh_clusters", "rtr_bvh_flat_nodes", "rtr_bvh_device_nodes",
    "rtr_bvh_device_triangles", "rtr_bvh_device_meshes", "rtr_bvh_adopt_dev",
    "rtr_trace_primary", "rtr_trace_primary_dev", "rtr_trace_rays", "rtr_trace_rays_dev", "rtr_render",
    "rtr_render_dev", "rtr_render_sharded_dev", "rtr_ctx_profile_enable", "rtr_ctx_profile_read",
    "rtr_comm_unique_id", "rtr_comm_init", "rtr_comm_destroy", "rtr_bvh_broadcast", "rtr_allgather_rows",
    "rtr_render_stripes_dev", "rtr_allgather_stripes", "rtr_ctx_switch_stream", "rtr_ctx_reserve_sms", "rtr_ctx_partition_sms", "rtr_bvh_broadcast_traversal",
    "rtr_dev_upload_async", "rtr_dev_download_async", "rtr_gather_stripes", "rtr_shade", "rtr_shade_dev",
    "rtr_gather_slices", "rtr_download_stripes_async",
    "rtr_bvh_build64", "rtr_bvh_build64_dev", "rtr_bvh_morton_codes64",
    "rtr_bvh_depth_overlay", "rtr_bvh_depth_overlay_dev",
    "rtr_obj_load", "rtr_obj_parse", "rtr_obj_free", "rtr_mesh_primitive", "rtr_mesh_init", "rtr_mesh_set_model",
    "rtr_mesh_set_position", "rtr_mesh_set_scale", "rtr_mesh_set_rotation", "rtr_mesh_set_material",
    "rtr_triangle_centroid", "rtr_bvh_stack_overflows", "rtr_camera_gpu_data",
]


class RtrError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("%s (%d): %s" % (ERROR_NAMES.get(code, "RTR_E_?"), code, message))
        self.code = code


_lib = None


def library_path() -> str:
    return _build.LIB_PATH


def load_library():
    """dlopen librtr_b200.so; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise FileNotFoundError(
            "%s is missing: run `python -m realtimeraytracing_b200.build` (there is no CPU fallback)" % path)
    L = C.CDLL(path)
    vp, u32, u64, i32, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_size_t
    pp = C.POINTER(C.c_void_p)
    L.rtr_version.restype = C.c_char_p
    L.rtr_ctx_create.argtypes = [i32, pp]
    L.rtr_ctx_destroy.argtypes = [vp]
    L.rtr_last_error.argtypes = [vp]
    L.rtr_last_error.restype = C.c_char_p
    L.rtr_ctx_sync.argtypes = [vp]
    L.rtr_ctx_stream.argtypes = [vp]
    L.rtr_ctx_stream.restype = vp
    L.rtr_ctx_set_stream.argtypes = [vp, vp]
    L.rtr_ctx_device.argtypes = [vp]
    L.rtr_ctx_sm_count.argtypes = [vp]
    L.rtr_ctx_launch_count.argtypes = [vp]
    L.rtr_ctx_launch_count.restype = u64
    L.rtr_host_alloc.argtypes = [sz, pp]
    L.rtr_host_register.argtypes = [vp, sz]
    L.rtr_host_unregister.argtypes = [vp]
    L.rtr_host_free.argtypes = [vp]
    L.rtr_dev_alloc.argtypes = [vp, sz, pp]
    L.rtr_dev_free.argtypes = [vp, vp]
    L.rtr_dev_upload.argtypes = [vp, vp, vp, sz]
    L.rtr_dev_download.argtypes = [vp, vp, vp, sz]
    L.rtr_dev_upload_async.argtypes = [vp, vp, vp, sz]
    L.rtr_dev_download_async.argtypes = [vp, vp, vp, sz]
    L.rtr_dev_zero.argtypes = [vp, vp, sz]
    L.rtr_bit_histogram32.argtypes = [vp, vp, u32, vp]
    L.rtr_bit_histogram32_dev.argtypes = [vp, vp, u32, vp]
    L.rtr_digitplace_exclusive_scan.argtypes = [vp, vp, vp]
    L.rtr_digitplace_exclusive_scan_dev.argtypes = [vp, vp, vp]
    L.rtr_sort_keys_u32.argtypes = [vp, vp, u32]
    L.rtr_sort_pairs_u32.argtypes = [vp, vp, vp, u32]
    L.rtr_sort_keys_u64.argtypes = [vp, vp, u32]
    L.rtr_sort_pairs_u64.argtypes = [vp, vp, vp, u32]
    L.rtr_sort_pairs_u32_dev.argtypes = [vp, vp, vp, u32, i32, i32]
    L.rtr_sort_pairs_u64_dev.argtypes = [vp, vp, vp, u32, i32, i32]
    L.rtr_morton_codes.argtypes = [vp, vp, u32, u32, vp, u32, vp]
    L.rtr_morton_codes_dev.argtypes = [vp, vp, u32, u32, vp, u32, vp]
    L.rtr_morton_codes64.argtypes = [vp, vp, u32, u32, vp, u32, vp]
    L.rtr_scene_bounds.argtypes = [vp, vp, u32, vp, u32, vp]
    L.rtr_bvh_build.argtypes = [vp, vp, u32, u32, vp, u32, u32, pp]
    L.rtr_bvh_build_dev.argtypes = [vp, vp, u32, u32, vp, u32, u32, pp]
    L.rtr_bvh_build64.argtypes = [vp, vp, u32, u32, vp, u32, u32, pp]
    L.rtr_bvh_build64_dev.argtypes = [vp, vp, u32, u32, vp, u32, u32, pp]
    L.rtr_bvh_morton_codes64.argtypes = [vp, vp]
    L.rtr_bvh_destroy.argtypes = [vp]
    L.rtr_bvh_nb_triangles.argtypes = [vp]
    L.rtr_bvh_nb_triangles.restype = u32
    L.rtr_bvh_nb_nodes.argtypes = [vp]
    L.rtr_bvh_nb_nodes.restype = u32
    L.rtr_bvh_iteration_trace.argtypes = [vp, vp, vp, u32, vp]
    L.rtr_bvh_iteration_times.argtypes = [vp, vp, u32, vp]
    L.rtr_bvh_enable_stage_timing.argtypes = [vp, i32]
    L.rtr_bvh_stage_ms.argtypes = [vp, vp]
    L.rtr_bvh_morton_codes.argtypes = [vp, vp]
    L.rtr_bvh_triangle_indices.argtypes = [vp, vp]
    L.rtr_bvh_clusters.argtypes = [vp, vp, vp, vp, vp, vp]
    L.rtr_bvh_flat_nodes.argtypes = [vp, vp]
    for name in ("rtr_bvh_device_nodes", "rtr_bvh_device_triangles", "rtr_bvh_device_meshes"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = vp
    L.rtr_bvh_adopt_dev.argtypes = [vp, vp, u32, vp, vp, u32, pp]
    L.rtr_trace_primary.argtypes = [vp, vp, vp, u32, u32, u32, u32, u32, vp]
    L.rtr_trace_primary_dev.argtypes = [vp, vp, vp, u32, u32, u32, u32, u32, u32, u32, vp]
    L.rtr_trace_rays.argtypes = [vp, vp, vp, u64, i32, vp, u32, vp]
    L.rtr_trace_rays_dev.argtypes = [vp, vp, vp, u64, i32, vp, u32, vp]
    L.rtr_render.argtypes = [vp, vp, vp, u32, u32, u32, u32, u32, u32, u32, i32, vp, u32, vp, vp, vp]
    L.rtr_render_dev.argtypes = [vp, vp, vp, u32, u32, u32, u32, u32, u32, u32, i32, vp, u32, vp, vp, vp]
    L.rtr_render_sharded_dev.argtypes = [vp, vp, vp, u32, u32, u32, u32, u32, u32, u32, u32, i32, vp, u32, vp, vp, vp]
    L.rtr_ctx_profile_enable.argtypes = [vp, i32]
    L.rtr_ctx_profile_read.argtypes = [vp, vp, sz, vp, vp, u32, vp]
    L.rtr_comm_unique_id.argtypes = [vp]
    L.rtr_comm_init.argtypes = [vp, vp, i32, i32]
    L.rtr_comm_destroy.argtypes = [vp]
    L.rtr_bvh_broadcast.argtypes = [vp, pp, i32]
    L.rtr_bvh_broadcast_traversal.argtypes = [vp, pp, i32, u32]
    L.rtr_allgather_rows.argtypes = [vp, vp, u32, u32, u32, u32]
    L.rtr_render_stripes_dev.argtypes = [vp, vp, vp, u32, u32, u32, u32, u32, vp, u32, u32, u32, i32, vp, u32, vp, vp, vp]
    L.rtr_allgather_stripes.argtypes = [vp, vp, u32, u32, u32, u32, vp]
    L.rtr_gather_stripes.argtypes = [vp, vp, u32, u32, u32, u32, vp, i32]
    L.rtr_gather_slices.argtypes = [vp, vp, vp, C.c_uint64, u32, i32]
    L.rtr_download_stripes_async.argtypes = [vp, vp, vp, u32, u32, u32, u32, vp]
    L.rtr_shade.argtypes = [vp, vp, C.c_uint64, vp, u32, vp, u32, vp, u32, u32, vp, vp]
    L.rtr_shade_dev.argtypes = [vp, vp, C.c_uint64, vp, vp, vp, u32, vp, vp]
    L.rtr_bvh_depth_overlay.argtypes = [vp, vp, vp, u32, u32, u32, u32, i32, vp]
    L.rtr_bvh_depth_overlay_dev.argtypes = [vp, vp, vp, u32, u32, u32, u32, i32, vp]
    L.rtr_ctx_switch_stream.argtypes = [vp, vp]
    L.rtr_ctx_reserve_sms.argtypes = [vp, u32]
    L.rtr_ctx_partition_sms.argtypes = [vp, u32, u32, vp, vp]
    f32 = C.c_float
    L.rtr_bvh_stack_overflows.argtypes = [vp, C.POINTER(u32)]
    L.rtr_camera_gpu_data.argtypes = [vp, C.c_float, C.c_float, C.c_float, C.c_float, i32, vp, vp, vp, vp]
    L.rtr_obj_load.argtypes = [C.c_char_p, u32, pp, C.POINTER(u64)]
    L.rtr_obj_parse.argtypes = [C.c_char_p, u64, u32, pp, C.POINTER(u64)]
    L.rtr_obj_free.argtypes = [vp]
    L.rtr_obj_free.restype = None
    L.rtr_mesh_primitive.argtypes = [i32, u32, vp, u64, C.POINTER(u64)]
    for name, args in (("rtr_mesh_init", [vp]), ("rtr_mesh_set_model", [vp, vp]), ("rtr_mesh_set_position", [vp, f32, f32, f32]),
                       ("rtr_mesh_set_scale", [vp, f32]), ("rtr_mesh_set_rotation", [vp, f32, f32, f32]),
                       ("rtr_mesh_set_material", [vp, u32]), ("rtr_triangle_centroid", [vp, vp, vp])):
        getattr(L, name).argtypes = args
        getattr(L, name).restype = None
    _lib = L
    return L


def _ptr(a):
    """Host numpy array, raw device address (int) or None -> c_void_p."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))


def _as(a, dtype):
    a = np.ascontiguousarray(a)
    if a.dtype != dtype:
        if dtype.fields is not None:
            raise TypeError("expected an array of dtype %s" % (dtype,))
        a = a.astype(dtype)
    return a


def _check_host(rc):
    """Host-only entry points (no ctx) report through rtr_last_error(NULL)."""
    if rc != 0:
        msg = load_library().rtr_last_error(None)
        raise RtrError(rc, msg.decode() if msg else "")


def host_register(array: np.ndarray):
    """Page-locks the memory of `array` (e.g. an np.memmap shared by all ranks) for asynchronous copies."""
    _check_host(load_library().rtr_host_register(C.c_void_p(array.ctypes.data), array.nbytes))


def host_unregister(array: np.ndarray):
    load_library().rtr_host_unregister(C.c_void_p(array.ctypes.data))


def _take_triangles(out, n) -> np.ndarray:
    L = load_library()
    try:
        tris = np.zeros(n.value, dtype=TRIANGLE)
        if n.value:
            C.memmove(tris.ctypes.data, out.value, n.value * TRIANGLE.itemsize)
        return tris
    finally:
        L.rtr_obj_free(out)


def load_obj(path: str, model_id: int = 0) -> np.ndarray:
    """cr::Mesh::load (mesh.cpp:186-263): TRIANGLE[] of a Wavefront OBJ file, in face order."""
    L = load_library()
    out, n = C.c_void_p(), C.c_uint64()
    _check_host(L.rtr_obj_load(os.fsencode(path), model_id, C.byref(out), C.byref(n)))
    return _take_triangles(out, n)


def parse_obj(text, model_id: int = 0) -> np.ndarray:
    """The same from the file's bytes."""
    L = load_library()
    data = text.encode() if isinstance(text, str) else bytes(text)
    out, n = C.c_void_p(), C.c_uint64()
    _check_host(L.rtr_obj_parse(data, len(data), model_id, C.byref(out), C.byref(n)))
    return _take_triangles(out, n)


def camera_gpu_data(eye, aspect: float, fov: float = 45.0, near: float = 0.1, far: float = 200.0, events=()) -> np.ndarray:
    """cr::Camera::getGpuData() (camera.cpp:23-34) after replaying `events` = [(kind, a, b), ...]; see rtr.h."""
    e = np.ascontiguousarray(eye, dtype=np.float32).reshape(3)
    kind = np.array([ev[0] for ev in events], dtype=np.int32)
    a = np.array([ev[1] for ev in events], dtype=np.float32)
    b = np.array([ev[2] for ev in events], dtype=np.float32)
    out = np.zeros(1, dtype=CAMERA)
    _check_host(load_library().rtr_camera_gpu_data(_ptr(e), aspect, fov, near, far, len(events), _ptr(kind), _ptr(a), _ptr(b), _ptr(out)))
    return out


PRIMITIVE_TRIANGLE, PRIMITIVE_SQUARE, PRIMITIVE_CUBE, PRIMITIVE_SPHERE = 0, 1, 2, 3


def mesh_primitive(which: int, model_id: int = 0) -> np.ndarray:
    """cr::Mesh::primitiveTriangle/Square/Cube/Sphere (mesh.cpp:64-183)."""
    L = load_library()
    tris = np.zeros(12, dtype=TRIANGLE)
    n = C.c_uint64()
    _check_host(L.rtr_mesh_primitive(which, model_id, _ptr(tris), tris.size, C.byref(n)))
    return tris[:n.value].copy()


def triangle_centroid(tri: np.ndarray, model) -> np.ndarray:
    """cr::Triangle::getCentroid(triangle, model) (triangle.cpp:30-32); model is the 16 floats, column-major."""
    t = np.ascontiguousarray(tri, dtype=TRIANGLE).reshape(1)
    m = np.ascontiguousarray(model, dtype=np.float32).reshape(16)
    out = np.zeros(3, dtype=np.float32)
    load_library().rtr_triangle_centroid(_ptr(t), _ptr(m), _ptr(out))
    return out


class Context:
    """rtr_ctx: one per (host thread, GPU)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.rtr_ctx_create(device, C.byref(h))
        if rc != RTR_OK:
            raise RtrError(rc, (self.lib.rtr_last_error(None) or b"").decode())
        self.handle = h

    def check(self, rc):
        if rc != RTR_OK:
            raise RtrError(rc, (self.lib.rtr_last_error(self.handle) or b"").decode())

    def close(self):
        if getattr(self, "handle", None):
            self.lib.rtr_ctx_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing ----
    def sync(self):
        self.check(self.lib.rtr_ctx_sync(self.handle))

    @property
    def stream(self) -> int:
        return int(self.lib.rtr_ctx_stream(self.handle) or 0)

    def set_stream(self, stream: int):
        self.check(self.lib.rtr_ctx_set_stream(self.handle, C.c_void_p(stream)))

    @property
    def sm_count(self) -> int:
        return int(self.lib.rtr_ctx_sm_count(self.handle))

    @property
    def launch_count(self) -> int:
        return int(self.lib.rtr_ctx_launch_count(self.handle))

    def profile_enable(self, enable: bool = True):
        self.check(self.lib.rtr_ctx_profile_enable(self.handle, 1 if enable else 0))

    def profile_read(self):
        """{kernel name: (total ms, launches)} since profile_enable(True)."""
        cap = 64
        names = C.create_string_buffer(8192)
        ms = np.zeros(cap, dtype=np.float32)
        cnt = np.zeros(cap, dtype=np.uint32)
        n = C.c_uint32(0)
        self.check(self.lib.rtr_ctx_profile_read(self.handle, names, 8192, _ptr(ms), _ptr(cnt), cap, C.byref(n)))
        keys = [k for k in names.value.decode().split("\n") if k]
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(keys)}

    def dev_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self.check(self.lib.rtr_dev_alloc(self.handle, nbytes, C.byref(p)))
        return int(p.value)

    def dev_free(self, ptr: int):
        self.check(self.lib.rtr_dev_free(self.handle, C.c_void_p(ptr)))

    def upload(self, dst_dev: int, src: np.ndarray):
        src = np.ascontiguousarray(src)
        self.check(self.lib.rtr_dev_upload(self.handle, C.c_void_p(dst_dev), _ptr(src), src.nbytes))

    def download(self, dst: np.ndarray, src_dev: int):
        self.check(self.lib.rtr_dev_download(self.handle, _ptr(dst), C.c_void_p(src_dev), dst.nbytes))

    def upload_async(self, dst_dev: int, src: np.ndarray):
        """Enqueue only; `src` must stay alive (and be pinned for the copy to overlap kernels)."""
        self.check(self.lib.rtr_dev_upload_async(self.handle, C.c_void_p(dst_dev), _ptr(src), src.nbytes))

    def download_async(self, dst: np.ndarray, src_dev: int):
        self.check(self.lib.rtr_dev_download_async(self.handle, _ptr(dst), C.c_void_p(src_dev), dst.nbytes))

    def zero(self, dst_dev: int, nbytes: int):
        self.check(self.lib.rtr_dev_zero(self.handle, C.c_void_p(dst_dev), nbytes))

    # ---- testsSortGPU stage entries ----
    def bit_histogram32(self, keys) -> np.ndarray:
        keys = _as(keys, np.dtype(np.uint32))
        out = np.zeros(32, dtype=np.uint32)
        self.check(self.lib.rtr_bit_histogram32(self.handle, _ptr(keys), keys.size, _ptr(out)))
        return out

    def bit_histogram32_dev(self, keys_dev: int, n: int, out_dev: int):
        self.check(self.lib.rtr_bit_histogram32_dev(self.handle, C.c_void_p(keys_dev), n, C.c_void_p(out_dev)))

    def digitplace_exclusive_scan(self, hist) -> np.ndarray:
        hist = _as(hist, np.dtype(np.uint32))
        if hist.size != 32:
            raise ValueError("expects the 32-bin histogram")
        out = np.zeros(32, dtype=np.uint32)
        self.check(self.lib.rtr_digitplace_exclusive_scan(self.handle, _ptr(hist), _ptr(out)))
        return out

    def digitplace_exclusive_scan_dev(self, in_dev: int, out_dev: int):
        self.check(self.lib.rtr_digitplace_exclusive_scan_dev(self.handle, C.c_void_p(in_dev), C.c_void_p(out_dev)))

    # ---- sort ----
    def sort_keys_u32(self, keys) -> np.ndarray:
        keys = np.array(keys, dtype=np.uint32)
        self.check(self.lib.rtr_sort_keys_u32(self.handle, _ptr(keys), keys.size))
        return keys

    def sort_pairs_u32(self, keys, values):
        keys = np.array(keys, dtype=np.uint32)
        values = np.array(values, dtype=np.uint32)
        if keys.size != values.size:
            raise ValueError("keys/values length mismatch")
        self.check(self.lib.rtr_sort_pairs_u32(self.handle, _ptr(keys), _ptr(values), keys.size))
        return keys, values

    def sort_keys_u64(self, keys) -> np.ndarray:
        keys = np.array(keys, dtype=np.uint64)
        self.check(self.lib.rtr_sort_keys_u64(self.handle, _ptr(keys), keys.size))
        return keys

    def sort_pairs_u64(self, keys, values):
        keys = np.array(keys, dtype=np.uint64)
        values = np.array(values, dtype=np.uint32)
        if keys.size != values.size:
            raise ValueError("keys/values length mismatch")
        self.check(self.lib.rtr_sort_pairs_u64(self.handle, _ptr(keys), _ptr(values), keys.size))
        return keys, values

    def sort_pairs_u32_dev(self, keys_dev: int, values_dev, n: int, begin_bit: int = 0, end_bit: int = 32):
        self.check(self.lib.rtr_sort_pairs_u32_dev(self.handle, C.c_void_p(keys_dev), _ptr(values_dev), n, begin_bit, end_bit))

    def sort_pairs_u64_dev(self, keys_dev: int, values_dev, n: int, begin_bit: int = 0, end_bit: int = 64):
        self.check(self.lib.rtr_sort_pairs_u64_dev(self.handle, C.c_void_p(keys_dev), _ptr(values_dev), n, begin_bit, end_bit))

    # ---- Morton ----
    def morton_codes(self, tris, meshes, n=None) -> np.ndarray:
        tris, meshes = _as(tris, TRIANGLE), _as(meshes, MESH)
        n = tris.size if n is None else n
        out = np.zeros(n, dtype=np.uint32)
        self.check(self.lib.rtr_morton_codes(self.handle, _ptr(tris), n, tris.size, _ptr(meshes), meshes.size, _ptr(out)))
        return out

    def morton_codes64(self, tris, meshes, n=None) -> np.ndarray:
        tris, meshes = _as(tris, TRIANGLE), _as(meshes, MESH)
        n = tris.size if n is None else n
        out = np.zeros(n, dtype=np.uint64)
        self.check(self.lib.rtr_morton_codes64(self.handle, _ptr(tris), n, tris.size, _ptr(meshes), meshes.size, _ptr(out)))
        return out

    def morton_codes_dev(self, tris_dev: int, n: int, array_len: int, meshes_dev: int, nb_meshes: int, codes_dev: int):
        self.check(self.lib.rtr_morton_codes_dev(self.handle, C.c_void_p(tris_dev), n, array_len, C.c_void_p(meshes_dev),
                                                 nb_meshes, C.c_void_p(codes_dev)))

    def scene_bounds(self, tris, meshes) -> np.ndarray:
        tris, meshes = _as(tris, TRIANGLE), _as(meshes, MESH)
        out = np.zeros(12, dtype=np.float32)
        self.check(self.lib.rtr_scene_bounds(self.handle, _ptr(tris), tris.size, _ptr(meshes), meshes.size, _ptr(out)))
        return out

    # ---- NCCL plumbing of the C ABI ----
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(NCCL_UNIQUE_ID_BYTES)
        rc = self.lib.rtr_comm_unique_id(buf)
        if rc != RTR_OK:
            raise RtrError(rc, (self.lib.rtr_last_error(None) or b"").decode())
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(bytes(unique_id), NCCL_UNIQUE_ID_BYTES)
        self.check(self.lib.rtr_comm_init(self.handle, buf, rank, nranks))

    def comm_destroy(self):
        self.check(self.lib.rtr_comm_destroy(self.handle))

    def shade(self, hits, tris, meshes, materials, wireframe: bool = False, bvh_rgba=None) -> np.ndarray:
        """getColor of raytracer.glsl (:159-179) on hit records: the reference's rgba32f pixels, [n, 4].
        bvh_rgba = Bvh.depth_overlay(...) when the BVH is displayed (uIsBVHDisplayed)."""
        hits = _as(hits, HIT); tris = _as(tris, TRIANGLE); meshes = _as(meshes, MESH)
        materials = np.ascontiguousarray(materials, dtype=np.float32).reshape(-1, 4)
        out = np.zeros((hits.size, 4), dtype=np.float32)
        flags = (SHADE_WIREFRAME if wireframe else 0) | (SHADE_BVH if bvh_rgba is not None else 0)
        b = None if bvh_rgba is None else np.ascontiguousarray(bvh_rgba, dtype=np.float32)
        self.check(self.lib.rtr_shade(self.handle, _ptr(hits), hits.size, _ptr(tris), tris.size, _ptr(meshes), meshes.size,
                                      _ptr(materials), materials.shape[0], flags, _ptr(b), _ptr(out)))
        return out

    def shade_dev(self, hits_dev: int, n: int, tris_dev: int, meshes_dev: int, materials_dev: int, rgba_dev: int,
                  wireframe: bool = False, bvh_rgba_dev=None):
        flags = (SHADE_WIREFRAME if wireframe else 0) | (SHADE_BVH if bvh_rgba_dev else 0)
        self.check(self.lib.rtr_shade_dev(self.handle, C.c_void_p(hits_dev), n, C.c_void_p(tris_dev), C.c_void_p(meshes_dev),
                                          C.c_void_p(materials_dev), flags, _ptr(bvh_rgba_dev), C.c_void_p(rgba_dev)))

    def gather_stripes(self, image_dev: int, width: int, height: int, bytes_per_pixel: int, rows_per_block: int,
                       stripes_of_rank, root: int = 0):
        """Like allgather_stripes, but only `root` receives the blocks (each travels once)."""
        st = np.ascontiguousarray(stripes_of_rank, dtype=np.uint32)
        self.check(self.lib.rtr_gather_stripes(self.handle, C.c_void_p(image_dev), width, height, bytes_per_pixel,
                                               rows_per_block, _ptr(st), root))

    def gather_slices(self, slice_dev: int, full_dev, n_elems: int, elem_bytes: int, root: int = 0):
        """Assemble on `root` an array whose contiguous slices (parallel.slice_range) the ranks uploaded themselves."""
        self.check(self.lib.rtr_gather_slices(self.handle, C.c_void_p(slice_dev), C.c_void_p(full_dev) if full_dev else None,
                                              n_elems, elem_bytes, root))

    def download_stripes_async(self, image_host_ptr: int, image_dev: int, width: int, height: int, bytes_per_pixel: int,
                               rows_per_block: int, stripes_of_rank):
        """This rank's row blocks of a striped frame, device -> the same rows of a full-size host image."""
        st = np.ascontiguousarray(stripes_of_rank, dtype=np.uint32)
        self.check(self.lib.rtr_download_stripes_async(self.handle, C.c_void_p(image_host_ptr), C.c_void_p(image_dev), width,
                                                       height, bytes_per_pixel, rows_per_block, _ptr(st)))

    def switch_stream(self, cuda_stream: int):
        """Enqueue later calls on `cuda_stream` without waiting for the work already enqueued (pipelined frames)."""
        self.check(self.lib.rtr_ctx_switch_stream(self.handle, C.c_void_p(cuda_stream)))

    def partition_sms(self, sms: int, n_streams: int = 2):
        """A green-context SM partition for the rays: returns (raw CUDA stream handles, SMs in the partition)."""
        out = (C.c_void_p * n_streams)()
        got = C.c_uint32(0)
        self.check(self.lib.rtr_ctx_partition_sms(self.handle, sms, n_streams, out, C.byref(got)))
        return [int(p) for p in out], got.value

    def reserve_sms(self, sms: int):
        self.check(self.lib.rtr_ctx_reserve_sms(self.handle, sms))

    def allgather_rows(self, image_dev: int, width: int, height: int, bytes_per_pixel: int, rows_per_block: int):
        self.check(self.lib.rtr_allgather_rows(self.handle, C.c_void_p(image_dev), width, height, bytes_per_pixel, rows_per_block))

    def allgather_stripes(self, image_dev: int, width: int, height: int, bytes_per_pixel: int, rows_per_block: int,
                          stripes_of_rank):
        """Gathers the blocks of a weighted dealing (parallel.stripe_layout)."""
        st = np.ascontiguousarray(stripes_of_rank, dtype=np.uint32)
        self.check(self.lib.rtr_allgather_stripes(self.handle, C.c_void_p(image_dev), width, height, bytes_per_pixel,
                                                  rows_per_block, _ptr(st)))


class Bvh:
    """rtr_bvh handle.  Build with Bvh.build / Bvh.build_dev, wrap received arrays with Bvh.adopt_dev."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.lib = ctx.lib
        self.handle = C.c_void_p()

    # -- construction --
    def build(self, tris, meshes, n=None, search_radius: int = 16, key_bits: int = 32):
        """key_bits 32: the reference's 30-bit Morton codes; 64: 63-bit codes (rtr_bvh_build64)."""
        tris, meshes = _as(tris, TRIANGLE), _as(meshes, MESH)
        n = tris.size if n is None else n
        fn = self.lib.rtr_bvh_build64 if key_bits == 64 else self.lib.rtr_bvh_build
        self.ctx.check(fn(self.ctx.handle, _ptr(tris), n, tris.size, _ptr(meshes), meshes.size,
                          search_radius, C.byref(self.handle)))
        return self

    def build_dev(self, tris_dev: int, n: int, array_len: int, meshes_dev: int, nb_meshes: int, search_radius: int = 16,
                  key_bits: int = 32):
        fn = self.lib.rtr_bvh_build64_dev if key_bits == 64 else self.lib.rtr_bvh_build_dev
        self.ctx.check(fn(self.ctx.handle, C.c_void_p(tris_dev), n, array_len,
                          C.c_void_p(meshes_dev), nb_meshes, search_radius, C.byref(self.handle)))
        return self

    def adopt_dev(self, nodes_dev: int, n: int, tris_dev: int, meshes_dev: int, nb_meshes: int):
        self.ctx.check(self.lib.rtr_bvh_adopt_dev(self.ctx.handle, C.c_void_p(nodes_dev), n, C.c_void_p(tris_dev),
                                                  C.c_void_p(meshes_dev), nb_meshes, C.byref(self.handle)))
        return self

    def broadcast(self, root: int = 0, traversal_only: bool = False, expected_triangles: int = 0):
        """Replicates the root's BVH on every rank; traversal_only sends just what the default traversal reads
        (and with expected_triangles set on every rank the call only enqueues)."""
        if traversal_only:
            self.ctx.check(self.lib.rtr_bvh_broadcast_traversal(self.ctx.handle, C.byref(self.handle), root, expected_triangles))
        else:
            self.ctx.check(self.lib.rtr_bvh_broadcast(self.ctx.handle, C.byref(self.handle), root))
        return self

    def close(self):
        if getattr(self, "handle", None) and self.handle.value and self.ctx.handle:
            self.lib.rtr_bvh_destroy(self.handle)
        self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- results --
    @property
    def nb_triangles(self) -> int:
        return int(self.lib.rtr_bvh_nb_triangles(self.handle))

    @property
    def nb_nodes(self) -> int:
        return int(self.lib.rtr_bvh_nb_nodes(self.handle))

    def enable_stage_timing(self, enable: bool = True):
        self.ctx.check(self.lib.rtr_bvh_enable_stage_timing(self.handle, 1 if enable else 0))

    def stage_ms(self) -> np.ndarray:
        out = np.zeros(6, dtype=np.float32)
        self.ctx.check(self.lib.rtr_bvh_stage_ms(self.handle, _ptr(out)))
        return out

    def iteration_trace(self):
        cap = 1 << 16
        active = np.zeros(cap, dtype=np.uint32)
        merges = np.zeros(cap, dtype=np.uint32)
        count = C.c_uint32(0)
        self.ctx.check(self.lib.rtr_bvh_iteration_trace(self.handle, _ptr(active), _ptr(merges), cap, C.byref(count)))
        k = min(cap, count.value)
        return active[:k].copy(), merges[:k].copy()

    def iteration_times(self) -> np.ndarray:
        """Device clock (ns) at the start of every PLOC iteration of the last build (+ the end of the loop)."""
        cap = (1 << 16) + 2
        t = np.zeros(cap, dtype=np.uint64)
        count = C.c_uint32(0)
        self.ctx.check(self.lib.rtr_bvh_iteration_times(self.handle, _ptr(t), cap, C.byref(count)))
        return t[:min(cap, count.value + 1)].copy()

    def morton_codes(self) -> np.ndarray:
        out = np.zeros(self.nb_triangles, dtype=np.uint32)
        self.ctx.check(self.lib.rtr_bvh_morton_codes(self.handle, _ptr(out)))
        return out

    def stack_overflows(self) -> int:
        """Rays that ran out of traversal stack since the last call (waits for the stream); see rtr.h."""
        n = C.c_uint32()
        self.ctx.check(self.lib.rtr_bvh_stack_overflows(self.handle, C.byref(n)))
        return n.value

    def depth_overlay(self, camera, width, height, depth: int, denom_w=0, denom_h=0) -> np.ndarray:
        """bvhColor of the shader's traversal (uDepthDisplayBVH = depth) per pixel, [height*width, 4]."""
        camera = _as(camera, CAMERA)
        out = np.zeros((width * height, 4), dtype=np.float32)
        self.ctx.check(self.lib.rtr_bvh_depth_overlay(self.ctx.handle, self.handle, _ptr(camera), width, height,
                                                      denom_w, denom_h, depth, _ptr(out)))
        return out

    def morton_codes64(self) -> np.ndarray:
        out = np.zeros(self.nb_triangles, dtype=np.uint64)
        self.ctx.check(self.lib.rtr_bvh_morton_codes64(self.handle, _ptr(out)))
        return out

    def triangle_indices(self) -> np.ndarray:
        out = np.zeros(self.nb_triangles, dtype=np.uint32)
        self.ctx.check(self.lib.rtr_bvh_triangle_indices(self.handle, _ptr(out)))
        return out

    def clusters(self):
        nc = self.nb_nodes
        clusters = np.zeros(nc, dtype=NODE)
        parent = np.zeros(nc, dtype=np.uint32)
        left = np.zeros(nc, dtype=np.uint32)
        right = np.zeros(nc, dtype=np.uint32)
        is_leaf = np.zeros(nc, dtype=np.uint8)
        self.ctx.check(self.lib.rtr_bvh_clusters(self.handle, _ptr(clusters), _ptr(parent), _ptr(left), _ptr(right), _ptr(is_leaf)))
        return clusters, parent, left, right, is_leaf

    def flat_nodes(self) -> np.ndarray:
        out = np.zeros(self.nb_nodes, dtype=NODE)
        self.ctx.check(self.lib.rtr_bvh_flat_nodes(self.handle, _ptr(out)))
        return out

    @property
    def device_nodes(self) -> int:
        return int(self.lib.rtr_bvh_device_nodes(self.handle) or 0)

    @property
    def device_triangles(self) -> int:
        return int(self.lib.rtr_bvh_device_triangles(self.handle) or 0)

    @property
    def device_meshes(self) -> int:
        return int(self.lib.rtr_bvh_device_meshes(self.handle) or 0)

    # -- traversal --
    def trace_primary(self, camera, width: int, height: int, denom_w: int = 0, denom_h: int = 0, flags: int = 0) -> np.ndarray:
        camera = _as(camera, CAMERA)
        hits = np.zeros(width * height, dtype=HIT)
        self.ctx.check(self.lib.rtr_trace_primary(self.ctx.handle, self.handle, _ptr(camera), width, height, denom_w,
                                                  denom_h, flags, _ptr(hits)))
        return hits

    def trace_primary_dev(self, camera, width, height, hits_dev: int, denom_w=0, denom_h=0, row0=0, row1=0, flags=0):
        camera = _as(camera, CAMERA)
        self.ctx.check(self.lib.rtr_trace_primary_dev(self.ctx.handle, self.handle, _ptr(camera), width, height, denom_w,
                                                      denom_h, row0, row1, flags, C.c_void_p(hits_dev)))

    def trace_rays(self, rays, any_hit: bool = False, t_max=None, flags: int = 0) -> np.ndarray:
        rays = _as(rays, RAY)
        hits = np.zeros(rays.size, dtype=HIT)
        tm = None if t_max is None else _as(t_max, np.dtype(np.float32))
        self.ctx.check(self.lib.rtr_trace_rays(self.ctx.handle, self.handle, _ptr(rays), rays.size, 1 if any_hit else 0,
                                               _ptr(tm), flags, _ptr(hits)))
        return hits

    def trace_rays_dev(self, rays_dev: int, n_rays: int, hits_dev: int, any_hit=False, t_max_dev=None, flags=0):
        self.ctx.check(self.lib.rtr_trace_rays_dev(self.ctx.handle, self.handle, C.c_void_p(rays_dev), n_rays,
                                                   1 if any_hit else 0, _ptr(t_max_dev), flags, C.c_void_p(hits_dev)))

    def render(self, camera, width, height, denom_w=0, denom_h=0, row0=0, row1=0, bounces=0, shadow=False,
               light=(0.0, 0.0, 0.0), flags=0, want_hits=True):
        camera = _as(camera, CAMERA)
        rows = (height if row1 == 0 else row1) - row0
        rgba = np.zeros((rows, width, 4), dtype=np.float32)
        hits = np.zeros(rows * width, dtype=HIT) if want_hits else None
        nrays = np.zeros(1, dtype=np.uint64)
        light = np.asarray(light, dtype=np.float32)
        self.ctx.check(self.lib.rtr_render(self.ctx.handle, self.handle, _ptr(camera), width, height, denom_w, denom_h,
                                           row0, row1, bounces, 1 if shadow else 0, _ptr(light), flags, _ptr(rgba),
                                           _ptr(hits), _ptr(nrays)))
        return rgba, hits, int(nrays[0])

    def render_sharded_dev(self, camera, width, height, rgba_dev, rows_per_block, shard_rank, shard_count, hits_dev=None,
                           rays_dev=None, denom_w=0, denom_h=0, bounces=0, shadow=False, light=(0.0, 0.0, 0.0), flags=0):
        camera = _as(camera, CAMERA)
        light = np.asarray(light, dtype=np.float32)
        self.ctx.check(self.lib.rtr_render_sharded_dev(self.ctx.handle, self.handle, _ptr(camera), width, height, denom_w,
                                                       denom_h, rows_per_block, shard_rank, shard_count, bounces,
                                                       1 if shadow else 0, _ptr(light), flags, _ptr(rgba_dev),
                                                       _ptr(hits_dev), _ptr(rays_dev)))

    def render_stripes_dev(self, camera, width, height, rgba_dev, rows_per_block, stripes_of_rank, rank,
                           hits_dev=None, rays_dev=None, denom_w=0, denom_h=0, bounces=0, shadow=False,
                           light=(0.0, 0.0, 0.0), flags=0):
        """Row blocks dealt with weights (parallel.stripe_layout): this rank renders the blocks of its stripes."""
        camera = _as(camera, CAMERA)
        light = np.asarray(light, dtype=np.float32)
        st = np.ascontiguousarray(stripes_of_rank, dtype=np.uint32)
        self.ctx.check(self.lib.rtr_render_stripes_dev(self.ctx.handle, self.handle, _ptr(camera), width, height, denom_w,
                                                       denom_h, rows_per_block, _ptr(st), st.size, rank,
                                                       bounces, 1 if shadow else 0, _ptr(light), flags, _ptr(rgba_dev),
                                                       _ptr(hits_dev), _ptr(rays_dev)))

    def render_dev(self, camera, width, height, rgba_dev, hits_dev=None, rays_dev=None, denom_w=0, denom_h=0, row0=0,
                   row1=0, bounces=0, shadow=False, light=(0.0, 0.0, 0.0), flags=0):
        camera = _as(camera, CAMERA)
        light = np.asarray(light, dtype=np.float32)
        self.ctx.check(self.lib.rtr_render_dev(self.ctx.handle, self.handle, _ptr(camera), width, height, denom_w, denom_h,
                                               row0, row1, bounces, 1 if shadow else 0, _ptr(light), flags,
                                               _ptr(rgba_dev), _ptr(hits_dev), _ptr(rays_dev)))
