"""ctypes bindings of the CPU oracle.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "librtr_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_ROOT = "/root/reference"

# numpy views of the struct layouts (kept local: the oracle does not import the product)
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


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def build_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "rtr_oracle.c")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return ORACLE_SO


def build_reference(caps=(65536, 1048576, 16777216)) -> bool:
    """Compile the reference's own bvh.cpp into oracle/_ref (only where /root/reference exists)."""
    if not os.path.isdir(REFERENCE_ROOT):
        return False
    subprocess.check_call(["make", "-s", "-C", HERE, "ref", "CAPS=" + " ".join(str(c) for c in caps)])
    return True


def oracle_available() -> bool:
    return os.path.exists(ORACLE_SO)


def reference_available(cap: int = 65536) -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libref_bvh_%d.so" % cap))


class OracleBvh:
    """Result of Oracle.bvh_build: numpy copies of every output array."""

    def __init__(self, n, morton_sorted, triangle_indices, clusters, parent, left, right, trace_active, trace_merges):
        self.n = n
        self.morton_sorted = morton_sorted
        self.triangle_indices = triangle_indices
        self.clusters = clusters
        self.parent = parent
        self.left = left
        self.right = right
        self.trace_active = trace_active
        self.trace_merges = trace_merges


class Oracle:
    def __init__(self):
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        self.lib = L
        vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
        L.orc_bit_histogram32.argtypes = [vp, u32, vp]
        L.orc_digitplace_exclusive_scan.argtypes = [vp, vp]
        L.orc_scene_aabb.argtypes = [vp, u32, vp, vp]
        L.orc_circumscribed_cube.argtypes = [vp, vp]
        L.orc_morton_codes.argtypes = [vp, u32, u32, vp, vp]
        L.orc_morton_codes64.argtypes = [vp, u32, u32, vp, vp]
        L.orc_sort_pairs.argtypes = [vp, vp, u32]
        L.orc_radix_sort_pairs.argtypes = [vp, vp, u32]
        L.orc_radix_sort_pairs_mt.argtypes = [vp, vp, u32, i32]
        L.orc_radix_sort_keys_u32.argtypes = [vp, u32]
        L.orc_radix_sort_keys_u64.argtypes = [vp, u32]
        L.orc_radix_sort_pairs_u64.argtypes = [vp, vp, u32]
        L.orc_bvh_build.argtypes = [vp, u32, u32, vp, u32, u32]
        L.orc_bvh_build.restype = vp
        L.orc_bvh_build64.argtypes = [vp, u32, u32, vp, u32, u32]
        L.orc_bvh_build64.restype = vp
        L.orc_bvh_destroy.argtypes = [vp]
        L.orc_bvh_nb_iterations.argtypes = [vp]
        L.orc_bvh_nb_iterations.restype = u32
        for name in ("morton_sorted", "triangle_indices", "clusters", "parent", "left", "right",
                     "trace_active", "trace_merges"):
            f = getattr(L, "orc_bvh_" + name)
            f.argtypes = [vp]
            f.restype = vp
        L.orc_flatten.argtypes = [vp, vp, vp, u32, vp]
        L.orc_hash_words.argtypes = [vp, u64]
        L.orc_hash_words.restype = u64
        L.orc_hash_flat_nodes.argtypes = [vp, u32]
        L.orc_hash_flat_nodes.restype = u64
        L.orc_get_ray.argtypes = [vp, u32, u32, u32, u32, vp]
        L.orc_get_rays.argtypes = [vp, u32, u32, u32, u32, vp]
        L.orc_ray_triangle.argtypes = [vp, vp, vp, u32, vp]
        L.orc_intersect_box.argtypes = [vp, vp]
        L.orc_intersect_box.restype = u32
        L.orc_intersect_box_edge.argtypes = [vp, vp, i32]
        L.orc_intersect_box_edge.restype = u32
        L.orc_closest_hit_bvh.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orc_closest_hit_brute.argtypes = [vp, vp, u32, vp, vp]
        L.orc_any_hit_bvh.argtypes = [vp, C.c_float, vp, vp, vp]
        L.orc_any_hit_bvh.restype = i32
        L.orc_trace_primary.argtypes = [vp, vp, vp, vp, u32, u32, u32, u32, vp, i32]
        L.orc_trace_rays.argtypes = [vp, vp, vp, vp, u64, vp, i32]
        L.orc_render.argtypes = [vp, vp, vp, vp, u32, u32, u32, u32, u32, u32, u32, i32, vp, vp, vp, vp, i32]
        L.orc_shade.argtypes = [vp, u64, vp, vp, vp, i32, vp, vp]
        L.orc_depth_overlay.argtypes = [vp, vp, u32, u32, u32, u32, i32, vp]
        L.orc_bounce_ray.argtypes = [vp, vp, vp, vp, vp]
        L.orc_bounce_ray.restype = i32
        L.orc_shadow_ray.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.orc_shadow_ray.restype = i32
        L.orc_num_threads.restype = i32

    # ---- sort pre-passes ----
    def bit_histogram32(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        out = np.zeros(32, dtype=np.uint32)
        self.lib.orc_bit_histogram32(_p(keys), keys.size, _p(out))
        return out

    def digitplace_exclusive_scan(self, hist):
        hist = np.ascontiguousarray(hist, dtype=np.uint32)
        out = np.zeros(32, dtype=np.uint32)
        self.lib.orc_digitplace_exclusive_scan(_p(hist), _p(out))
        return out

    # ---- morton / sort ----
    def scene_aabb(self, tris, meshes):
        out = np.zeros(6, dtype=np.float32)
        self.lib.orc_scene_aabb(_p(tris), tris.size, _p(meshes), _p(out))
        return out

    def circumscribed_cube(self, scene6):
        scene6 = np.ascontiguousarray(scene6, dtype=np.float32)
        out = np.zeros(6, dtype=np.float32)
        self.lib.orc_circumscribed_cube(_p(scene6), _p(out))
        return out

    def morton_codes(self, tris, meshes, n=None):
        n = tris.size if n is None else n
        out = np.zeros(n, dtype=np.uint32)
        self.lib.orc_morton_codes(_p(tris), n, tris.size, _p(meshes), _p(out))
        return out

    def morton_codes64(self, tris, meshes, n=None):
        n = tris.size if n is None else n
        out = np.zeros(n, dtype=np.uint64)
        self.lib.orc_morton_codes64(_p(tris), n, tris.size, _p(meshes), _p(out))
        return out

    def sort_pairs(self, codes, indices=None):
        codes = np.array(codes, dtype=np.uint32)
        idx = np.arange(codes.size, dtype=np.uint32) if indices is None else np.array(indices, dtype=np.uint32)
        self.lib.orc_sort_pairs(_p(codes), _p(idx), codes.size)
        return codes, idx

    def radix_sort_pairs(self, keys, vals):
        keys = np.array(keys, dtype=np.uint32)
        vals = np.array(vals, dtype=np.uint32)
        self.lib.orc_radix_sort_pairs(_p(keys), _p(vals), keys.size)
        return keys, vals

    def radix_sort_pairs_mt(self, keys, vals, threads=0):
        keys = np.array(keys, dtype=np.uint32)
        vals = np.array(vals, dtype=np.uint32)
        self.lib.orc_radix_sort_pairs_mt(_p(keys), _p(vals), keys.size, threads)
        return keys, vals

    def radix_sort_keys_u32(self, keys):
        keys = np.array(keys, dtype=np.uint32)
        self.lib.orc_radix_sort_keys_u32(_p(keys), keys.size)
        return keys

    def radix_sort_keys_u64(self, keys):
        keys = np.array(keys, dtype=np.uint64)
        self.lib.orc_radix_sort_keys_u64(_p(keys), keys.size)
        return keys

    def radix_sort_pairs_u64(self, keys, vals):
        keys = np.array(keys, dtype=np.uint64)
        vals = np.array(vals, dtype=np.uint32)
        self.lib.orc_radix_sort_pairs_u64(_p(keys), _p(vals), keys.size)
        return keys, vals

    # ---- PLOC + flatten ----
    def bvh_build(self, tris, meshes, n=None, search_radius=16, key_bits=32) -> OracleBvh:
        n = tris.size if n is None else n
        fn = self.lib.orc_bvh_build64 if key_bits == 64 else self.lib.orc_bvh_build
        h = fn(_p(tris), n, tris.size, _p(meshes), meshes.size, search_radius)
        if not h:
            raise ValueError("orc_bvh_build rejected its arguments")
        try:
            nc = 2 * n - 1
            its = self.lib.orc_bvh_nb_iterations(h)

            def arr(name, dtype, count):
                ptr = getattr(self.lib, "orc_bvh_" + name)(h)
                if count == 0:
                    return np.zeros(0, dtype=dtype)
                buf = (C.c_char * (np.dtype(dtype).itemsize * count)).from_address(ptr)
                return np.frombuffer(buf, dtype=dtype, count=count).copy()

            return OracleBvh(n, arr("morton_sorted", np.uint32, n), arr("triangle_indices", np.uint32, n),
                             arr("clusters", NODE, nc), arr("parent", np.uint32, nc),
                             arr("left", np.uint32, nc), arr("right", np.uint32, nc),
                             arr("trace_active", np.uint32, its), arr("trace_merges", np.uint32, its))
        finally:
            self.lib.orc_bvh_destroy(h)

    def flatten(self, clusters, left, right):
        n = (clusters.size + 1) // 2
        clusters = np.ascontiguousarray(clusters)
        left = np.ascontiguousarray(left, dtype=np.uint32)
        right = np.ascontiguousarray(right, dtype=np.uint32)
        flat = np.zeros(2 * n - 1, dtype=NODE)
        self.lib.orc_flatten(_p(clusters), _p(left), _p(right), n, _p(flat))
        return flat

    def hash_flat_nodes(self, flat):
        flat = np.ascontiguousarray(flat)
        return int(self.lib.orc_hash_flat_nodes(_p(flat), flat.size))

    def hash_words(self, words):
        words = np.ascontiguousarray(words, dtype=np.uint32)
        return int(self.lib.orc_hash_words(_p(words), words.size))

    # ---- traversal ----
    def get_rays(self, cam, width, height, denom_w, denom_h):
        rays = np.zeros(width * height, dtype=RAY)
        self.lib.orc_get_rays(_p(cam), width, height, denom_w, denom_h, _p(rays))
        return rays

    def trace_primary(self, flat, tris, meshes, cam, width, height, denom_w=None, denom_h=None, threads=0):
        denom_w = width if denom_w is None else denom_w
        denom_h = height if denom_h is None else denom_h
        hits = np.zeros(width * height, dtype=HIT)
        self.lib.orc_trace_primary(_p(flat), _p(tris), _p(meshes), _p(cam), width, height,
                                   denom_w, denom_h, _p(hits), threads)
        return hits

    def trace_rays(self, flat, tris, meshes, rays, threads=0):
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(rays.size, dtype=HIT)
        self.lib.orc_trace_rays(_p(flat), _p(tris), _p(meshes), _p(rays), rays.size, _p(hits), threads)
        return hits

    def ray_triangle(self, tris, meshes, rays, tri_index):
        rays = np.ascontiguousarray(rays)
        out = np.zeros(rays.size, dtype=HIT)
        for i in range(rays.size):
            self.lib.orc_ray_triangle(_p(rays[i:i + 1]), _p(tris), _p(meshes), int(tri_index[i]), _p(out[i:i + 1]))
        return out

    def intersect_box_edge(self, flat, rays, node_index, display_depth):
        rays = np.ascontiguousarray(rays)
        flat = np.ascontiguousarray(flat)
        out = np.zeros(rays.size, dtype=np.uint32)
        for i in range(rays.size):
            out[i] = self.lib.orc_intersect_box_edge(_p(rays[i:i + 1]), _p(flat[int(node_index[i]):int(node_index[i]) + 1]),
                                                     display_depth)
        return out

    def closest_hit_brute(self, tris, meshes, rays):
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(rays.size, dtype=HIT)
        for i in range(rays.size):
            self.lib.orc_closest_hit_brute(_p(rays[i:i + 1]), _p(tris), tris.size, _p(meshes), _p(hits[i:i + 1]))
        return hits

    def any_hit(self, flat, tris, meshes, rays, t_max):
        rays = np.ascontiguousarray(rays)
        t_max = np.ascontiguousarray(t_max, dtype=np.float32)
        out = np.zeros(rays.size, dtype=np.uint32)
        for i in range(rays.size):
            out[i] = self.lib.orc_any_hit_bvh(_p(rays[i:i + 1]), C.c_float(float(t_max[i])), _p(flat), _p(tris), _p(meshes))
        return out

    def render(self, flat, tris, meshes, cam, width, height, denom_w=None, denom_h=None, row0=0, row1=None,
               bounces=0, shadow=False, light=(0.0, 0.0, 0.0), threads=0):
        denom_w = width if denom_w is None else denom_w
        denom_h = height if denom_h is None else denom_h
        row1 = height if row1 is None else row1
        rows = row1 - row0
        rgba = np.zeros((rows, width, 4), dtype=np.float32)
        hits = np.zeros(rows * width, dtype=HIT)
        nrays = np.zeros(1, dtype=np.uint64)
        light = np.asarray(light, dtype=np.float32)
        self.lib.orc_render(_p(flat), _p(tris), _p(meshes), _p(cam), width, height, denom_w, denom_h,
                            row0, row1, bounces, 1 if shadow else 0, _p(light), _p(rgba), _p(hits), _p(nrays), threads)
        return rgba, hits, int(nrays[0])

    def shade(self, hits, tris, meshes, materials, wireframe=False, bvh_rgba=None):
        """getColor of raytracer.glsl: rgba [n, 4]; bvh_rgba = depth_overlay(...) when the BVH is displayed."""
        hits = np.ascontiguousarray(hits)
        materials = np.ascontiguousarray(materials, dtype=np.float32)
        out = np.zeros((hits.size, 4), dtype=np.float32)
        b = None if bvh_rgba is None else np.ascontiguousarray(bvh_rgba, dtype=np.float32)
        self.lib.orc_shade(_p(hits), hits.size, _p(tris), _p(meshes), _p(materials), 1 if wireframe else 0,
                           None if b is None else _p(b), _p(out))
        return out

    def depth_overlay(self, flat, cam, width, height, depth, denom_w=None, denom_h=None):
        """bvhColor of the shader's traversal per pixel, [height*width, 4]."""
        out = np.zeros((width * height, 4), dtype=np.float32)
        self.lib.orc_depth_overlay(_p(np.ascontiguousarray(flat)), _p(cam), width, height,
                                   width if denom_w is None else denom_w, height if denom_h is None else denom_h,
                                   depth, _p(out))
        return out

    def num_threads(self):
        return int(self.lib.orc_num_threads())


class ReferenceBvh:
    def __init__(self, n, morton_unsorted, triangle_indices, clusters, parent, left, right, is_leaf, build_ms):
        self.n = n
        self.morton_unsorted = morton_unsorted
        self.triangle_indices = triangle_indices
        self.clusters = clusters
        self.parent = parent
        self.left = left
        self.right = right
        self.is_leaf = is_leaf
        self.build_ms = build_ms


class Reference:
    """The reference's own cr::BVH (bvh.cpp) behind oracle/ref_driver.cpp."""

    def __init__(self, cap: int = 65536):
        path = os.path.join(REF_DIR, "libref_bvh_%d.so" % cap)
        if not os.path.exists(path):
            if not build_reference((cap,)):
                raise FileNotFoundError(path)
        L = C.CDLL(path)
        self.lib = L
        vp, u32 = C.c_void_p, C.c_uint32
        L.ref_max_triangles.restype = C.c_uint64
        L.ref_bvh_build.argtypes = [vp, u32, u32, vp, u32]
        L.ref_bvh_build.restype = vp
        L.ref_bvh_build_ms.argtypes = [vp]
        L.ref_bvh_build_ms.restype = C.c_double
        L.ref_bvh_destroy.argtypes = [vp]
        L.ref_bvh_morton_codes.argtypes = [vp, vp]
        L.ref_bvh_triangle_indices.argtypes = [vp, vp]
        L.ref_bvh_clusters.argtypes = [vp, vp, vp, vp, vp, vp]
        L.ref_bvh_clusters.restype = u32
        L.ref_set_threads.argtypes = [C.c_int]
        L.ref_max_threads.restype = C.c_int
        L.ref_sort_pairs_ms.argtypes = [vp, u32, vp, vp]
        L.ref_sort_pairs_ms.restype = C.c_double
        self.cap = int(L.ref_max_triangles())
        assert self.cap == cap

    def set_threads(self, n: int):
        """OpenMP threads of the next builds (cluster ids are only reproducible with 1, Q3)."""
        self.lib.ref_set_threads(int(n))

    def sort_pairs_ms(self, codes):
        """The reference's std::sort of (code, index) pairs (bvh.cpp:223-231): (sorted codes, indices, milliseconds)."""
        codes = np.ascontiguousarray(codes, dtype=np.uint32)
        k = np.zeros(codes.size, np.uint32)
        v = np.zeros(codes.size, np.uint32)
        ms = float(self.lib.ref_sort_pairs_ms(_p(codes), codes.size, _p(k), _p(v)))
        return k, v, ms

    def bvh_build(self, tris, meshes, n=None, want_morton=True, timing_only=False) -> ReferenceBvh:
        n = tris.size if n is None else n
        tris = np.ascontiguousarray(tris)
        meshes = np.ascontiguousarray(meshes)
        h = self.lib.ref_bvh_build(_p(tris), n, tris.size, _p(meshes), meshes.size)
        if not h:
            raise ValueError("ref_bvh_build rejected its arguments (n=%d, cap=%d)" % (n, self.cap))
        try:
            ms = float(self.lib.ref_bvh_build_ms(h))
            if timing_only:
                return ReferenceBvh(n, None, None, None, None, None, None, None, ms)
            nc = 2 * n - 1
            codes = np.zeros(n, dtype=np.uint32)
            if want_morton:
                self.lib.ref_bvh_morton_codes(h, _p(codes))
            idx = np.zeros(n, dtype=np.uint32)
            self.lib.ref_bvh_triangle_indices(h, _p(idx))
            clusters = np.zeros(nc, dtype=NODE)
            parent = np.zeros(nc, dtype=np.uint32)
            left = np.zeros(nc, dtype=np.uint32)
            right = np.zeros(nc, dtype=np.uint32)
            is_leaf = np.zeros(nc, dtype=np.uint8)
            present = self.lib.ref_bvh_clusters(h, _p(clusters), _p(parent), _p(left), _p(right), _p(is_leaf))
            assert present == nc, (present, nc)
            return ReferenceBvh(n, codes, idx, clusters, parent, left, right, is_leaf, ms)
        finally:
            self.lib.ref_bvh_destroy(h)


def raytracer_available(variant: str = "glm") -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libref_raytracer_%s.so" % variant))


class ReferenceRaytracer:
    """The reference's own raytracer.glsl, compiled as C++ through its vendored GLM (oracle/ref_raytracer.cpp).

    variant: "glm"  -- glm::normalize, undefined Hit w := +inf (guarded misses, Q6)
             "div"  -- normalize := v / sqrt(dot(v, v)) (what oracle + kernels pin): bit-equal to rtr_oracle.c
             "zero" -- glm::normalize, undefined := 0 (the letter of raytracer.glsl:279 on zeroed registers)
    One scene / camera is bound per instance (the shader's SSBOs and uniforms are globals of the library)."""

    def __init__(self, variant: str = "glm"):
        path = os.path.join(REF_DIR, "libref_raytracer_%s.so" % variant)
        if not os.path.exists(path):
            if not os.path.isdir(REFERENCE_ROOT):
                raise FileNotFoundError(path)
            subprocess.check_call(["make", "-s", "-C", HERE, path])
        L = C.CDLL(path)
        self.lib = L
        self.variant = variant
        vp, u32, u64, i32, f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_float
        L.ref_rt_variant.restype = C.c_char_p
        L.ref_rt_undefined_w.restype = f32
        L.ref_rt_set_threads.argtypes = [i32]
        L.ref_rt_set_threads(os.cpu_count() or 1)   # the tests pin OMP_NUM_THREADS=1 for the reference's bvh.cpp (Q3)
        L.ref_rt_set_scene.argtypes = [vp, u32, vp, u32, u32, vp, u32, vp, u32]
        L.ref_rt_set_scene.restype = i32
        L.ref_rt_set_camera.argtypes = [vp]
        L.ref_rt_set_flags.argtypes = [i32, i32, i32]
        L.ref_rt_get_ray.argtypes = [f32, f32, vp]
        L.ref_rt_trace_primary.argtypes = [u32, u32, u32, u32, vp, vp]
        L.ref_rt_ray_triangle.argtypes = [vp, vp, u64, vp]
        L.ref_rt_intersect_bvh.argtypes = [vp, vp, u64, vp]
        L.ref_rt_closest_hit_bvh.argtypes = [vp, u64, vp, vp]
        L.ref_rt_all_hits.argtypes = [vp, u64, vp]
        L.ref_rt_get_color.argtypes = [vp, vp, u64, vp]
        L.ref_rt_dispatch.argtypes = [u32, u32, u32, u32, vp]
        self._keep = None

    def set_scene(self, tris, meshes, flat, materials=None, model_stride=68):
        tris = np.ascontiguousarray(tris)
        meshes = np.ascontiguousarray(meshes)
        flat = np.ascontiguousarray(flat)
        if materials is None:
            materials = np.ones((1, 4), dtype=np.float32)
        materials = np.ascontiguousarray(materials, dtype=np.float32).reshape(-1, 4)
        assert tris.dtype.itemsize == 64 and meshes.dtype.itemsize == 68 and flat.dtype.itemsize == 48
        rc = self.lib.ref_rt_set_scene(_p(tris), tris.size, _p(meshes), meshes.size, model_stride,
                                       _p(materials), materials.shape[0], _p(flat), flat.size)
        assert rc == 0
        self.nb_tris = tris.size

    def set_camera(self, cam):
        cam = np.ascontiguousarray(cam)
        assert cam.dtype.itemsize == 284
        self.lib.ref_rt_set_camera(_p(cam))

    def set_flags(self, depth_display_bvh=-1, is_bvh_displayed=False, wireframe=False):
        self.lib.ref_rt_set_flags(int(depth_display_bvh), int(bool(is_bvh_displayed)), int(bool(wireframe)))

    def get_ray(self, pos_x, pos_y):
        out = np.zeros(1, dtype=RAY)
        self.lib.ref_rt_get_ray(C.c_float(pos_x), C.c_float(pos_y), _p(out))
        return out[0]

    def primary_rays(self, width, height, denom_w=None, denom_h=None):
        """The rays main() builds: pos = (x / denom_w, y / denom_h), float division as at raytracer.glsl:304-305."""
        denom_w = width if denom_w is None else denom_w
        denom_h = height if denom_h is None else denom_h
        rays = np.zeros(width * height, dtype=RAY)
        one = np.zeros(1, dtype=RAY)
        for y in range(height):
            py = np.float32(y) / np.float32(denom_h)
            for x in range(width):
                px = np.float32(x) / np.float32(denom_w)
                self.lib.ref_rt_get_ray(C.c_float(px), C.c_float(py), _p(one))
                rays[y * width + x] = one[0]
        return rays

    def trace_primary(self, width, height, denom_w=None, denom_h=None, want_rays=False):
        """getClosestHitBVH of every pixel's getRay (whole frame, OpenMP over rows)."""
        denom_w = (width // 16) * 16 if denom_w is None else denom_w
        denom_h = (height // 16) * 16 if denom_h is None else denom_h
        hits = np.zeros(width * height, dtype=HIT)
        rays = np.zeros(width * height, dtype=RAY) if want_rays else None
        self.lib.ref_rt_trace_primary(width, height, denom_w, denom_h, _p(hits), None if rays is None else _p(rays))
        return (hits, rays) if want_rays else hits

    def ray_triangle(self, rays, tri_index):
        rays = np.ascontiguousarray(rays)
        tri_index = np.ascontiguousarray(tri_index, dtype=np.uint32)
        out = np.zeros(rays.size, dtype=HIT)
        self.lib.ref_rt_ray_triangle(_p(rays), _p(tri_index), rays.size, _p(out))
        return out

    def intersect_bvh(self, rays, node_index):
        rays = np.ascontiguousarray(rays)
        node_index = np.ascontiguousarray(node_index, dtype=np.uint32)
        out = np.zeros(rays.size, dtype=np.uint32)
        self.lib.ref_rt_intersect_bvh(_p(rays), _p(node_index), rays.size, _p(out))
        return out

    def closest_hit_bvh(self, rays, want_bvh_color=False):
        rays = np.ascontiguousarray(rays)
        out = np.zeros(rays.size, dtype=HIT)
        col = np.zeros((rays.size, 4), dtype=np.float32) if want_bvh_color else None
        self.lib.ref_rt_closest_hit_bvh(_p(rays), rays.size, _p(out), None if col is None else _p(col))
        return (out, col) if want_bvh_color else out

    def all_hits(self, rays):
        rays = np.ascontiguousarray(rays)
        out = np.zeros(rays.size, dtype=HIT)
        self.lib.ref_rt_all_hits(_p(rays), rays.size, _p(out))
        return out

    def get_color(self, hits, bvh_color=None):
        hits = np.ascontiguousarray(hits)
        out = np.zeros((hits.size, 4), dtype=np.float32)
        b = None if bvh_color is None else np.ascontiguousarray(bvh_color, dtype=np.float32)
        self.lib.ref_rt_get_color(_p(hits), None if b is None else _p(b), hits.size, _p(out))
        return out

    def dispatch(self, width, height, groups_x=None, groups_y=None, fill=np.nan):
        """glDispatchCompute(floor(W/16), floor(H/16), 1) (application.cpp:225-245) -> rgba [H, W, 4];
        pixels no invocation stores keep `fill`."""
        groups_x = width // 16 if groups_x is None else groups_x
        groups_y = height // 16 if groups_y is None else groups_y
        img = np.full((height, width, 4), fill, dtype=np.float32)
        self.lib.ref_rt_dispatch(groups_x, groups_y, width, height, _p(img))
        return img
