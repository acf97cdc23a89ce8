"""Pins the traversal half of the oracle (oracle/rtr_oracle.c) to the reference's OWN shader.

raytracer.glsl is compiled as C++ through the reference's vendored GLM by oracle/Makefile (ref_raytracer.cpp +
glsl2cpp.sed: a syntactic mapping, no shader text in the repo) into oracle/_ref/libref_raytracer_<variant>.so;
tests/golden/raytracer_golden.npz holds outputs of that library (gen_raytracer_golden.py).  Every test here runs
against the committed fixture; where the library itself is present (this container, and the GPU box: oracle/_ref
travels) the same comparisons run live as well.

What is bit-exact and what is not:
 * variant "div" (normalize := v / sqrt(dot(v, v)), the formula oracle + kernels pin): every function of the shader
   equals the restatement BIT FOR BIT, except rays with a direction component of exactly 0 whose origin lies on a
   slab plane (0 * inf = NaN): GLSL leaves min/max of a NaN to the implementation, GLM's `y < x ? y : x` and the
   IEEE fminf/fmaxf of oracle + kernels then differ (documented deviation).
 * variant "glm" (glm::normalize = v * (1 / sqrt(dot))): directions differ by <= 1 ulp, hit ids are equal and
   t / barycentrics agree within the 1e-5 relative tolerance BASELINE.json states.
 * variant "zero" is the letter of raytracer.glsl:279 with the undefined Hit._Coords.w read as 0: a miss replaces the
   closest hit (Q6).  The oracle follows the guarded semantics of the shader's own getAllHits (:152-153) instead;
   the test records how many pixels that changes.
"""
import os

import numpy as np
import pytest

import raytracer_cases as rc
from oracle import ReferenceRaytracer, raytracer_available
from realtimeraytracing_b200.layouts import HIT, RAY

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "raytracer_golden.npz")
OVERLAY_DEPTH = 3
BATCH = 4000
REL_TOL = 1e-5  # BASELINE.json north_star: "hit distances/barycentrics within 1e-5 relative"


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def built(oracle):
    cache = {}

    def get(name):
        if name not in cache:
            tris, meshes, materials, cam, w, h = rc.case(name)
            b = oracle.bvh_build(tris, meshes)
            cache[name] = (tris, meshes, materials, cam, w, h, oracle.flatten(b.clusters, b.left, b.right))
        return cache[name]
    return get


def words(a):
    return np.ascontiguousarray(a).view(np.uint32).reshape(a.size, -1)


def oracle_rays(oracle, cam, w, h):
    return oracle.get_rays(cam, w, h, (w // 16) * 16, (h // 16) * 16)


def traced_pixels(w, h):
    """Q5: only floor(W/16)*16 x floor(H/16)*16 pixels have an invocation."""
    m = np.zeros((h, w), bool)
    m[:(h // 16) * 16, :(w // 16) * 16] = True
    return m.reshape(-1)


def assert_hits_close(a, b, tris, meshes, what):
    """Two hit lists of (up to 1 ulp) different rays.  t within REL_TOL relative; the hit POINT named by the
    barycentrics within REL_TOL * t (a 1-ulp change of the direction moves the point by ~1e-7 * t, which is many ulps
    of a barycentric when the triangle is small against the distance).  The triangle id is equal except on ties
    (BASELINE.json: "exact except for documented ties"): the two records may name different triangles only if they
    name the same point at the same distance, i.e. the ray passes through a shared edge; the hit / miss decision may
    differ only where the hit lies on an edge of its triangle (a silhouette)."""
    both = (a["did_hit"] == 1) & (b["did_hit"] == 1)
    one = a["did_hit"] != b["did_hit"]
    for h in (a[one & (a["did_hit"] == 1)], b[one & (b["did_hit"] == 1)]):
        assert np.all(np.minimum(np.minimum(h["b0"], h["b1"]), h["b2"]) < 1e-4), what
    assert one.mean() < 2e-3, what
    x, y = a[both], b[both]
    assert np.all(np.abs(x["t"] - y["t"]) <= REL_TOL * np.abs(y["t"])), what
    px, py = rc.hit_points(tris, meshes, x), rc.hit_points(tris, meshes, y)
    assert np.all(np.linalg.norm(px - py, axis=1) <= REL_TOL * np.abs(y["t"].astype(np.float64))), what
    assert (x["tri"] != y["tri"]).mean() < 5e-3, what


# ------------------------------------------------------------------ against the committed shader outputs
@pytest.mark.parametrize("name", rc.CASE_NAMES)
def test_get_ray_and_closest_hit_equal_the_shader(oracle, golden, built, name):
    tris, meshes, materials, cam, w, h, flat = built(name)
    px = traced_pixels(w, h)
    rays = oracle_rays(oracle, cam, w, h)[px]
    assert np.array_equal(rays["d"].view(np.uint32), golden[name + "/rays_div"][px])      # getRay + :303-305, bit-exact
    ulp = np.abs(rays["d"].view(np.int32).astype(np.int64) - golden[name + "/rays_glm"][px].view(np.int32).astype(np.int64))
    assert ulp.max() <= 1                                                                    # glm::normalize: 1 ulp
    hits = oracle.trace_rays(flat, tris, meshes, rays)
    ref = golden[name + "/hits_div"].view(HIT).reshape(-1)[px]
    same = (words(hits) == words(ref)).all(axis=1)
    q11 = rc.has_zero_component(rays)
    assert same[~q11].all()                                                                  # bit-exact
    if name == "grid":
        assert q11.sum() > 0 and (~same[q11]).sum() > 0     # the documented deviation exists and is confined to Q11 rays
    else:
        assert same.all()
    glm = golden[name + "/hits_glm"].view(HIT).reshape(-1)[px]
    assert_hits_close(hits[~q11], glm[~q11], tris, meshes, name)


@pytest.mark.parametrize("name", rc.CASE_NAMES)
def test_all_hits_brute_force_equals_the_shader(oracle, golden, built, name):
    """getAllHits (:149-157) is the shader's own guarded closest hit: no undefined read, lowest index wins ties."""
    tris, meshes, materials, cam, w, h, flat = built(name)
    px = traced_pixels(w, h)[::7]
    rays = oracle_rays(oracle, cam, w, h)[::7][px]
    brute = oracle.closest_hit_brute(tris, meshes, rays)
    ref = golden[name + "/allhits_div"].view(HIT).reshape(-1)[px]
    assert np.array_equal(words(brute), words(ref))
    # and the BVH walk returns the same hit except on exact-t ties (first leaf of the right-first DFS vs lowest index)
    walk = oracle.trace_rays(flat, tris, meshes, rays)
    q11 = rc.has_zero_component(rays)
    d = ~(words(walk) == words(ref)).all(axis=1) & ~q11
    assert np.array_equal(walk["did_hit"][~q11], ref["did_hit"][~q11])
    assert np.array_equal(walk["t"][d], ref["t"][d])        # a different id only where t is equal


@pytest.mark.parametrize("name", rc.CASE_NAMES)
def test_ray_batches_equal_the_shader(oracle, golden, built, name):
    tris, meshes, materials, cam, w, h, flat = built(name)
    extent = float(np.abs(np.concatenate([tris["p0"][:, :3], tris["p1"][:, :3], tris["p2"][:, :3]])).max()) * 2
    batch = rc.random_rays(BATCH, extent, seed=len(name))
    assert not rc.has_zero_component(batch).any()
    assert np.array_equal(words(oracle.trace_rays(flat, tris, meshes, batch)), golden[name + "/batch_hits_div"])
    rng = np.random.RandomState(5)
    ti = rng.randint(0, tris.size, size=BATCH).astype(np.uint32)
    ni = rng.randint(0, flat.size, size=BATCH).astype(np.uint32)
    tri_hits = oracle.ray_triangle(tris, meshes, rc.rays_at_triangles(tris, meshes, ti, 6), ti)
    ref = golden[name + "/batch_tri_div"]
    assert np.array_equal(words(tri_hits), ref)                                   # rayTriangleIntersection :102-147
    assert 200 < int((ref[:, 4] != 0).sum()) < BATCH - 200                        # hits and misses both present
    codes = oracle.intersect_box_edge(flat, rc.rays_at_boxes(flat, ni, 7), ni, OVERLAY_DEPTH)
    ref = golden[name + "/batch_box_div"]
    assert np.array_equal(codes, ref)                                             # intersectBVH :182-237 incl. code 2
    assert all(int((ref == c).sum()) > 100 for c in (0, 1, 2))


@pytest.mark.parametrize("name", rc.CASE_NAMES)
def test_frame_equals_the_shader_main(oracle, golden, built, name):
    """main() through glDispatchCompute(floor(W/16), floor(H/16)) (application.cpp:225-245): pixel mapping with the
    Q5 divisors, getClosestHitBVH with the depth overlay (:269-275), getColor with wireframe (:159-179), imageStore."""
    tris, meshes, materials, cam, w, h, flat = built(name)
    dw, dh = (w // 16) * 16, (h // 16) * 16
    ref = golden[name + "/image_div"]
    assert ref.shape == (h, w, 4)
    hits = oracle.trace_primary(flat, tris, meshes, cam, w, h, dw, dh)
    overlay = oracle.depth_overlay(flat, cam, w, h, OVERLAY_DEPTH, dw, dh)
    img = oracle.shade(hits, tris, meshes, materials, wireframe=True, bvh_rgba=overlay).reshape(h, w, 4)
    traced = np.zeros((h, w), bool)
    traced[:dh, :dw] = True
    assert np.isnan(ref[~traced]).all()                     # no invocation exists for those pixels (Q5)
    rays = oracle.get_rays(cam, w, h, dw, dh)
    ok = traced & ~rc.has_zero_component(rays).reshape(h, w)
    assert np.array_equal(img[ok].view(np.uint32), ref[ok].view(np.uint32))
    assert (overlay.reshape(h, w, 4)[ok][:, 3] > 0).any()   # the overlay is really exercised


# ------------------------------------------------------------------ live, against the library itself
needs_lib = pytest.mark.skipif(not (raytracer_available("div") and raytracer_available("glm") and raytracer_available("zero")),
                               reason="oracle/_ref/libref_raytracer_*.so not built (needs /root/reference)")


@needs_lib
def test_fixture_is_what_the_library_returns(oracle, golden, built):
    name = "two_mesh"
    tris, meshes, materials, cam, w, h, flat = built(name)
    r = ReferenceRaytracer("div")
    r.set_scene(tris, meshes, flat, materials)
    r.set_camera(cam)
    r.set_flags(-1, False, False)
    rays = r.primary_rays(w, h, (w // 16) * 16, (h // 16) * 16)
    assert np.array_equal(rays["d"].view(np.uint32), golden[name + "/rays_div"])
    assert np.array_equal(words(r.closest_hit_bvh(rays)), golden[name + "/hits_div"])


@needs_lib
def test_larger_scene_live(oracle):
    """20 k-triangle soup, 160x96 primary rays + 20 000 random rays: restatement == compiled shader, bit for bit."""
    import scenes
    from realtimeraytracing_b200 import synth
    tris, meshes, L = scenes.soup(20000)
    b = oracle.bvh_build(tris, meshes)
    flat = oracle.flatten(b.clusters, b.left, b.right)
    w, h = 160, 96
    cam = synth.soup_camera(L, w, h)
    r = ReferenceRaytracer("div")
    r.set_scene(tris, meshes, flat)
    r.set_camera(cam)
    rays = np.concatenate([oracle.get_rays(cam, w, h, w, h), rc.random_rays(20000, L, seed=3)])
    q11 = rc.has_zero_component(rays)
    a, b_ = oracle.trace_rays(flat, tris, meshes, rays), r.closest_hit_bvh(rays)
    assert np.array_equal(words(a)[~q11], words(b_)[~q11])
    assert a["did_hit"].mean() > 0.1


@needs_lib
def test_model_stride_q7(oracle, built):
    """Q7: the host uploads 68-byte MeshModelGPU records, the shader's std430 Model has stride 80.  Model 0 reads the
    same bytes either way (every scene the reference application builds has one mesh); with a second mesh the shader
    fetches garbage -- the ABI consumes the 68-byte host layout (documented deviation)."""
    tris, meshes, materials, cam, w, h, flat = built("two_mesh")
    rays = oracle_rays(oracle, cam, w, h)
    out = {}
    for stride in (68, 80):
        r = ReferenceRaytracer("div")
        r.set_scene(tris, meshes, flat, materials, model_stride=stride)
        out[stride] = r.closest_hit_bvh(rays)
    assert not np.array_equal(words(out[68]), words(out[80]))
    one = tris["model_id"] == 0
    t1, m1 = tris[one].copy(), meshes[:1].copy()
    b = oracle.bvh_build(t1, m1)
    f1 = oracle.flatten(b.clusters, b.left, b.right)
    for stride in (68, 80):
        r = ReferenceRaytracer("div")
        r.set_scene(t1, m1, f1, materials, model_stride=stride)
        out[stride] = r.closest_hit_bvh(rays)
    assert np.array_equal(words(out[68]), words(out[80]))


@needs_lib
def test_unguarded_miss_q6_is_a_real_difference(oracle, built):
    """With the undefined w read as 0 a missed leaf resets the closest hit (raytracer.glsl:279): the frame loses hits.
    The oracle, the kernels and variant "div"/"glm" follow the shader's own guarded getAllHits instead."""
    tris, meshes, materials, cam, w, h, flat = built("soup3000")
    rays = oracle_rays(oracle, cam, w, h)
    z = ReferenceRaytracer("zero")
    z.set_scene(tris, meshes, flat, materials)
    g = ReferenceRaytracer("glm")
    g.set_scene(tris, meshes, flat, materials)
    hz, hg = z.closest_hit_bvh(rays), g.closest_hit_bvh(rays)
    assert hg["did_hit"].sum() > hz["did_hit"].sum() > 0
    assert np.array_equal(words(g.all_hits(rays[::5])), words(z.all_hits(rays[::5])))  # getAllHits has no undefined read
