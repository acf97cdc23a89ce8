"""CPU tests of the oracle itself: the reference's own known-answer vectors (tests/testsSortGPU),
the survey's known-answer hashes of the reference build, internal consistency."""
import numpy as np
import pytest

from realtimeraytracing_b200 import synth
from realtimeraytracing_b200.layouts import hash_words, node_words

import scenes


# ---- tests/testsSortGPU/testHistogramCreation.cpp ----
def test_hist_powers_of_two(oracle):  # :43-53, :145-153
    keys = np.array([1 << (i % 32) for i in range(130)], dtype=np.uint32)
    expected = np.zeros(32, dtype=np.uint32)
    for i in range(130):
        expected[i % 32] += 1
    assert expected[0] == 5 and expected[1] == 5 and expected[2] == 4
    assert np.array_equal(oracle.bit_histogram32(keys), expected)


def test_hist_zeros_ones_twos(oracle):  # :55-73, :155-181
    assert np.array_equal(oracle.bit_histogram32(np.zeros(130, np.uint32)), np.zeros(32, np.uint32))
    e = np.zeros(32, np.uint32); e[0] = 130
    assert np.array_equal(oracle.bit_histogram32(np.full(130, 1, np.uint32)), e)
    e = np.zeros(32, np.uint32); e[1] = 130
    assert np.array_equal(oracle.bit_histogram32(np.full(130, 2, np.uint32)), e)


def test_hist_random_matches_numpy(oracle):  # :17-41 host recurrence, also the Hard variant's size
    for n in (130, 65536):
        keys = synth.random_keys_u32(n, seed=n, lo=0, hi=8192)
        expected = np.array([(keys >> b & 1).sum() for b in range(32)], dtype=np.uint32)
        assert np.array_equal(oracle.bit_histogram32(keys), expected)


# ---- tests/testsSortGPU/testHistogramPrefixSum.cpp ----
KNOWN_IN = [37, 41, 49, 53, 37, 48, 44, 51, 35, 51, 53, 41] + [0] * 20   # :53-68
KNOWN_OUT = [0, 37, 78, 127, 0, 37, 85, 129, 0, 35, 86, 139] + [0] * 20  # :70-86


def test_prefix_known_values(oracle):
    assert np.array_equal(oracle.digitplace_exclusive_scan(np.array(KNOWN_IN, np.uint32)), np.array(KNOWN_OUT, np.uint32))


def test_prefix_random_recurrence(oracle):  # :43-51
    keys = synth.random_keys_u32(32, seed=9, lo=0, hi=8192)
    hist = oracle.bit_histogram32(keys)
    out = np.zeros(32, np.uint32)
    for j in range(8):
        for i in range(1, 4):
            out[4 * j + i] = hist[4 * j + i - 1] + out[4 * j + i - 1]
    assert np.array_equal(oracle.digitplace_exclusive_scan(hist), out)


# ---- sort ----
@pytest.mark.parametrize("n", [0, 1, 2, 130, 4099, 100000])
def test_radix_equals_comparison_sort(oracle, n):
    keys = synth.random_keys_u32(n, seed=1) if n else np.zeros(0, np.uint32)
    keys = keys & np.uint32(0x3FF) if n > 1000 else keys  # many duplicates: stability matters
    idx = np.arange(n, dtype=np.uint32)
    k1, v1 = oracle.sort_pairs(keys, idx)
    k2, v2 = oracle.radix_sort_pairs(keys, idx)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k1, keys[order]) and np.array_equal(v1, idx[order])
    assert np.array_equal(k2, k1) and np.array_equal(v2, v1)


def test_radix_u64(oracle):
    rng = np.random.RandomState(2)
    keys = (rng.randint(0, 1 << 31, size=5000).astype(np.uint64) << np.uint64(33)) | rng.randint(0, 1 << 31, size=5000).astype(np.uint64)
    keys[::7] = keys[0]
    vals = np.arange(5000, dtype=np.uint32)
    k, v = oracle.radix_sort_pairs_u64(keys, vals)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order]) and np.array_equal(v, vals[order])
    assert np.array_equal(oracle.radix_sort_keys_u64(keys), keys[order])


# ---- known answers recorded from the reference's own bvh.cpp ----
def test_survey_known_answer_scene(oracle):
    tris, meshes = synth.survey_known_answer_scene()
    aabb = oracle.scene_aabb(tris, meshes)
    assert np.allclose(aabb, [-0.349036, -2.04506, -1.04801, 0.348272, 2.047, 1.04835], rtol=0, atol=1e-5)  # survey prints 6 significant digits
    cube = oracle.circumscribed_cube(aabb)
    side = cube[3:] - cube[:3]
    assert np.allclose(side, 0.697308, atol=2e-6)  # Q1: x extent on every axis
    b = oracle.bvh_build(tris, meshes)
    flat = oracle.flatten(b.clusters, b.left, b.right)
    assert flat.size == 39999 and flat[0]["left"] == 1 and flat[0]["right"] == 20100
    assert "%016x" % oracle.hash_flat_nodes(flat) == "6c2aabd40444c4ec"
    assert "%016x" % oracle.hash_words(b.triangle_indices) == "cc1bb62a0ae92f47"
    assert hash_words(node_words(flat)) == oracle.hash_flat_nodes(flat)  # python and C hashes agree


def test_tree_invariants(oracle):
    tris, meshes, _ = scenes.soup(3000)
    b = oracle.bvh_build(tris, meshes)
    n = b.n
    assert b.trace_merges.sum() == n - 1
    assert b.trace_active[0] == n
    assert np.all(b.parent[: 2 * n - 2] != 0xFFFFFFFF) and b.parent[2 * n - 2] == 0xFFFFFFFF
    flat = oracle.flatten(b.clusters, b.left, b.right)
    leaves = (flat["left"] == 0) & (flat["right"] == 0)
    assert leaves.sum() == n
    assert np.array_equal(np.sort(flat["tri"][leaves]), np.arange(n))
    inner = ~leaves
    idx = np.nonzero(inner)[0]
    assert np.all(flat["left"][inner] == idx + 1)
    # parent box contains both children
    for side in ("left", "right"):
        ch = flat[side][inner]
        assert np.all(flat["bmin"][inner] <= flat["bmin"][ch]) and np.all(flat["bmax"][inner] >= flat["bmax"][ch])


def test_single_and_two_triangles(oracle):
    tris, meshes, _ = scenes.soup(2, seed=4)
    b1 = oracle.bvh_build(tris[:1], meshes)
    f1 = oracle.flatten(b1.clusters, b1.left, b1.right)
    assert f1.size == 1 and f1[0]["left"] == 0 and f1[0]["right"] == 0 and f1[0]["tri"] == 0
    b2 = oracle.bvh_build(tris, meshes)
    f2 = oracle.flatten(b2.clusters, b2.left, b2.right)
    assert f2.size == 3 and f2[0]["left"] == 1 and f2[0]["right"] == 2


# ---- traversal: BVH walk == brute force getAllHits semantics (raytracer.glsl:149-157) ----
def test_bvh_traversal_equals_brute_force(oracle):
    tris, meshes, L = scenes.soup(2000)
    b = oracle.bvh_build(tris, meshes)
    flat = oracle.flatten(b.clusters, b.left, b.right)
    cam = synth.soup_camera(L, 64, 48)
    hits = oracle.trace_primary(flat, tris, meshes, cam, 64, 48)
    rays = oracle.get_rays(cam, 64, 48, 64, 48)
    brute = oracle.closest_hit_brute(tris, meshes, rays)
    assert hits["did_hit"].sum() > 200
    assert np.array_equal(hits["did_hit"], brute["did_hit"])
    h = hits["did_hit"] == 1
    assert np.array_equal(hits["t"][h], brute["t"][h])
    # ids may differ only on exact-t ties (first leaf in right-first DFS vs lowest index)
    diff = hits["tri"][h] != brute["tri"][h]
    assert diff.sum() == 0


def test_mesh_traversal_ties_and_render(oracle):
    tris, meshes = synth.grid_mesh(24, 24)
    b = oracle.bvh_build(tris, meshes)
    flat = oracle.flatten(b.clusters, b.left, b.right)
    cam = synth.reference_camera(aspect=64 / 64)
    hits = oracle.trace_primary(flat, tris, meshes, cam, 64, 64)
    assert hits["did_hit"].mean() > 0.5
    rgba, first, nrays = oracle.render(flat, tris, meshes, cam, 64, 64, bounces=2, shadow=True, light=(0.0, 3.0, -4.0))
    assert np.array_equal(first["tri"], hits["tri"]) and np.array_equal(first["t"], hits["t"])
    assert nrays >= 64 * 64 + hits["did_hit"].sum()
    assert np.all(rgba[..., 3] == 1.0) and rgba[..., 0].max() > 0.1
    barys = np.stack([hits["b0"], hits["b1"], hits["b2"]], 1)[hits["did_hit"] == 1]
    assert np.all(barys >= 0) and np.allclose(barys.sum(1), 1.0, atol=1e-5)


def test_shade_follows_getcolor(oracle):
    """raytracer.glsl:159-179 by hand: a miss is (0,0,0,1), a hit adds the material colour of the triangle's model,
    wireframe blackens hits whose barycentric is below 0.02."""
    from realtimeraytracing_b200.layouts import HIT, MESH, TRIANGLE
    tris = np.zeros(3, TRIANGLE); tris["model_id"] = [0, 1, 1]
    meshes = np.zeros(2, MESH); meshes["material_id"] = [1, 0]
    materials = np.array([[0.25, 0.5, 0.75, 1.0], [1.0, 0.0, 0.5, 0.125]], np.float32)
    hits = np.zeros(4, HIT)
    hits[1] = (0.3, 0.3, 0.4, 2.0, 1, 0)     # triangle 0 -> model 0 -> material 1
    hits[2] = (0.01, 0.49, 0.5, 1.0, 1, 2)   # near an edge; triangle 2 -> model 1 -> material 0
    hits[3] = (0.5, 0.49, 0.019, 1.0, 1, 1)
    plain = oracle.shade(hits, tris, meshes, materials)
    assert np.array_equal(plain, np.array([[0, 0, 0, 1], [1, 0, 0.5, 1.125], [0.25, 0.5, 0.75, 2], [0.25, 0.5, 0.75, 2]], np.float32))
    wire = oracle.shade(hits, tris, meshes, materials, wireframe=True)
    assert np.array_equal(wire, np.array([[0, 0, 0, 1], [1, 0, 0.5, 1.125], [0, 0, 0, 1], [0, 0, 0, 1]], np.float32))


def test_oracle_build64_refines_the_32_bit_order(oracle):
    """orc_bvh_build64 is the same builder over 63-bit keys: a valid tree whose leaf order, truncated to 30 bits,
    is the reference's sorted code sequence."""
    import scenes
    tris, meshes, _ = scenes.soup(3000, seed=21)
    b64 = oracle.bvh_build(tris, meshes, key_bits=64)
    b32 = oracle.bvh_build(tris, meshes)
    assert np.array_equal(b64.morton_sorted, b32.morton_sorted)          # code64 >> 33 == code32, both sorted
    assert sorted(b64.triangle_indices.tolist()) == list(range(tris.size))
    c64 = oracle.morton_codes64(tris, meshes)
    assert np.array_equal(c64[b64.triangle_indices], np.sort(c64, kind="stable"))
    assert int(b64.left[-1]) != 0xFFFFFFFF and b64.trace_merges.sum() == tris.size - 1


def test_depth_overlay_oracle_on_a_two_leaf_tree(oracle):
    """Hand-checkable overlay: depth 0 colours every pixel whose ray meets the root box, depth 1 the children; a ray
    entering near an edge of the box gets the line colour (raytracer.glsl:222-233, :269-275)."""
    import scenes
    from realtimeraytracing_b200 import synth
    tris, meshes, L = scenes.soup(2, seed=1)
    b = oracle.bvh_build(tris, meshes)
    flat = oracle.flatten(b.clusters, b.left, b.right)
    W, H = 64, 48
    cam = synth.soup_camera(L, W, H)
    d0 = oracle.depth_overlay(flat, cam, W, H, 0)
    d1 = oracle.depth_overlay(flat, cam, W, H, 1)
    d2 = oracle.depth_overlay(flat, cam, W, H, 2)
    box, line = (0.5, 0.0, 0.5, 0.1), (0.7, 0.0, 0.7, 0.1)
    def kinds(a):
        return {tuple(np.float32(c).tolist()) for c in np.unique(a, axis=0)}
    f32 = lambda t: tuple(np.float32(x).item() for x in t)
    assert kinds(d0) <= {f32(box), f32(line), (0.0, 0.0, 0.0, 0.0)} and f32(box) in kinds(d0)
    hit0 = (d0[:, 3] > 0); hit1 = (d1[:, 3] > 0)
    assert hit1.sum() > 0 and not (hit1 & ~hit0).any()       # children are inside the root box
    assert not (d2[:, 3] > 0).any()                          # no node at depth 2 in a two-leaf tree
    hits = oracle.trace_primary(flat, tris, meshes, cam, W, H, W, H)
    img = oracle.shade(hits, tris, meshes, np.ones((1, 4), np.float32), bvh_rgba=d0)
    miss = hits["did_hit"] == 0
    assert np.array_equal(img[miss], d0[miss])               # a miss shows the overlay colour alone
