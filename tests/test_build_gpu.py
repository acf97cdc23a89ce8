"""Morton codes, sorted indices, PLOC topology and the flattened node array: bit-exact against the
reference's outputs (tests/golden/, generated from the reference's own bvh.cpp) and the oracle."""
import numpy as np
import pytest

import golden_util
import scenes
from realtimeraytracing_b200 import capi, scene as rscene, synth
from realtimeraytracing_b200.layouts import node_words

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    return golden_util.load()


@pytest.mark.parametrize("name", golden_util.SCENE_NAMES)
def test_morton_codes_match_reference(ctx, oracle, golden, name):
    tris, meshes, n = golden_util.scene(name)
    codes = ctx.morton_codes(tris, meshes, n=n)
    assert oracle.hash_words(codes) == int(golden[name + "/morton_hash"])
    assert np.array_equal(codes, oracle.morton_codes(tris, meshes, n=n))


def test_scene_bounds_and_q1_cube(ctx, oracle):
    tris, meshes = synth.survey_known_answer_scene()
    b = ctx.scene_bounds(tris, meshes)
    aabb = oracle.scene_aabb(tris, meshes)
    assert np.array_equal(b[:6], aabb)
    assert np.array_equal(b[6:], oracle.circumscribed_cube(aabb))
    assert np.allclose(b[9:] - b[6:9], 0.697308, atol=2e-6)  # Q1: the x extent on every axis


def test_morton64_matches_oracle(ctx, oracle):
    tris, meshes, _ = scenes.soup(20000)
    assert np.array_equal(ctx.morton_codes64(tris, meshes), oracle.morton_codes64(tris, meshes))


@pytest.mark.parametrize("name", golden_util.SCENE_NAMES)
def test_build_matches_reference_golden(ctx, oracle, golden, name):
    tris, meshes, n = golden_util.scene(name)
    bvh = capi.Bvh(ctx).build(tris, meshes, n=n)
    try:
        idx = bvh.triangle_indices()
        assert oracle.hash_words(idx) == int(golden[name + "/tri_idx_hash"]), "sort permutation"
        clusters, parent, left, right, is_leaf = bvh.clusters()
        ch = oracle.hash_words(np.concatenate([node_words(clusters).ravel(), left, right, parent]))
        assert ch == int(golden[name + "/cluster_hash"]), "PLOC topology by cluster id"
        flat = bvh.flat_nodes()
        assert oracle.hash_flat_nodes(flat) == int(golden[name + "/flat_hash"]), "flattened node array"
        assert is_leaf[:n].all() and not is_leaf[n:].any()
        if name == "soup512":
            assert np.array_equal(node_words(flat), golden[name + "/flat_words"])
            assert np.array_equal(left, golden[name + "/left"]) and np.array_equal(right, golden[name + "/right"])
            assert np.array_equal(parent, golden[name + "/parent"])
    finally:
        bvh.close()


@pytest.mark.parametrize("n", [1, 2, 3, 17, 33, 479, 480, 481, 1023, 1024, 1025, 2049, 30011, 200000])
def test_build_matches_oracle_sizes(ctx, oracle, n):
    """Edge sizes around the tile (480), the tail threshold (1024) and the window (16/32)."""
    tris, meshes, _ = scenes.soup(n, seed=1000 + n)
    ob = oracle.bvh_build(tris, meshes)
    bvh = capi.Bvh(ctx).build(tris, meshes)
    try:
        assert np.array_equal(bvh.morton_codes(), ob.morton_sorted)
        assert np.array_equal(bvh.triangle_indices(), ob.triangle_indices)
        clusters, parent, left, right, _ = bvh.clusters()
        assert np.array_equal(left, ob.left) and np.array_equal(right, ob.right)
        assert np.array_equal(parent, ob.parent)
        assert np.array_equal(node_words(clusters), node_words(ob.clusters))
        assert np.array_equal(node_words(bvh.flat_nodes()), node_words(oracle.flatten(ob.clusters, ob.left, ob.right)))
        active, merges = bvh.iteration_trace()
        assert np.array_equal(active, ob.trace_active) and np.array_equal(merges, ob.trace_merges)
    finally:
        bvh.close()


@pytest.mark.parametrize("radius", [1, 4, 8, 16])
def test_build_search_radius(ctx, oracle, radius):
    tris, meshes, _ = scenes.soup(5000, seed=77)
    ob = oracle.bvh_build(tris, meshes, search_radius=radius)
    bvh = capi.Bvh(ctx).build(tris, meshes, search_radius=radius)
    try:
        assert np.array_equal(node_words(bvh.flat_nodes()), node_words(oracle.flatten(ob.clusters, ob.left, ob.right)))
    finally:
        bvh.close()


def test_rebuild_reuses_object_and_is_deterministic(ctx, oracle):
    tris, meshes, _ = scenes.soup(40000, seed=5)
    bvh = capi.Bvh(ctx).build(tris, meshes)
    try:
        h0 = oracle.hash_flat_nodes(bvh.flat_nodes())
        for _ in range(3):
            bvh.build(tris, meshes)
            assert oracle.hash_flat_nodes(bvh.flat_nodes()) == h0
        small, m2, _ = scenes.soup(700, seed=6)
        bvh.build(small, m2)  # smaller build in the same (larger) object
        ob = oracle.bvh_build(small, m2)
        assert np.array_equal(node_words(bvh.flat_nodes()), node_words(oracle.flatten(ob.clusters, ob.left, ob.right)))
    finally:
        bvh.close()


def test_cr_bvh_mirror_and_scene_padding(ctx, oracle):
    """cr::BVH-shaped host mirror + glr::Scene with the reference's 65 536/64 padded vectors (Q2)."""
    tris, meshes = scenes.two_mesh_scene()
    sc = rscene.Scene(ctx=ctx, reference_padding=True)
    sc.addMesh(tris[:700], model=meshes["m"][0].reshape(4, 4).T, material_id=1)
    sc.addMesh(tris[700:], model=meshes["m"][1].reshape(4, 4).T, material_id=2)
    flat = sc.sendDataToGpu()
    padded_t, padded_m = sc.getTriangleToGPUData(), sc.getMeshModelToGPUData()
    assert padded_t.size == 65536 and padded_m.size == 64
    ob = oracle.bvh_build(padded_t, padded_m, n=1400)
    assert np.array_equal(node_words(flat), node_words(oracle.flatten(ob.clusters, ob.left, ob.right)))
    p = sc._BVH._InternalStruct
    assert np.array_equal(p._TriangleIndices, ob.triangle_indices)
    assert np.array_equal(p._LeftChild, ob.left) and np.array_equal(p._Parent, ob.parent)


def test_build_rejects_bad_arguments(ctx):
    tris, meshes, _ = scenes.soup(10)
    with pytest.raises(capi.RtrError) as e:
        capi.Bvh(ctx).build(tris, meshes, n=0)
    assert e.value.code == -1
    with pytest.raises(capi.RtrError):
        capi.Bvh(ctx).build(tris, meshes, n=11)       # n > array length
    with pytest.raises(capi.RtrError):
        capi.Bvh(ctx).build(tris, meshes, search_radius=17)
    with pytest.raises(capi.RtrError) as e:
        capi.Bvh(ctx).flat_nodes()                     # nothing built
    assert e.value.code in (-1, -6)


def test_build_one_million_properties(ctx):
    """Config 3 size: structural invariants that do not need the oracle's full arrays."""
    n = 1_000_000
    tris, meshes, _ = scenes.soup(n)
    bvh = capi.Bvh(ctx).build(tris, meshes)
    try:
        flat = bvh.flat_nodes()
        leaves = (flat["left"] == 0) & (flat["right"] == 0)
        assert leaves.sum() == n
        seen = np.zeros(n, bool); seen[flat["tri"][leaves]] = True
        assert seen.all()
        inner = np.nonzero(~leaves)[0]
        assert np.array_equal(flat["left"][inner], inner + 1)
        for side in ("left", "right"):
            ch = flat[side][inner]
            assert np.all(flat["bmin"][inner] <= flat["bmin"][ch]) and np.all(flat["bmax"][inner] >= flat["bmax"][ch])
        active, merges = bvh.iteration_trace()
        assert merges.sum() == n - 1 and active[0] == n
        codes = bvh.morton_codes()
        assert np.all(codes[1:] >= codes[:-1])
    finally:
        bvh.close()


def test_build_one_million_matches_oracle(ctx, oracle):
    n = 1_000_000
    tris, meshes, _ = scenes.soup(n)
    ob = oracle.bvh_build(tris, meshes)
    bvh = capi.Bvh(ctx).build(tris, meshes)
    try:
        assert np.array_equal(bvh.triangle_indices(), ob.triangle_indices)
        assert oracle.hash_flat_nodes(bvh.flat_nodes()) == oracle.hash_flat_nodes(oracle.flatten(ob.clusters, ob.left, ob.right))
    finally:
        bvh.close()


@pytest.mark.parametrize("n", [1, 2, 33, 1025, 30011, 200000])
def test_build_over_64_bit_keys_matches_oracle(ctx, oracle, n):
    """rtr_bvh_build64: leaves ordered by 63-bit Morton keys (21 bits per axis), 8 Onesweep passes over
    (u64, u32) pairs, then the same PLOC + flatten -- bit-exact against oracle orc_bvh_build64."""
    tris, meshes, L = scenes.soup(n, seed=4000 + n)
    ob = oracle.bvh_build(tris, meshes, key_bits=64)
    bvh = capi.Bvh(ctx).build(tris, meshes, key_bits=64)
    try:
        codes64 = bvh.morton_codes64()
        assert np.array_equal(codes64, np.sort(oracle.morton_codes64(tris, meshes), kind="stable"))
        assert np.array_equal(bvh.triangle_indices(), ob.triangle_indices)
        clusters, parent, left, right, _ = bvh.clusters()
        assert np.array_equal(left, ob.left) and np.array_equal(right, ob.right) and np.array_equal(parent, ob.parent)
        assert np.array_equal(node_words(clusters), node_words(ob.clusters))
        assert np.array_equal(node_words(bvh.flat_nodes()), node_words(oracle.flatten(ob.clusters, ob.left, ob.right)))
        # the 64-bit order refines the 32-bit one: code64 >> 33 is the sorted sequence of the reference's codes
        assert np.array_equal((codes64 >> np.uint64(33)).astype(np.uint32), np.sort(oracle.morton_codes(tris, meshes)))
        with pytest.raises(capi.RtrError):
            bvh.morton_codes()
        # the same object goes back to the reference's keys
        bvh.build(tris, meshes)
        ob32 = oracle.bvh_build(tris, meshes)
        assert np.array_equal(bvh.triangle_indices(), ob32.triangle_indices)
        assert np.array_equal(node_words(bvh.flat_nodes()), node_words(oracle.flatten(ob32.clusters, ob32.left, ob32.right)))
    finally:
        bvh.close()


def test_64_bit_keys_trace_like_the_oracle(ctx, oracle):
    tris, meshes, L = scenes.soup(20000, seed=9)
    bvh = capi.Bvh(ctx).build(tris, meshes, key_bits=64)
    try:
        W, H = 96, 64
        cam = synth.soup_camera(L, W, H)
        got = bvh.trace_primary(cam, W, H, W, H)
        exp = oracle.trace_primary(bvh.flat_nodes(), tris, meshes, cam, W, H, W, H)
        assert np.array_equal(got.view(np.uint8), exp.view(np.uint8))
    finally:
        bvh.close()
