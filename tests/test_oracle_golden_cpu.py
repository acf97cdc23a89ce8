"""The C restatement (oracle/rtr_oracle.c) against the committed outputs of the reference's own
bvh.cpp (tests/golden/) and, where oracle/_ref was built, against the reference library itself."""
import numpy as np
import pytest

import golden_util
from realtimeraytracing_b200.layouts import node_words


@pytest.fixture(scope="module")
def golden():
    return golden_util.load()


@pytest.mark.parametrize("name", golden_util.SCENE_NAMES)
def test_oracle_matches_reference_golden(oracle, golden, name):
    tris, meshes, n = golden_util.scene(name)
    assert int(golden[name + "/n"]) == n
    codes = oracle.morton_codes(tris, meshes, n=n)
    assert oracle.hash_words(codes) == int(golden[name + "/morton_hash"])
    b = oracle.bvh_build(tris, meshes, n=n)
    assert oracle.hash_words(b.triangle_indices) == int(golden[name + "/tri_idx_hash"])
    ch = oracle.hash_words(np.concatenate([node_words(b.clusters).ravel(), b.left, b.right, b.parent]))
    assert ch == int(golden[name + "/cluster_hash"])
    flat = oracle.flatten(b.clusters, b.left, b.right)
    assert oracle.hash_flat_nodes(flat) == int(golden[name + "/flat_hash"])
    if name == "soup512":
        assert np.array_equal(codes, golden[name + "/morton"])
        assert np.array_equal(b.triangle_indices, golden[name + "/tri_idx"])
        assert np.array_equal(node_words(flat), golden[name + "/flat_words"])
        assert np.array_equal(b.left, golden[name + "/left"])
        assert np.array_equal(b.right, golden[name + "/right"])
        assert np.array_equal(b.parent, golden[name + "/parent"])


def test_sorted_codes_are_stable_sort_of_codes(oracle):
    tris, meshes, n = golden_util.scene("dupcodes")
    codes = oracle.morton_codes(tris, meshes, n=n)
    assert np.unique(codes).size < n // 2  # the scene really has ties
    b = oracle.bvh_build(tris, meshes, n=n)
    order = np.argsort(codes, kind="stable")
    assert np.array_equal(b.triangle_indices, order.astype(np.uint32))
    assert np.array_equal(b.morton_sorted, codes[order])


def test_oracle_matches_live_reference_library(oracle):
    """Direct comparison with oracle/_ref/libref_bvh_65536.so (the unmodified reference sources)."""
    from oracle import Reference, reference_available
    if not reference_available(65536):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    ref = Reference(65536)
    for name in ("two_mesh", "padded", "soup512"):
        tris, meshes, n = golden_util.scene(name)
        rb = ref.bvh_build(tris, meshes, n=n)
        b = oracle.bvh_build(tris, meshes, n=n)
        assert np.array_equal(oracle.morton_codes(tris, meshes, n=n), rb.morton_unsorted)
        assert np.array_equal(b.triangle_indices, rb.triangle_indices)
        assert np.array_equal(node_words(b.clusters), node_words(rb.clusters))
        assert np.array_equal(b.left, rb.left) and np.array_equal(b.right, rb.right)
        assert np.array_equal(b.parent, rb.parent)
        assert np.array_equal(rb.is_leaf[:n], np.ones(n, np.uint8)) and not rb.is_leaf[n:].any()  # Q9
