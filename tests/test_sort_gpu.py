"""testsSortGPU re-hosted on the C ABI (same cases, same asserts) + the Onesweep sort against the
oracle (std::sort over pairs == stable sort by key).  Bit-exact."""
import numpy as np
import pytest

from realtimeraytracing_b200 import synth

from test_oracle_cpu import KNOWN_IN, KNOWN_OUT

pytestmark = pytest.mark.gpu


# ---- tests/testsSortGPU/testHistogramCreation.cpp:145-195 ----
def test_histogram_powers_of_two(ctx):
    keys = np.array([1 << (i % 32) for i in range(130)], dtype=np.uint32)
    expected = np.zeros(32, dtype=np.uint32)
    for i in range(130):
        expected[i % 32] += 1
    assert np.array_equal(ctx.bit_histogram32(keys), expected)


def test_histogram_zeros_ones_twos(ctx):
    assert np.array_equal(ctx.bit_histogram32(np.zeros(130, np.uint32)), np.zeros(32, np.uint32))
    e = np.zeros(32, np.uint32); e[0] = 130
    assert np.array_equal(ctx.bit_histogram32(np.full(130, 1, np.uint32)), e)
    e = np.zeros(32, np.uint32); e[1] = 130
    assert np.array_equal(ctx.bit_histogram32(np.full(130, 2, np.uint32)), e)


@pytest.mark.parametrize("n", [130, 65536, 1000003])  # random :183-191, Hard variant (65 536 keys), beyond
def test_histogram_random(ctx, oracle, n):
    keys = synth.random_keys_u32(n, seed=n, lo=0, hi=8192 if n <= 65536 else 0xFFFFFFFF)
    assert np.array_equal(ctx.bit_histogram32(keys), oracle.bit_histogram32(keys))


def test_histogram_dev_accumulates_like_the_shader(ctx, oracle):
    """SSBO 3 is pre-zeroed by the test and the shader atomically adds (testHistogramCreation.cpp:113-116)."""
    keys = synth.random_keys_u32(5000, seed=3, lo=0, hi=8192)
    d_keys = ctx.dev_alloc(keys.nbytes)
    d_out = ctx.dev_alloc(128)
    try:
        ctx.upload(d_keys, keys)
        ctx.zero(d_out, 128)
        ctx.bit_histogram32_dev(d_keys, keys.size, d_out)
        ctx.bit_histogram32_dev(d_keys, keys.size, d_out)
        out = np.zeros(32, np.uint32)
        ctx.download(out, d_out)
        assert np.array_equal(out, 2 * oracle.bit_histogram32(keys))
    finally:
        ctx.dev_free(d_keys); ctx.dev_free(d_out)


# ---- tests/testsSortGPU/testHistogramPrefixSum.cpp ----
def test_prefix_known_values(ctx):
    assert np.array_equal(ctx.digitplace_exclusive_scan(np.array(KNOWN_IN, np.uint32)), np.array(KNOWN_OUT, np.uint32))


def test_prefix_random(ctx, oracle):
    keys = synth.random_keys_u32(32, seed=9, lo=0, hi=8192)
    hist = oracle.bit_histogram32(keys)
    assert np.array_equal(ctx.digitplace_exclusive_scan(hist), oracle.digitplace_exclusive_scan(hist))


# ---- Onesweep sort ----
SIZES = [0, 1, 2, 31, 130, 8191, 8192, 8193, 65536, 250001]


@pytest.mark.parametrize("n", SIZES)
def test_sort_keys_u32(ctx, oracle, n):
    keys = synth.random_keys_u32(n, seed=1) if n else np.zeros(0, np.uint32)
    got = ctx.sort_keys_u32(keys)
    assert np.array_equal(got, np.sort(keys, kind="stable"))
    assert np.array_equal(got, oracle.radix_sort_keys_u32(keys))


@pytest.mark.parametrize("n", SIZES)
def test_sort_pairs_u32_is_stable(ctx, oracle, n):
    keys = (synth.random_keys_u32(n, seed=2) & np.uint32(0x3FF003FF)) if n else np.zeros(0, np.uint32)
    vals = np.arange(n, dtype=np.uint32)
    k, v = ctx.sort_pairs_u32(keys, vals)
    ek, ev = oracle.sort_pairs(keys, vals)  # bvh.cpp:227 semantics
    assert np.array_equal(k, ek)
    assert np.array_equal(v, ev)


def test_sort_config1_one_million_keys(ctx):
    """BASELINE config 1: 1M random uint32 keys, mt19937(1), bit-exact vs a stable host sort."""
    keys = synth.random_keys_u32(1_000_000, seed=1)
    assert np.array_equal(ctx.sort_keys_u32(keys), np.sort(keys, kind="stable"))
    harness = synth.random_keys_u32(1_000_000, seed=1, lo=0, hi=8192)  # the harness distribution: heavy duplicates
    vals = np.arange(harness.size, dtype=np.uint32)
    k, v = ctx.sort_pairs_u32(harness, vals)
    order = np.argsort(harness, kind="stable")
    assert np.array_equal(k, harness[order]) and np.array_equal(v, order.astype(np.uint32))


def test_sort_edge_distributions(ctx):
    n = 100_000
    for keys in (np.zeros(n, np.uint32), np.full(n, 0xFFFFFFFF, np.uint32), np.arange(n, dtype=np.uint32)[::-1].copy(),
                 np.arange(n, dtype=np.uint32), (np.arange(n, dtype=np.uint32) % 3) << 29):
        vals = np.arange(n, dtype=np.uint32)
        k, v = ctx.sort_pairs_u32(keys, vals)
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(k, keys[order]) and np.array_equal(v, order.astype(np.uint32))


@pytest.mark.parametrize("n", [0, 1, 130, 4097, 100003])
def test_sort_u64(ctx, oracle, n):
    rng = np.random.RandomState(7)
    keys = (rng.randint(0, 1 << 31, size=n).astype(np.uint64) << np.uint64(33)) ^ rng.randint(0, 1 << 31, size=n).astype(np.uint64)
    if n > 10:
        keys[::5] = keys[1]
    vals = np.arange(n, dtype=np.uint32)
    k, v = ctx.sort_pairs_u64(keys, vals)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order]) and np.array_equal(v, order.astype(np.uint32))
    assert np.array_equal(ctx.sort_keys_u64(keys), keys[order])


def test_sort_dev_bit_range_and_odd_passes(ctx):
    """Device entry point, 30-bit Morton range [0,30) (4 passes) and a 3-pass range (odd: copy back)."""
    n = 50_000
    keys = synth.random_keys_u32(n, seed=5)
    vals = np.arange(n, dtype=np.uint32)
    d_k = ctx.dev_alloc(keys.nbytes); d_v = ctx.dev_alloc(vals.nbytes)
    try:
        for (b0, b1) in ((0, 30), (0, 24), (8, 20), (3, 4)):
            ctx.upload(d_k, keys); ctx.upload(d_v, vals)
            ctx.sort_pairs_u32_dev(d_k, d_v, n, b0, b1)
            k = np.zeros_like(keys); v = np.zeros_like(vals)
            ctx.download(k, d_k); ctx.download(v, d_v)
            field = (keys >> np.uint32(b0)) & np.uint32((1 << (b1 - b0)) - 1)
            order = np.argsort(field, kind="stable")
            assert np.array_equal(k, keys[order]), (b0, b1)
            assert np.array_equal(v, order.astype(np.uint32)), (b0, b1)
        # unaligned key pointer: the non-TMA load path
        ctx.upload(d_k, keys)
        ctx.sort_pairs_u32_dev(d_k + 4, None, n - 1, 0, 32)
        k = np.zeros(n, np.uint32); ctx.download(k, d_k)
        assert k[0] == keys[0] and np.array_equal(k[1:], np.sort(keys[1:]))
    finally:
        ctx.dev_free(d_k); ctx.dev_free(d_v)


def test_sort_full_size_properties(ctx):
    """10M (key,index) pairs -- the headline size: sortedness, permutation, stability (size-independent checks)."""
    n = 10_000_000
    keys = synth.random_keys_u32(n, seed=10) & np.uint32(0x3FFFFFFF)
    vals = np.arange(n, dtype=np.uint32)
    k, v = ctx.sort_pairs_u32(keys, vals)
    assert np.all(k[1:] >= k[:-1])
    assert np.array_equal(keys[v], k)                       # values travelled with their keys
    seen = np.zeros(n, dtype=bool); seen[v] = True
    assert seen.all()                                       # a permutation
    ties = k[1:] == k[:-1]
    assert np.all(v[1:][ties] > v[:-1][ties])               # stable
