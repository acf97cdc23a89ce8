// BASELINE.json configs[0]: 1M random uint32 keys, LSD radix sort checked bit-exact against
// std::stable_sort on the host -- the test the reference's harness was building towards
// (chainedScanDigitBinning.glsl is an empty main).  Also the (code, index) pair sort that replaces
// std::sort at srcCommon/scene/geometry/bvh.cpp:223-231, edge sizes, and 64-bit keys.
#include <algorithm>
#include <numeric>

#include "harness.hpp"

static void testKeys(rtr_ctx* ctx, size_t n, uint32_t lo, uint32_t hi, uint32_t seed) {
    std::vector<uint32_t> keys(n), expected;
    harness::initRandomValuesToSort(keys.data(), n, lo, hi, seed);
    expected = keys;
    std::stable_sort(expected.begin(), expected.end());
    HARNESS_CHECK(ctx, rtr_sort_keys_u32(ctx, keys.data(), static_cast<uint32_t>(n)));
    for (size_t i = 0; i < n; ++i) assert(keys[i] == expected[i]);
    std::fprintf(stderr, "keys n=%zu range [%u,%u]: passed\n", n, lo, hi);
}

static void testPairs(rtr_ctx* ctx, size_t n, uint32_t hi, uint32_t seed) {
    std::vector<uint32_t> keys(n), idx(n), order(n);
    harness::initRandomValuesToSort(keys.data(), n, 0, hi, seed);
    std::iota(idx.begin(), idx.end(), 0u);
    std::iota(order.begin(), order.end(), 0u);
    // what bvh.cpp:223-231 computes: pairs (code, index) sorted ascending == stable sort by code
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    std::vector<uint32_t> sortedKeys(n);
    for (size_t i = 0; i < n; ++i) sortedKeys[i] = keys[order[i]];
    HARNESS_CHECK(ctx, rtr_sort_pairs_u32(ctx, keys.data(), idx.data(), static_cast<uint32_t>(n)));
    for (size_t i = 0; i < n; ++i) assert(keys[i] == sortedKeys[i] && idx[i] == order[i]);
    std::fprintf(stderr, "pairs n=%zu keys<=%u: passed\n", n, hi);
}

static void testKeys64(rtr_ctx* ctx, size_t n) {
    std::mt19937_64 gen(5);
    std::vector<uint64_t> keys(n), expected;
    for (auto& k : keys) k = gen();
    expected = keys;
    std::stable_sort(expected.begin(), expected.end());
    HARNESS_CHECK(ctx, rtr_sort_keys_u64(ctx, keys.data(), static_cast<uint32_t>(n)));
    for (size_t i = 0; i < n; ++i) assert(keys[i] == expected[i]);
    std::fprintf(stderr, "keys64 n=%zu: passed\n", n);
}

int main() {
    rtr_ctx* ctx = harness::dummyApplication();
    testKeys(ctx, 1000000, 0, 0xFFFFFFFFu, 1);  // config 1
    testKeys(ctx, 1000000, 0, 8192, 1);         // the harness distribution (many duplicates)
    for (size_t n : {size_t(1), size_t(2), size_t(130), size_t(6143), size_t(6144), size_t(6145), size_t(65536)})
        testKeys(ctx, n, 0, 0xFFFFFFFFu, 3);
    testPairs(ctx, 1000000, 0x3FFFFFFFu, 2);    // 30-bit Morton-code range
    testPairs(ctx, 100000, 255, 4);             // heavy ties: stability is what is being checked
    testKeys64(ctx, 300000);
    uint32_t none = 0;
    HARNESS_CHECK(ctx, rtr_sort_keys_u32(ctx, &none, 0));  // empty input is a no-op
    rtr_ctx_destroy(ctx);
    return EXIT_SUCCESS;
}
