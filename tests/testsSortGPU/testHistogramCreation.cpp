// Re-host of tests/testsSortGPU/testHistogramCreation.cpp and testHistogramCreationHard.cpp: the per-bit
// 32-bin histogram of histogramOfGlobalDigitCounts.glsl (bin b = number of keys with bit b set), same
// five cases plus the 65 536-key "hard" case, same element-wise asserts.
#include "harness.hpp"

static const uint32_t HISTOGRAM_SIZE = 32;

static void expectedHistogram(const uint32_t* values, size_t n, uint32_t* out) {
    for (uint32_t b = 0; b < HISTOGRAM_SIZE; ++b) out[b] = 0;
    for (size_t i = 0; i < n; ++i)
        for (uint32_t b = 0; b < HISTOGRAM_SIZE; ++b)
            if (values[i] & (1u << b)) out[b]++;
}

static void runTest(rtr_ctx* ctx, const char* name, const std::vector<uint32_t>& valuesToSort, const uint32_t* expected) {
    std::fprintf(stderr, "\nBegin test: %s...\n", name);
    const uint32_t n = static_cast<uint32_t>(valuesToSort.size());
    void* in = harness::initBuffer(ctx, sizeof(uint32_t) * n, valuesToSort.data());   // binding 2
    uint32_t histogram[HISTOGRAM_SIZE] = {0};
    void* out = harness::initBuffer(ctx, sizeof(histogram), histogram);               // binding 3, pre-zeroed
    HARNESS_CHECK(ctx, rtr_bit_histogram32_dev(ctx, static_cast<const uint32_t*>(in), n, static_cast<uint32_t*>(out)));
    HARNESS_CHECK(ctx, rtr_dev_download(ctx, histogram, out, sizeof(histogram)));
    if (n <= 256) harness::displayBuffer("input", valuesToSort.data(), n);
    harness::displayBuffer("expected", expected, HISTOGRAM_SIZE);
    harness::displayBuffer("results", histogram, HISTOGRAM_SIZE);
    for (uint32_t i = 0; i < HISTOGRAM_SIZE; ++i) assert(histogram[i] == expected[i]);
    // the host-pointer form must agree
    uint32_t again[HISTOGRAM_SIZE];
    HARNESS_CHECK(ctx, rtr_bit_histogram32(ctx, valuesToSort.data(), n, again));
    for (uint32_t i = 0; i < HISTOGRAM_SIZE; ++i) assert(again[i] == expected[i]);
    HARNESS_CHECK(ctx, rtr_dev_free(ctx, in));
    HARNESS_CHECK(ctx, rtr_dev_free(ctx, out));
    std::fprintf(stderr, "Test %s passed\n", name);
}

int main() {
    rtr_ctx* ctx = harness::dummyApplication();
    const uint32_t NB = 130;  // testHistogramCreation.cpp:10
    uint32_t expected[HISTOGRAM_SIZE];
    {   // powers of two: value i is 1 << (i % 32) => bins 0 and 1 hold 5, the others 4
        std::vector<uint32_t> v(NB);
        for (uint32_t i = 0; i < NB; ++i) v[i] = 1u << (i % 32);
        for (uint32_t b = 0; b < HISTOGRAM_SIZE; ++b) expected[b] = NB / 32 + (b < NB % 32 ? 1 : 0);
        runTest(ctx, "powers of two", v, expected);
    }
    for (uint32_t value = 0; value <= 2; ++value) {  // zeros, ones (bin 0 = 130), twos (bin 1 = 130)
        std::vector<uint32_t> v(NB, value);
        for (uint32_t b = 0; b < HISTOGRAM_SIZE; ++b) expected[b] = 0;
        if (value) expected[value - 1] = NB;
        runTest(ctx, value == 0 ? "zeros" : value == 1 ? "ones" : "twos", v, expected);
    }
    {   // random in [0, 8192], expectation recomputed on the host
        std::vector<uint32_t> v(NB);
        harness::initRandomValuesToSort(v.data(), NB);
        expectedHistogram(v.data(), NB, expected);
        runTest(ctx, "random", v, expected);
    }
    {   // testHistogramCreationHard.cpp: 65 536 keys
        std::vector<uint32_t> v(65536);
        harness::initRandomValuesToSort(v.data(), v.size(), 0, 8192, 2);
        expectedHistogram(v.data(), v.size(), expected);
        runTest(ctx, "random hard (65536)", v, expected);
    }
    rtr_ctx_destroy(ctx);
    return EXIT_SUCCESS;
}
