// Re-host of tests/testsSortGPU/testHistogramPrefixSum.cpp: exclusive prefix sum inside each group of
// NB_DIGIT = 4 bins of the 32-bin histogram (prefixSumOfGlobalDigitCounts.glsl), known-answer vector
// (:53-86) and a random case checked against the host recurrence (:43-51).
#include "harness.hpp"

static const uint32_t NB_DIGIT = 4, NB_DIGIT_PLACE = 8, BUFFER_SIZE = NB_DIGIT * NB_DIGIT_PLACE;

static void hostRecurrence(const uint32_t* in, uint32_t* out) {
    for (uint32_t j = 0; j < NB_DIGIT_PLACE; ++j) {
        out[j * NB_DIGIT] = 0;
        for (uint32_t i = 1; i < NB_DIGIT; ++i) {
            const uint32_t cur = j * NB_DIGIT + i;
            out[cur] = in[cur - 1] + out[cur - 1];
        }
    }
}

static void runTest(rtr_ctx* ctx, const char* name, const uint32_t* input, const uint32_t* expected) {
    std::fprintf(stderr, "\nBegin test: %s...\n", name);
    void* in = harness::initBuffer(ctx, sizeof(uint32_t) * BUFFER_SIZE, input);
    uint32_t result[BUFFER_SIZE] = {0};
    void* out = harness::initBuffer(ctx, sizeof(result), result);
    HARNESS_CHECK(ctx, rtr_digitplace_exclusive_scan_dev(ctx, static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out)));
    HARNESS_CHECK(ctx, rtr_dev_download(ctx, result, out, sizeof(result)));
    harness::displayBuffer("input", input, BUFFER_SIZE);
    harness::displayBuffer("expected", expected, BUFFER_SIZE);
    harness::displayBuffer("results", result, BUFFER_SIZE);
    for (uint32_t i = 0; i < BUFFER_SIZE; ++i) assert(result[i] == expected[i]);
    uint32_t again[BUFFER_SIZE];
    HARNESS_CHECK(ctx, rtr_digitplace_exclusive_scan(ctx, input, again));
    for (uint32_t i = 0; i < BUFFER_SIZE; ++i) assert(again[i] == expected[i]);
    HARNESS_CHECK(ctx, rtr_dev_free(ctx, in));
    HARNESS_CHECK(ctx, rtr_dev_free(ctx, out));
    std::fprintf(stderr, "Test %s passed\n", name);
}

int main() {
    rtr_ctx* ctx = harness::dummyApplication();
    {   // known values: [37,41,49,53 | 37,48,44,51 | 35,51,53,41 | 0...] -> [0,37,78,127 | 0,37,85,129 | 0,35,86,139 | 0...]
        uint32_t in[BUFFER_SIZE] = {37, 41, 49, 53, 37, 48, 44, 51, 35, 51, 53, 41};
        uint32_t expected[BUFFER_SIZE] = {0, 37, 78, 127, 0, 37, 85, 129, 0, 35, 86, 139};
        runTest(ctx, "known values", in, expected);
    }
    {   // random: histogram of 32 random keys, expectation from the host recurrence
        uint32_t keys[32], in[BUFFER_SIZE] = {0}, expected[BUFFER_SIZE];
        harness::initRandomValuesToSort(keys, 32);
        for (uint32_t i = 0; i < 32; ++i)
            for (uint32_t b = 0; b < BUFFER_SIZE; ++b)
                if (keys[i] & (1u << b)) in[b]++;
        hostRecurrence(in, expected);
        runTest(ctx, "random", in, expected);
    }
    rtr_ctx_destroy(ctx);
    return EXIT_SUCCESS;
}
