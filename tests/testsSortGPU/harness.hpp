// harness.hpp -- what the reference's tests/testsSortGPU get from glr::Application::dummyApplication(),
// glr::Shader/Program and the SSBO helpers (tests/testsSortGPU/testHistogramCreation.cpp:75-98), re-hosted
// on the C ABI of librtr_b200.so: a device context instead of a GL context, device buffers instead of
// SSBOs 2 and 3, a kernel launch instead of glDispatchCompute.  Same test shape: fill input, zero output,
// run, read back, assert element by element.
#pragma once
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "rtr.h"

#define HARNESS_CHECK(ctx, call)                                                                        \
    do {                                                                                                \
        const int _rc = (call);                                                                         \
        if (_rc != RTR_OK) {                                                                            \
            std::fprintf(stderr, "%s:%d %s -> %d: %s\n", __FILE__, __LINE__, #call, _rc, rtr_last_error(ctx)); \
            std::exit(EXIT_FAILURE);                                                                    \
        }                                                                                               \
    } while (0)

namespace harness {

// dummyApplication() (srcOpenGL/application.cpp:416-423): "give me a device to run compute on"
inline rtr_ctx* dummyApplication() {
    rtr_ctx* ctx = nullptr;
    HARNESS_CHECK(nullptr, rtr_ctx_create(0, &ctx));
    return ctx;
}

// initBuffer (testHistogramCreation.cpp:75-85): create a device buffer holding `data`
inline void* initBuffer(rtr_ctx* ctx, size_t bytes, const void* data) {
    void* p = nullptr;
    HARNESS_CHECK(ctx, rtr_dev_alloc(ctx, bytes, &p));
    HARNESS_CHECK(ctx, rtr_dev_upload(ctx, p, data, bytes));
    return p;
}

inline void displayBuffer(const std::string& name, const uint32_t* buffer, size_t n) {
    std::fprintf(stderr, "%s:\n[", name.c_str());
    for (size_t i = 0; i + 1 < n; ++i) std::fprintf(stderr, "%u, ", buffer[i]);
    std::fprintf(stderr, "%u]\n", buffer[n - 1]);
}

// the reference seeds from std::random_device (not reproducible); a fixed seed keeps failures replayable
inline void initRandomValuesToSort(uint32_t* array, size_t n, uint32_t lo = 0, uint32_t hi = 8192, uint32_t seed = 1) {
    std::mt19937 gen(seed);
    std::uniform_int_distribution<uint32_t> distrib(lo, hi);
    for (size_t i = 0; i < n; ++i) array[i] = distrib(gen);
}

}  // namespace harness
