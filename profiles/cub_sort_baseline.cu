// Library yardstick for DESIGN.md §4.2 (NOT part of the product, not linked into librtr_b200): CUB's
// DeviceRadixSort::SortPairs (its own Onesweep implementation) on the same 10 M / 1 M (u32 key, u32 value) pairs, timed
// with CUDA events.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a profiles/cub_sort_baseline.cu -o gpurun_out/cub_sort_baseline
#include <cub/device/device_radix_sort.cuh>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <random>

static float time_sort(uint32_t n, int end_bit, int reps) {
    std::vector<uint32_t> h(n);
    std::mt19937 rng(1);
    for (auto& x : h) x = rng() >> (32 - end_bit);
    uint32_t *k0, *k1, *v0, *v1;
    cudaMalloc(&k0, 4ull * n); cudaMalloc(&k1, 4ull * n); cudaMalloc(&v0, 4ull * n); cudaMalloc(&v1, 4ull * n);
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k0, k1, v0, v1, n, 0, end_bit);
    void* tmp; cudaMalloc(&tmp, tmp_bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaMemcpy(k0, h.data(), 4ull * n, cudaMemcpyHostToDevice);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, n, 0, end_bit);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(tmp);
    return best;
}

int main() {
    for (uint32_t n : {10000000u, 1000000u})
        for (int bits : {32, 30}) {
            const float ms = time_sort(n, bits, 6);
            std::printf("cub::DeviceRadixSort::SortPairs  n=%u  bits 0..%d: %.3f ms = %.1f Gkeys/s\n", n, bits, ms, n / ms / 1e6);
        }
    return 0;
}
