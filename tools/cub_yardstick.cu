// Yardstick only (not part of the product): the toolkit's cub::DeviceRadixSort on the
// same input shape as the event sort, to see how far the hand-written onesweep is
// from a tuned library on this GPU.   nvcc -O3 -arch=sm_100a tools/cub_yardstick.cu
#include <cstdio>
#include <vector>
#include <random>
#include <cub/device/device_radix_sort.cuh>
int main() {
    const size_t n = 16000000;
    const int bits = 27;
    std::vector<unsigned> h(n);
    std::mt19937 rng(1);
    for (auto& x : h) x = rng() & ((1u << bits) - 1);
    unsigned *a, *b, *src;
    cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4); cudaMalloc(&src, n * 4);
    cudaMemcpy(src, h.data(), n * 4, cudaMemcpyHostToDevice);
    size_t tmp_bytes = 0; void* tmp = nullptr;
    cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, a, b, (int)n, 0, bits);
    cudaMalloc(&tmp, tmp_bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 8; ++it) {
        cudaMemcpy(a, src, n * 4, cudaMemcpyDeviceToDevice);
        cudaEventRecord(e0);
        cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, a, b, (int)n, 0, bits);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    printf("cub::DeviceRadixSort::SortKeys  n=%zu bits=%d: %.1f us total (%.1f G keys/s)\n", n, bits, best * 1e3, n / best / 1e6);
    return 0;
}
