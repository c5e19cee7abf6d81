// How long does the host wait to learn a count a kernel just produced?
#include <cstdio>
#include <chrono>
#include <cuda_runtime.h>
__global__ void produce(unsigned long long* c, unsigned long long v, int spin) {
    long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    *c = v;
}
__global__ void produce_mapped(volatile unsigned long long* host, unsigned long long v, int spin) {
    long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    *host = v;
    __threadfence_system();
}
__global__ void consumer(unsigned long long* c) { if (*c == 123456789) printf("x"); }
static double now() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    unsigned long long *d, *pinned, *mapped, *mapped_dev, pageable = 0;
    cudaMalloc(&d, 8);
    cudaHostAlloc(&pinned, 8, cudaHostAllocDefault);
    cudaHostAlloc(&mapped, 8, cudaHostAllocMapped);
    cudaHostGetDevicePointer(&mapped_dev, mapped, 0);
    cudaStream_t s; cudaStreamCreate(&s);
    const int spin = 40000;   // ~20 us kernel
    for (int mode = 0; mode < 4; ++mode) {
        double tot = 0; int N = 200;
        for (int it = 0; it < N + 20; ++it) {
            cudaStreamSynchronize(s);
            double t0 = now();
            if (mode == 0) { produce<<<1, 1, 0, s>>>(d, it + 1, spin); cudaMemcpyAsync(&pageable, d, 8, cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); }
            if (mode == 1) { produce<<<1, 1, 0, s>>>(d, it + 1, spin); cudaMemcpyAsync(pinned, d, 8, cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); }
            if (mode == 2) { *mapped = 0; produce_mapped<<<1, 1, 0, s>>>(mapped_dev, it + 1, spin); while (*(volatile unsigned long long*)mapped != (unsigned long long)(it + 1)) {} }
            if (mode == 3) { produce<<<1, 1, 0, s>>>(d, it + 1, spin); cudaStreamSynchronize(s); }
            // next launch right after the count is known
            consumer<<<1, 1, 0, s>>>(d);
            double t1 = now();
            if (it >= 20) tot += t1 - t0;
        }
        const char* names[] = {"pageable memcpyAsync+sync", "pinned memcpyAsync+sync", "mapped host word, polled", "sync only (no count)"};
        printf("%-28s %.1f us per launch+count+next launch (kernel itself ~%.1f us)\n", names[mode], tot / N, spin / 1965.0);
    }
    return 0;
}
