// TEST INFRASTRUCTURE: compiles the __host__ __device__ p-value function of
// graph_peak_caller_b200/csrc/poisson.cuh for the CPU so that the CPU test
// suite can compare the exact code the kernels run with scipy.
#include "../../graph_peak_caller_b200/csrc/poisson.cuh"

extern "C" void poisson_eval(const double* k, const double* lam, double* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = gpc::neglog10_poisson_sf(k[i], lam[i]);
}
