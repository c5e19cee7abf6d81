// Exported device-wide primitives (used by the host side for small jobs and
// by the parity tests to exercise the sort / scan building blocks directly).
#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

using namespace gpc;

extern "C" int gpc_sort_u32(uint32_t* keys, uint32_t* keys_tmp, uint64_t n, int32_t begin_bit,
                            int32_t end_bit, int32_t* result_in_tmp, void* stream) {
    GPC_STAGE();
    bool in_tmp = false;
    int s = radix_sort<uint32_t, false>(keys, keys_tmp, nullptr, nullptr, n, begin_bit, end_bit,
                                        false, &in_tmp, (cudaStream_t)stream);
    *result_in_tmp = in_tmp ? 1 : 0;
    return s;
}

extern "C" int gpc_sort_u64_pairs(uint64_t* keys, uint64_t* keys_tmp, uint32_t* vals,
                                  uint32_t* vals_tmp, uint64_t n, int32_t begin_bit,
                                  int32_t end_bit, int32_t* result_in_tmp, void* stream) {
    GPC_STAGE();
    bool in_tmp = false;
    int s;
    if (vals)
        s = radix_sort<unsigned long long, true>(
            reinterpret_cast<unsigned long long*>(keys),
            reinterpret_cast<unsigned long long*>(keys_tmp), vals, vals_tmp, n, begin_bit, end_bit,
            true, &in_tmp, (cudaStream_t)stream);
    else
        s = radix_sort<unsigned long long, false>(
            reinterpret_cast<unsigned long long*>(keys),
            reinterpret_cast<unsigned long long*>(keys_tmp), nullptr, nullptr, n, begin_bit,
            end_bit, true, &in_tmp, (cudaStream_t)stream);
    *result_in_tmp = in_tmp ? 1 : 0;
    return s;
}

extern "C" int gpc_exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint64_t n,
                                      uint64_t* total, void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    Scratch t;
    GPC_CUDA(t.alloc(8, stream));
    GPC_TRY(exclusive_scan<uint32_t>(in, out, n, t.as<unsigned long long>(), stream));
    unsigned long long h = 0;
    GPC_READ_BACK({t.p, &h, (unsigned)(8)});
    *total = h;
    return GPC_OK;
}
