// Poisson p-values and Benjamini-Hochberg q-values on sm_100a.
//
// Replaces the reference's
//   PValuesFinder.get_p_values_pileup  sparsepvalues.py:16-31 (via
//       SparseDiffs.apply_binary_func, sparsediffs.py:208-224)
//   PToQValuesMapper._from_subcounts / get_p_to_q_values   sparsepvalues.py:50-59, 95-103
//   QValuesFinder.get_q_values         sparsepvalues.py:116-125
//   SparseValues.threshold_copy        sparsediffs.py:58-62
//
// p-values: merge path over the sample runs (uint32 position, int32 count) and
// the control entries (uint32 position, float64 lambda, positions may repeat:
// the last one wins); one Poisson evaluation per distinct position, output
// canonicalised (neighbouring equal p dropped) through a chained scan.
// q-values: the (p -> base pairs) table is a hash aggregation with warp-level
// pre-combining, the distinct p are radix sorted descending, ranked, turned
// into q with the running minimum, and looked up per run by exact key.
#include "common.cuh"
#include "poisson.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace gpc {

constexpr int PV_THREADS = 256;
constexpr int PV_ITEMS = 8;
constexpr int PV_TILE = PV_THREADS * PV_ITEMS;

__global__ void __launch_bounds__(PV_THREADS)
pvalues_kernel(const uint32_t* __restrict__ spos, const int32_t* __restrict__ sval, unsigned ns,
               double sample_scale, const uint32_t* __restrict__ cpos,
               const double* __restrict__ cval, unsigned nc, double control_scale,
               uint32_t* __restrict__ out_pos,
               double* __restrict__ out_p, unsigned long long* __restrict__ status,
               unsigned* __restrict__ ticket, unsigned long long* __restrict__ out_count) {
    __shared__ unsigned s_slot;
    __shared__ int s_scan[33];
    __shared__ unsigned long long s_obase;
    const unsigned tile = claim_tile(ticket, &s_slot);
    const unsigned total = ns + nc;
    const unsigned d0 = min(total, tile * PV_TILE + threadIdx.x * PV_ITEMS);
    const unsigned d1 = min(total, d0 + PV_ITEMS);
    uint32_t e_pos[PV_ITEMS];
    double e_p[PV_ITEMS];
    int cnt = 0;
    if (d0 < d1) {
        // merge-path split of diagonal d0 (ties: sample entry first)
        unsigned lo = d0 > nc ? d0 - nc : 0, hi = min(d0, ns);
        while (lo < hi) {
            unsigned mid = (lo + hi) >> 1;
            if (spos[mid] <= cpos[d0 - 1 - mid]) lo = mid + 1; else hi = mid;
        }
        unsigned i = lo, j = d0 - lo;
        // state in force before my first element's position
        uint32_t first_pos;
        {
            bool take_s = i < ns && (j >= nc || spos[i] <= cpos[j]);
            first_pos = take_s ? spos[i] : cpos[j];
        }
        unsigned ib = i, jb = j;
        while (ib > 0 && spos[ib - 1] == first_pos) --ib;
        while (jb > 0 && cpos[jb - 1] == first_pos) --jb;
        bool have_prev = (ib > 0) || (jb > 0);
        double prev_p = 0.0;
        if (have_prev) {
            double k = ib > 0 ? (double)sval[ib - 1] : 0.0;
            double lam = jb > 0 ? __dmul_rn(cval[jb - 1], control_scale) : 0.0;
            prev_p = neglog10_poisson_sf(k == 0.0 ? 0.0 : k * sample_scale, lam);
        }
        double k_cur = i > 0 ? (double)sval[i - 1] : 0.0;
        double lam_cur = j > 0 ? __dmul_rn(cval[j - 1], control_scale) : 0.0;
        for (unsigned d = d0; d < d1; ++d) {
            bool take_s = i < ns && (j >= nc || spos[i] <= cpos[j]);
            uint32_t pos;
            if (take_s) { pos = spos[i]; k_cur = (double)sval[i]; ++i; }
            else { pos = cpos[j]; lam_cur = __dmul_rn(cval[j], control_scale); ++j; }
            bool more = i < ns || j < nc;
            uint32_t next_pos = 0;
            if (more) {
                bool ts = i < ns && (j >= nc || spos[i] <= cpos[j]);
                next_pos = ts ? spos[i] : cpos[j];
            }
            if (!more || next_pos != pos) {
                double p = neglog10_poisson_sf(k_cur == 0.0 ? 0.0 : k_cur * sample_scale, lam_cur);
                if (!have_prev || p != prev_p) {
                    e_pos[cnt] = pos;
                    e_p[cnt] = p;
                    ++cnt;
                }
                prev_p = p;
                have_prev = true;
            }
        }
    }
    int tile_out;
    int o = block_excl_sum<int>(cnt, s_scan, tile_out);
    if (threadIdx.x == 0) {
        s_obase = lookback_excl(status, tile, (unsigned long long)tile_out);
        if (tile == gridDim.x - 1) *out_count = s_obase + tile_out;
    }
    __syncthreads();
    unsigned long long w = s_obase + o;
#pragma unroll
    for (int t = 0; t < PV_ITEMS; ++t)
        if (t < cnt) { out_pos[w + t] = e_pos[t]; out_p[w + t] = e_p[t]; }
}

// ---- (p -> base pairs) table --------------------------------------------------

__device__ __forceinline__ unsigned long long hash64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

__global__ void ptable_insert_kernel(const uint32_t* __restrict__ pos, const double* __restrict__ p,
                                     unsigned n, uint32_t track_size,
                                     unsigned long long* __restrict__ tkeys,
                                     unsigned long long* __restrict__ tcount, unsigned mask,
                                     int* __restrict__ fail) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long key = 0, len = 0;
    if (i < n) {
        double v = p[i];
        if (v != 0.0) {
            key = (unsigned long long)__double_as_longlong(v);
            uint32_t next = (i + 1 < n) ? pos[i + 1] : track_size;
            len = next - pos[i];
        }
    }
    // combine equal keys inside the warp before touching the table
    unsigned peers = __match_any_sync(FULL, key);
    unsigned long long sum = 0;
#pragma unroll
    for (int src = 0; src < 32; ++src) {
        unsigned long long l = __shfl_sync(FULL, len, src);
        if ((peers >> src) & 1u) sum += l;
    }
    bool leader = (__ffs(peers) - 1) == (int)lane_id();
    if (key == 0 || !leader || sum == 0) return;
    unsigned h = (unsigned)hash64(key) & mask;
    for (unsigned probe = 0; probe <= mask; ++probe) {
        unsigned long long cur = tkeys[h];
        if (cur == 0) {
            unsigned long long old = atomicCAS(&tkeys[h], 0ull, key);
            cur = old == 0 ? key : old;
        }
        if (cur == key) {
            atomicAdd(&tcount[h], sum);
            return;
        }
        h = (h + 1) & mask;
        if (probe > 4096) break;
    }
    atomicExch(fail, 1);
}

__global__ void ptable_flag_kernel(const unsigned long long* __restrict__ tkeys, unsigned cap,
                                   uint8_t* __restrict__ flag) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) flag[i] = tkeys[i] != 0;
}

__global__ void ptable_compact_kernel(const unsigned long long* __restrict__ tkeys,
                                      const unsigned long long* __restrict__ tcount,
                                      const uint32_t* __restrict__ offs, unsigned cap,
                                      double* __restrict__ out_p,
                                      unsigned long long* __restrict__ out_count) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap && tkeys[i] != 0) {
        out_p[offs[i]] = __longlong_as_double((long long)tkeys[i]);
        out_count[offs[i]] = tcount[i];
    }
}

// ---- q table -------------------------------------------------------------------

__global__ void qtable_keys_kernel(const double* __restrict__ p, unsigned n,
                                   unsigned long long* __restrict__ keys,
                                   uint32_t* __restrict__ idx) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        keys[i] = ~f64_to_key(p[i]);      // ascending sort of ~key == descending p
        idx[i] = i;
    }
}

// Single block: running base-pair count in descending-p order, then per
// distinct p:  q = p + log10(1 + bp with larger p) - log10(N), first clamped at
// 0, running minimum (sparsepvalues.py:95-103).  Writes the distinct p
// (descending) and their q, returns how many.
constexpr int QT_THREADS = 1024;
__global__ void __launch_bounds__(QT_THREADS)
qtable_kernel(const unsigned long long* __restrict__ sorted_keys,
              const uint32_t* __restrict__ sorted_idx,
              const unsigned long long* __restrict__ counts, unsigned n, double log_n,
              double* __restrict__ out_p, double* __restrict__ out_q,
              unsigned long long* __restrict__ n_out) {
    __shared__ unsigned long long s_cnt[33];
    __shared__ int s_int[33];
    __shared__ double s_min[33];
    unsigned long long carry_count = 0;   // bp in all earlier chunks
    double carry_min = INFINITY;          // running minimum over earlier chunks
    unsigned carry_out = 0;
    for (unsigned base = 0; base < n; base += QT_THREADS) {
        unsigned i = base + threadIdx.x;
        bool valid = i < n;
        unsigned long long key = valid ? sorted_keys[i] : 0;
        unsigned long long c = valid ? counts[sorted_idx[i]] : 0;
        unsigned long long tot_c;
        unsigned long long ex_c = block_excl_sum<unsigned long long>(c, s_cnt, tot_c);
        bool last = valid && (i + 1 >= n || sorted_keys[i + 1] != key);
        // bp with strictly larger p = running count before the first entry of this p
        // (computed at the group's first element, consumed at its last)
        double p = valid ? key_to_f64(~key) : 0.0;
        // every element of a group needs the count before the group's first
        // element; groups are tiny, walk back
        double q = INFINITY;
        if (last) {
            unsigned long long before = carry_count + ex_c;
            unsigned g = i;
            while (g > 0 && sorted_keys[g - 1] == key) {
                --g;
                before -= counts[sorted_idx[g]];
            }
            q = p + log10(1.0 + (double)before) - log_n;
            if (g == 0) q = fmax(0.0, q);
        }
        // inclusive running minimum over the block
        double m = q;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            double o = __shfl_up_sync(FULL, m, d);
            if (lane_id() >= (unsigned)d) m = fmin(m, o);
        }
        if (lane_id() == 31) s_min[warp_id()] = m;
        __syncthreads();
        double pre = carry_min;
        for (unsigned w = 0; w < warp_id(); ++w) pre = fmin(pre, s_min[w]);
        m = fmin(m, pre);
        double chunk_min = carry_min;
        for (unsigned w = 0; w < QT_THREADS / 32; ++w) chunk_min = fmin(chunk_min, s_min[w]);
        int tot_o;
        int ex_o = block_excl_sum<int>(last ? 1 : 0, s_int, tot_o);
        if (last) {
            out_p[carry_out + ex_o] = p;
            out_q[carry_out + ex_o] = m;
        }
        carry_count += tot_c;
        carry_min = chunk_min;
        carry_out += tot_o;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = carry_out;
}

__global__ void qvalues_apply_kernel(const double* __restrict__ p, unsigned n,
                                     const double* __restrict__ table_p,
                                     const double* __restrict__ table_q, unsigned n_table,
                                     double* __restrict__ q, int* __restrict__ missing) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = p[i];
    if (v == 0.0) { q[i] = 0.0; return; }       // d[0] = 0
    unsigned lo = 0, hi = n_table;               // table_p is descending
    while (lo < hi) {
        unsigned mid = (lo + hi) >> 1;
        if (table_p[mid] > v) lo = mid + 1; else hi = mid;
    }
    if (lo < n_table && table_p[lo] == v) q[i] = table_q[lo];
    else { q[i] = 0.0; atomicExch(missing, 1); }
}

// threshold: value >= cutoff, canonical boolean runs
__global__ void threshold_flag_kernel(const double* __restrict__ q, unsigned n, double cutoff,
                                      uint8_t* __restrict__ flag) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool cur = q[i] >= cutoff;
    flag[i] = (i == 0) || (cur != (q[i - 1] >= cutoff));
}

__global__ void threshold_emit_kernel(const uint32_t* __restrict__ pos, const double* __restrict__ q,
                                      unsigned n, double cutoff, const uint8_t* __restrict__ flag,
                                      const uint32_t* __restrict__ offs,
                                      uint32_t* __restrict__ out_pos, uint8_t* __restrict__ out_val) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) {
        out_pos[offs[i]] = pos[i];
        out_val[offs[i]] = q[i] >= cutoff;
    }
}

}  // namespace gpc

using namespace gpc;

extern "C" int gpc_pvalues(const uint32_t* sample_pos, const int32_t* sample_val, uint64_t n_sample,
                           double sample_scale, const uint32_t* control_pos,
                           const double* control_val, uint64_t n_control, double control_scale,
                           uint32_t* out_pos, double* out_p, uint64_t* n_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    *n_out = 0;
    uint64_t total = n_sample + n_control;
    if (total == 0) return GPC_OK;
    if (total >= (1ull << 31)) { set_error("tracks too long"); return GPC_ERR_ARG; }
    unsigned tiles = div_up(total, PV_TILE);
    Scratch sc;
    size_t bytes = (size_t)tiles * 8 + 32;
    GPC_CUDA(sc.alloc(bytes, stream));
    GPC_CUDA(cudaMemsetAsync(sc.p, 0, bytes, stream));
    unsigned long long* count = sc.as<unsigned long long>();
    unsigned* ticket = reinterpret_cast<unsigned*>(count + 1);
    unsigned long long* status = count + 4;
    GPC_PROF_BYTES("pvalues", 8.0 * n_sample + 12.0 * n_control);
    pvalues_kernel<<<tiles, PV_THREADS, 0, stream>>>(sample_pos, sample_val, (unsigned)n_sample,
                                                    sample_scale, control_pos, control_val,
                                                    (unsigned)n_control, control_scale, out_pos,
                                                    out_p, status,
                                                    ticket, count);
    GPC_LAUNCH_CHECK();
    unsigned long long h = 0;
    GPC_CUDA(cudaMemcpyAsync(&h, count, 8, cudaMemcpyDeviceToHost, stream));
    GPC_CUDA(cudaStreamSynchronize(stream));
    *n_out = h;
    prof_add_bytes("pvalues", 12.0 * (double)h);
    return GPC_OK;
}

extern "C" int gpc_ptable_local(const uint32_t* pos, const double* p, uint64_t n_runs,
                                uint32_t track_size, double* table_p, uint64_t* table_count,
                                uint64_t table_capacity, uint64_t* n_distinct, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    *n_distinct = 0;
    if (n_runs == 0) return GPC_OK;
    unsigned n = (unsigned)n_runs;
    unsigned long long cap = 1ull << 16;
    while (cap < 2ull * n && cap < (1ull << 22)) cap <<= 1;
    while (true) {
        Scratch keys, counts, fail, flag, offs, total;
        GPC_CUDA(keys.alloc(cap * 8, stream));
        GPC_CUDA(counts.alloc(cap * 8, stream));
        GPC_CUDA(fail.alloc(8, stream));
        GPC_CUDA(cudaMemsetAsync(keys.p, 0, cap * 8, stream));
        GPC_CUDA(cudaMemsetAsync(counts.p, 0, cap * 8, stream));
        GPC_CUDA(cudaMemsetAsync(fail.p, 0, 8, stream));
        GPC_PROF("ptable_insert");
        ptable_insert_kernel<<<div_up(n, 256), 256, 0, stream>>>(
            pos, p, n, track_size, keys.as<unsigned long long>(), counts.as<unsigned long long>(),
            (unsigned)(cap - 1), fail.as<int>());
        GPC_LAUNCH_CHECK();
        GPC_CUDA(flag.alloc(cap, stream));
        GPC_CUDA(offs.alloc(cap * 4, stream));
        GPC_CUDA(total.alloc(8, stream));
        GPC_PROF("ptable_flag");
        ptable_flag_kernel<<<div_up(cap, 256), 256, 0, stream>>>(keys.as<unsigned long long>(),
                                                                (unsigned)cap, flag.as<uint8_t>());
        GPC_LAUNCH_CHECK();
        GPC_TRY(exclusive_scan<uint8_t>(flag.as<uint8_t>(), offs.as<uint32_t>(), cap,
                                        total.as<unsigned long long>(), stream));
        unsigned long long h = 0;
        int hfail = 0;
        GPC_CUDA(cudaMemcpyAsync(&h, total.p, 8, cudaMemcpyDeviceToHost, stream));
        GPC_CUDA(cudaMemcpyAsync(&hfail, fail.p, 4, cudaMemcpyDeviceToHost, stream));
        GPC_CUDA(cudaStreamSynchronize(stream));
        if (hfail || h * 2 > cap) {               // too full: grow and redo
            if (cap >= 4ull * n + (1ull << 16)) { set_error("p table overflow"); return GPC_ERR_INTERNAL; }
            cap <<= 2;
            continue;
        }
        *n_distinct = h;
        if (h > table_capacity) return GPC_ERR_CAPACITY;
        GPC_PROF("ptable_compact");
        ptable_compact_kernel<<<div_up(cap, 256), 256, 0, stream>>>(
            keys.as<unsigned long long>(), counts.as<unsigned long long>(), offs.as<uint32_t>(),
            (unsigned)cap, table_p, reinterpret_cast<unsigned long long*>(table_count));
        GPC_LAUNCH_CHECK();
        GPC_CUDA(cudaStreamSynchronize(stream));
        return GPC_OK;
    }
}

extern "C" int gpc_qtable_build(const double* table_p, const uint64_t* table_count,
                                uint64_t n_entries, uint64_t total_bp, double* out_p, double* out_q,
                                uint64_t* n_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    *n_out = 0;
    if (n_entries == 0) return GPC_OK;
    unsigned n = (unsigned)n_entries;
    Scratch k0, k1, i0, i1, cnt;
    GPC_CUDA(k0.alloc((size_t)n * 8, stream));
    GPC_CUDA(k1.alloc((size_t)n * 8, stream));
    GPC_CUDA(i0.alloc((size_t)n * 4, stream));
    GPC_CUDA(i1.alloc((size_t)n * 4, stream));
    GPC_CUDA(cnt.alloc(8, stream));
    GPC_PROF("qtable_keys");
    qtable_keys_kernel<<<div_up(n, 256), 256, 0, stream>>>(table_p, n, k0.as<unsigned long long>(),
                                                          i0.as<uint32_t>());
    GPC_LAUNCH_CHECK();
    bool in_tmp = false;
    GPC_TRY((radix_sort<unsigned long long, true>(k0.as<unsigned long long>(),
                                                  k1.as<unsigned long long>(), i0.as<uint32_t>(),
                                                  i1.as<uint32_t>(), n, 0, 64, true, &in_tmp,
                                                  stream)));
    GPC_PROF("qtable");
    qtable_kernel<<<1, QT_THREADS, 0, stream>>>(
        in_tmp ? k1.as<unsigned long long>() : k0.as<unsigned long long>(),
        in_tmp ? i1.as<uint32_t>() : i0.as<uint32_t>(),
        reinterpret_cast<const unsigned long long*>(table_count), n, log10((double)total_bp),
        out_p, out_q, cnt.as<unsigned long long>());
    GPC_LAUNCH_CHECK();
    unsigned long long h = 0;
    GPC_CUDA(cudaMemcpyAsync(&h, cnt.p, 8, cudaMemcpyDeviceToHost, stream));
    GPC_CUDA(cudaStreamSynchronize(stream));
    *n_out = h;
    return GPC_OK;
}

extern "C" int gpc_qvalues_apply(const double* p, uint64_t n_runs, const double* table_p,
                                 const double* table_q, uint64_t n_table, double* out_q,
                                 void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_runs == 0) return GPC_OK;
    Scratch miss;
    GPC_CUDA(miss.alloc(8, stream));
    GPC_CUDA(cudaMemsetAsync(miss.p, 0, 8, stream));
    GPC_PROF("qvalues_apply");
    qvalues_apply_kernel<<<div_up(n_runs, 256), 256, 0, stream>>>(
        p, (unsigned)n_runs, table_p, table_q, (unsigned)n_table, out_q, miss.as<int>());
    GPC_LAUNCH_CHECK();
    int h = 0;
    GPC_CUDA(cudaMemcpyAsync(&h, miss.p, 4, cudaMemcpyDeviceToHost, stream));
    GPC_CUDA(cudaStreamSynchronize(stream));
    if (h) { set_error("p-value without an entry in the p->q table"); return GPC_ERR_ARG; }
    return GPC_OK;
}

extern "C" int gpc_threshold(const uint32_t* pos, const double* q, uint64_t n_runs, double cutoff,
                             uint32_t* out_pos, uint8_t* out_val, uint64_t* n_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    *n_out = 0;
    if (n_runs == 0) return GPC_OK;
    unsigned n = (unsigned)n_runs;
    Scratch flag, offs, total;
    GPC_CUDA(flag.alloc(n, stream));
    GPC_CUDA(offs.alloc((size_t)n * 4, stream));
    GPC_CUDA(total.alloc(8, stream));
    GPC_PROF("threshold_flag");
    threshold_flag_kernel<<<div_up(n, 256), 256, 0, stream>>>(q, n, cutoff, flag.as<uint8_t>());
    GPC_LAUNCH_CHECK();
    GPC_TRY(exclusive_scan<uint8_t>(flag.as<uint8_t>(), offs.as<uint32_t>(), n,
                                    total.as<unsigned long long>(), stream));
    GPC_PROF("threshold_emit");
    threshold_emit_kernel<<<div_up(n, 256), 256, 0, stream>>>(pos, q, n, cutoff, flag.as<uint8_t>(),
                                                             offs.as<uint32_t>(), out_pos, out_val);
    GPC_LAUNCH_CHECK();
    unsigned long long h = 0;
    GPC_CUDA(cudaMemcpyAsync(&h, total.p, 8, cudaMemcpyDeviceToHost, stream));
    GPC_CUDA(cudaStreamSynchronize(stream));
    *n_out = h;
    return GPC_OK;
}
