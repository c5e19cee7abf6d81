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
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "poisson.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace gpc {

// Opportunistic memo of p(count, lambda): tracks take few distinct (count,
// lambda) pairs (tens of thousands) compared to runs (tens of millions).
// Whoever evaluates a pair publishes it; later visitors reuse it.  Nobody ever
// waits: a pair that is not published (yet) is simply evaluated again.
//
// One 32-byte, write-once entry per pair.  The fast probe uses ordinary
// (L1-cached) loads: an entry whose tag says READY was complete when the line
// was fetched, and a stale line can only say "empty"/"claimed", which sends the
// pair to the tile's miss queue, where the probe is repeated against L2.
// The payload validates itself: check = bits(count) ^ bits(lam) ^ bits(p).  A reader whose
// tag and payload loads were served by different fills of the line (the fast probe has no
// acquire ordering; the compiler may also split the entry into two 16-byte loads) cannot pair
// a READY tag with a stale p: the check fails and the pair goes to the miss queue.  The only
// stale payload that passes is all zeroes for (count 0, lambda 0), whose p is 0 anyway.
struct __align__(32) MemoEntry {
    unsigned long long tag;    // 0 empty | hash|1 claimed | hash|3 ready (hash has bits 0-1 clear)
    double lam, p;
    unsigned long long check;
};
__device__ __forceinline__ unsigned long long memo_check(double count, double lam, double p) {
    return (unsigned long long)__double_as_longlong(count) ^ (unsigned long long)__double_as_longlong(lam) ^
           (unsigned long long)__double_as_longlong(p);
}
struct PMemo {
    MemoEntry* e;              // [mask + 1] or nullptr
    unsigned mask;
};

__device__ __forceinline__ unsigned long long memo_hash(double count, double lam) {
    unsigned long long x = (unsigned long long)__double_as_longlong(lam);
    x ^= (unsigned long long)__double_as_longlong(count) * 0x9e3779b97f4a7c15ull;
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x & ~3ull;
}

constexpr int MEMO_PROBES = 16;

// Read-only probe.  FROM_L2 bypasses L1 (used right before an evaluation).
template <bool FROM_L2>
__device__ __forceinline__ bool memo_find(const PMemo& m, unsigned long long h, double count,
                                          double lam, double& p) {
    unsigned s = (unsigned)(h >> 20) & m.mask;
    for (int probe = 0; probe < MEMO_PROBES; ++probe) {
        const MemoEntry* e = &m.e[s];
        unsigned long long tag = FROM_L2 ? __ldcg(&e->tag) : e->tag;
        if (tag == (h | 3ull)) {
            const double l = FROM_L2 ? __ldcg(&e->lam) : e->lam;
            const double v = FROM_L2 ? __ldcg(&e->p) : e->p;
            const unsigned long long chk = FROM_L2 ? __ldcg(&e->check) : e->check;
            if (l == lam && chk == memo_check(count, lam, v)) {
                p = v;
                return true;
            }
        } else if (tag == 0ull || (tag | 3ull) == (h | 3ull)) {
            return false;          // end of the chain, or the pair is being evaluated right now
        }
        s = (s + 1) & m.mask;      // another pair lives (or is about to live) here
    }
    return false;
}

// Publishes an evaluated pair in the first free slot of its chain (best effort).
__device__ __forceinline__ void memo_publish(const PMemo& m, unsigned long long h, double count,
                                             double lam, double p) {
    unsigned s = (unsigned)(h >> 20) & m.mask;
    for (int probe = 0; probe < MEMO_PROBES; ++probe) {
        MemoEntry* e = &m.e[s];
        unsigned long long tag = __ldcg(&e->tag);
        if (tag == 0ull) tag = atomicCAS(&e->tag, 0ull, h | 1ull);
        if (tag == 0ull) {                                     // mine
            e->lam = lam;
            e->p = p;
            e->check = memo_check(count, lam, p);
            __threadfence();
            *reinterpret_cast<volatile unsigned long long*>(&e->tag) = h | 3ull;
            return;
        }
        if ((tag | 3ull) == (h | 3ull)) return;                // somebody else publishes this hash
        s = (s + 1) & m.mask;
    }
}

constexpr int PV_THREADS = 256;
constexpr int PV_ITEMS = 8;
constexpr int PV_TILE = PV_THREADS * PV_ITEMS;

// Coarse merge-path partition: number of sample entries among the first
// t * PV_TILE merged elements (ties: sample entry first), one thread per tile edge.
__global__ void pv_partition_kernel(const uint32_t* __restrict__ spos, unsigned ns,
                                    const uint32_t* __restrict__ cpos, unsigned nc,
                                    unsigned n_tiles, uint32_t* __restrict__ split) {
    unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    unsigned long long dd = (unsigned long long)t * PV_TILE;
    unsigned d = (unsigned)min(dd, (unsigned long long)ns + nc);
    unsigned lo = d > nc ? d - nc : 0, hi = min(d, ns);
    while (lo < hi) {
        unsigned mid = (lo + hi) >> 1;
        if (spos[mid] <= cpos[d - 1 - mid]) lo = mid + 1; else hi = mid;
    }
    split[t] = lo;
}

// Step 1: merge the two tracks and evaluate p at every distinct position.
//
// The number of positions a tile emits is known from the merge alone, so the
// tile publishes its share of the chained scan BEFORE any Poisson evaluation:
// a tile that has to evaluate new (count, lambda) pairs (tens of microseconds)
// delays nobody else.  (The first version compacted equal neighbouring p in
// this kernel; the output count then depended on p, every tile waited in the
// look-back for the slowest tile of its wave, and 56 % of all stall samples
// sat on that barrier -- profiles/pvalues_r01_before_raw.csv.)
// Pairs that the L1 probe does not find go to a per-tile queue that the whole
// block drains, one pair per thread, so evaluations run on full warps.  (The two block
// barriers around the queue hold 20 % of the kernel's stall samples,
// profiles/pv_eval_r02_final_raw.csv; resolving the misses in their own threads instead --
// no queue, no barriers -- was measured at 0.485 ms against 0.397: eight inlined copies of the
// chain walk and the Poisson tail per thread cost far more than the barriers.)
//
// (Measured and dropped: collecting a tile's DISTINCT pairs in a shared-memory table first --
// one CAS per new pair, one memo probe per occupied slot, every position reading its p from
// shared memory -- cut the L2 probes twentyfold and made the kernel SLOWER, 0.485 -> 0.638 ms
// on the chr22-scale tracks: three more block barriers and same-address shared-memory traffic
// cost more than the probes they replaced.)
//
// Three changes of round 2 (A/B on one box, chr22-scale tracks: 0.487 ms as it stood, 0.463
// with the first, 0.419 with the first two, 0.397 with all three; without any memo probe the
// kernel takes 0.29 ms):
//  * a thread's share of both slices (8-9 elements) is fetched by loads issued back to back
//    and stored to shared memory afterwards; written as two load / store loops with a dynamic
//    trip count, every iteration waited for its own DRAM round trip;
//  * a tile's entries go to shared memory at their rank in the tile and leave as whole
//    sectors (a thread's eight entries written straight to global memory are eight
//    instructions of 32 scattered partial sectors each);
//  * the first memo slot of four pairs at a time is fetched whole (two 16-byte loads per
//    entry, all eight in flight together) and checked afterwards; called one after the other
//    the probes were eight dependent pairs of round trips per thread.  A pair whose first slot
//    holds something else goes to the miss queue, where the chain is walked.
__global__ void __launch_bounds__(PV_THREADS)
pv_eval_kernel(const uint32_t* __restrict__ spos, const int32_t* __restrict__ sval, unsigned ns,
               double sample_scale, const uint32_t* __restrict__ cpos,
               const double* __restrict__ cval, unsigned nc, double control_scale, PMemo memo,
               const uint32_t* __restrict__ split, uint32_t* __restrict__ out_pos,
               double* __restrict__ out_p, unsigned long long* __restrict__ status,
               unsigned* __restrict__ ticket, unsigned long long* __restrict__ out_count) {
    // slices of both tracks that this tile merges, plus one element before
    // (state in force at the tile's left edge) and one after (look-ahead);
    // after the merge a_val / b_val are reused as the miss queue
    __shared__ __align__(16) uint32_t a_pos[PV_TILE + 2];
    __shared__ __align__(16) int32_t a_val[PV_TILE + 2];
    __shared__ __align__(16) uint32_t b_pos[PV_TILE + 2];
    __shared__ __align__(16) double b_val[PV_TILE + 2];
    __shared__ unsigned s_slot;
    __shared__ unsigned s_nq;
    __shared__ int s_scan[33];
    __shared__ unsigned long long s_obase;
    const unsigned tile = claim_tile(ticket, &s_slot);
    if (threadIdx.x == 0) s_nq = 0;
    const unsigned total = ns + nc;
    const unsigned t0 = min(total, tile * PV_TILE), t1 = min(total, t0 + PV_TILE);
    const unsigned i_lo = split[tile], i_hi = split[tile + 1];
    const unsigned j_lo = t0 - i_lo, j_hi = t1 - i_hi;
    const unsigned na = i_hi - i_lo, nb = j_hi - j_lo;
    // slot k of a_* holds global sample entry i_lo - 1 + k  (k = 0 .. na + 1), slot k of b_* control
    // entry j_lo - 1 + k
    {
        // one index space over both slices: [0, na + 2) sample, [na + 2, na + nb + 4) control
        constexpr int ROUNDS = (PV_TILE + 4 + PV_THREADS - 1) / PV_THREADS;      // 9
        uint32_t r_pos[ROUNDS];
        unsigned long long r_val[ROUNDS];
#pragma unroll
        for (int it = 0; it < ROUNDS; ++it) {
            const unsigned k = threadIdx.x + it * PV_THREADS;
            r_pos[it] = 0xffffffffu;
            r_val[it] = 0ull;
            if (k < na + 2) {
                const long long gi = (long long)i_lo - 1 + k;
                if (gi >= 0 && gi < (long long)ns) {
                    r_pos[it] = spos[gi];
                    r_val[it] = (unsigned long long)(unsigned)sval[gi];
                }
            } else if (k < na + nb + 4) {
                const long long gj = (long long)j_lo - 1 + (k - (na + 2));
                if (gj >= 0 && gj < (long long)nc) {
                    r_pos[it] = cpos[gj];
                    r_val[it] = (unsigned long long)__double_as_longlong(cval[gj]);
                }
            }
        }
#pragma unroll
        for (int it = 0; it < ROUNDS; ++it) {
            const unsigned k = threadIdx.x + it * PV_THREADS;
            if (k < na + 2) {
                a_pos[k] = r_pos[it];
                a_val[k] = (int32_t)(unsigned)r_val[it];
            } else if (k < na + nb + 4) {
                const unsigned kb = k - (na + 2);
                b_pos[kb] = r_pos[it];
                b_val[kb] = r_pos[it] == 0xffffffffu && r_val[it] == 0ull
                                ? 0.0 : __dmul_rn(__longlong_as_double((long long)r_val[it]), control_scale);
            }
        }
    }
    __syncthreads();
    // local coordinates: sample entry x <-> a_*[x + 1], control entry y <-> b_*[y + 1]
    const unsigned m = na + nb;
    const unsigned d0 = min(m, threadIdx.x * PV_ITEMS), d1 = min(m, d0 + PV_ITEMS);
    uint32_t e_pos[PV_ITEMS];
    int32_t e_k[PV_ITEMS];
    double e_lam[PV_ITEMS];
    int cnt = 0;
    if (d0 < d1) {
        // merge-path split of the local diagonal d0 (ties: sample entry first)
        unsigned lo = d0 > nb ? d0 - nb : 0, hi = min(d0, na);
        while (lo < hi) {
            unsigned mid = (lo + hi) >> 1;
            if (a_pos[mid + 1] <= b_pos[d0 - 1 - mid + 1]) lo = mid + 1; else hi = mid;
        }
        unsigned i = lo, j = d0 - lo;
        // values in force before my first element (slot 0 = entry before the tile; empty -> 0)
        int32_t k_cur = a_val[i];
        double lam_cur = b_val[j];
        for (unsigned d = d0; d < d1; ++d) {
            bool take_s = i < na && (j >= nb || a_pos[i + 1] <= b_pos[j + 1]);
            uint32_t pos;
            if (take_s) { pos = a_pos[i + 1]; k_cur = a_val[i + 1]; ++i; }
            else { pos = b_pos[j + 1]; lam_cur = b_val[j + 1]; ++j; }
            // look-ahead: slots na + 1 / nb + 1 hold the first entries after the tile;
            // the last entry of a position group carries the values in force there
            uint32_t next_pos = min(a_pos[i + 1], b_pos[j + 1]);     // 0xffffffff when exhausted
            if (next_pos != pos) {
                e_pos[cnt] = pos;
                e_k[cnt] = k_cur;
                e_lam[cnt] = lam_cur;
                ++cnt;
            }
        }
    }
    int tile_out;
    int o = block_excl_sum<int>(cnt, s_scan, tile_out);     // (barriers: the merge is over)
    if (threadIdx.x < 32) {
        unsigned long long b = lookback_excl_warp(status, tile, (unsigned long long)tile_out);
        if (threadIdx.x == 0) {
            s_obase = b;
            if (tile == gridDim.x - 1) *out_count = b + tile_out;
        }
    }
    double e_p[PV_ITEMS];
    // fast path: the first memo slot of every pair, four pairs per round trip
    unsigned miss = 0;
    if (memo.e) {
#pragma unroll
        for (int b = 0; b < PV_ITEMS; b += 4) {
            ulonglong2 v0[4], v1[4];
            unsigned long long hs[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int t = b + u;
                e_p[t] = 0.0;
                hs[u] = 0ull;
                v0[u] = make_ulonglong2(0ull, 0ull);
                v1[u] = v0[u];
                if (t < cnt && e_k[t] != 0) {
                    hs[u] = memo_hash((double)e_k[t] * sample_scale, e_lam[t]);
                    const ulonglong2* e = reinterpret_cast<const ulonglong2*>(
                        &memo.e[(unsigned)(hs[u] >> 20) & memo.mask]);
                    v0[u] = e[0];                     // tag, lambda
                    v1[u] = e[1];                     // p, check
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int t = b + u;
                if (t < cnt && e_k[t] != 0) {
                    const double count = (double)e_k[t] * sample_scale;
                    const double l = __longlong_as_double((long long)v0[u].y);
                    const double v = __longlong_as_double((long long)v1[u].x);
                    if (v0[u].x == (hs[u] | 3ull) && l == e_lam[t] && v1[u].y == memo_check(count, e_lam[t], v))
                        e_p[t] = v;
                    else
                        miss |= 1u << t;
                }
            }
        }
    } else {
#pragma unroll
        for (int t = 0; t < PV_ITEMS; ++t) {
            e_p[t] = 0.0;
            if (t < cnt && e_k[t] != 0) miss |= 1u << t;       // (small inputs: no memo)
        }
    }
    unsigned q0 = 0;
    if (miss) {
        q0 = atomicAdd(&s_nq, (unsigned)__popc(miss));
        unsigned q = q0;
#pragma unroll
        for (int t = 0; t < PV_ITEMS; ++t)
            if ((miss >> t) & 1u) { a_val[q] = e_k[t]; b_val[q] = e_lam[t]; ++q; }
    }
    __syncthreads();
    // slow path: the block drains the queue; L2 probe first (pairs published meanwhile)
    const unsigned nq = s_nq;
    for (unsigned q = threadIdx.x; q < nq; q += PV_THREADS) {
        double count = (double)a_val[q] * sample_scale, lam = b_val[q], p;
        unsigned long long h = memo_hash(count, lam);
        if (!(memo.e && memo_find<true>(memo, h, count, lam, p))) {
            p = neglog10_poisson_sf(count, lam);
            if (memo.e) memo_publish(memo, h, count, lam, p);
        }
        b_val[q] = p;
    }
    __syncthreads();
    if (miss) {
        unsigned q = q0;
#pragma unroll
        for (int t = 0; t < PV_ITEMS; ++t)
            if ((miss >> t) & 1u) e_p[t] = b_val[q++];
    }
    {
        // (the miss queue is drained: a_pos / b_val are free; readers of b_val passed the
        //  barrier above only if there were misses, so one more barrier before reuse)
        __syncthreads();
#pragma unroll
        for (int t = 0; t < PV_ITEMS; ++t)
            if (t < cnt) { a_pos[o + t] = e_pos[t]; b_val[o + t] = e_p[t]; }
        __syncthreads();
        const unsigned long long w0 = s_obase;
        for (unsigned q = threadIdx.x; q < (unsigned)tile_out; q += PV_THREADS) {
            out_pos[w0 + q] = a_pos[q];
            out_p[w0 + q] = b_val[q];
        }
    }
}

// Step 2: canonical runs -- entries whose p equals the previous position's p
// are dropped (SparseDiffs keeps only positions where the value changes).
struct CanonOp {
    const uint32_t* in_pos;
    const double* in_p;
    uint32_t* out_pos;
    double* out_p;
    __device__ __forceinline__ bool keep(unsigned i) const { return i == 0 || in_p[i] != in_p[i - 1]; }
    __device__ __forceinline__ void emit(unsigned i, unsigned long long w) const {
        out_pos[w] = in_pos[i];
        out_p[w] = in_p[i];
    }
    __device__ __forceinline__ void finish(unsigned long long) const {}
};

// ---- (p -> base pairs) table --------------------------------------------------

__device__ __forceinline__ unsigned long long hash64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

// Adds `len` base pairs to key in the global table (linear probing).
__device__ __forceinline__ void ptable_global_add(unsigned long long key, unsigned long long len,
                                                  unsigned long long h,
                                                  unsigned long long* __restrict__ tkeys,
                                                  unsigned long long* __restrict__ tcount,
                                                  unsigned mask, int* __restrict__ fail) {
    unsigned s = (unsigned)(h >> 16) & mask;
    for (unsigned probe = 0; probe < 256; ++probe) {
        unsigned long long cur = tkeys[s];
        if (cur == 0) {
            unsigned long long old = atomicCAS(&tkeys[s], 0ull, key);
            cur = old == 0 ? key : old;
        }
        if (cur == key) {
            atomicAdd(&tcount[s], len);
            return;
        }
        s = (s + 1) & mask;
    }
    atomicExch(fail, 1);      // table too full: the host grows it and starts over
}

// Neighbouring runs never share a p (the track is canonical) but a stretch of
// a few thousand runs holds few distinct values: each block first aggregates
// its PT_TILE runs in a shared-memory table and then adds one entry per
// distinct value to the global table.  (The first version combined equal keys
// inside a warp with match_any + 32 shuffles per run and was MIO bound.)
constexpr int PT_THREADS = 256;
constexpr int PT_ITEMS = 8;
constexpr int PT_TILE = PT_THREADS * PT_ITEMS;
constexpr int PT_SLOTS = 2048;

__global__ void __launch_bounds__(PT_THREADS)
ptable_insert_kernel(const uint32_t* __restrict__ pos, const double* __restrict__ p, unsigned n,
                     uint32_t track_size, unsigned long long* __restrict__ tkeys,
                     unsigned long long* __restrict__ tcount, unsigned mask,
                     int* __restrict__ fail) {
    __shared__ unsigned long long s_key[PT_SLOTS];
    __shared__ unsigned s_cnt[PT_SLOTS];     // a tile's lengths sum to less than the track size (< 2^32): native 32-bit atomics
    __shared__ int s_failed;
    for (int i = threadIdx.x; i < PT_SLOTS; i += PT_THREADS) { s_key[i] = 0; s_cnt[i] = 0; }
    if (threadIdx.x == 0) s_failed = *reinterpret_cast<volatile int*>(fail);
    __syncthreads();
    if (s_failed) return;                    // an earlier block gave up already (uniform exit)
    const unsigned base = blockIdx.x * PT_TILE;
#pragma unroll
    for (int j = 0; j < PT_ITEMS; ++j) {
        unsigned i = base + j * PT_THREADS + threadIdx.x;
        if (i >= n) break;
        double v = p[i];
        if (v == 0.0) continue;
        unsigned long long key = (unsigned long long)__double_as_longlong(v);
        uint32_t next = (i + 1 < n) ? pos[i + 1] : track_size;
        unsigned len = next - pos[i];
        unsigned long long h = hash64(key);
        unsigned s = (unsigned)h & (PT_SLOTS - 1);
        bool done = false;
        for (int probe = 0; probe < 8 && !done; ++probe) {
            unsigned long long cur = s_key[s];
            if (cur == 0) {
                unsigned long long old = atomicCAS(&s_key[s], 0ull, key);
                cur = old == 0 ? key : old;
            }
            if (cur == key) {
                atomicAdd(&s_cnt[s], len);
                done = true;
            }
            s = (s + 1) & (PT_SLOTS - 1);
        }
        if (!done) ptable_global_add(key, len, h, tkeys, tcount, mask, fail);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PT_SLOTS; i += PT_THREADS) {
        unsigned long long key = s_key[i];
        if (key) ptable_global_add(key, s_cnt[i], hash64(key), tkeys, tcount, mask, fail);
    }
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

// Running base-pair count in descending-p order, then per distinct p:
//   q = p + log10(1 + bp with larger p) - log10(N), first clamped at 0, running minimum
// (sparsepvalues.py:95-103).  Writes the distinct p (descending) and their q.
//
// Three launches over tiles of 2048 sorted entries, tiles independent of each other (two-level
// tile totals, scan.cuh): counts and group ends per tile; q of every group end + tile minima;
// running minimum + compaction.  The first version was ONE block walking the table in chunks
// (0.10 ms for the 35 k entries of one chr22-scale chromosome -- and 0.80 ms on every rank for
// the joined table of eight of them, the largest kernel of the 8-GPU pass).
constexpr int QT_THREADS = 256;
constexpr int QT_ITEMS = 8;               // consecutive entries per thread (their loads overlap)
constexpr int QT_TILE = QT_THREADS * QT_ITEMS;

struct QtScratch {
    unsigned long long* cnt_sums;   unsigned long long* cnt_supers;    // base pairs per tile / group of tiles
    unsigned long long* end_sums;   unsigned long long* end_supers;    // distinct p per tile / group
    unsigned long long* min_tile;   unsigned long long* min_super;     // ordered key of the smallest q
    double* q;                                                          // q at group ends, +inf elsewhere
};

__device__ __forceinline__ unsigned long long qt_min_lookup(const unsigned long long* __restrict__ tiles,
                                                            const unsigned long long* __restrict__ supers,
                                                            unsigned tile, unsigned long long* s_red) {
    const unsigned group = tile >> RTS_GROUP_LOG2, first = group << RTS_GROUP_LOG2;
    unsigned long long v = ~0ull;
    for (unsigned i = threadIdx.x; i < group; i += blockDim.x) v = min(v, supers[i]);
    for (unsigned i = first + threadIdx.x; i < tile; i += blockDim.x) v = min(v, tiles[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = min(v, __shfl_xor_sync(FULL, v, d));
    if (lane_id() == 0) s_red[warp_id()] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = ~0ull;
        for (unsigned w = 0; w < (blockDim.x >> 5); ++w) t = min(t, s_red[w]);
        s_red[32] = t;
    }
    __syncthreads();
    const unsigned long long r = s_red[32];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(QT_THREADS)
qtable_count_kernel(const unsigned long long* __restrict__ sorted_keys,
                    const uint32_t* __restrict__ sorted_idx,
                    const unsigned long long* __restrict__ counts, unsigned n, QtScratch sc) {
    __shared__ unsigned long long s_red[2][QT_THREADS / 32];
    const unsigned i0 = blockIdx.x * QT_TILE + threadIdx.x * QT_ITEMS;
    unsigned long long key[QT_ITEMS + 1];
#pragma unroll
    for (int j = 0; j <= QT_ITEMS; ++j) key[j] = (i0 + j < n) ? sorted_keys[i0 + j] : 0ull;
    unsigned long long tc = 0, ends = 0;
#pragma unroll
    for (int j = 0; j < QT_ITEMS; ++j) {
        const unsigned i = i0 + j;
        if (i < n) {
            tc += counts[sorted_idx[i]];
            ends += (i + 1 >= n || key[j + 1] != key[j]) ? 1u : 0u;
        }
    }
    tc = warp_sum(tc);
    ends = warp_sum(ends);
    if (lane_id() == 0) { s_red[0][warp_id()] = tc; s_red[1][warp_id()] = ends; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long a = 0, b = 0;
#pragma unroll
        for (int w = 0; w < QT_THREADS / 32; ++w) { a += s_red[0][w]; b += s_red[1][w]; }
        tile_total_publish(sc.cnt_sums, sc.cnt_supers, blockIdx.x, a);
        tile_total_publish(sc.end_sums, sc.end_supers, blockIdx.x, b);
    }
}

__global__ void __launch_bounds__(QT_THREADS)
qtable_q_kernel(const unsigned long long* __restrict__ sorted_keys,
                const uint32_t* __restrict__ sorted_idx,
                const unsigned long long* __restrict__ counts, unsigned n, double log_n, QtScratch sc) {
    __shared__ unsigned long long s_cnt[33];
    __shared__ unsigned long long s_red[33];
    const unsigned i0 = blockIdx.x * QT_TILE + threadIdx.x * QT_ITEMS;
    unsigned long long key[QT_ITEMS + 1], c[QT_ITEMS];
#pragma unroll
    for (int j = 0; j <= QT_ITEMS; ++j) key[j] = (i0 + j < n) ? sorted_keys[i0 + j] : 0ull;
    unsigned long long tc = 0;
#pragma unroll
    for (int j = 0; j < QT_ITEMS; ++j) {
        c[j] = (i0 + j < n) ? counts[sorted_idx[i0 + j]] : 0ull;
        tc += c[j];
    }
    const unsigned long long before_tile = tile_prefix_lookup(sc.cnt_sums, sc.cnt_supers, blockIdx.x, s_red);
    unsigned long long tot_c;
    unsigned long long run = before_tile + block_excl_sum<unsigned long long>(tc, s_cnt, tot_c);
    // q at the last entry of every distinct p; bp with strictly larger p =
    // running count before the group's first entry (groups are tiny: walk back)
    double tmin = INFINITY;
#pragma unroll
    for (int j = 0; j < QT_ITEMS; ++j) {
        const unsigned i = i0 + j;
        if (i < n) {
            double q = INFINITY;
            if (i + 1 >= n || key[j + 1] != key[j]) {
                unsigned long long before = run;            // count before entry i
                unsigned g = i;
                while (g > 0 && sorted_keys[g - 1] == key[j]) {
                    --g;
                    before -= counts[sorted_idx[g]];
                }
                q = key_to_f64(~key[j]) + log10(1.0 + (double)before) - log_n;
                if (g == 0) q = fmax(0.0, q);
            }
            sc.q[i] = q;
            tmin = fmin(tmin, q);
        }
        run += c[j];
    }
    unsigned long long kmin = f64_to_key(tmin);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) kmin = min(kmin, __shfl_xor_sync(FULL, kmin, d));
    if (lane_id() == 0) s_red[warp_id()] = kmin;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = ~0ull;
#pragma unroll
        for (int w = 0; w < QT_THREADS / 32; ++w) t = min(t, s_red[w]);
        sc.min_tile[blockIdx.x] = t;
        atomicMin(sc.min_super + (blockIdx.x >> RTS_GROUP_LOG2), t);
    }
}

__global__ void __launch_bounds__(QT_THREADS)
qtable_emit_kernel(const unsigned long long* __restrict__ sorted_keys, unsigned n, QtScratch sc,
                   double* __restrict__ out_p, double* __restrict__ out_q,
                   unsigned long long* __restrict__ n_out) {
    __shared__ unsigned long long s_red[33];
    __shared__ int s_int[33];
    __shared__ double s_min[QT_THREADS / 32];
    const unsigned i0 = blockIdx.x * QT_TILE + threadIdx.x * QT_ITEMS;
    const unsigned long long kmin = qt_min_lookup(sc.min_tile, sc.min_super, blockIdx.x, s_red);
    const double carry_min = kmin == ~0ull ? INFINITY : key_to_f64(kmin);     // (all ones: nothing before)
    const unsigned long long carry_out = tile_prefix_lookup(sc.end_sums, sc.end_supers, blockIdx.x, s_red);
    double q[QT_ITEMS];
    unsigned last = 0;
    double tmin = INFINITY;
#pragma unroll
    for (int j = 0; j < QT_ITEMS; ++j) {
        const unsigned i = i0 + j;
        const double v = i < n ? sc.q[i] : INFINITY;
        if (i < n && (i + 1 >= n || sorted_keys[i + 1] != sorted_keys[i])) last |= 1u << j;
        tmin = fmin(tmin, v);
        q[j] = tmin;                                  // inclusive running minimum inside the thread
    }
    double m = tmin;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double o = __shfl_up_sync(FULL, m, d);
        if (lane_id() >= (unsigned)d) m = fmin(m, o);
    }
    if (lane_id() == 31) s_min[warp_id()] = m;
    __syncthreads();
    double pre = carry_min;                           // everything before this thread
    for (unsigned w = 0; w < warp_id(); ++w) pre = fmin(pre, s_min[w]);
    {
        const double up = __shfl_up_sync(FULL, m, 1);
        if (lane_id() > 0) pre = fmin(pre, up);
    }
    int tot_o;
    const int ex_o = block_excl_sum<int>(__popc(last), s_int, tot_o);
    unsigned long long w_out = carry_out + (unsigned long long)ex_o;
#pragma unroll
    for (int j = 0; j < QT_ITEMS; ++j)
        if ((last >> j) & 1u) {
            out_p[w_out] = key_to_f64(~sorted_keys[i0 + j]);
            out_q[w_out] = fmin(pre, q[j]);
            ++w_out;
        }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *n_out = carry_out + (unsigned long long)tot_o;
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
struct ThresholdOp {
    const uint32_t* pos;
    const double* q;
    double cutoff;
    uint32_t* out_pos;
    uint8_t* out_val;
    __device__ __forceinline__ bool keep(unsigned i) const {
        return i == 0 || ((q[i] >= cutoff) != (q[i - 1] >= cutoff));
    }
    __device__ __forceinline__ void emit(unsigned i, unsigned long long w) const {
        out_pos[w] = pos[i];
        out_val[w] = q[i] >= cutoff;
    }
    __device__ __forceinline__ void finish(unsigned long long) const {}
};

}  // namespace gpc

using namespace gpc;

extern "C" int gpc_pvalues(const uint32_t* sample_pos, const int32_t* sample_val, uint64_t n_sample,
                           double sample_scale, const uint32_t* control_pos,
                           const double* control_val, uint64_t n_control, double control_scale,
                           uint32_t* out_pos, double* out_p, uint64_t* n_out, void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    *n_out = 0;
    uint64_t total = n_sample + n_control;
    if (total == 0) return GPC_OK;
    if (total >= (1ull << 31)) { set_error("tracks too long"); return GPC_ERR_ARG; }
    unsigned tiles = div_up(total, PV_TILE);
    Scratch sc;
    size_t bytes = (size_t)tiles * 8 + 64;
    GPC_CUDA(sc.alloc(bytes, stream));
    GPC_CUDA(cudaMemsetAsync(sc.p, 0, bytes, stream));
    unsigned long long* count = sc.as<unsigned long long>();      // [0] positions, [1] runs
    unsigned* ticket = reinterpret_cast<unsigned*>(count + 2);
    unsigned long long* status = count + 4;
    // memo table: 2^20 entries of 32 B (stays in L2); small inputs do without
    Scratch memo_buf;
    PMemo memo;
    memo.e = nullptr;
    memo.mask = 0;
    if (total >= (1u << 16)) {
        int lg = 20;
        if (const char* v = getenv("GPC_MEMO_LOG2")) lg = atoi(v);
        lg = lg < 10 ? 10 : (lg > 26 ? 26 : lg);               // 32 KB .. 2 GB of entries
        const size_t cap = (size_t)1 << lg;
        GPC_CUDA(memo_buf.alloc(cap * sizeof(MemoEntry), stream));
        GPC_CUDA(cudaMemsetAsync(memo_buf.p, 0, cap * sizeof(MemoEntry), stream));
        memo.e = memo_buf.as<MemoEntry>();
        memo.mask = (unsigned)(cap - 1);
    }
    Scratch split, tmp_pos, tmp_p;
    GPC_CUDA(split.alloc((size_t)(tiles + 1) * 4, stream));
    GPC_CUDA(tmp_pos.alloc(total * 4, stream));
    GPC_CUDA(tmp_p.alloc(total * 8, stream));
    GPC_PROF("pv_partition");
    pv_partition_kernel<<<div_up(tiles + 1, 256), 256, 0, stream>>>(
        sample_pos, (unsigned)n_sample, control_pos, (unsigned)n_control, tiles, split.as<uint32_t>());
    GPC_LAUNCH_CHECK();
    GPC_PROF_BYTES("pvalues", 8.0 * n_sample + 12.0 * n_control);
    pv_eval_kernel<<<tiles, PV_THREADS, 0, stream>>>(sample_pos, sample_val, (unsigned)n_sample,
                                                    sample_scale, control_pos, control_val,
                                                    (unsigned)n_control, control_scale, memo,
                                                    split.as<uint32_t>(), tmp_pos.as<uint32_t>(),
                                                    tmp_p.as<double>(), status, ticket, count);
    GPC_LAUNCH_CHECK();
    // (the number of positions stays on the device: no host round trip in between)
    GPC_TRY(select_if("pv_canon", CanonOp{tmp_pos.as<uint32_t>(), tmp_p.as<double>(), out_pos, out_p},
                      total, count + 1, stream, count));
    unsigned long long h[2] = {0, 0};
    GPC_READ_BACK({count, h, (unsigned)(16)});
    *n_out = h[1];
    prof_add_bytes("pvalues", 12.0 * (double)h[1]);
    prof_add_bytes("pv_canon", 12.0 * (double)h[0] + 12.0 * (double)h[1]);
    return GPC_OK;
}

extern "C" int gpc_ptable_local(const uint32_t* pos, const double* p, uint64_t n_runs,
                                uint32_t track_size, double* table_p, uint64_t* table_count,
                                uint64_t table_capacity, uint64_t* n_distinct, void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    *n_distinct = 0;
    if (n_runs == 0) return GPC_OK;
    unsigned n = (unsigned)n_runs;
    // tracks hold few distinct p (tens of thousands) when the linear map is all-integer,
    // millions when it is not (dense variation: every node scale its own lambda values):
    // start from the size the last call ended with (a failed size costs a whole insert pass;
    // the dense-variation workload went through four of them per call), grow when the table
    // turns out more than half full
    static std::atomic<unsigned long long> cap_hint{1ull << 17};
    unsigned long long cap = cap_hint.load();
    while (true) {
        Scratch keys, counts, fail, flag, offs, total;
        GPC_CUDA(keys.alloc(cap * 8, stream));
        GPC_CUDA(counts.alloc(cap * 8, stream));
        GPC_CUDA(fail.alloc(8, stream));
        GPC_CUDA(cudaMemsetAsync(keys.p, 0, cap * 8, stream));
        GPC_CUDA(cudaMemsetAsync(counts.p, 0, cap * 8, stream));
        GPC_CUDA(cudaMemsetAsync(fail.p, 0, 8, stream));
        GPC_PROF("ptable_insert");
        ptable_insert_kernel<<<div_up(n, PT_TILE), PT_THREADS, 0, stream>>>(
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
        GPC_READ_BACK({total.p, &h, (unsigned)(8)}, {fail.p, &hfail, (unsigned)(4)});
        if (hfail || h * 2 > cap) {               // too full: grow and redo
            if (cap >= 4ull * n + (1ull << 17)) { set_error("p table overflow"); return GPC_ERR_INTERNAL; }
            cap <<= 3;
            continue;
        }
        *n_distinct = h;
        {   // next call: the smallest power of two that keeps this many keys below a quarter
            unsigned long long next = 1ull << 17;
            while (next < 4 * h) next <<= 1;
            cap_hint.store(next);
        }
        if (h > table_capacity) return GPC_ERR_CAPACITY;
        GPC_PROF("ptable_compact");
        ptable_compact_kernel<<<div_up(cap, 256), 256, 0, stream>>>(
            keys.as<unsigned long long>(), counts.as<unsigned long long>(), offs.as<uint32_t>(),
            (unsigned)cap, table_p, reinterpret_cast<unsigned long long*>(table_count));
        GPC_LAUNCH_CHECK();
        return GPC_OK;                            // (outputs complete in stream order)
    }
}

extern "C" int gpc_qtable_build(const double* table_p, const uint64_t* table_count,
                                uint64_t n_entries, uint64_t total_bp, double* out_p, double* out_q,
                                uint64_t* n_out, void* stream_) {
    GPC_STAGE();
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
    const unsigned long long* sk = in_tmp ? k1.as<unsigned long long>() : k0.as<unsigned long long>();
    const uint32_t* si = in_tmp ? i1.as<uint32_t>() : i0.as<uint32_t>();
    const unsigned long long* tcount = reinterpret_cast<const unsigned long long*>(table_count);
    const unsigned tiles = div_up(n, QT_TILE), groups = (tiles >> RTS_GROUP_LOG2) + 1;
    Scratch qs, qq;
    GPC_CUDA(qs.alloc(((size_t)tiles + groups) * 3 * 8, stream));
    GPC_CUDA(qq.alloc((size_t)n * 8, stream));
    unsigned long long* w = qs.as<unsigned long long>();
    QtScratch sc{w, w + tiles, w + tiles + groups, w + 2 * tiles + groups,
                 w + 2 * (tiles + groups), w + 2 * (tiles + groups) + tiles, qq.as<double>()};
    GPC_CUDA(cudaMemsetAsync(w, 0, ((size_t)tiles + groups) * 2 * 8, stream));
    GPC_CUDA(cudaMemsetAsync(sc.min_tile, 0xff, ((size_t)tiles + groups) * 8, stream));
    GPC_PROF("qtable_count");
    qtable_count_kernel<<<tiles, QT_THREADS, 0, stream>>>(sk, si, tcount, n, sc);
    GPC_LAUNCH_CHECK();
    GPC_PROF("qtable_q");
    qtable_q_kernel<<<tiles, QT_THREADS, 0, stream>>>(sk, si, tcount, n, log10((double)total_bp), sc);
    GPC_LAUNCH_CHECK();
    GPC_PROF("qtable");
    qtable_emit_kernel<<<tiles, QT_THREADS, 0, stream>>>(sk, n, sc, out_p, out_q, cnt.as<unsigned long long>());
    GPC_LAUNCH_CHECK();
    unsigned long long h = 0;
    GPC_READ_BACK({cnt.p, &h, (unsigned)(8)});
    *n_out = h;
    return GPC_OK;
}

// Exact-key lookup through a hash table built per call (16-byte entries {bits of p, q}, load
// <= 1/4, key 0 = empty: the table holds no p = 0).  The binary search it replaces made ~16
// dependent loads per run, the last six or so to a different line in every lane: 0.174 ms for
// the 15.7 M runs of the chr22-scale track, bound by L1 wavefronts.  One probe is one
// scattered 16-byte load: 0.083 ms + 0.010 ms for building the table (A/B on one box,
// profiles/ab_r02_pvalues.txt, call 30).
struct QSlot { unsigned long long key; double q; };

__global__ void qhash_build_kernel(const double* __restrict__ table_p, const double* __restrict__ table_q,
                                   unsigned n_table, QSlot* __restrict__ slots, unsigned mask) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_table) return;
    const unsigned long long key = (unsigned long long)__double_as_longlong(table_p[i]);
    if (key == 0ull) return;                      // (p = 0 is answered without the table)
    unsigned s = (unsigned)(hash64(key) >> 20) & mask;
    while (true) {
        const unsigned long long old = atomicCAS(&slots[s].key, 0ull, key);
        if (old == 0ull || old == key) {          // (equal p from several rows carry the same q)
            slots[s].q = table_q[i];
            return;
        }
        s = (s + 1) & mask;
    }
}

__global__ void qhash_apply_kernel(const double* __restrict__ p, unsigned n, const QSlot* __restrict__ slots,
                                   unsigned mask, double* __restrict__ q, int* __restrict__ missing) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = p[i];
    if (v == 0.0) { q[i] = 0.0; return; }         // d[0] = 0
    const unsigned long long key = (unsigned long long)__double_as_longlong(v);
    unsigned s = (unsigned)(hash64(key) >> 20) & mask;
    for (unsigned probe = 0; probe <= mask; ++probe) {
        const ulonglong2 e = *reinterpret_cast<const ulonglong2*>(&slots[s]);
        if (e.x == key) { q[i] = __longlong_as_double((long long)e.y); return; }
        if (e.x == 0ull) break;
        s = (s + 1) & mask;
    }
    q[i] = 0.0;
    atomicExch(missing, 1);
}

extern "C" int gpc_qvalues_apply(const double* p, uint64_t n_runs, const double* table_p,
                                 const double* table_q, uint64_t n_table, double* out_q,
                                 void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_runs == 0) return GPC_OK;
    Scratch miss;
    GPC_CUDA(miss.alloc(8, stream));
    GPC_CUDA(cudaMemsetAsync(miss.p, 0, 8, stream));
    // small jobs keep the binary search (no table to build); GPC_Q_SEARCH=1 forces it
    static const bool force_search = getenv("GPC_Q_SEARCH") != nullptr;
    // (and tables of more than 2^20 distinct p -- dense variation graphs -- whose hash table
    //  would leave the L2)
    if (n_runs < (1u << 16) || n_table == 0 || n_table > (1ull << 20) || force_search) {
        GPC_PROF("qvalues_apply");
        qvalues_apply_kernel<<<div_up(n_runs, 256), 256, 0, stream>>>(
            p, (unsigned)n_runs, table_p, table_q, (unsigned)n_table, out_q, miss.as<int>());
        GPC_LAUNCH_CHECK();
    } else {
        unsigned long long cap = 1024;
        while (cap < 4 * n_table) cap <<= 1;
        Scratch slots;
        GPC_CUDA(slots.alloc(cap * sizeof(QSlot), stream));
        GPC_CUDA(cudaMemsetAsync(slots.p, 0, cap * sizeof(QSlot), stream));
        GPC_PROF("qhash_build");
        qhash_build_kernel<<<div_up(n_table, 256), 256, 0, stream>>>(
            table_p, table_q, (unsigned)n_table, slots.as<QSlot>(), (unsigned)(cap - 1));
        GPC_LAUNCH_CHECK();
        GPC_PROF("qvalues_apply");
        qhash_apply_kernel<<<div_up(n_runs, 256), 256, 0, stream>>>(
            p, (unsigned)n_runs, slots.as<QSlot>(), (unsigned)(cap - 1), out_q, miss.as<int>());
        GPC_LAUNCH_CHECK();
    }
    int h = 0;
    GPC_READ_BACK({miss.p, &h, (unsigned)(4)});
    if (h) { set_error("p-value without an entry in the p->q table"); return GPC_ERR_ARG; }
    return GPC_OK;
}

extern "C" int gpc_threshold(const uint32_t* pos, const double* q, uint64_t n_runs, double cutoff,
                             uint32_t* out_pos, uint8_t* out_val, uint64_t* n_out, void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    *n_out = 0;
    if (n_runs == 0) return GPC_OK;
    unsigned n = (unsigned)n_runs;
    Scratch total;
    GPC_CUDA(total.alloc(8, stream));
    GPC_TRY(select_if("threshold", ThresholdOp{pos, q, cutoff, out_pos, out_val}, n,
                      total.as<unsigned long long>(), stream));
    unsigned long long h = 0;
    GPC_READ_BACK({total.p, &h, (unsigned)(8)});
    *n_out = h;
    return GPC_OK;
}
