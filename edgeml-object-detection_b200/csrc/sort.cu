// Device primitive of the dataset index build: a stable LSD radix sort that runs
// ALL of its passes in one persistent cooperative kernel (coop.cuh).
//
// The sort is the engine's "sort by (class, confidence)" — done ONCE per dataset
// instead of once per target as lib/metrics.py:101 does (np.argsort(-conf) inside
// every ap_per_class call) — and the regrouping of that order by 32-image batch
// (index.cu).
//
// One CTA per SM owns a contiguous range of the items.  Per 8-bit pass:
//   A  digit histogram of the CTA's range (shared-memory atomics) -> table[cta][digit]
//      -- grid barrier --
//   B  every CTA reads the whole table (148 x 256 words, L2): its digit bases
//      = exclusive scan of the digit totals + counts of the CTAs before it;
//      a pass whose digit is the same for all items is skipped by every CTA
//   C  stable scatter, tile by tile: ballot-ranked digits inside a 32-item step,
//      per-warp digit counters carry ranks across steps, warps and tiles
//      -- grid barrier --
// The digit of a pass is either 8 bits of a 64-bit key or a function of the item's
// value (class / batch of the detection id), so the regrouping sort moves 4-byte
// values only and needs no re-keying kernel.
#include <stdlib.h>

#include <algorithm>
#include <mutex>

#include "coop.cuh"

namespace orie {

namespace {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortThreads = 1024;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortSteps = 8;                          // 32-item steps per warp and tile
constexpr int kSortTile = kSortThreads * kSortSteps;   // items per tile
constexpr int kSortSplit = kSortThreads / kRadix;      // threads per digit in phase B

// class / image of a combined detection id (see SortJob)
__device__ __forceinline__ uint32_t digit_source(const SortJob &j, int kind, uint32_t u) {
    switch (kind) {
    case kDigitClass:
        return (uint32_t)(u < j.split ? __ldg(j.cls_lo + u) : __ldg(j.cls_hi + (u - j.split)));
    default:  // kDigitBatch
        return __ldg(j.img + u) >> 5;
    }
}

__device__ __forceinline__ int item_digit(const SortJob &j, const SortPass &p, uint64_t key, uint32_t u) {
    const uint32_t src = p.kind == kDigitKey ? (uint32_t)(key >> p.shift) : (digit_source(j, p.kind, u) >> p.shift);
    return (int)(src & p.mask);
}

// lanes holding the same digit as this lane: 9 ballots (8 digit bits + validity) instead of MATCH.ANY,
// whose latency grows with the number of distinct values in the warp
__device__ __forceinline__ unsigned digit_peers(int dg, bool ok) {
    unsigned peers = __ballot_sync(kFull, ok);
#pragma unroll
    for (int b = 0; b < kRadixBits; ++b) {
        const bool bit = (dg >> b) & 1;
        const unsigned vote = __ballot_sync(kFull, bit);
        peers &= bit ? vote : ~vote;
    }
    return peers;
}

// One tile of the stable scatter: the calling warp owns items [wbase, wbase + 32 * kSortSteps) (those < e are live),
// already loaded into k / u with digits dg.
struct ScatterSmem {
    uint32_t cnt[kSortWarps][kRadix];      // per-warp digit counters of the current tile
    uint32_t run[kRadix];                  // histogram (A), then running output position per digit (C)
};

__device__ __forceinline__ void scatter_tile(ScatterSmem &sm, const uint64_t (&k)[kSortSteps], const uint32_t (&u)[kSortSteps],
                                             const int (&dg)[kSortSteps], int64_t wbase, int64_t e, bool store_keys,
                                             uint64_t *kout, uint32_t *vout) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kSortWarps * kRadix; i += kSortThreads) (&sm.cnt[0][0])[i] = 0;
    __syncthreads();
    unsigned peers[kSortSteps];
#pragma unroll
    for (int s = 0; s < kSortSteps; ++s) {
        const bool ok = wbase + s * 32 + lane < e;
        peers[s] = digit_peers(dg[s], ok);
        if (ok && lane == (__ffs(peers[s]) - 1)) sm.cnt[warp][dg[s]] += __popc(peers[s]);
        __syncwarp();
    }
    __syncthreads();
    if (tid < kRadix) {
        uint32_t r = sm.run[tid];
#pragma unroll 8
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t c = sm.cnt[w][tid];
            sm.cnt[w][tid] = r;
            r += c;
        }
        sm.run[tid] = r;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < kSortSteps; ++s) {
        const bool ok = wbase + s * 32 + lane < e;
        uint32_t pos = 0;
        if (ok) pos = sm.cnt[warp][dg[s]] + __popc(peers[s] & ((1u << lane) - 1u));
        __syncwarp();
        if (ok) {
            if (store_keys) kout[pos] = k[s];
            vout[pos] = u[s];
            if (lane == (__ffs(peers[s]) - 1)) sm.cnt[warp][dg[s]] += __popc(peers[s]);
        }
        __syncwarp();
    }
    __syncthreads();
}

// SINGLE: the CTA's range fits in one tile; its items are loaded once per pass and stay in registers from the
// histogram to the scatter.
struct LsdSmem {
    ScatterSmem sm;
    uint32_t part[kSortSplit][2][kRadix];  // phase B partial sums
    uint32_t ws[kSortWarps];
    int s_skip;
};

template <bool SINGLE>
__device__ __forceinline__ void lsd_sort_body(const SortJob &j, LsdSmem &S, unsigned &epoch) {
    ScatterSmem &sm = S.sm;
    uint32_t (&part)[kSortSplit][2][kRadix] = S.part;
    uint32_t (&ws)[kSortWarps] = S.ws;
    int &s_skip = S.s_skip;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t b0 = (int64_t)blockIdx.x * j.per;
    const int64_t e = b0 + j.per < j.n ? b0 + j.per : j.n;   // b0 >= n: empty range

    const uint64_t *kin = j.keys0;
    uint64_t *kout = j.keys1;
    const uint32_t *vin = j.vals_in;
    uint32_t *vout = j.vals_b, *vnext = j.vals_a;
    int last_key_pass = -1;
    for (int p = 0; p < j.npass; ++p)
        if (j.pass[p].kind == kDigitKey) last_key_pass = p;

    for (int p = 0; p < j.npass; ++p) {
        const SortPass ps = j.pass[p];
        const bool use_keys = ps.kind == kDigitKey;
        uint32_t *table = j.table + (size_t)(p & 1) * gridDim.x * kRadix;
        uint64_t k[kSortSteps];
        uint32_t u[kSortSteps];
        int dg[kSortSteps];
        auto load_tile = [&](int64_t wbase) {
#pragma unroll
            for (int s = 0; s < kSortSteps; ++s) {
                const int64_t i = wbase + s * 32 + lane;
                const bool ok = i < e;
                k[s] = (ok && use_keys) ? __ldcg(kin + i) : 0ull;
                u[s] = ok ? (vin ? __ldcg(vin + i) : (uint32_t)i) : 0u;
            }
#pragma unroll
            for (int s = 0; s < kSortSteps; ++s) dg[s] = wbase + s * 32 + lane < e ? item_digit(j, ps, k[s], u[s]) : 0;
        };

        // ---- A: histogram of this CTA's range
        if (tid < kRadix) sm.run[tid] = 0;
        if (tid == 0) s_skip = 0;
        __syncthreads();
        if (SINGLE) {
            const int64_t wbase = b0 + (int64_t)warp * (32 * kSortSteps);
            load_tile(wbase);
#pragma unroll
            for (int s = 0; s < kSortSteps; ++s)
                if (wbase + s * 32 + lane < e) atomicAdd(&sm.run[dg[s]], 1u);
        } else {
            for (int64_t i = b0 + tid; i < e; i += kSortThreads) {
                const uint64_t key = use_keys ? __ldcg(kin + i) : 0ull;
                const uint32_t v = vin ? __ldcg(vin + i) : (uint32_t)i;
                atomicAdd(&sm.run[item_digit(j, ps, key, v)], 1u);
            }
        }
        __syncthreads();
        if (tid < kRadix) __stcg(table + (size_t)blockIdx.x * kRadix + tid, sm.run[tid]);
        grid_sync(j.bar, epoch);

        // ---- B: digit bases of this CTA.  All loads of the table are issued before the first use.
        {
            const int d = tid & (kRadix - 1), q = tid / kRadix;
            uint32_t below = 0, tot = 0;
            constexpr int kBatch = 40;      // covers 160 CTAs in one round trip
            for (int base = q; base < (int)gridDim.x; base += kBatch * kSortSplit) {
                uint32_t v[kBatch];
#pragma unroll
                for (int it = 0; it < kBatch; ++it) {
                    const int b = base + it * kSortSplit;
                    v[it] = b < (int)gridDim.x ? __ldcg(table + (size_t)b * kRadix + d) : 0u;
                }
#pragma unroll
                for (int it = 0; it < kBatch; ++it) {
                    tot += v[it];
                    below += base + it * kSortSplit < (int)blockIdx.x ? v[it] : 0u;
                }
            }
            part[q][0][d] = below;
            part[q][1][d] = tot;
            __syncthreads();
            uint32_t my_tot = 0, my_below = 0;
            if (tid < kRadix) {
#pragma unroll
                for (int i = 0; i < kSortSplit; ++i) {
                    my_below += part[i][0][tid];
                    my_tot += part[i][1][tid];
                }
                if ((int64_t)my_tot == j.n) s_skip = 1;      // every item has this digit: the pass is the identity
            }
            const uint32_t excl = block_exclusive_scan<kSortThreads>(my_tot, ws, nullptr);
            if (tid < kRadix) sm.run[tid] = excl + my_below;
            __syncthreads();
        }
        if (s_skip) continue;   // same decision in every CTA (same table); the tables alternate, so no barrier is needed

        // ---- C: stable scatter
        const bool store_keys = use_keys && p < last_key_pass;
        if (SINGLE) {
            scatter_tile(sm, k, u, dg, b0 + (int64_t)warp * (32 * kSortSteps), e, store_keys, kout, vout);
        } else {
            for (int64_t t0 = b0; t0 < e; t0 += kSortTile) {
                const int64_t wbase = t0 + (int64_t)warp * (32 * kSortSteps);
                load_tile(wbase);
                scatter_tile(sm, k, u, dg, wbase, e, store_keys, kout, vout);
            }
        }
        grid_sync(j.bar, epoch);
        if (store_keys) {
            const uint64_t *t = kin;
            kin = kout;
            kout = const_cast<uint64_t *>(t);
        }
        vin = vout;
        {
            uint32_t *t = vout;
            vout = vnext;
            vnext = t;
        }
    }

    // ---- the result always ends in vals_a
    if (vin != j.vals_a)
        for (int64_t i = b0 + tid; i < e; i += kSortThreads) j.vals_a[i] = vin ? __ldcg(vin + i) : (uint32_t)i;

    // ---- optional epilogue: rank_out[v] = number of sorted values < rank_split before position v
    if (j.rank_out) {
        __syncthreads();
        uint32_t c = 0;
        for (int64_t i = b0 + tid; i < e; i += kSortThreads) c += __ldcg(j.vals_a + i) < j.rank_split ? 1u : 0u;
        uint32_t total;
        block_exclusive_scan<kSortThreads>(c, ws, &total);
        uint32_t *table = j.table + (size_t)(j.npass & 1) * gridDim.x * kRadix;
        if (tid == 0) __stcg(table + blockIdx.x, total);
        grid_sync(j.bar, epoch);
        uint32_t before = 0;
        for (int b = tid; b < (int)blockIdx.x; b += kSortThreads) before += __ldcg(table + b);
        uint32_t carry;
        block_exclusive_scan<kSortThreads>(before, ws, &carry);
        for (int64_t t0 = b0; t0 < e; t0 += kSortThreads) {
            const int64_t i = t0 + tid;
            const uint32_t f = (i < e && __ldcg(j.vals_a + i) < j.rank_split) ? 1u : 0u;
            uint32_t tile_total;
            const uint32_t x = block_exclusive_scan<kSortThreads>(f, ws, &tile_total);
            if (i < e) j.rank_out[i] = carry + x;
            carry += tile_total;
        }
    }
}

template <bool SINGLE>
__global__ void __launch_bounds__(kSortThreads, 1) coop_radix_kernel(const SortJob j) {
    __shared__ LsdSmem S;
    unsigned epoch = 0;
    lsd_sort_body<SINGLE>(j, S, epoch);
}

// ----------------------------------------------------------------------------
// bucket sort (see common.cuh: BucketSortJob)
// ----------------------------------------------------------------------------
constexpr int kWarpBucket = 256;         // largest bucket one warp sorts in registers (8 items per lane)
constexpr int kCtaBucket = 8192;         // largest bucket one CTA sorts in shared memory
constexpr int kMaxBuckets = 8192;        // bucket tables kept in shared memory in the last phase
constexpr int kMaxBig = 2048;            // buckets beyond a warp's reach the partition kernel can list
constexpr int kLocalThreads = 256;       // block size of the local-sort kernel

struct Elem {
    uint32_t c;      // class
    uint64_t k;      // key
    uint32_t i;      // id
};
__device__ __forceinline__ bool elem_less(const Elem &a, const Elem &b) {      // branch-free: lanes never diverge
    return (a.c < b.c) | ((a.c == b.c) & ((a.k < b.k) | ((a.k == b.k) & (a.i < b.i))));
}
__device__ __forceinline__ Elem elem_shfl_xor(const Elem &e, int lanemask) {
    Elem o;
    o.c = __shfl_xor_sync(kFull, e.c, lanemask);
    o.k = __shfl_xor_sync(kFull, e.k, lanemask);
    o.i = __shfl_xor_sync(kFull, e.i, lanemask);
    return o;
}

// Bitonic network over 256 elements held by one warp: lane l holds the elements of index 8 l + r, r = 0..7.
// Compare distances below 8 stay inside a thread, the others are one shuffle per element.
template <int K, int J>
__device__ __forceinline__ void warp_bitonic_stage(Elem (&e)[8], int lane) {
    if (J >= 8) {
        constexpr int lj = J >> 3;
        const bool lower = (lane & lj) == 0;             // this lane holds the lower index of each pair
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const bool up = (((lane << 3) | r) & K) == 0;
            const Elem o = elem_shfl_xor(e[r], lj);
            // elements are distinct (ids are), except padding, which is all alike: taking "o < mine" for the minimum
            // and "not (o < mine)" for the maximum is then a total rule
            const bool o_less = elem_less(o, e[r]);
            if (o_less == (lower == up)) e[r] = o;
        }
    } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (r & J) continue;
            const bool up = (((lane << 3) | r) & K) == 0;
            Elem &a = e[r], &b = e[r | J];
            if (elem_less(b, a) == up) { const Elem t = a; a = b; b = t; }
        }
    }
}
template <int K, int J>
struct WarpBitonicInner {
    static __device__ __forceinline__ void run(Elem (&e)[8], int lane) {
        warp_bitonic_stage<K, J>(e, lane);
        WarpBitonicInner<K, J / 2>::run(e, lane);
    }
};
template <int K>
struct WarpBitonicInner<K, 0> {
    static __device__ __forceinline__ void run(Elem (&)[8], int) {}
};
template <int K>
struct WarpBitonicOuter {
    static __device__ __forceinline__ void run(Elem (&e)[8], int lane) {
        WarpBitonicOuter<K / 2>::run(e, lane);
        WarpBitonicInner<K, K / 2>::run(e, lane);
    }
};
template <>
struct WarpBitonicOuter<1> {
    static __device__ __forceinline__ void run(Elem (&)[8], int) {}
};

struct BucketSmem {                      // local-sort kernel
    uint64_t key[kCtaBucket];
    uint32_t id[kCtaBucket];
    uint32_t cls[kCtaBucket];
    uint32_t ws[kLocalThreads / 32];
};

__device__ __forceinline__ uint32_t class_of(const SortJob &j, uint32_t u, int64_t C) {
    const int32_t c = u < j.split ? __ldg(j.cls_lo + u) : __ldg(j.cls_hi + (u - j.split));
    return (uint32_t)(c < 0 ? 0 : (c >= C ? C - 1 : c));      // out-of-range ids are flagged by prep_kernel; stay in bounds
}

// Phases 1-3 (binning, bucket layout, scatter) and the bucket tables the local sorts need; falls back to the radix
// passes when a bucket is too large for a CTA.
__global__ void __launch_bounds__(kSortThreads, 1) bucket_partition_kernel(const BucketSortJob q) {
    __shared__ LsdSmem L;
    __shared__ uint32_t s_flag, s_nbig;
    uint32_t (&ws)[kSortWarps] = L.ws;
    const SortJob &j = q.lsd;
    const int tid = threadIdx.x;
    const int64_t n = q.n;
    const int64_t b0 = min((int64_t)blockIdx.x * j.per, n), e = min(b0 + j.per, n);
    unsigned epoch = 0;

    // the key bits that vary across the dataset: the leading common bits carry no order
    const uint64_t diff = q.key_and_or[0] ^ q.key_and_or[1];
    const int nbits = diff ? 64 - __clzll((long long)diff) : 0;
    const int kb = nbits < q.bits ? nbits : q.bits;
    const int shift = nbits - kb;
    const uint32_t kmask = (1u << kb) - 1u;
    const uint32_t cap = (uint32_t)q.cap;

    // ---- 1: histogram over (class, leading varying key bits)
    for (int64_t i = b0 + tid; i < e; i += kSortThreads) {
        const uint32_t bin = (class_of(j, (uint32_t)i, q.C) << q.bits) | ((uint32_t)(j.keys0[i] >> shift) & kmask);
        atomicAdd(q.binhist + bin, 1u);
    }
    grid_sync(j.bar, epoch);

    // ---- 2: every CTA owns a contiguous slice of the bins: slice totals, then the position of every bin in the
    //         sorted order; bins are packed into buckets by position / cap (a bucket = whole bins, so at most
    //         cap - 1 + its last bin's count items)
    const int64_t TB = q.C << q.bits;
    const int64_t slice = (TB + gridDim.x - 1) / gridDim.x;
    const int64_t s0 = min((int64_t)blockIdx.x * slice, TB), s1 = min(s0 + slice, TB);
    {
        uint32_t mine = 0;
        for (int64_t b = s0 + tid; b < s1; b += kSortThreads) mine += __ldcg(q.binhist + b);
        uint32_t total;
        block_exclusive_scan<kSortThreads>(mine, ws, &total);
        if (tid == 0) __stcg(q.slice_sum + blockIdx.x, total);
    }
    grid_sync(j.bar, epoch);
    {
        uint32_t before = 0;
        for (int b = tid; b < (int)blockIdx.x; b += kSortThreads) before += __ldcg(q.slice_sum + b);
        uint32_t carry;
        block_exclusive_scan<kSortThreads>(before, ws, &carry);
        for (int64_t t0 = s0; t0 < s1; t0 += kSortThreads) {
            const int64_t b = t0 + tid;
            const uint32_t cnt = b < s1 ? __ldcg(q.binhist + b) : 0u;
            uint32_t tile_total;
            const uint32_t pre = carry + block_exclusive_scan<kSortThreads>(cnt, ws, &tile_total);
            carry += tile_total;
            if (cnt) {
                const uint32_t bucket = pre / cap;
                atomicMax(q.bucket_start + bucket, ~pre);          // start of a bucket = lowest position of its bins
                __stcg(q.binhist + b, bucket);                    // the table now maps bin -> bucket
            }
        }
    }
    grid_sync(j.bar, epoch);

    // ---- 3: scatter every item to its bucket (order inside a bucket does not matter: the bucket is sorted next)
    for (int64_t i = b0 + tid; i < e; i += kSortThreads) {
        const uint64_t key = j.keys0[i];
        const uint32_t bin = (class_of(j, (uint32_t)i, q.C) << q.bits) | ((uint32_t)(key >> shift) & kmask);
        const uint32_t bucket = __ldcg(q.binhist + bin);
        const uint32_t pos = ~__ldcg(q.bucket_start + bucket) + atomicAdd(q.bucket_fill + bucket, 1u);
        q.keys_part[pos] = key;
        q.vals_part[pos] = (uint32_t)i;
        if ((uint32_t)i < j.rank_split) atomicAdd(q.bucket_weak + bucket, 1u);
    }
    grid_sync(j.bar, epoch);

    // ---- 4: bucket tables for the local sorts.  Every CTA checks the sizes (they all see the same numbers, so
    //         they all take the same decision); CTA 0 writes the tables: ids < split in front of each bucket and
    //         the list of buckets beyond a warp's reach.
    const int NB = (int)q.nbuckets;
    if (tid == 0) { s_flag = 0; s_nbig = 0; }
    __syncthreads();
    {
        uint32_t carry = 0;
        for (int t0 = 0; t0 < NB; t0 += kSortThreads) {
            const int k = t0 + tid;
            const uint32_t w = k < NB ? __ldcg(q.bucket_weak + k) : 0u;
            const uint32_t sz = k < NB ? __ldcg(q.bucket_fill + k) : 0u;
            if (sz > (uint32_t)kCtaBucket) s_flag = 1;
            if (sz > (uint32_t)kWarpBucket) {
                const uint32_t at = atomicAdd(&s_nbig, 1u);
                if (at >= (uint32_t)kMaxBig) s_flag = 1;
                else if (blockIdx.x == 0) q.big[at] = (uint32_t)k;
            }
            uint32_t tile_total;
            const uint32_t x = block_exclusive_scan<kSortThreads>(w, ws, &tile_total);
            if (k < NB && blockIdx.x == 0) q.wbefore[k] = carry + x;
            carry += tile_total;
        }
        __syncthreads();
        if (blockIdx.x == 0 && tid == 0) { q.ctl[0] = s_flag; q.ctl[1] = s_flag ? 0u : s_nbig; }
    }
    if (s_flag) {
        // a bucket beyond what a CTA sorts (heavy exact ties): sort from scratch with the radix passes
        __syncthreads();
        if (j.per <= kSortTile) lsd_sort_body<true>(j, L, epoch);
        else lsd_sort_body<false>(j, L, epoch);
    }
}

// Local sorts: one warp per bucket of up to 256 items (registers + shuffles), then the listed larger buckets by
// whole CTAs in shared memory.  Writes the sorted ids and, for every position, the number of ids < split in front.
__global__ void __launch_bounds__(kLocalThreads) bucket_local_kernel(const BucketSortJob q) {
    const SortJob &j = q.lsd;
    if (q.ctl[0]) return;                    // the partition kernel fell back to the radix passes
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NB = (int)q.nbuckets;
    const int gw = blockIdx.x * (kLocalThreads / 32) + warp, nw = gridDim.x * (kLocalThreads / 32);
    for (int k = gw; k < NB; k += nw) {
        const int sz = (int)q.bucket_fill[k];
        if (sz == 0 || sz > kWarpBucket) continue;
        const uint32_t start = ~q.bucket_start[k];
        Elem el[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int idx = (lane << 3) | r;
            if (idx < sz) {
                el[r].k = q.keys_part[start + idx];
                el[r].i = q.vals_part[start + idx];
                el[r].c = class_of(j, el[r].i, q.C);
            } else {
                el[r].k = ~0ull; el[r].i = 0xffffffffu; el[r].c = 0xffffffffu;      // padding sorts last
            }
        }
        WarpBitonicOuter<kWarpBucket>::run(el, lane);
        uint32_t mine = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) mine += el[r].i < j.rank_split ? 1u : 0u;
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += y;
        }
        uint32_t run = q.wbefore[k] + incl - mine;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int idx = (lane << 3) | r;
            if (idx < sz) {
                j.vals_a[start + idx] = el[r].i;
                if (j.rank_out) j.rank_out[start + idx] = run;
            }
            run += el[r].i < j.rank_split ? 1u : 0u;
        }
    }
}

// The listed larger buckets (rare: a bin with hundreds of items, i.e. many exact ties), one CTA per bucket in shared
// memory.  A separate launch so that its 128 KB of shared memory do not limit the occupancy of the warp sorts.
__global__ void __launch_bounds__(kLocalThreads) bucket_local_cta_kernel(const BucketSortJob q) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BucketSmem &S = *reinterpret_cast<BucketSmem *>(smem_raw);
    const SortJob &j = q.lsd;
    if (q.ctl[0]) return;
    const int tid = threadIdx.x;
    const int nbig = (int)q.ctl[1];
    for (int b = blockIdx.x; b < nbig; b += gridDim.x) {
        const int k = (int)q.big[b];
        const int sz = (int)q.bucket_fill[k];
        const uint32_t start = ~q.bucket_start[k];
        int P = 512;
        while (P < sz) P <<= 1;
        __syncthreads();
        for (int idx = tid; idx < P; idx += kLocalThreads) {
            if (idx < sz) {
                const uint32_t id = q.vals_part[start + idx];
                S.key[idx] = q.keys_part[start + idx];
                S.id[idx] = id;
                S.cls[idx] = class_of(j, id, q.C);
            } else {
                S.key[idx] = ~0ull; S.id[idx] = 0xffffffffu; S.cls[idx] = 0xffffffffu;
            }
        }
        __syncthreads();
        for (int kk = 2; kk <= P; kk <<= 1) {
            for (int jj = kk >> 1; jj >= 1; jj >>= 1) {
                for (int t = tid; t < P / 2; t += kLocalThreads) {
                    const int lo = ((t & ~(jj - 1)) << 1) | (t & (jj - 1)), hi = lo | jj;
                    const bool up = (lo & kk) == 0;
                    Elem a{S.cls[lo], S.key[lo], S.id[lo]}, c{S.cls[hi], S.key[hi], S.id[hi]};
                    if (elem_less(c, a) == up) {
                        S.cls[lo] = c.c; S.key[lo] = c.k; S.id[lo] = c.i;
                        S.cls[hi] = a.c; S.key[hi] = a.k; S.id[hi] = a.i;
                    }
                }
                __syncthreads();
            }
        }
        uint32_t carry = q.wbefore[k];
        for (int t0 = 0; t0 < sz; t0 += kLocalThreads) {
            const int idx = t0 + tid;
            const uint32_t id = idx < sz ? S.id[idx] : 0xffffffffu;
            const uint32_t f = id < j.rank_split ? 1u : 0u;
            uint32_t tile_total;
            const uint32_t x = block_exclusive_scan<kLocalThreads>(f, S.ws, &tile_total);
            if (idx < sz) {
                j.vals_a[start + idx] = id;
                if (j.rank_out) j.rank_out[start + idx] = carry + x;
            }
            carry += tile_total;
        }
    }
}

}  // namespace

size_t sort_scratch_bytes(int max_blocks) { return 2 * (size_t)max_blocks * kRadix * 4 + 256; }

int sort_max_blocks(int *out) {
    int a = 0, b = 0;
    ORIE_TRY(coop_max_blocks(coop_radix_kernel<true>, kSortThreads, 0, &a));
    ORIE_TRY(coop_max_blocks(coop_radix_kernel<false>, kSortThreads, 0, &b));
    *out = a < b ? a : b;
    return ORIE_OK;
}

int sort_add_passes(SortJob *job, int kind, int bit_lo, int bit_hi) {
    for (int lo = bit_lo; lo < bit_hi; lo += kRadixBits) {
        if (job->npass >= kMaxSortPasses) {
            set_error("radix sort: more than %d passes", kMaxSortPasses);
            return ORIE_ELIMIT;
        }
        const int bits = bit_hi - lo < kRadixBits ? bit_hi - lo : kRadixBits;
        job->pass[job->npass++] = SortPass{kind, lo, (1u << bits) - 1u};
    }
    return ORIE_OK;
}


// ---------------------------------------------------------------------------- bucket sort, host side
bool bucket_sort_applicable(int64_t n, int64_t C) {
    // the (class, key-bits) table must stay small: at least 4 key bits per class inside 2^18 bins
    return n > 0 && n < ((int64_t)1 << 31) && C >= 1 && (C << 4) <= ((int64_t)1 << 18);
}

static int bucket_bits(int64_t C) {
    int bits = 12;
    while (bits > 4 && (C << bits) > ((int64_t)1 << 18)) --bits;
    return bits;
}
static int bucket_cap(int64_t n) {
    // three quarters of what a warp sorts, so that a bucket (whole bins) usually stays within one warp's 256 items;
    // larger on big inputs so that the bucket tables fit the last phase's shared memory
    return (int)std::max<int64_t>(192, ceil_div(n, kMaxBuckets - 192));
}

size_t bucket_sort_scratch_bytes(int64_t n, int64_t C, int max_blocks) {
    const int64_t nb = (std::max<int64_t>(n, 1) - 1) / bucket_cap(n) + 1;
    const size_t lsd = sort_scratch_bytes(max_blocks);
    return lsd + (size_t)round_up(((C << bucket_bits(C)) + 4 * nb + max_blocks + kMaxBig + 2) * 4, 256);
}

int bucket_sort_max_blocks(int *out) {
    static std::mutex mu;
    static bool attr_done[64] = {};
    int dev = 0;
    ORIE_CUDA(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(mu);
        if (dev >= 0 && dev < 64 && !attr_done[dev]) {
            ORIE_CUDA(cudaFuncSetAttribute(bucket_local_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BucketSmem)));
            attr_done[dev] = true;
        }
    }
    return coop_max_blocks(bucket_partition_kernel, kSortThreads, 0, out);
}

int bucket_sort_run(BucketSortJob job, int64_t C, int max_blocks, void *scratch, cudaStream_t st) {
    if (job.n <= 0) return ORIE_OK;
    if (!bucket_sort_applicable(job.n, C)) {
        set_error("bucket sort: not applicable to n=%lld, C=%lld", (long long)job.n, (long long)C);
        return ORIE_EINVAL;
    }
    job.C = C;
    job.bits = bucket_bits(C);
    job.cap = bucket_cap(job.n);
    job.nbuckets = (job.n - 1) / job.cap + 1;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(max_blocks, ceil_div(job.n, kSortTile / 4)));
    job.lsd.n = job.n;
    job.lsd.per = round_up(ceil_div(job.n, blocks), 32);
    job.lsd.bar = (unsigned *)scratch;
    job.lsd.table = (uint32_t *)((char *)scratch + 256);
    uint32_t *extra = (uint32_t *)((char *)scratch + sort_scratch_bytes(max_blocks));
    job.binhist = extra;
    job.bucket_start = job.binhist + (C << job.bits);
    job.bucket_fill = job.bucket_start + job.nbuckets;
    job.bucket_weak = job.bucket_fill + job.nbuckets;
    job.wbefore = job.bucket_weak + job.nbuckets;
    job.slice_sum = job.wbefore + job.nbuckets;
    job.big = job.slice_sum + max_blocks;
    job.ctl = job.big + kMaxBig;
    job.keys_part = job.lsd.keys1;
    job.vals_part = job.lsd.vals_b;
    ORIE_CUDA(cudaMemsetAsync(job.lsd.bar, 0, 4, st));
    ORIE_CUDA(cudaMemsetAsync(extra, 0, (size_t)((C << job.bits) + 3 * job.nbuckets) * 4, st));
    void *args[] = {&job};
    ORIE_CUDA(cudaLaunchCooperativeKernel((const void *)bucket_partition_kernel, dim3(blocks), dim3(kSortThreads), args, 0, st));
    ORIE_LAUNCH_CHECK();
    // one warp per bucket, at most a few waves of blocks
    const unsigned local_blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(job.nbuckets, kLocalThreads / 32), 148 * 16));
    bucket_local_kernel<<<local_blocks, kLocalThreads, 0, st>>>(job);
    bucket_local_cta_kernel<<<148, kLocalThreads, sizeof(BucketSmem), st>>>(job);
    ORIE_LAUNCH_CHECK_N(2);
    return ORIE_OK;
}

int sort_run(SortJob job, int max_blocks, void *scratch, cudaStream_t st) {
    if (job.n <= 0) return ORIE_OK;
    if (job.n >= (int64_t)1 << 31) {
        set_error("radix sort: n=%lld exceeds 2^31-1", (long long)job.n);
        return ORIE_ELIMIT;
    }
    // a quarter tile per CTA at least; one CTA per SM at most
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(max_blocks, ceil_div(job.n, kSortTile / 4)));
    job.per = round_up(ceil_div(job.n, blocks), 32);
    const bool single = job.per <= kSortTile;
    job.bar = (unsigned *)scratch;
    job.table = (uint32_t *)((char *)scratch + 256);
    ORIE_CUDA(cudaMemsetAsync(job.bar, 0, 4, st));
    void *args[] = {&job};
    ORIE_CUDA(cudaLaunchCooperativeKernel(single ? (const void *)coop_radix_kernel<true> : (const void *)coop_radix_kernel<false>,
                                          dim3(blocks), dim3(kSortThreads), args, 0, st));
    ORIE_LAUNCH_CHECK();
    return ORIE_OK;
}

}  // namespace orie
