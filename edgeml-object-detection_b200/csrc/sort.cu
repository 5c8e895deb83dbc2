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
template <bool SINGLE>
__global__ void __launch_bounds__(kSortThreads, 1) coop_radix_kernel(const SortJob j) {
    __shared__ ScatterSmem sm;
    __shared__ uint32_t part[kSortSplit][2][kRadix];  // phase B partial sums
    __shared__ uint32_t ws[kSortWarps];
    __shared__ int s_skip;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t b0 = (int64_t)blockIdx.x * j.per;
    const int64_t e = b0 + j.per < j.n ? b0 + j.per : j.n;   // b0 >= n: empty range
    unsigned epoch = 0;

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
