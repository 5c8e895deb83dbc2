// Device primitives: stable LSD radix sort of (u64 key, u32 value) pairs and an
// exclusive u32 scan.  Hand-written for sm_100a; used by the one-off dataset
// index build (index.cu).  The sort is the engine's "sort by (class,
// confidence)" — done ONCE per dataset instead of once per target as
// lib/metrics.py:101 does (np.argsort(-conf) inside every ap_per_class call).
#include "common.cuh"

namespace orie {

// ----------------------------------------------------------------------------
// exclusive scan
// ----------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096

__global__ void __launch_bounds__(kScanThreads)
scan_tile_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int64_t n,
                 uint32_t *__restrict__ tile_sums, uint32_t *__restrict__ total) {
    __shared__ uint32_t warp_sums[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        int64_t k = base + i;
        v[i] = k < n ? in[k] : 0u;
        sum += v[i];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += y;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < kScanThreads / 32 ? warp_sums[lane] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(kFull, wi, d);
            if (lane >= d) wi += y;
        }
        if (lane < kScanThreads / 32) warp_sums[lane] = wi - w;  // exclusive
        if (lane == kScanThreads / 32 - 1) {
            if (tile_sums) tile_sums[blockIdx.x] = wi;
            if (total && gridDim.x == 1) *total = wi;
        }
    }
    __syncthreads();
    uint32_t run = warp_sums[warp] + (incl - sum);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        int64_t k = base + i;
        if (k < n) out[k] = run;
        run += v[i];
    }
}

__global__ void scan_add_kernel(uint32_t *__restrict__ out, int64_t n, const uint32_t *__restrict__ tile_offs) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] += tile_offs[k / kScanTile];
}

size_t scan_scratch_bytes(int64_t n) {
    size_t bytes = 256;
    while (n > kScanTile) {
        n = ceil_div(n, kScanTile);
        bytes += round_up(n * 4, 256);
    }
    return bytes;
}

int exclusive_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, uint32_t *total, void *scratch, cudaStream_t st) {
    if (n <= 0) {
        if (total) ORIE_CUDA(cudaMemsetAsync(total, 0, 4, st));
        return ORIE_OK;
    }
    const int64_t tiles = ceil_div(n, kScanTile);
    if (tiles == 1) {
        scan_tile_kernel<<<1, kScanThreads, 0, st>>>(in, out, n, nullptr, total);
        ORIE_LAUNCH_CHECK();
        return ORIE_OK;
    }
    uint32_t *sums = (uint32_t *)scratch;
    void *rest = (char *)scratch + round_up(tiles * 4, 256);
    scan_tile_kernel<<<(unsigned)tiles, kScanThreads, 0, st>>>(in, out, n, sums, nullptr);
    ORIE_LAUNCH_CHECK();
    ORIE_TRY(exclusive_scan_u32(sums, sums, tiles, total, rest, st));
    scan_add_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(out, n, sums);
    ORIE_LAUNCH_CHECK();
    return ORIE_OK;
}

// ----------------------------------------------------------------------------
// radix sort (8-bit digits, stable): three launches per pass
//   radix_hist_kernel     per-block digit counts            hist[block][digit]
//   radix_offsets_kernel  one block: digit-major exclusive offsets over (digit, block), in place
//   radix_scatter_kernel  stable scatter
// ----------------------------------------------------------------------------
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortSteps = 8;                       // 32-item steps per warp
constexpr int kSortTile = kSortThreads * kSortSteps;  // 4096 items per block

__device__ __forceinline__ int digit_of(uint64_t key, int shift, uint32_t mask) {
    return (int)((uint32_t)(key >> shift) & mask);
}

__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, int shift, uint32_t mask, uint32_t *__restrict__ hist) {
    __shared__ uint32_t h[kRadix];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll 4
    for (int s = 0; s < kSortSteps; ++s) {
        int64_t k = base + s * kSortThreads + threadIdx.x;
        if (k < n) atomicAdd(&h[digit_of(keys[k], shift, mask)], 1u);
    }
    __syncthreads();
    hist[(int64_t)blockIdx.x * kRadix + threadIdx.x] = h[threadIdx.x];
}

// One warp per digit: lane l owns a contiguous run of blocks; exclusive prefix over the blocks of that digit
// (offs[block][digit]) and the digit's total.  The digit bases (exclusive scan over the 256 totals) are added
// by the scatter kernel itself.
constexpr int kOffsetsThreads = 256;
__global__ void __launch_bounds__(kOffsetsThreads)
radix_offsets_kernel(const uint32_t *__restrict__ hist, uint32_t *__restrict__ offs, uint32_t *__restrict__ totals,
                     int nblocks) {
    const int d = (blockIdx.x * kOffsetsThreads + threadIdx.x) >> 5;
    if (d >= kRadix) return;
    const int lane = threadIdx.x & 31;
    const int per = (nblocks + 31) / 32;
    const int b0 = lane * per, b1 = min(b0 + per, nblocks);
    uint32_t sum = 0;
#pragma unroll 8
    for (int b = b0; b < b1; ++b) sum += hist[(int64_t)b * kRadix + d];
    uint32_t incl = sum;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
        uint32_t y = __shfl_up_sync(kFull, incl, k);
        if (lane >= k) incl += y;
    }
    uint32_t run = incl - sum;
#pragma unroll 8
    for (int b = b0; b < b1; ++b) {
        const uint32_t c = hist[(int64_t)b * kRadix + d];
        offs[(int64_t)b * kRadix + d] = run;
        run += c;
    }
    if (lane == 31) totals[d] = incl;
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

// Stable scatter.  Warp w of a block owns items [w*32*kSortSteps, (w+1)*32*kSortSteps) of the tile, walked in kSortSteps steps of 32
// consecutive items; ballots rank equal digits inside a step, per-warp digit counters carry the rank across
// steps and warps.
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                     uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n,
                     int shift, uint32_t mask, const uint32_t *__restrict__ offs, const uint32_t *__restrict__ totals) {
    __shared__ uint32_t cnt[kSortWarps][kRadix];
    __shared__ uint32_t wsum[kSortWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kSortWarps * kRadix; i += kSortThreads) (&cnt[0][0])[i] = 0;
    // digit base = exclusive scan of the digit totals (thread d owns digit d)
    const uint32_t my_total = totals[threadIdx.x];
    uint32_t incl = my_total;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
        uint32_t y = __shfl_up_sync(kFull, incl, k);
        if (lane >= k) incl += y;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t digit_base = incl - my_total;
    for (int w = 0; w < warp; ++w) digit_base += wsum[w];
    const int64_t wbase = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * (32 * kSortSteps);
    uint64_t k[kSortSteps];
    unsigned peers[kSortSteps];
#pragma unroll
    for (int s = 0; s < kSortSteps; ++s) {
        const int64_t i = wbase + s * 32 + lane;
        const bool ok = i < n;
        k[s] = ok ? keys[i] : 0ull;
        const int dg = digit_of(k[s], shift, mask);
        peers[s] = digit_peers(dg, ok);
        if (ok && lane == (__ffs(peers[s]) - 1)) cnt[warp][dg] += __popc(peers[s]);
        __syncwarp();
    }
    __syncthreads();
    {
        const int d = threadIdx.x;  // kSortThreads == kRadix
        uint32_t run = offs[(int64_t)blockIdx.x * kRadix + d] + digit_base;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t c = cnt[w][d];
            cnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < kSortSteps; ++s) {
        const int64_t i = wbase + s * 32 + lane;
        const bool ok = i < n;
        const int dg = digit_of(k[s], shift, mask);
        uint32_t pos = 0;
        if (ok) pos = cnt[warp][dg] + __popc(peers[s] & ((1u << lane) - 1u));
        __syncwarp();
        if (ok) {
            keys_out[pos] = k[s];
            vals_out[pos] = vals[i];
            if (lane == (__ffs(peers[s]) - 1)) cnt[warp][dg] += __popc(peers[s]);
        }
        __syncwarp();
    }
}

static_assert(kSortThreads == kRadix, "one thread per digit in the carry pass");

size_t radix_scratch_bytes(int64_t n) {
    int64_t nblocks = ceil_div(n > 0 ? n : 1, kSortTile);
    return 2 * (size_t)round_up(nblocks * kRadix * 4, 256) + kRadix * 4;
}

int radix_sort_pairs(uint64_t *keys, uint32_t *vals, uint64_t *keys_tmp, uint32_t *vals_tmp, int64_t n,
                     int bit_lo, int bit_hi, void *scratch, cudaStream_t st) {
    if (n <= 1 || bit_hi <= bit_lo) return ORIE_OK;
    if (n >= (int64_t)1 << 31) {
        set_error("radix_sort_pairs: n=%lld exceeds 2^31-1", (long long)n);
        return ORIE_ELIMIT;
    }
    const int nblocks = (int)ceil_div(n, kSortTile);
    uint32_t *hist = (uint32_t *)scratch;
    uint32_t *offs = (uint32_t *)((char *)scratch + round_up((int64_t)nblocks * kRadix * 4, 256));
    uint32_t *totals = (uint32_t *)((char *)scratch + 2 * round_up((int64_t)nblocks * kRadix * 4, 256));
    uint64_t *kin = keys, *kout = keys_tmp;
    uint32_t *vin = vals, *vout = vals_tmp;
    for (int lo = bit_lo; lo < bit_hi; lo += kRadixBits) {
        const int bits = (bit_hi - lo) < kRadixBits ? (bit_hi - lo) : kRadixBits;
        const uint32_t mask = (1u << bits) - 1u;
        radix_hist_kernel<<<nblocks, kSortThreads, 0, st>>>(kin, n, lo, mask, hist);
        ORIE_LAUNCH_CHECK();
        radix_offsets_kernel<<<kRadix * 32 / kOffsetsThreads, kOffsetsThreads, 0, st>>>(hist, offs, totals, nblocks);
        ORIE_LAUNCH_CHECK();
        radix_scatter_kernel<<<nblocks, kSortThreads, 0, st>>>(kin, vin, kout, vout, n, lo, mask, offs, totals);
        ORIE_LAUNCH_CHECK();
        uint64_t *tk = kin; kin = kout; kout = tk;
        uint32_t *tv = vin; vin = vout; vout = tv;
    }
    if (kin != keys) {
        ORIE_CUDA(cudaMemcpyAsync(keys, kin, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
        ORIE_CUDA(cudaMemcpyAsync(vals, vin, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    }
    return ORIE_OK;
}

}  // namespace orie
