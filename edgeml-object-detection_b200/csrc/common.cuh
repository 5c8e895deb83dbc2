// Shared helpers for the orie_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/orie_b200.h"

namespace orie {

constexpr int kChunk = 32;           // slots per chunk == warp width
constexpr unsigned kFull = 0xffffffffu;

void set_error(const char *fmt, ...);

#define ORIE_CUDA(expr)                                                                        \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::orie::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return ORIE_ECUDA;                                                                 \
        }                                                                                      \
    } while (0)

// every kernel launch of the library is followed by this (n = launches since the last check)
void count_launches(int n);
#define ORIE_LAUNCH_CHECK_N(n) \
    do {                       \
        ::orie::count_launches(n); \
        ORIE_CUDA(cudaGetLastError()); \
    } while (0)
#define ORIE_LAUNCH_CHECK() ORIE_LAUNCH_CHECK_N(1)

#define ORIE_TRY(expr)                \
    do {                              \
        int _r = (expr);              \
        if (_r != ORIE_OK) return _r; \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---- device primitive (sort.cu) ---------------------------------------------
// Stable LSD radix sort of uint32 values (optionally carrying 64-bit keys), all passes in one cooperative
// launch.  The digit of a pass is 8 bits (or fewer) of
//   kDigitKey            the item's 64-bit key                  (key passes must come first)
//   kDigitClass          the class of the value u, a combined row id: u < split ? cls_lo[u] : cls_hi[u - split]
//   kDigitBatch          img[u] / 32
// vals_in == nullptr means the identity (value = position).  The sorted values always end in vals_a
// (vals_a, vals_b: two buffers of n values, distinct from vals_in); keys0 holds the input keys and is
// overwritten, keys1 is a second buffer of n keys; the sorted keys are not kept.
// rank_out (optional): rank_out[v] = number of sorted values < rank_split in front of position v.
enum { kDigitKey = 0, kDigitClass = 1, kDigitBatch = 2 };
constexpr int kMaxSortPasses = 12;
struct SortPass {
    int kind, shift;
    uint32_t mask;
};
struct SortJob {
    int64_t n = 0;
    int npass = 0;
    SortPass pass[kMaxSortPasses] = {};
    uint64_t *keys0 = nullptr, *keys1 = nullptr;
    const uint32_t *vals_in = nullptr;
    uint32_t *vals_a = nullptr, *vals_b = nullptr;
    const int32_t *cls_lo = nullptr, *cls_hi = nullptr;
    const uint32_t *img = nullptr;
    uint32_t split = 0xffffffffu;
    uint32_t *rank_out = nullptr;
    uint32_t rank_split = 0;
    // filled in by sort_run
    uint32_t *table = nullptr;
    unsigned *bar = nullptr;
    int64_t per = 0;
};
// ---- bucket sort (sort.cu): the dataset sort of the index build -------------
// Sorts the ids [0, n) by (class, 64-bit key, id) without a pass per key byte: ONE pass bins every item by
// (class, leading varying key bits), consecutive bins are packed into buckets of a few hundred items, items are
// scattered to their bucket with atomics, and every bucket is sorted on its own by a bitonic network — by one warp
// in registers (<= 256 items) or by one CTA in shared memory (<= 8192).  Four grid barriers in all; if the data
// defeats the binning (a bucket beyond 8192 items: heavy exact ties) the kernel falls back to the LSD radix passes.
struct BucketSortJob {
    int64_t n = 0, C = 1;
    int bits = 0;                      // key bins per class = 1 << bits
    int cap = 0;                       // items per bucket the packing aims at
    int64_t nbuckets = 0;              // (n - 1) / cap + 1
    const uint64_t *key_and_or = nullptr;   // device [2]: AND and OR of all keys (common leading bits are skipped)
    uint64_t *keys_part = nullptr;     // [n] keys partitioned by bucket (= lsd.keys1)
    uint32_t *vals_part = nullptr;     // [n] ids partitioned by bucket (= lsd.vals_b)
    uint32_t *binhist = nullptr;       // [C << bits]   zero on entry
    uint32_t *bucket_start = nullptr;  // [nbuckets]    zero on entry (holds ~start, maximised)
    uint32_t *bucket_fill = nullptr;   // [nbuckets]    zero on entry
    uint32_t *bucket_weak = nullptr;   // [nbuckets]    zero on entry
    uint32_t *wbefore = nullptr;       // [nbuckets]    ids < split in front of each bucket
    uint32_t *slice_sum = nullptr;     // [grid]
    uint32_t *big = nullptr;           // buckets of more than 256 items (sorted by a whole CTA)
    uint32_t *ctl = nullptr;           // [2]: fell back to the radix passes, number of entries of `big`
    SortJob lsd;                       // keys0 = input keys, vals_a = sorted ids out, rank_out / rank_split, classes,
                                       // and the fallback's passes; bar / table / per are filled in by bucket_sort_run
};
bool bucket_sort_applicable(int64_t n, int64_t C);
size_t bucket_sort_scratch_bytes(int64_t n, int64_t C, int max_blocks);
int bucket_sort_max_blocks(int *out);
int bucket_sort_run(BucketSortJob job, int64_t C, int max_blocks, void *scratch, cudaStream_t st);

int sort_max_blocks(int *out);                    // co-resident CTAs of the sort kernel on the current device
size_t sort_scratch_bytes(int max_blocks);
int sort_add_passes(SortJob *job, int kind, int bit_lo, int bit_hi);
int sort_run(SortJob job, int max_blocks, void *scratch, cudaStream_t st);

}  // namespace orie
