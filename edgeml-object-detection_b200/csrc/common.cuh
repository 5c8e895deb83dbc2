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

// ---- device primitives (sort.cu) -------------------------------------------
// Stable LSD radix sort of (key, value) pairs on bits [bit_lo, bit_hi) of the key.
// keys/vals are ping-ponged with the *_tmp buffers; the result always ends in keys/vals.
// scratch: at least radix_scratch_bytes(n) bytes.
size_t radix_scratch_bytes(int64_t n);
int radix_sort_pairs(uint64_t *keys, uint32_t *vals, uint64_t *keys_tmp, uint32_t *vals_tmp, int64_t n,
                     int bit_lo, int bit_hi, void *scratch, cudaStream_t st);
// Exclusive prefix sum of uint32 (in place allowed); total (optional, device) receives the sum.
size_t scan_scratch_bytes(int64_t n);
int exclusive_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, uint32_t *total, void *scratch, cudaStream_t st);

}  // namespace orie
