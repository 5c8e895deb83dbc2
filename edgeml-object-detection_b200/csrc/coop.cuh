// Persistent cooperative kernels: grid-wide barrier and launch helper.
//
// The index build's sort / scan steps run as ONE launch each, with every CTA
// resident for the whole kernel (cudaLaunchCooperativeKernel guarantees it) and
// phases separated by a counter barrier in global memory.  At the sizes of this
// path (10^5..10^7 keys) a pass is latency-bound, so a ~1.5 us barrier replaces
// a kernel boundary (launch + drain + ramp-up) of several microseconds and the
// per-pass launch count drops from three to a fraction of one.
#pragma once
#include "common.cuh"

namespace orie {

// Monotonic counter barrier.  `bar` is zero before the launch; `epoch` is a per-thread running target
// (start at 0, same sequence of calls in every CTA).  Data written before the barrier by any CTA is
// visible after it to loads that do not go through the non-coherent path (use __ldcg / plain loads on
// non-const, non-restrict pointers for such data).
__device__ __forceinline__ void grid_sync(unsigned *bar, unsigned &epoch) {
    epoch += gridDim.x;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while (v < epoch);
        __threadfence();
    }
    __syncthreads();
}

// Exclusive prefix of `v` over the threads of the block (THREADS = multiple of 32, <= 1024); `total` (optional)
// receives the block sum.  `ws` is shared scratch of THREADS / 32 words; ends with a __syncthreads.
template <int THREADS>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *ws, uint32_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += y;
    }
    __syncthreads();            // ws may still be read from a previous call
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    uint32_t base = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) {
        const uint32_t x = ws[w];
        if (w < warp) base += x;
        sum += x;
    }
    if (total) *total = sum;
    return base + incl - v;
}

// Number of CTAs of `kernel` that can be co-resident on the current device.
template <typename K>
static inline int coop_max_blocks(K kernel, int threads, size_t smem, int *out) {
    int dev = 0, sms = 0, per = 0;
    ORIE_CUDA(cudaGetDevice(&dev));
    ORIE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ORIE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kernel, threads, smem));
    *out = sms * per;
    if (*out < 1) {
        set_error("cooperative kernel does not fit on the device");
        return ORIE_ECUDA;
    }
    return ORIE_OK;
}

}  // namespace orie
