// Rank normalisation of a reward vector for one cross-validation fold (orie_rank_normalize): what the reference's
// regression.py:439-441 does on the host with argsort(argsort(.)) and an O(n^2) comparison loop, here as one radix
// sort (sort.cu) plus two small kernels.  First consumer-side row of SURVEY 8f-4.
#include "common.cuh"

namespace orie {

// order-isomorphic to "ascending double" with -0.0 == +0.0 (numpy compares them equal)
__device__ __forceinline__ uint64_t asc_key(double x) {
    if (x == 0.0) x = 0.0;
    const uint64_t b = (uint64_t)__double_as_longlong(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

// validation rows sort behind every train row (a train reward is never NaN: its key is below 0xfff0...)
__global__ void rank_keys_kernel(const double *__restrict__ reward, const uint8_t *__restrict__ val_mask, int64_t M,
                                 uint64_t *__restrict__ keys, uint32_t *__restrict__ n_train) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const bool val = val_mask && val_mask[i];
    keys[i] = val ? ~0ull : asc_key(reward[i]);
    if (!val) atomicAdd(n_train, 1u);
}

// sorted position i < n_train holds train row order[i]: its rank is i + 1
__global__ void rank_train_kernel(const double *__restrict__ reward, const uint32_t *__restrict__ order, int64_t M,
                                  const uint32_t *__restrict__ n_train, double *__restrict__ out, double *__restrict__ sorted) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = *n_train;
    if (i >= M || i >= n) return;
    const uint32_t row = order[i];
    out[row] = __ddiv_rn((double)(i + 1), (double)n);
    sorted[i] = reward[row];
}

// validation row: share of the train rewards <= its own (upper bound in the sorted train rewards)
__global__ void rank_val_kernel(const double *__restrict__ reward, const uint8_t *__restrict__ val_mask, int64_t M,
                                const uint32_t *__restrict__ n_train, const double *__restrict__ sorted, double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M || !val_mask || !val_mask[i]) return;
    const uint32_t n = *n_train;
    const double x = reward[i];
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sorted[mid] <= x) lo = mid + 1; else hi = mid;
    }
    out[i] = __ddiv_rn((double)lo, (double)n);      // n == 0: 0/0 = NaN, as upstream
}

struct RankLayout {
    size_t keys0, keys1, vals_a, vals_b, sorted, n_train, scratch, total;
};

static int rank_layout(int64_t M, RankLayout *L) {
    int blocks = 0;
    ORIE_TRY(sort_max_blocks(&blocks));
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (size_t)round_up((int64_t)(bytes ? bytes : 1), 256); return at; };
    const size_t m = (size_t)(M > 0 ? M : 1);
    L->keys0 = take(m * 8); L->keys1 = take(m * 8);
    L->vals_a = take(m * 4); L->vals_b = take(m * 4);
    L->sorted = take(m * 8);
    L->n_train = take(4);
    L->scratch = take(sort_scratch_bytes(blocks));
    L->total = o;
    return ORIE_OK;
}

}  // namespace orie

using namespace orie;

extern "C" size_t orie_rank_workspace_bytes(int64_t M) {
    RankLayout L;
    if (M < 0 || rank_layout(M, &L) != ORIE_OK) return 0;
    return L.total;
}

extern "C" int orie_rank_normalize(const double *reward, const uint8_t *val_mask, int64_t M, double *out, void *workspace,
                                   size_t workspace_bytes, orie_stream_t stream) {
    if (M < 0 || (M > 0 && (!reward || !out || !workspace))) {
        set_error("orie_rank_normalize: null buffer or negative length");
        return ORIE_EINVAL;
    }
    if (M == 0) return ORIE_OK;
    if (M >= ((int64_t)1 << 31)) {
        set_error("orie_rank_normalize: M=%lld exceeds 2^31-1", (long long)M);
        return ORIE_ELIMIT;
    }
    RankLayout L;
    ORIE_TRY(rank_layout(M, &L));
    if (workspace_bytes < L.total || ((uintptr_t)workspace & 255)) {
        set_error("orie_rank_normalize: workspace needs %zu bytes, 256-byte aligned (got %zu)", L.total, workspace_bytes);
        return ORIE_EWORKSPACE;
    }
    char *ws = (char *)workspace;
    uint32_t *n_train = (uint32_t *)(ws + L.n_train);
    double *sorted = (double *)(ws + L.sorted);
    const unsigned grid = (unsigned)ceil_div(M, 256);
    ORIE_CUDA(cudaMemsetAsync(n_train, 0, 4, stream));
    rank_keys_kernel<<<grid, 256, 0, stream>>>(reward, val_mask, M, (uint64_t *)(ws + L.keys0), n_train);
    ORIE_LAUNCH_CHECK();
    SortJob j;
    j.n = M;
    j.keys0 = (uint64_t *)(ws + L.keys0); j.keys1 = (uint64_t *)(ws + L.keys1);
    j.vals_a = (uint32_t *)(ws + L.vals_a); j.vals_b = (uint32_t *)(ws + L.vals_b);
    ORIE_TRY(sort_add_passes(&j, kDigitKey, 0, 64));
    int blocks = 0;
    ORIE_TRY(sort_max_blocks(&blocks));
    ORIE_TRY(sort_run(j, blocks, ws + L.scratch, stream));
    rank_train_kernel<<<grid, 256, 0, stream>>>(reward, j.vals_a, M, n_train, out, sorted);
    ORIE_LAUNCH_CHECK();
    if (val_mask) {
        rank_val_kernel<<<grid, 256, 0, stream>>>(reward, val_mask, M, n_train, sorted, out);
        ORIE_LAUNCH_CHECK();
    }
    return ORIE_OK;
}
