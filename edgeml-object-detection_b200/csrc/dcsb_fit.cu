// DCSB baseline: threshold search and offloading decisions (orie_dcsb_fit).
//
// Replaces baseline.py:67-152 (fit_dcsb) as driven by baseline.py:161-206 for --baseline dcsb: a consumer of the
// reward vector (the rewards are binarised, > 0 -> 1) and of the weak detector's outputs (second half of SURVEY 8f-4).
//   1. confidence threshold by bisection of [0, 1] until the number of train detections above it matches the number
//      of train labels within 1e-4 (baseline.py:95-106);
//   2. per image: detections above that threshold, the smallest box area among them, detections above 0.5
//      (filter_box, baseline.py:79-88; area = (x2 - x1) * (y2 - y1) as get_area, baseline.py:155-158);
//   3. grid search over (object-count threshold 1..10) x (area thresholds handed in by the host, upstream
//      np.arange(0.2, 0.9, 0.01)): train accuracy of "difficult case = counts differ and (count > n or area < a)",
//      first maximum over a, strictly better over n (baseline.py:110-126);
//   4. decisions for every image (train and validation rows alike, baseline.py:128-141).
// Accuracies are compared as integer hit counts (same denominator), everything else is IEEE float64 in upstream's order.
#include "common.cuh"

namespace orie {

constexpr int kFitThreads = 1024;
constexpr int kMaxAreaSteps = 256;
constexpr int kMaxCountSteps = 32;

struct FitArgs {
    int64_t M, D;
    const double *box, *conf;      // weak detections: xyxy f64[D,4], f64[D]
    const int64_t *off;            // [M+1]
    const uint8_t *val_mask;       // [M] nullable: 1 = validation row
    const int64_t *label_num;      // [M] ground-truth objects per image
    const int64_t *reward;         // [M] binarised rewards (0 / 1)
    int n_area, n_count;
    double *tconf;                 // [D] confidence of train detections, -1 for validation rows' detections
    int32_t *num, *det;            // [M] detections above the fitted threshold / above 0.5
    double *area;                  // [M] smallest area among the former (0 if none)
    double *model;                 // [4] out: conf threshold, count threshold, area threshold, train hits of the best pair
    int64_t *est;                  // [M] out: decisions
    int32_t *iters;                // [1] out: bisection steps taken
};
struct AreaSteps {
    double v[kMaxAreaSteps];
};

// confidence of every detection of a train image, -1 elsewhere (never above a threshold in [0, 1])
__global__ void fit_train_conf_kernel(const FitArgs a) {
    const int64_t img = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (img >= a.M) return;
    const bool val = a.val_mask && a.val_mask[img];
    for (int64_t d = a.off[img] + (threadIdx.x & 31); d < a.off[img + 1]; d += 32) a.tconf[d] = val ? -1.0 : a.conf[d];
}

// One CTA: the bisection is a chain of dependent counts over the train detections.
__global__ void __launch_bounds__(kFitThreads) fit_conf_threshold_kernel(const FitArgs a, int max_iters) {
    __shared__ unsigned long long s_cnt[kFitThreads / 32];
    __shared__ unsigned long long s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto block_sum = [&](unsigned long long v) -> unsigned long long {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
        __syncthreads();
        if (lane == 0) s_cnt[warp] = v;
        __syncthreads();
        if (tid == 0) {
            unsigned long long t = 0;
            for (int w = 0; w < kFitThreads / 32; ++w) t += s_cnt[w];
            s_total = t;
        }
        __syncthreads();
        return s_total;
    };
    unsigned long long lab = 0;
    for (int64_t i = tid; i < a.M; i += kFitThreads)
        if (!(a.val_mask && a.val_mask[i])) lab += (unsigned long long)a.label_num[i];
    const long long labels = (long long)block_sum(lab);
    double low = 0.0, high = 1.0, mid = 0.0;
    int it = 0;
    for (; it < max_iters; ++it) {
        mid = __ddiv_rn(__dadd_rn(low, high), 2.0);
        unsigned long long c = 0;
        for (int64_t d = tid; d < a.D; d += kFitThreads) c += a.tconf[d] > mid ? 1ull : 0ull;
        const long long diff = (long long)block_sum(c) - labels;
        if (diff >= 0) low = mid; else high = mid;
        // abs(num_diff) / np.sum(train_label) < 1e-4; no labels: x / 0 is inf or nan, never below the tolerance
        const double ad = (double)(diff < 0 ? -diff : diff);
        if (labels != 0 && __ddiv_rn(ad, (double)labels) < 1e-4) { ++it; break; }
    }
    if (tid == 0) { a.model[0] = mid; *a.iters = it; }
}

// filter_box at the fitted threshold and at 0.5: one warp per image
__global__ void fit_filter_kernel(const FitArgs a) {
    const int64_t img = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (img >= a.M) return;
    const int lane = threadIdx.x & 31;
    const double thr = a.model[0];
    int n = 0, n5 = 0;
    double amin = __longlong_as_double(0x7ff0000000000000ll);       // +inf
    for (int64_t d = a.off[img] + lane; d < a.off[img + 1]; d += 32) {
        const double c = a.conf[d];
        n5 += c > 0.5;
        if (c > thr) {
            ++n;
            const double *b = a.box + d * 4;
            amin = fmin(amin, __dmul_rn(__dsub_rn(b[2], b[0]), __dsub_rn(b[3], b[1])));
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        n += __shfl_xor_sync(kFull, n, d);
        n5 += __shfl_xor_sync(kFull, n5, d);
        amin = fmin(amin, __shfl_xor_sync(kFull, amin, d));
    }
    if (lane == 0) {
        a.num[img] = n;
        a.det[img] = n5;
        a.area[img] = n ? amin : 0.0;
    }
}

// one thread per (count threshold, area threshold) pair over the train rows, then upstream's sequential choice
__global__ void __launch_bounds__(kFitThreads) fit_grid_kernel(const FitArgs a, const __grid_constant__ AreaSteps steps) {
    __shared__ unsigned hits[kMaxCountSteps * kMaxAreaSteps > 8192 ? 8192 : kMaxCountSteps * kMaxAreaSteps];
    const int pairs = a.n_count * a.n_area;
    for (int p = threadIdx.x; p < pairs; p += kFitThreads) {
        const int nth = p / a.n_area + 1;
        const double ath = steps.v[p % a.n_area];
        unsigned h = 0;
        for (int64_t i = 0; i < a.M; ++i) {
            if (a.val_mask && a.val_mask[i]) continue;
            const int num = a.num[i];
            const bool hard = num != a.det[i] && (num > nth || a.area[i] < ath);
            h += (hard ? 1 : 0) == a.reward[i] ? 1u : 0u;
        }
        hits[p] = h;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned best = 0;
        double best_n = 0.0, best_a = 0.0;
        for (int n = 0; n < a.n_count; ++n) {
            int arg = 0;
            for (int k = 1; k < a.n_area; ++k)
                if (hits[n * a.n_area + k] > hits[n * a.n_area + arg]) arg = k;      // np.argmax: first maximum
            if (hits[n * a.n_area + arg] > best) {                                  // strictly better only
                best = hits[n * a.n_area + arg];
                best_n = (double)(n + 1);
                best_a = steps.v[arg];
            }
        }
        a.model[1] = best_n; a.model[2] = best_a; a.model[3] = (double)best;
    }
}

__global__ void fit_decide_kernel(const FitArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.M) return;
    const int num = a.num[i];
    a.est[i] = (num != a.det[i] && ((double)num > a.model[1] || a.area[i] < a.model[2])) ? 1 : 0;
}

static size_t fit_layout(int64_t M, int64_t D, size_t *tconf, size_t *num, size_t *det, size_t *area, size_t *iters) {
    size_t o = 0;
    auto take = [&](int64_t bytes) { size_t at = o; o += (size_t)round_up(bytes > 0 ? bytes : 1, 256); return at; };
    *tconf = take(D * 8); *num = take(M * 4); *det = take(M * 4); *area = take(M * 8); *iters = take(4);
    return o;
}

}  // namespace orie

using namespace orie;

extern "C" size_t orie_dcsb_fit_workspace_bytes(int64_t M, int64_t D) {
    if (M < 0 || D < 0) return 0;
    size_t a, b, c, d, e;
    return fit_layout(M, D, &a, &b, &c, &d, &e);
}

extern "C" int orie_dcsb_fit(const double *w_box, const double *w_conf, const int64_t *w_off, int64_t M, int64_t D,
                             const uint8_t *val_mask, const int64_t *label_num, const int64_t *reward01,
                             const double *area_steps_host, int n_area, int n_count,
                             double *model, int64_t *est, void *workspace, size_t workspace_bytes, orie_stream_t stream) {
    if (M < 1 || D < 0 || !w_off || !label_num || !reward01 || !area_steps_host || !model || !est || !workspace ||
        (D > 0 && (!w_box || !w_conf))) {
        set_error("orie_dcsb_fit: null buffer or empty dataset");
        return ORIE_EINVAL;
    }
    if (n_area < 1 || n_area > kMaxAreaSteps || n_count < 1 || n_count > kMaxCountSteps || n_area * n_count > 8192) {
        set_error("orie_dcsb_fit: grid of %d x %d thresholds outside the supported %d x %d", n_count, n_area, kMaxCountSteps, kMaxAreaSteps);
        return ORIE_ELIMIT;
    }
    size_t o_tconf, o_num, o_det, o_area, o_iters;
    const size_t need = fit_layout(M, D, &o_tconf, &o_num, &o_det, &o_area, &o_iters);
    if (workspace_bytes < need || ((uintptr_t)workspace & 255)) {
        set_error("orie_dcsb_fit: workspace needs %zu bytes, 256-byte aligned (got %zu)", need, workspace_bytes);
        return ORIE_EWORKSPACE;
    }
    char *ws = (char *)workspace;
    FitArgs a;
    memset(&a, 0, sizeof(a));
    a.M = M; a.D = D; a.box = w_box; a.conf = w_conf; a.off = w_off; a.val_mask = val_mask; a.label_num = label_num;
    a.reward = reward01; a.n_area = n_area; a.n_count = n_count;
    a.tconf = (double *)(ws + o_tconf); a.num = (int32_t *)(ws + o_num); a.det = (int32_t *)(ws + o_det);
    a.area = (double *)(ws + o_area); a.iters = (int32_t *)(ws + o_iters);
    a.model = model; a.est = est;
    AreaSteps steps;
    for (int k = 0; k < kMaxAreaSteps; ++k) steps.v[k] = k < n_area ? area_steps_host[k] : 0.0;
    const unsigned img_grid = (unsigned)ceil_div(M, 8);
    fit_train_conf_kernel<<<img_grid, 256, 0, stream>>>(a);
    // upstream loops until the tolerance is met; a double interval cannot be halved more than ~1100 times
    fit_conf_threshold_kernel<<<1, kFitThreads, 0, stream>>>(a, 1100);
    fit_filter_kernel<<<img_grid, 256, 0, stream>>>(a);
    fit_grid_kernel<<<1, kFitThreads, 0, stream>>>(a, steps);
    fit_decide_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, stream>>>(a);
    ORIE_LAUNCH_CHECK_N(5);
    return ORIE_OK;
}
