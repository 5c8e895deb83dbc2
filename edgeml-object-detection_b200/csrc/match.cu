// Batched FP64 box-IoU + true-positive matching (orie_match) and the DCSB count
// (orie_dcsb).
//
// Replaces lib/metrics.py:67-86 (box_iou) and :38-64 (box_correct) as driven per
// image by lib/data.py:63-83.  Upstream's sort / unique / unique procedure is
// equivalent to the data-parallel rule implemented here (SURVEY.md §8a row A4,
// oracle/orie_oracle.py:match_detections):
//   best[d] = same-class label with the largest IoU (independent of the
//             threshold; exact ties -> highest label row),
//   d is a TP at threshold t  <=>  IoU[best[d], d] >= t and no earlier row d' < d
//             of the same image has best[d'] == best[d] with IoU >= t.
// One CTA per (image, detector).  Label boxes are staged in shared memory with
// 16-byte loads; "first detection per label" is a shared-memory atomicMin — for all
// thresholds in one pass when labels x thresholds fit the table (the usual case),
// otherwise per threshold, issued by the lowest lane of each __match_any_sync group.
// All IoU arithmetic is IEEE float64 in upstream's operation order; the library
// is compiled with -fmad=false so nothing is contracted into an FMA.
#include "common.cuh"

namespace orie {

constexpr int kMatchThreads = 128;
constexpr int kLabTile = 256;
constexpr int kFirstCap = ORIE_MAX_LABELS_PER_IMAGE;

struct Thresholds {
    double v[ORIE_MAX_THRESHOLDS];
};

__device__ __forceinline__ double iou_f64(double lx1, double ly1, double lx2, double ly2, double la,
                                          double dx1, double dy1, double dx2, double dy2, double da) {
    const double ix1 = fmax(lx1, dx1), iy1 = fmax(ly1, dy1);
    const double ix2 = fmin(lx2, dx2), iy2 = fmin(ly2, dy2);
    const double inter = __dmul_rn(fmax(0.0, __dsub_rn(ix2, ix1)), fmax(0.0, __dsub_rn(iy2, iy1)));
    return __ddiv_rn(inter, __dsub_rn(__dadd_rn(la, da), inter));
}

__global__ void __launch_bounds__(kMatchThreads)
match_kernel(const double *__restrict__ det_box, const int32_t *__restrict__ det_cls, const int64_t *__restrict__ det_off,
             const double *__restrict__ lab_box, const int32_t *__restrict__ lab_cls, const int64_t *__restrict__ lab_off,
             const __grid_constant__ Thresholds thr, int T, uint16_t *__restrict__ tp_mask, int32_t *__restrict__ match_idx,
             double *__restrict__ best_iou) {
    __shared__ double2 s_box[kLabTile][2];   // (x1,y1),(x2,y2)
    __shared__ double s_area[kLabTile];
    __shared__ int32_t s_cls[kLabTile];
    __shared__ int32_t s_first[kFirstCap];
    __shared__ double s_thr[ORIE_MAX_THRESHOLDS];

    const int64_t img = blockIdx.x;
    const int64_t d0 = det_off[img];
    const int nd = (int)(det_off[img + 1] - d0);
    const int64_t l0 = lab_off[img];
    const int nl = (int)(lab_off[img + 1] - l0);
    if (nd == 0) return;
    const int tid = threadIdx.x;
    if (tid < ORIE_MAX_THRESHOLDS) s_thr[tid] = thr.v[tid];      // read again only after a __syncthreads
    if (nl == 0) {
        for (int d = tid; d < nd; d += kMatchThreads) {
            tp_mask[d0 + d] = 0;
            match_idx[d0 + d] = -1;
            best_iou[d0 + d] = 0.0;
        }
        return;
    }
    const double2 *lab2 = reinterpret_cast<const double2 *>(lab_box) + l0 * 2;
    const double2 *det2 = reinterpret_cast<const double2 *>(det_box) + d0 * 2;

    // ---- phase 1: best same-class label of every detection
    for (int dbase = 0; dbase < nd; dbase += kMatchThreads) {
        const int d = dbase + tid;
        const bool live = d < nd;
        double dx1 = 0, dy1 = 0, dx2 = 0, dy2 = 0, da = 0;
        int dc = -1;
        if (live) {
            const double2 a = det2[d * 2], b = det2[d * 2 + 1];
            dx1 = a.x; dy1 = a.y; dx2 = b.x; dy2 = b.y;
            da = __dmul_rn(__dsub_rn(dx2, dx1), __dsub_rn(dy2, dy1));
            dc = det_cls[d0 + d];
        }
        int best = -1;
        double bv = -1.0;
        for (int lt = 0; lt < nl; lt += kLabTile) {
            const int cnt = min(kLabTile, nl - lt);
            if (lt > 0 || dbase == 0 || nl > kLabTile) {
                __syncthreads();
                for (int i = tid; i < cnt; i += kMatchThreads) {
                    const double2 a = lab2[(lt + i) * 2], b = lab2[(lt + i) * 2 + 1];
                    s_box[i][0] = a;
                    s_box[i][1] = b;
                    s_area[i] = __dmul_rn(__dsub_rn(b.x, a.x), __dsub_rn(b.y, a.y));
                    s_cls[i] = lab_cls[l0 + lt + i];
                }
                __syncthreads();
            }
            if (live) {
                for (int i = 0; i < cnt; ++i) {
                    if (s_cls[i] != dc) continue;
                    const double2 a = s_box[i][0], b = s_box[i][1];
                    const double v = iou_f64(a.x, a.y, b.x, b.y, s_area[i], dx1, dy1, dx2, dy2, da);
                    if (v >= bv) { bv = v; best = lt + i; }
                }
            }
        }
        if (live) {
            match_idx[d0 + d] = best;
            best_iou[d0 + d] = best >= 0 ? bv : 0.0;
            tp_mask[d0 + d] = 0;
        }
    }
    __syncthreads();

    // ---- phase 2: first (lowest row) candidate per label and threshold
    const int lane = tid & 31;
    if (nl * T <= kFirstCap) {
        // all thresholds in one pass: first[t][label] = lowest row that claims the label at threshold t
        for (int i = tid; i < nl * T; i += kMatchThreads) s_first[i] = 0x7fffffff;
        __syncthreads();
        for (int d = tid; d < nd; d += kMatchThreads) {
            const int best = match_idx[d0 + d];
            if (best < 0) continue;
            const double bv = best_iou[d0 + d];
            for (int t = 0; t < T; ++t)
                if (bv >= s_thr[t]) atomicMin(&s_first[t * nl + best], d);
        }
        __syncthreads();
        for (int d = tid; d < nd; d += kMatchThreads) {
            const int best = match_idx[d0 + d];
            unsigned mask = 0;
            if (best >= 0) {
                const double bv = best_iou[d0 + d];
                for (int t = 0; t < T; ++t)
                    if (bv >= s_thr[t] && s_first[t * nl + best] == d) mask |= 1u << t;
            }
            tp_mask[d0 + d] = (uint16_t)mask;
            if (mask == 0) match_idx[d0 + d] = -1;
        }
    } else if (nl <= kFirstCap) {
        for (int t = 0; t < T; ++t) {
            const double th = s_thr[t];
            __syncthreads();
            for (int i = tid; i < nl; i += kMatchThreads) s_first[i] = 0x7fffffff;
            __syncthreads();
            for (int dbase = 0; dbase < nd; dbase += kMatchThreads) {   // uniform trip count
                const int d = dbase + tid;
                const bool live = d < nd;
                const int best = live ? match_idx[d0 + d] : -1;
                const bool cand = best >= 0 && best_iou[d0 + d] >= th;
                const int key = cand ? best : -1 - lane;               // non-candidates never pair up
                const unsigned peers = __match_any_sync(kFull, key);
                if (cand && lane == (__ffs(peers) - 1)) atomicMin(&s_first[best], d);
            }
            __syncthreads();
            for (int d = tid; d < nd; d += kMatchThreads) {
                const int best = match_idx[d0 + d];
                if (best >= 0 && best_iou[d0 + d] >= th && s_first[best] == d)
                    tp_mask[d0 + d] |= (uint16_t)(1u << t);             // d is owned by this thread
            }
        }
        for (int d = tid; d < nd; d += kMatchThreads)
            if (tp_mask[d0 + d] == 0) match_idx[d0 + d] = -1;
    } else {
        // more labels than the shared table holds: quadratic scan over earlier rows
        for (int d = tid; d < nd; d += kMatchThreads) {
            const int best = match_idx[d0 + d];
            const double bv = best_iou[d0 + d];
            unsigned mask = 0;
            if (best >= 0) {
                for (int t = 0; t < T; ++t) {
                    if (!(bv >= s_thr[t])) continue;
                    bool first = true;
                    for (int e = 0; e < d && first; ++e)
                        if (match_idx[d0 + e] == best && best_iou[d0 + e] >= s_thr[t]) first = false;
                    if (first) mask |= 1u << t;
                }
            }
            tp_mask[d0 + d] = (uint16_t)mask;
        }
        __syncthreads();
        for (int d = tid; d < nd; d += kMatchThreads)
            if (tp_mask[d0 + d] == 0) match_idx[d0 + d] = -1;
    }
}

__global__ void dcsb_kernel(const double *__restrict__ w_conf, const int64_t *__restrict__ w_off,
                            const double *__restrict__ s_conf, const int64_t *__restrict__ s_off, int64_t M,
                            int64_t *__restrict__ out) {
    const int64_t img = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (img >= M) return;
    const int lane = threadIdx.x & 31;
    int nw = 0, ns = 0;
    for (int64_t i = w_off[img] + lane; i < w_off[img + 1]; i += 32) nw += w_conf[i] > 0.5;
    for (int64_t i = s_off[img] + lane; i < s_off[img + 1]; i += 32) ns += s_conf[i] > 0.5;
    int diff = ns - nw;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) diff += __shfl_xor_sync(kFull, diff, d);
    if (lane == 0) out[img] = diff;
}

}  // namespace orie

using namespace orie;

extern "C" int orie_match(const double *det_box, const int32_t *det_cls, const int64_t *det_off,
                          const double *lab_box, const int32_t *lab_cls, const int64_t *lab_off,
                          const double *iouv_host, int T, int64_t M,
                          uint16_t *tp_mask, int32_t *match_idx, double *best_iou, orie_stream_t stream) {
    if (T < 1 || T > ORIE_MAX_THRESHOLDS) {
        set_error("orie_match: T=%d outside [1,%d]", T, ORIE_MAX_THRESHOLDS);
        return ORIE_ELIMIT;
    }
    if (M < 0 || !det_off || !lab_off || !iouv_host || !tp_mask || !match_idx || !best_iou) {
        set_error("orie_match: null pointer or negative image count");
        return ORIE_EINVAL;
    }
    if (M == 0) return ORIE_OK;
    Thresholds thr;
    for (int t = 0; t < ORIE_MAX_THRESHOLDS; ++t) thr.v[t] = t < T ? iouv_host[t] : 2.0;
    match_kernel<<<(unsigned)M, kMatchThreads, 0, stream>>>(det_box, det_cls, det_off, lab_box, lab_cls, lab_off, thr, T,
                                                           tp_mask, match_idx, best_iou);
    ORIE_LAUNCH_CHECK();
    return ORIE_OK;
}

extern "C" int orie_dcsb(const double *w_conf, const int64_t *w_off, const double *s_conf, const int64_t *s_off,
                         int64_t M, int64_t *out, orie_stream_t stream) {
    if (M < 0 || !w_off || !s_off || !out) {
        set_error("orie_dcsb: null pointer or negative image count");
        return ORIE_EINVAL;
    }
    if (M == 0) return ORIE_OK;
    const int warps = 8;
    dcsb_kernel<<<(unsigned)ceil_div(M, warps), warps * 32, 0, stream>>>(w_conf, w_off, s_conf, s_off, M, out);
    ORIE_LAUNCH_CHECK();
    return ORIE_OK;
}
