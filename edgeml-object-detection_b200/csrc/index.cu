// Dataset index build (orie_index_build): everything about the ORIE computation
// that does not depend on the target or on its ensemble is computed here once.
//
// reward.py:40-49 gathers N+1 cached records and lib/metrics.py:100-104 re-sorts
// them by confidence for EVERY target.  The relative order of two detections never
// changes between targets, so the engine sorts the detections of the whole dataset
// once — weak and strong together, by (class asc, confidence desc, weak before
// strong, image, row) with the radix sort in sort.cu — and lays the weak ones out in
// "slots"; an ensemble then only selects a subset of slots (reward.cu), and a strong
// detection's place among the weak ones is the number of weak detections sorted
// before it.  See DESIGN.md §3 for the layout.
#include <algorithm>
#include <vector>

#include "index.cuh"

namespace orie {

// ----------------------------------------------------------------------------
// small kernels.  "u" is a combined detection id: [0, Dw) weak rows, [Dw, Dw+Ds) strong rows.
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint64_t conf_desc_key(double c) {
    // order-isomorphic to "descending double": smaller key <=> larger confidence
    uint64_t b = (uint64_t)__double_as_longlong(c);
    uint64_t asc = (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
    return ~asc;
}

struct Dets {
    int64_t Dw, Ds;
    const int32_t *w_cls, *s_cls;
    const double *w_conf, *s_conf;
    const uint16_t *w_tp, *s_tp;
    __device__ __forceinline__ int cls(uint32_t u) const { return u < Dw ? w_cls[u] : s_cls[u - Dw]; }
    __device__ __forceinline__ double conf(uint32_t u) const { return u < Dw ? w_conf[u] : s_conf[u - Dw]; }
};

// image of every row of one CSR block (largest i with off[i] <= k)
__global__ void image_of_row_kernel(const int64_t *__restrict__ off, int64_t M, int64_t n, uint32_t *__restrict__ img,
                                    int32_t *__restrict__ status) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int64_t lo = 0, hi = M;
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (off[mid] <= k) lo = mid; else hi = mid;
    }
    img[k] = (uint32_t)lo;
    if (off[lo + 1] - off[lo] > 65535) atomicOr(status, 1);
}

__global__ void conf_keys_kernel(const Dets d, int64_t n, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    keys[u] = conf_desc_key(d.conf((uint32_t)u));
    vals[u] = (uint32_t)u;
}

enum KeyKind { kKeyClass = 0, kKeyImageDetector = 1, kKeyBatchDetector = 2 };

// re-key the current order for the next stable pass
template <int KIND>
__global__ void rekey_kernel(const Dets d, const uint32_t *__restrict__ img_all, const uint32_t *__restrict__ vals, int64_t n,
                             uint64_t *__restrict__ keys) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint32_t u = vals[v];
    uint64_t k;
    if (KIND == kKeyClass) k = (uint64_t)(uint32_t)d.cls(u);
    else if (KIND == kKeyImageDetector) k = ((uint64_t)img_all[u] << 1) | (u >= d.Dw);
    else k = ((uint64_t)(img_all[u] >> 5) << 1) | (u >= d.Dw);
    keys[v] = k;
}

// class histogram: per-block shared-memory bins (C <= kHistSmemBins), flushed with one global atomic per
// non-empty bin; plain global atomics beyond that
constexpr int kHistSmemBins = 8192;
constexpr int kHistItemsPerBlock = 8192;
__global__ void class_hist_kernel(const int32_t *__restrict__ cls, int64_t n, int64_t C, uint32_t *__restrict__ hist,
                                  int32_t *__restrict__ status) {
    extern __shared__ uint32_t bins[];
    const bool local = C <= kHistSmemBins;
    if (local) {
        for (int i = threadIdx.x; i < C; i += blockDim.x) bins[i] = 0;
        __syncthreads();
    }
    const int64_t base = (int64_t)blockIdx.x * kHistItemsPerBlock;
    const int64_t end = base + kHistItemsPerBlock < n ? base + kHistItemsPerBlock : n;
    for (int64_t k = base + threadIdx.x; k < end; k += blockDim.x) {
        const int c = cls[k];
        if (c < 0 || c >= C) { atomicOr(status, 2); continue; }
        atomicAdd(local ? &bins[c] : &hist[c], 1u);
    }
    if (local) {
        __syncthreads();
        for (int i = threadIdx.x; i < C; i += blockDim.x)
            if (bins[i]) atomicAdd(&hist[i], bins[i]);
    }
}

__global__ void fill_u32_kernel(uint32_t *__restrict__ p, int64_t n, uint32_t v) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) p[k] = v;
}

__global__ void weak_flag_kernel(const uint32_t *__restrict__ order, int64_t n, uint32_t Dw, uint32_t *__restrict__ flag) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) flag[v] = order[v] < Dw ? 1u : 0u;
}

// position v of the combined (class, conf desc) order -> slot (weak) / insertion slot (strong).
// wpre[v] = number of weak detections sorted before v.
__global__ void place_slots_kernel(const Dets d, const uint32_t *__restrict__ order, const uint32_t *__restrict__ wpre,
                                   const uint32_t *__restrict__ img_all, int64_t n, const uint32_t *__restrict__ cls_off,
                                   const uint32_t *__restrict__ pad_off, uint32_t *__restrict__ slot_img,
                                   uint16_t *__restrict__ slot_tp, uint32_t *__restrict__ q_of_det) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint32_t u = order[v];
    const int c = d.cls(u);
    const uint32_t slot = pad_off[c] + (wpre[v] - cls_off[c]);
    q_of_det[u] = slot;       // weak: its own slot; strong: the slot it would be inserted in front of
    if (u < d.Dw) {
        slot_img[slot] = img_all[u];
        slot_tp[slot] = d.w_tp[u];
    }
}

// one warp per chunk
__global__ void event_bits_kernel(const uint16_t *__restrict__ slot_tp, int64_t nchunks, uint32_t *__restrict__ evbits,
                                  uint32_t *__restrict__ evcnt) {
    int64_t ch = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ch >= nchunks) return;
    const int lane = threadIdx.x & 31;
    unsigned b = __ballot_sync(kFull, slot_tp[ch * 32 + lane] != 0);
    if (lane == 0) { evbits[ch] = b; evcnt[ch] = __popc(b); }
}

__global__ void event_mask_kernel(const uint16_t *__restrict__ slot_tp, int64_t nchunks,
                                  const uint32_t *__restrict__ evbits, const uint32_t *__restrict__ evbase,
                                  uint16_t *__restrict__ evmask) {
    int64_t ch = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ch >= nchunks) return;
    const int lane = threadIdx.x & 31;
    const unsigned b = evbits[ch];
    if ((b >> lane) & 1u) evmask[evbase[ch] + __popc(b & ((1u << lane) - 1u))] = slot_tp[ch * 32 + lane];
}

__global__ void gather_seg_ev0_kernel(const int32_t *__restrict__ seg_chunk0, int64_t S, const uint32_t *__restrict__ evbase,
                                      uint32_t *__restrict__ seg_ev0) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) seg_ev0[s] = evbase[seg_chunk0[s]];
}

// Image-major order (key = image * 2 + is_strong): image i occupies [w_off[i] + s_off[i], ...), its weak
// rows first.  Fills the own lists (aligned with w_off / s_off) and remembers each row's own position.
__global__ void own_fill_kernel(const Dets d, const uint32_t *__restrict__ order, const uint32_t *__restrict__ img_all,
                                int64_t n, const int64_t *__restrict__ w_off, const int64_t *__restrict__ s_off,
                                const uint32_t *__restrict__ q_of_det, uint32_t *__restrict__ own_w_q,
                                uint16_t *__restrict__ own_w_m, uint16_t *__restrict__ own_w_c, uint32_t *__restrict__ own_s_q,
                                uint16_t *__restrict__ own_s_m, uint16_t *__restrict__ own_s_c, uint32_t *__restrict__ ownpos) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint32_t u = order[v];
    const uint32_t im = img_all[u];
    const int64_t local = v - (w_off[im] + s_off[im]);
    const int64_t nw = w_off[im + 1] - w_off[im];
    if (u < d.Dw) {
        const int64_t pos = w_off[im] + local;
        own_w_q[pos] = q_of_det[u]; own_w_m[pos] = d.w_tp[u]; own_w_c[pos] = (uint16_t)d.w_cls[u];
        ownpos[u] = (uint32_t)pos;
    } else {
        const int64_t pos = s_off[im] + (local - nw);
        own_s_q[pos] = q_of_det[u]; own_s_m[pos] = d.s_tp[u - d.Dw]; own_s_c[pos] = (uint16_t)d.s_cls[u - d.Dw];
        ownpos[u] = (uint32_t)pos;
    }
}

// class start table of every image's own list: cs[img][c] = first local index with class >= c
__global__ void own_class_start_kernel(const uint16_t *__restrict__ own_c, const uint32_t *__restrict__ img_of_row, int64_t n,
                                       const int64_t *__restrict__ off, int64_t C, uint16_t *__restrict__ cs) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t im = img_of_row[i];
    const int c = own_c[i];
    const int64_t a = off[im], b = off[im + 1];
    uint16_t *row = cs + (int64_t)im * (C + 1);
    const int local = (int)(i - a);
    if (i == a)
        for (int cc = 0; cc <= c; ++cc) row[cc] = 0;
    const int cnext = (i + 1 < b) ? (int)own_c[i + 1] : (int)C;
    for (int cc = c + 1; cc <= cnext; ++cc) row[cc] = (uint16_t)(local + 1);
}

// Batch-major order (key = (image / 32) * 2 + is_strong): batch b occupies [w_off[32b] + s_off[32b], ...),
// its weak rows first, each part ascending by query slot.
__global__ void batch_query_kernel(const Dets d, const uint32_t *__restrict__ order, const uint32_t *__restrict__ img_all,
                                   int64_t n, int64_t M, const int64_t *__restrict__ w_off, const int64_t *__restrict__ s_off,
                                   const uint32_t *__restrict__ q_of_det, const uint32_t *__restrict__ ownpos,
                                   uint2 *__restrict__ bq_w, uint2 *__restrict__ bq_s) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint32_t u = order[v];
    const uint32_t im = img_all[u];
    const int64_t i0 = (int64_t)(im >> 5) * 32;
    const int64_t i1 = i0 + 32 < M ? i0 + 32 : M;
    const int64_t local = v - (w_off[i0] + s_off[i0]);
    const int64_t nw = w_off[i1] - w_off[i0];
    const uint2 e = make_uint2(q_of_det[u], ((im & 31u) << 27) | ownpos[u]);
    if (u < d.Dw) bq_w[w_off[i0] + local] = e;
    else bq_s[s_off[i0] + (local - nw)] = e;
}

// bqoff[b][s] = first entry of batch b with q >= first slot of segment s  (s == S: end of the batch)
__global__ void batch_query_offsets_kernel(const uint2 *__restrict__ bq, const int64_t *__restrict__ off, int64_t M,
                                           int64_t nbatch, const int32_t *__restrict__ seg_chunk0, int64_t S,
                                           uint32_t *__restrict__ bqoff) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nbatch * (S + 1)) return;
    const int64_t b = k / (S + 1), s = k % (S + 1);
    const int64_t i0 = b * 32 < M ? b * 32 : M, i1 = (b + 1) * 32 < M ? (b + 1) * 32 : M;
    int64_t lo = off[i0], hi = off[i1];
    if (s < S) {
        const uint32_t slot0 = (uint32_t)seg_chunk0[s] * 32u;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (bq[mid].x < slot0) lo = mid + 1; else hi = mid;
        }
    } else {
        lo = hi;
    }
    bqoff[k] = (uint32_t)lo;
}

__global__ void place_labels_kernel(const uint64_t *__restrict__ keys_sorted, const uint32_t *__restrict__ img_sorted,
                                    int64_t n, const uint32_t *__restrict__ cls_off, const uint32_t *__restrict__ pad_off,
                                    uint32_t *__restrict__ slot_img) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int c = (int)keys_sorted[r];
    slot_img[pad_off[c] + ((uint32_t)r - cls_off[c])] = img_sorted[r];
}

__global__ void label_keys_kernel(const int32_t *__restrict__ cls, const uint32_t *__restrict__ img, int64_t n, int64_t C,
                                  uint64_t *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t *__restrict__ gtcnt) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int c = cls[k];
    keys[k] = (uint64_t)(uint32_t)c;
    vals[k] = img[k];
    if (c >= 0 && c < C) atomicAdd(&gtcnt[(int64_t)img[k] * C + c], 1u);
}

static int bits_for(int64_t n) {  // bits needed to represent values in [0, n)
    int b = 1;
    while (((int64_t)1 << b) < n) ++b;
    return b;
}

// Allocations are stream-ordered (cudaMallocAsync on the device's default pool, whose release
// threshold is raised once so that rebuilding an index does not go back to the OS every time).
struct Builder {
    orie_index *ix;
    cudaStream_t st;
    std::vector<void *> temps;
    ~Builder() {
        cudaStreamSynchronize(st);
        for (void *p : temps) cudaFreeAsync(p, st);
    }
    template <typename Tp>
    int temp(Tp **p, int64_t count) {
        void *q = nullptr;
        ORIE_CUDA(cudaMallocAsync(&q, (size_t)std::max<int64_t>(count, 1) * sizeof(Tp), st));
        temps.push_back(q);
        *p = (Tp *)q;
        return ORIE_OK;
    }
    template <typename Tp>
    int keep(Tp **p, int64_t count) {
        void *q = nullptr;
        size_t bytes = (size_t)std::max<int64_t>(count, 1) * sizeof(Tp);
        ORIE_CUDA(cudaMallocAsync(&q, bytes, st));
        if (ix->n_allocs >= (int)(sizeof(ix->allocs) / sizeof(ix->allocs[0]))) {
            cudaFreeAsync(q, st);
            set_error("orie_index_build: allocation table full");
            return ORIE_EINVAL;
        }
        ix->allocs[ix->n_allocs++] = q;
        ix->device_bytes += (int64_t)bytes;
        *p = (Tp *)q;
        return ORIE_OK;
    }
};

static inline unsigned grid_for(int64_t n, int threads = 256) { return (unsigned)std::max<int64_t>(ceil_div(n, threads), 1); }

// Padded layout of one class-sorted stream on the host: per class padded length, then segments.
struct StreamLayout {
    std::vector<uint32_t> cls_off, pad_off;
    std::vector<int32_t> seg_chunk0, seg_nch, cls_seg0;
    int64_t P = 0;
};

static StreamLayout make_layout(const uint32_t *cnt, int64_t C, int extra_pad, int seg_chunks) {
    StreamLayout L;
    L.cls_off.resize(C + 1);
    L.pad_off.resize(C + 1);
    L.cls_seg0.resize(C + 1);
    uint32_t a = 0, p = 0;
    for (int64_t c = 0; c < C; ++c) {
        L.cls_off[c] = a;
        L.pad_off[c] = p;
        L.cls_seg0[c] = (int32_t)L.seg_chunk0.size();
        const int64_t padlen = round_up((int64_t)cnt[c] + extra_pad, kChunk);
        const int nch = (int)(padlen / kChunk);
        for (int k = 0; k < nch; k += seg_chunks) {
            L.seg_chunk0.push_back((int32_t)(p / kChunk) + k);
            L.seg_nch.push_back(std::min(seg_chunks, nch - k));
        }
        a += cnt[c];
        p += (uint32_t)padlen;
    }
    L.cls_off[C] = a;
    L.pad_off[C] = p;
    L.cls_seg0[C] = (int32_t)L.seg_chunk0.size();
    L.P = p;
    return L;
}

template <typename Tp>
static int upload(Builder &B, Tp **dst, const std::vector<Tp> &src, bool keep) {
    if (keep) ORIE_TRY(B.keep(dst, (int64_t)src.size()));
    else ORIE_TRY(B.temp(dst, (int64_t)src.size()));
    if (!src.empty())
        ORIE_CUDA(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(Tp), cudaMemcpyHostToDevice, B.st));
    return ORIE_OK;
}

static int build(orie_index *ix, const int64_t *w_off, const int32_t *w_cls, const double *w_conf, const uint16_t *w_tp,
                 const int64_t *s_off, const int32_t *s_cls, const double *s_conf, const uint16_t *s_tp,
                 const int64_t *l_off, const int32_t *l_cls, int seg_chunks_req, cudaEvent_t tp_ready, cudaStream_t st) {
    // host staging buffers are declared before the Builder so that they outlive its destructor, which
    // synchronises the stream (asynchronous copies from / into them may still be in flight on error paths)
    const int64_t M = ix->M, C = ix->C;
    std::vector<uint32_t> h_hist(2 * C);
    std::vector<int32_t> cls_order(C);
    StreamLayout LD, LL;
    int32_t h_status = 0;
    uint32_t h_total = 0;
    int64_t tails[3];
    Builder B{ix, st};
    {
        int dev = 0;
        cudaMemPool_t pool;
        uint64_t keep_all = UINT64_MAX;
        ORIE_CUDA(cudaGetDevice(&dev));
        ORIE_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        ORIE_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep_all));
    }
    // ---- sizes: given by the caller; off[M] of each block is read back with the class counts and cross-checked
    const int64_t Dw = ix->Dw, Ds = ix->Ds, G = ix->G;
    if (Dw < 0 || Ds < 0 || G < 0 || Dw >= ((int64_t)1 << 27) || Ds >= ((int64_t)1 << 27) || G >= ((int64_t)1 << 31)) {
        set_error("orie_index_build: row counts out of range (weak %lld, strong %lld, labels %lld; limit 2^27-1 detections per detector)",
                  (long long)Dw, (long long)Ds, (long long)G);
        return ORIE_ELIMIT;
    }
    const int64_t n = Dw + Ds;
    const int64_t Nmax = std::max<int64_t>(std::max(n, G), 1);
    const Dets dets{Dw, Ds, w_cls, s_cls, w_conf, s_conf, w_tp, s_tp};

    ORIE_TRY(B.keep(&ix->w_off, M + 1));
    ORIE_TRY(B.keep(&ix->s_off, M + 1));
    ORIE_CUDA(cudaMemcpyAsync(ix->w_off, w_off, (size_t)(M + 1) * 8, cudaMemcpyDeviceToDevice, st));
    ORIE_CUDA(cudaMemcpyAsync(ix->s_off, s_off, (size_t)(M + 1) * 8, cudaMemcpyDeviceToDevice, st));

    // ---- temporaries
    uint64_t *keys, *keys_tmp;
    uint32_t *vals, *vals_tmp, *img_all, *img_l, *order, *hist, *wpre, *q_of_det, *ownpos, *evcnt, *d_total;
    uint16_t *slot_tp, *own_w_c, *own_s_c;
    int32_t *status;
    char *rscratch, *sscratch;
    ORIE_TRY(B.temp(&keys, Nmax));
    ORIE_TRY(B.temp(&keys_tmp, Nmax));
    ORIE_TRY(B.temp(&vals, Nmax));
    ORIE_TRY(B.temp(&vals_tmp, Nmax));
    ORIE_TRY(B.temp(&img_all, n));
    ORIE_TRY(B.temp(&img_l, G));
    ORIE_TRY(B.temp(&order, n));
    ORIE_TRY(B.temp(&hist, 3 * C));
    ORIE_TRY(B.temp(&wpre, n));
    ORIE_TRY(B.temp(&q_of_det, n));
    ORIE_TRY(B.temp(&ownpos, n));
    ORIE_TRY(B.temp(&own_w_c, Dw));
    ORIE_TRY(B.temp(&own_s_c, Ds));
    ORIE_TRY(B.temp(&status, 1));
    ORIE_TRY(B.temp(&d_total, 1));
    ORIE_TRY(B.temp(&rscratch, (int64_t)radix_scratch_bytes(Nmax)));
    ORIE_TRY(B.temp(&sscratch, (int64_t)scan_scratch_bytes(Nmax)));
    ORIE_CUDA(cudaMemsetAsync(status, 0, 4, st));
    ORIE_CUDA(cudaMemsetAsync(hist, 0, (size_t)(3 * C) * 4, st));

    const int cbits = bits_for(C), ibits = bits_for(M) + 1, bbits = bits_for(ix->nbatch) + 1;

    // ---- image of every row; class histograms (weak / labels drive the layouts, strong is range-checked)
    const size_t hist_smem = C <= kHistSmemBins ? (size_t)C * 4 : 0;
    if (Dw) {
        image_of_row_kernel<<<grid_for(Dw), 256, 0, st>>>(w_off, M, Dw, img_all, status);
        ORIE_LAUNCH_CHECK();
        class_hist_kernel<<<grid_for(Dw, kHistItemsPerBlock), 256, hist_smem, st>>>(w_cls, Dw, C, hist, status);
        ORIE_LAUNCH_CHECK();
    }
    if (Ds) {
        image_of_row_kernel<<<grid_for(Ds), 256, 0, st>>>(s_off, M, Ds, img_all + Dw, status);
        ORIE_LAUNCH_CHECK();
        class_hist_kernel<<<grid_for(Ds, kHistItemsPerBlock), 256, hist_smem, st>>>(s_cls, Ds, C, hist + 2 * C, status);
        ORIE_LAUNCH_CHECK();
    }
    if (G) {
        image_of_row_kernel<<<grid_for(G), 256, 0, st>>>(l_off, M, G, img_l, status);
        ORIE_LAUNCH_CHECK();
        class_hist_kernel<<<grid_for(G, kHistItemsPerBlock), 256, hist_smem, st>>>(l_cls, G, C, hist + C, status);
        ORIE_LAUNCH_CHECK();
    }

    // ---- ONE sort of all detections: confidence desc (64-bit key), then class (stable).  Weak rows precede
    //      strong rows in the input, so on exact confidence ties weak sorts first (stable concatenation order).
    if (n) {
        conf_keys_kernel<<<grid_for(n), 256, 0, st>>>(dets, n, keys, vals);
        ORIE_LAUNCH_CHECK();
        ORIE_TRY(radix_sort_pairs(keys, vals, keys_tmp, vals_tmp, n, 0, 64, rscratch, st));
        rekey_kernel<kKeyClass><<<grid_for(n), 256, 0, st>>>(dets, img_all, vals, n, keys);
        ORIE_LAUNCH_CHECK();
        ORIE_TRY(radix_sort_pairs(keys, vals, keys_tmp, vals_tmp, n, 0, cbits, rscratch, st));
        ORIE_CUDA(cudaMemcpyAsync(order, vals, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        weak_flag_kernel<<<grid_for(n), 256, 0, st>>>(order, n, (uint32_t)Dw, wpre);
        ORIE_LAUNCH_CHECK();
        ORIE_TRY(exclusive_scan_u32(wpre, wpre, n, nullptr, sscratch, st));
    }

    // ---- host: class counts -> padded layouts and segment tables
    ORIE_CUDA(cudaMemcpyAsync(h_hist.data(), hist, (size_t)(2 * C) * 4, cudaMemcpyDeviceToHost, st));
    ORIE_CUDA(cudaMemcpyAsync(&h_status, status, 4, cudaMemcpyDeviceToHost, st));
    ORIE_CUDA(cudaMemcpyAsync(&tails[0], w_off + M, 8, cudaMemcpyDeviceToHost, st));
    ORIE_CUDA(cudaMemcpyAsync(&tails[1], s_off + M, 8, cudaMemcpyDeviceToHost, st));
    ORIE_CUDA(cudaMemcpyAsync(&tails[2], l_off + M, 8, cudaMemcpyDeviceToHost, st));
    ORIE_CUDA(cudaStreamSynchronize(st));
    if (tails[0] != Dw || tails[1] != Ds || tails[2] != G) {
        set_error("orie_index_build: row counts (%lld, %lld, %lld) do not match the offset arrays (%lld, %lld, %lld)",
                  (long long)Dw, (long long)Ds, (long long)G, (long long)tails[0], (long long)tails[1], (long long)tails[2]);
        return ORIE_EDATA;
    }
    if (h_status & 2) {
        set_error("orie_index_build: class id outside [0, %lld)", (long long)C);
        return ORIE_EDATA;
    }
    if (h_status & 1) {
        set_error("orie_index_build: an image has more than 65535 rows in one file");
        return ORIE_ELIMIT;
    }
    int64_t raw_chunks = 0;
    for (int64_t c = 0; c < C; ++c) raw_chunks += ceil_div((int64_t)h_hist[c] + 1, kChunk);
    // Segment length: about two segments per average class (measured best for both the walk's load balance
    // and the AP sweep on COCO-shaped data, profiles/), enough (batch, segment) warp items for >= 4 waves of
    // 148 SMs x 8 warps, and no segment so long that a single warp becomes the tail of the walk.
    const int64_t seg_target = std::max<int64_t>(2 * C, ceil_div(4 * 148 * 8, ix->nbatch));
    int seg_chunks = seg_chunks_req > 0 ? seg_chunks_req
                                        : (int)std::min<int64_t>(512, std::max<int64_t>(16, ceil_div(raw_chunks, seg_target)));
    ix->seg_chunks = seg_chunks;
    LD = make_layout(h_hist.data(), C, 1, seg_chunks);
    LL = make_layout(h_hist.data() + C, C, 0, seg_chunks);
    ix->P = LD.P; ix->nchunks = LD.P / kChunk; ix->S = (int64_t)LD.seg_chunk0.size();
    ix->PL = LL.P; ix->nchunksL = LL.P / kChunk; ix->SL = (int64_t)LL.seg_chunk0.size();
    if (ix->P >= ((int64_t)1 << 31)) {
        set_error("orie_index_build: %lld slots exceed 2^31-1", (long long)ix->P);
        return ORIE_ELIMIT;
    }
    for (int64_t c = 0; c < C; ++c) cls_order[c] = (int32_t)c;
    std::stable_sort(cls_order.begin(), cls_order.end(), [&](int32_t a, int32_t b) { return h_hist[a] > h_hist[b]; });
    uint32_t *d_cls_off, *d_pad_off, *d_lcls_off, *d_lpad_off;
    ORIE_TRY(upload(B, &d_cls_off, LD.cls_off, false));
    ORIE_TRY(upload(B, &d_pad_off, LD.pad_off, false));
    ORIE_TRY(upload(B, &d_lcls_off, LL.cls_off, false));
    ORIE_TRY(upload(B, &d_lpad_off, LL.pad_off, false));
    ORIE_TRY(upload(B, &ix->seg_chunk0, LD.seg_chunk0, true));
    ORIE_TRY(upload(B, &ix->seg_nch, LD.seg_nch, true));
    ORIE_TRY(upload(B, &ix->cls_seg0, LD.cls_seg0, true));
    ORIE_TRY(upload(B, &ix->cls_order, cls_order, true));
    ORIE_TRY(upload(B, &ix->lseg_chunk0, LL.seg_chunk0, true));
    ORIE_TRY(upload(B, &ix->lseg_nch, LL.seg_nch, true));
    ORIE_TRY(upload(B, &ix->lcls_seg0, LL.cls_seg0, true));
    // the host vectors stay alive until the end of this function, which ends with a stream synchronisation

    // ---- slots, strong insertion slots
    ORIE_TRY(B.keep(&ix->slot_img, ix->P));
    ORIE_TRY(B.temp(&slot_tp, ix->P));
    ORIE_TRY(B.temp(&evcnt, ix->nchunks));
    ORIE_TRY(B.keep(&ix->evbits, ix->nchunks));
    ORIE_TRY(B.keep(&ix->evbase, ix->nchunks));
    ORIE_TRY(B.keep(&ix->seg_ev0, ix->S));
    fill_u32_kernel<<<grid_for(ix->P), 256, 0, st>>>(ix->slot_img, ix->P, (uint32_t)M);
    ORIE_LAUNCH_CHECK();
    ORIE_CUDA(cudaMemsetAsync(slot_tp, 0, (size_t)ix->P * 2, st));
    if (tp_ready) ORIE_CUDA(cudaStreamWaitEvent(st, tp_ready, 0));     // first reader of the true-positive masks
    if (n) {
        place_slots_kernel<<<grid_for(n), 256, 0, st>>>(dets, order, wpre, img_all, n, d_cls_off, d_pad_off, ix->slot_img,
                                                      slot_tp, q_of_det);
        ORIE_LAUNCH_CHECK();
    }

    // ---- events (the total is read back together with the final synchronisation; evmask is sized by its bound)
    event_bits_kernel<<<grid_for(ix->nchunks * 32), 256, 0, st>>>(slot_tp, ix->nchunks, ix->evbits, evcnt);
    ORIE_LAUNCH_CHECK();
    ORIE_TRY(exclusive_scan_u32(evcnt, ix->evbase, ix->nchunks, d_total, sscratch, st));
    ORIE_CUDA(cudaMemcpyAsync(&h_total, d_total, 4, cudaMemcpyDeviceToHost, st));
    ORIE_TRY(B.keep(&ix->evmask, Dw));             // events <= weak detections
    event_mask_kernel<<<grid_for(ix->nchunks * 32), 256, 0, st>>>(slot_tp, ix->nchunks, ix->evbits, ix->evbase, ix->evmask);
    ORIE_LAUNCH_CHECK();
    gather_seg_ev0_kernel<<<grid_for(ix->S), 256, 0, st>>>(ix->seg_chunk0, ix->S, ix->evbase, ix->seg_ev0);
    ORIE_LAUNCH_CHECK();

    // ---- own lists (image-major) and batch query lists (batch-major), both detectors in one pass each
    ORIE_TRY(B.keep(&ix->own_w_q, Dw));
    ORIE_TRY(B.keep(&ix->own_w_m, Dw));
    ORIE_TRY(B.keep(&ix->own_s_q, Ds));
    ORIE_TRY(B.keep(&ix->own_s_m, Ds));
    ORIE_TRY(B.keep(&ix->own_w_cs, M * (C + 1)));
    ORIE_TRY(B.keep(&ix->own_s_cs, M * (C + 1)));
    ORIE_TRY(B.keep(&ix->bq_w, Dw));
    ORIE_TRY(B.keep(&ix->bq_s, Ds));
    ORIE_TRY(B.keep(&ix->bqoff_w, ix->nbatch * (ix->S + 1)));
    ORIE_TRY(B.keep(&ix->bqoff_s, ix->nbatch * (ix->S + 1)));
    ORIE_CUDA(cudaMemsetAsync(ix->own_w_cs, 0, (size_t)(M * (C + 1)) * 2, st));
    ORIE_CUDA(cudaMemsetAsync(ix->own_s_cs, 0, (size_t)(M * (C + 1)) * 2, st));
    if (n) {
        ORIE_CUDA(cudaMemcpyAsync(vals, order, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        rekey_kernel<kKeyImageDetector><<<grid_for(n), 256, 0, st>>>(dets, img_all, vals, n, keys);
        ORIE_LAUNCH_CHECK();
        ORIE_TRY(radix_sort_pairs(keys, vals, keys_tmp, vals_tmp, n, 0, ibits, rscratch, st));
        own_fill_kernel<<<grid_for(n), 256, 0, st>>>(dets, vals, img_all, n, ix->w_off, ix->s_off, q_of_det, ix->own_w_q,
                                                   ix->own_w_m, own_w_c, ix->own_s_q, ix->own_s_m, own_s_c, ownpos);
        ORIE_LAUNCH_CHECK();
        if (Dw) {
            own_class_start_kernel<<<grid_for(Dw), 256, 0, st>>>(own_w_c, img_all, Dw, ix->w_off, C, ix->own_w_cs);
            ORIE_LAUNCH_CHECK();
        }
        if (Ds) {
            own_class_start_kernel<<<grid_for(Ds), 256, 0, st>>>(own_s_c, img_all + Dw, Ds, ix->s_off, C, ix->own_s_cs);
            ORIE_LAUNCH_CHECK();
        }
        ORIE_CUDA(cudaMemcpyAsync(vals, order, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        rekey_kernel<kKeyBatchDetector><<<grid_for(n), 256, 0, st>>>(dets, img_all, vals, n, keys);
        ORIE_LAUNCH_CHECK();
        ORIE_TRY(radix_sort_pairs(keys, vals, keys_tmp, vals_tmp, n, 0, bbits, rscratch, st));
        batch_query_kernel<<<grid_for(n), 256, 0, st>>>(dets, vals, img_all, n, M, ix->w_off, ix->s_off, q_of_det, ownpos,
                                                      ix->bq_w, ix->bq_s);
        ORIE_LAUNCH_CHECK();
    }
    batch_query_offsets_kernel<<<grid_for(ix->nbatch * (ix->S + 1)), 256, 0, st>>>(ix->bq_w, ix->w_off, M, ix->nbatch,
                                                                                 ix->seg_chunk0, ix->S, ix->bqoff_w);
    ORIE_LAUNCH_CHECK();
    batch_query_offsets_kernel<<<grid_for(ix->nbatch * (ix->S + 1)), 256, 0, st>>>(ix->bq_s, ix->s_off, M, ix->nbatch,
                                                                                 ix->seg_chunk0, ix->S, ix->bqoff_s);
    ORIE_LAUNCH_CHECK();

    // ---- label stream
    ORIE_TRY(B.keep(&ix->lab_slot_img, ix->PL));
    ORIE_TRY(B.keep(&ix->gtcnt, M * C));
    ORIE_CUDA(cudaMemsetAsync(ix->gtcnt, 0, (size_t)(M * C) * 4, st));
    if (ix->PL) {
        fill_u32_kernel<<<grid_for(ix->PL), 256, 0, st>>>(ix->lab_slot_img, ix->PL, (uint32_t)M);
        ORIE_LAUNCH_CHECK();
    }
    if (G) {
        label_keys_kernel<<<grid_for(G), 256, 0, st>>>(l_cls, img_l, G, C, keys, vals, ix->gtcnt);
        ORIE_LAUNCH_CHECK();
        ORIE_TRY(radix_sort_pairs(keys, vals, keys_tmp, vals_tmp, G, 0, cbits, rscratch, st));
        place_labels_kernel<<<grid_for(G), 256, 0, st>>>(keys, vals, G, d_lcls_off, d_lpad_off, ix->lab_slot_img);
        ORIE_LAUNCH_CHECK();
    }
    ORIE_CUDA(cudaStreamSynchronize(st));
    ix->Ev = h_total;
    return ORIE_OK;
}

}  // namespace orie

using namespace orie;

extern "C" int orie_index_build(int64_t M, int64_t C, int T, int64_t num_weak, int64_t num_strong, int64_t num_labels,
                                const int64_t *w_off, const int32_t *w_cls, const double *w_conf, const uint16_t *w_tp,
                                const int64_t *s_off, const int32_t *s_cls, const double *s_conf, const uint16_t *s_tp,
                                const int64_t *l_off, const int32_t *l_cls, int seg_chunks, orie_event_t tp_ready,
                                orie_stream_t stream, orie_index_t **out) {
    if (!out) {
        set_error("orie_index_build: out is NULL");
        return ORIE_EINVAL;
    }
    *out = nullptr;
    if (M < 1 || C < 1 || !w_off || !s_off || !l_off) {
        set_error("orie_index_build: need M >= 1, C >= 1 and the three offset arrays");
        return ORIE_EINVAL;
    }
    if (T < 1 || T > ORIE_MAX_THRESHOLDS) {
        set_error("orie_index_build: T=%d outside [1,%d]", T, ORIE_MAX_THRESHOLDS);
        return ORIE_ELIMIT;
    }
    if (M >= ((int64_t)1 << 27) || C > 65535) {
        set_error("orie_index_build: M=%lld or C=%lld exceeds the engine limits (2^27-1 images, 65535 classes)",
                  (long long)M, (long long)C);
        return ORIE_ELIMIT;
    }
    orie_index *ix = new orie_index();
    ix->M = M; ix->C = C; ix->T = T;
    ix->Dw = num_weak; ix->Ds = num_strong; ix->G = num_labels;
    ix->stream = stream;
    ix->nbatch = ceil_div(M, 32);
    ix->ens_words = ceil_div(M + 1, 32);
    ix->cls_per_warp = 32 / T;
    ix->class_groups = ceil_div(C, ix->cls_per_warp);
    int rc = build(ix, w_off, w_cls, w_conf, w_tp, s_off, s_cls, s_conf, s_tp, l_off, l_cls, seg_chunks, tp_ready, stream);
    if (rc != ORIE_OK) {
        orie_index_destroy(ix);
        return rc;
    }
    *out = ix;
    return ORIE_OK;
}

extern "C" void orie_index_destroy(orie_index_t *ix) {
    if (!ix) return;
    for (int i = 0; i < ix->n_allocs; ++i) cudaFreeAsync(ix->allocs[i], ix->stream);
    delete ix;
}

extern "C" int orie_index_info(const orie_index_t *ix, orie_index_info_t *info) {
    if (!ix || !info) {
        set_error("orie_index_info: null argument");
        return ORIE_EINVAL;
    }
    info->num_images = ix->M; info->num_classes = ix->C;
    info->num_thresholds = ix->T; info->seg_chunks = ix->seg_chunks;
    info->num_weak = ix->Dw; info->num_strong = ix->Ds; info->num_labels = ix->G;
    info->slots = ix->P; info->segments = ix->S; info->events = ix->Ev;
    info->label_slots = ix->PL; info->label_segments = ix->SL;
    info->class_groups = ix->class_groups;
    info->ens_words = ix->ens_words;
    info->device_bytes = ix->device_bytes;
    return ORIE_OK;
}
