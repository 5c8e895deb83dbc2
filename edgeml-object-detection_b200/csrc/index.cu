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
#include <mutex>
#include <vector>

#include "coop.cuh"
#include "index.cuh"

namespace orie {

// ----------------------------------------------------------------------------
// kernels.  "u" is a combined detection id: [0, Dw) weak rows, [Dw, Dw+Ds) strong rows.
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
};

// ---- prep: one warp per (image, block) with block in {weak, strong, labels}.  Writes the image of every
// row, the confidence sort keys, the class histograms (weak / labels drive the layouts, strong is only
// range-checked), the per-image ground-truth class counts, and validates the inputs.
//   status bits (IndexMeta::status): kStatusRows, kStatusClass, kStatusCounts
struct PrepArgs {
    int64_t M, C, Dw, Ds, G;
    const int64_t *w_off, *s_off, *l_off;
    const int32_t *w_cls, *s_cls, *l_cls;
    const double *w_conf, *s_conf;
    uint64_t *keys;      // [Dw + Ds]
    uint32_t *img_all;   // [Dw + Ds]
    uint32_t *img_l;     // [G]
    uint32_t *hist;      // [3][C]: weak, labels, strong
    uint32_t *gtcnt;     // [M][C], zero on entry
    unsigned long long *key_and_or;   // [2]: AND (starts all ones) and OR (starts zero) of all keys
    uint32_t *status;
    int smem_bins;       // 3 * C if the block-local histogram fits in shared memory, else 0
};
constexpr int kPrepThreads = 256;
constexpr int kPrepMaxSmemBins = 12288;   // 48 KB

__global__ void __launch_bounds__(kPrepThreads) prep_kernel(const PrepArgs a) {
    extern __shared__ uint32_t bins[];
    __shared__ unsigned long long s_and[kPrepThreads / 32], s_or[kPrepThreads / 32];
    const int lane = threadIdx.x & 31;
    unsigned long long k_and = ~0ull, k_or = 0ull;
    for (int i = threadIdx.x; i < a.smem_bins; i += kPrepThreads) bins[i] = 0;
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0 && (a.w_off[a.M] != a.Dw || a.s_off[a.M] != a.Ds || a.l_off[a.M] != a.G))
        atomicOr(a.status, kStatusCounts);
    const int64_t nwarps = (int64_t)gridDim.x * (kPrepThreads / 32);
    for (int64_t w = (int64_t)blockIdx.x * (kPrepThreads / 32) + (threadIdx.x >> 5); w < 3 * a.M; w += nwarps) {
        const int blk = (int)(w / a.M);            // 0 weak, 1 labels, 2 strong (the order of the histograms)
        const int64_t im = w - blk * a.M;
        const int64_t *off = blk == 0 ? a.w_off : (blk == 1 ? a.l_off : a.s_off);
        const int32_t *cls = blk == 0 ? a.w_cls : (blk == 1 ? a.l_cls : a.s_cls);
        const int64_t rows = blk == 0 ? a.Dw : (blk == 1 ? a.G : a.Ds);
        int64_t r0 = off[im], r1 = off[im + 1];
        if (r0 < 0 || r1 < r0 || r1 > rows) {        // inconsistent offsets: flag, and never read outside the arrays
            if (lane == 0) atomicOr(a.status, kStatusCounts);
            r1 = r1 < 0 ? 0 : (r1 > rows ? rows : r1);
            r0 = r0 < 0 ? 0 : (r0 > r1 ? r1 : r0);
        }
        if (r1 - r0 > 65535 && lane == 0) atomicOr(a.status, kStatusRows);
        for (int64_t r = r0 + lane; r < r1; r += 32) {
            const int c = cls[r];
            const bool ok = c >= 0 && c < a.C;
            if (!ok) atomicOr(a.status, kStatusClass);
            else if (a.smem_bins) atomicAdd(&bins[blk * a.C + c], 1u);
            else atomicAdd(&a.hist[blk * a.C + c], 1u);
            if (blk == 1) {
                a.img_l[r] = (uint32_t)im;
                if (ok) atomicAdd(&a.gtcnt[im * a.C + c], 1u);
            } else {
                const int64_t u = blk == 0 ? r : a.Dw + r;
                const uint64_t key = conf_desc_key(blk == 0 ? a.w_conf[r] : a.s_conf[r]);
                a.keys[u] = key;
                k_and &= key; k_or |= key;
                a.img_all[u] = (uint32_t)im;
            }
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        k_and &= __shfl_xor_sync(kFull, k_and, d);
        k_or |= __shfl_xor_sync(kFull, k_or, d);
    }
    if (lane == 0) { s_and[threadIdx.x >> 5] = k_and; s_or[threadIdx.x >> 5] = k_or; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kPrepThreads / 32; ++w) { k_and &= s_and[w]; k_or |= s_or[w]; }
        if (k_and != ~0ull) atomicAnd(a.key_and_or, k_and);
        if (k_or != 0ull) atomicOr(a.key_and_or + 1, k_or);
    }
    for (int i = threadIdx.x; i < a.smem_bins; i += kPrepThreads)
        if (bins[i]) atomicAdd(&a.hist[i], bins[i]);
}

// ---- layout: everything that follows from the class histograms, on the device (one CTA; the host used to read the
// histograms back, lay the classes out and upload the tables — a synchronisation in the middle of every build).
//   per class and stream: first row (cls_off), first slot of its padded range (pad_off), first segment (cls_seg0);
//   per segment: first chunk and length in chunks; IndexMeta: slots / chunks / segments of both streams.
// Detection stream: every class is padded to whole 32-slot chunks with at least one padding slot; label stream:
// padded to whole chunks (a class without labels has no chunk and no segment).
struct LayoutArgs {
    int64_t C, S_cap, SL_cap;
    int seg_chunks;
    const uint32_t *hist;      // [3][C]: weak, labels, strong
    uint32_t *cls_off, *pad_off, *lcls_off, *lpad_off;          // [C+1]
    int32_t *cls_seg0, *lcls_seg0;                              // [C+1]
    int32_t *seg_chunk0, *seg_nch, *lseg_chunk0, *lseg_nch;     // [S_cap], [SL_cap]
    int32_t *seg_order;        // [S_cap] detection segments, longest first (the walk hands them out in this order)
    IndexMeta *meta;
};
constexpr int kLayoutThreads = 1024;

__global__ void __launch_bounds__(kLayoutThreads) layout_kernel(const LayoutArgs a) {
    __shared__ uint32_t ws[kLayoutThreads / 32];
    const int tid = threadIdx.x;
    uint32_t carry[2][3] = {{0, 0, 0}, {0, 0, 0}};      // rows, slots, segments in front of this tile, per stream
    for (int64_t base = 0; base < a.C; base += kLayoutThreads) {
        const int64_t c = base + tid;
        const bool live = c < a.C;
#pragma unroll
        for (int st = 0; st < 2; ++st) {                // 0: detections (extra padding slot), 1: labels
            const uint32_t cnt = live ? a.hist[st * a.C + c] : 0u;
            const uint32_t nch = live ? (cnt + (st == 0 ? 1u : 0u) + 31u) / 32u : 0u;
            const uint32_t segs = (nch + (uint32_t)a.seg_chunks - 1u) / (uint32_t)a.seg_chunks;
            const uint32_t seglen = segs ? (nch + segs - 1u) / segs : 0u;     // equal-length segments: no short leftover
            uint32_t t_cnt, t_pad, t_seg;
            const uint32_t x_cnt = carry[st][0] + block_exclusive_scan<kLayoutThreads>(cnt, ws, &t_cnt);
            const uint32_t x_pad = carry[st][1] + block_exclusive_scan<kLayoutThreads>(nch * 32u, ws, &t_pad);
            const uint32_t x_seg = carry[st][2] + block_exclusive_scan<kLayoutThreads>(segs, ws, &t_seg);
            carry[st][0] += t_cnt; carry[st][1] += t_pad; carry[st][2] += t_seg;
            if (live) {
                (st == 0 ? a.cls_off : a.lcls_off)[c] = x_cnt;
                (st == 0 ? a.pad_off : a.lpad_off)[c] = x_pad;
                (st == 0 ? a.cls_seg0 : a.lcls_seg0)[c] = (int32_t)x_seg;
                int32_t *chunk0 = st == 0 ? a.seg_chunk0 : a.lseg_chunk0, *len = st == 0 ? a.seg_nch : a.lseg_nch;
                const int64_t cap = st == 0 ? a.S_cap : a.SL_cap;
                for (uint32_t k = 0; k < segs; ++k) {
                    if ((int64_t)(x_seg + k) >= cap) break;        // cannot happen: the capacities are upper bounds
                    chunk0[x_seg + k] = (int32_t)(x_pad / 32u + k * seglen);
                    len[x_seg + k] = (int32_t)min(seglen, nch - k * seglen);
                }
            }
        }
    }
    if (tid == 0) {
        a.cls_off[a.C] = carry[0][0]; a.pad_off[a.C] = carry[0][1]; a.cls_seg0[a.C] = (int32_t)carry[0][2];
        a.lcls_off[a.C] = carry[1][0]; a.lpad_off[a.C] = carry[1][1]; a.lcls_seg0[a.C] = (int32_t)carry[1][2];
        a.meta->P = carry[0][1]; a.meta->nchunks = carry[0][1] / 32u; a.meta->S = carry[0][2];
        a.meta->PL = carry[1][1]; a.meta->nchunksL = carry[1][1] / 32u; a.meta->SL = carry[1][2];
    }
    // detection segments by descending length (counting sort over the <= 2047 possible lengths): the walk's warps take
    // segments from the front of this list, so the long ones are under way first and the short ones fill the end
    __shared__ uint32_t bins[2048];
    const int64_t S = min((int64_t)carry[0][2], a.S_cap);
    for (int b = tid; b < 2048; b += kLayoutThreads) bins[b] = 0;
    __syncthreads();                               // also orders the segment tables written above before the reads below
    for (int64_t sgm = tid; sgm < S; sgm += kLayoutThreads) atomicAdd(&bins[2047 - min(a.seg_nch[sgm], 2047)], 1u);
    __syncthreads();
    {
        const uint32_t c0 = bins[2 * tid], c1 = bins[2 * tid + 1];
        const uint32_t x = block_exclusive_scan<kLayoutThreads>(c0 + c1, ws, nullptr);
        bins[2 * tid] = x; bins[2 * tid + 1] = x + c0;
    }
    __syncthreads();
    for (int64_t sgm = tid; sgm < S; sgm += kLayoutThreads)
        a.seg_order[atomicAdd(&bins[2047 - min(a.seg_nch[sgm], 2047)], 1u)] = (int32_t)sgm;
}

// ----------------------------------------------------------------------------
// Everything after the sorts, in ONE cooperative launch (four phases separated by grid barriers; eight dependent
// launches of a few microseconds each before, whose enqueue cost exceeded their run time):
//   A  padding slots of both streams (image M = never a member, no true positive); position v of the combined
//      (class, conf desc) order -> slot (weak) / insertion slot (strong) [wpre[v] = weak detections sorted before v];
//      labels into the class-major label stream (only the grouping by class matters — the label walk counts members
//      per class — so a label takes the next free slot of its class); classes ranked by weak-detection count
//      (cls_order, by counting)
//   B  event counts per CTA range of chunks (event = slot holding a true positive);
//      own lists: the rows of one image and detector in (class, conf) order, i.e. ascending by their position in the
//      global order — an image has a few hundred rows at most in practice, so one warp ranks them by counting instead
//      of a dataset-wide regrouping sort — and the class-start table of the list: cs[img][c] = first entry with
//      class >= c, c in [0, C];
//      bqoff[b][s] = first query of batch b with slot >= first slot of segment s (searched through the batch-major
//      order and the slots of phase A)
//   C  events in front of every chunk / segment (exclusive scan) and their total; per-batch query list: position v
//      of the batch-major order IS the entry, ascending by slot, weak and strong interleaved
//   D  per image the classes in which it has a detection from either detector, deepest AP sweep first (act_cls)
// The sizes that depend on the data (slots, chunks, segments) are read from IndexMeta, written by layout_kernel.
// Arrays written in one phase and read in a later one are read with __ldcg (L2), never through the read-only path.
// ----------------------------------------------------------------------------
struct PostArgs {
    Dets d;
    int64_t n, M, C, G, nbatch, S_cap;
    const uint32_t *order, *wpre, *img_all, *ord_bat, *img_l, *hist;
    const int32_t *l_cls, *seg_chunk0;
    const uint32_t *cls_off, *pad_off, *lcls_off, *lpad_off;
    const int64_t *w_off, *s_off;
    uint32_t *lcursor;
    uint32_t *slot_img, *lab_slot_img, *slot_pk, *ev_img;
    uint16_t *slot_tp, *ev_mask;
    uint32_t *q_of_det, *pos_of_det, *ownpos;
    uint32_t *own_w_q, *own_s_q;
    uint16_t *own_w_m, *own_s_m, *own_w_c, *own_s_c, *own_w_cs, *own_s_cs;
    int32_t *cls_order;
    uint16_t *act_cls;
    uint32_t *nact, *evbase, *act_key;
    uint16_t *act_tmp;
    uint2 *bq;
    uint32_t *bqoff, *seg_ev0, *table;
    IndexMeta *meta;
    unsigned *bar;
};
constexpr int kPostThreads = 256;
constexpr int kPostWarps = kPostThreads / 32;
constexpr int kRankStage = 512;          // rows of an image staged in shared memory for the ranking

__global__ void __launch_bounds__(kPostThreads) post_kernel(const PostArgs a) {
    __shared__ uint32_t ws[kPostWarps];
    __shared__ uint32_t staged[kPostWarps][kRankStage];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t gtid = (int64_t)blockIdx.x * kPostThreads + threadIdx.x;
    const int64_t gsize = (int64_t)gridDim.x * kPostThreads;
    const int64_t gwarp = gtid >> 5, nwarps = gsize >> 5;
    unsigned epoch = 0;
    // invalid input (flagged by prep_kernel): class ids may be out of range, nothing below is safe — every CTA leaves
    if (a.meta->status & (kStatusRows | kStatusClass | kStatusCounts)) return;
    const int64_t nchunks = a.meta->nchunks, S = a.meta->S;

    // ---- A
    for (int64_t k = gtid; k < a.C * 64; k += gsize) {           // at most 32 padding slots per class and stream
        const int64_t c = k >> 6, i = k & 63;
        const uint32_t slot = a.pad_off[c] + (a.cls_off[c + 1] - a.cls_off[c]) + (uint32_t)i;
        if (slot < a.pad_off[c + 1]) {
            a.slot_img[slot] = (uint32_t)a.M;
            a.slot_tp[slot] = 0;
            if (a.slot_pk) a.slot_pk[slot] = (uint32_t)a.M;
        }
        const uint32_t lslot = a.lpad_off[c] + (a.lcls_off[c + 1] - a.lcls_off[c]) + (uint32_t)i;
        if (lslot < a.lpad_off[c + 1]) a.lab_slot_img[lslot] = (uint32_t)a.M;
    }
    for (int64_t v = gtid; v < a.n; v += gsize) {
        const uint32_t u = a.order[v];
        const int c = a.d.cls(u);
        const uint32_t slot = a.pad_off[c] + (a.wpre[v] - a.cls_off[c]);
        a.q_of_det[u] = slot;       // weak: its own slot; strong: the slot it would be inserted in front of
        a.pos_of_det[u] = (uint32_t)v;
        if (u < a.d.Dw) {
            const uint32_t im = a.img_all[u];
            const uint16_t tp = a.d.w_tp[u];
            a.slot_img[slot] = im;
            a.slot_tp[slot] = tp;
            if (a.slot_pk) a.slot_pk[slot] = im | ((uint32_t)tp << 16);
        }
    }
    for (int64_t g = gtid; g < a.G; g += gsize) {
        const int c = a.l_cls[g];
        a.lab_slot_img[a.lpad_off[c] + atomicAdd(&a.lcursor[c], 1u)] = a.img_l[g];
    }
    for (int64_t c = gtid; c < a.C; c += gsize) {                // stable rank by descending weak-detection count
        const uint32_t h = a.hist[c];
        uint32_t rank = 0;
        for (int64_t o = 0; o < a.C; ++o) {
            const uint32_t ho = a.hist[o];
            rank += (ho > h || (ho == h && o < c)) ? 1u : 0u;
        }
        a.cls_order[rank] = (int32_t)c;
    }
    grid_sync(a.bar, epoch);

    // ---- B: events of this CTA's chunk range, split over its warps
    const int64_t ev_per = (nchunks + gridDim.x - 1) / gridDim.x;
    const int64_t c0 = min((int64_t)blockIdx.x * ev_per, nchunks), c1 = min(c0 + ev_per, nchunks);
    const int64_t cw = (ev_per + kPostWarps - 1) / kPostWarps;
    const int64_t w0 = min(c0 + warp * cw, c1), w1 = min(w0 + cw, c1);
    uint32_t mine = 0;
    for (int64_t ch = w0; ch < w1; ++ch) mine += __popc(__ballot_sync(kFull, __ldcg(a.slot_tp + ch * 32 + lane) != 0));
    uint32_t cta_total;
    const uint32_t excl = block_exclusive_scan<kPostThreads>(lane == 0 ? mine : 0u, ws, &cta_total);
    const uint32_t warp_base = __shfl_sync(kFull, excl, 0);
    if (threadIdx.x == 0) __stcg(a.table + blockIdx.x, cta_total);
    // own lists and their class-start tables: one warp per (detector, image)
    for (int64_t w = gwarp; w < 2 * a.M; w += nwarps) {
        const bool strong = w >= a.M;
        const int64_t im = strong ? w - a.M : w;
        const int64_t *off = strong ? a.s_off : a.w_off;
        const int64_t r0 = off[im];
        const int n = (int)(off[im + 1] - r0);
        const uint32_t u0 = (uint32_t)(strong ? a.d.Dw + r0 : r0);
        const uint32_t *pos = a.pos_of_det + u0;
        uint32_t *own_q = strong ? a.own_s_q : a.own_w_q;
        uint16_t *own_m = strong ? a.own_s_m : a.own_w_m, *own_c = strong ? a.own_s_c : a.own_w_c;
        const uint16_t *tp = strong ? a.d.s_tp : a.d.w_tp;
        const int32_t *cls = strong ? a.d.s_cls : a.d.w_cls;
        const bool fits = n <= kRankStage;      // the usual case: the image's positions are staged once
        __syncwarp();
        if (fits)
            for (int k = lane; k < n; k += 32) staged[warp][k] = __ldcg(pos + k);
        __syncwarp();
        for (int rb = 0; rb < n; rb += 32) {     // uniform trip count: the staging below is warp-wide
            const int r = rb + lane;
            const uint32_t me = r < n ? (fits ? staged[warp][r] : __ldcg(pos + r)) : 0u;
            int rank = 0;
            if (fits) {
                for (int k = 0; k < n; ++k) rank += staged[warp][k] < me;
            } else {
                for (int t0 = 0; t0 < n; t0 += kRankStage) {          // larger images: tile by tile through the stage
                    const int cnt = n - t0 < kRankStage ? n - t0 : kRankStage;
                    __syncwarp();
                    for (int k = lane; k < cnt; k += 32) staged[warp][k] = __ldcg(pos + t0 + k);
                    __syncwarp();
                    for (int k = 0; k < cnt; ++k) rank += staged[warp][k] < me;
                }
            }
            if (r < n) {
                const int64_t at = r0 + rank;
                own_q[at] = __ldcg(a.q_of_det + u0 + r);
                own_m[at] = tp[r0 + r];
                own_c[at] = (uint16_t)cls[r0 + r];
                a.ownpos[u0 + r] = (uint32_t)at;
            }
        }
        __syncwarp();                      // the list of this image is complete (written by this warp)
        const uint16_t *list = own_c + r0;
        uint16_t *row = (strong ? a.own_s_cs : a.own_w_cs) + im * (a.C + 1);
        for (int c = lane; c <= a.C; c += 32) {
            int lo = 0, hi = n;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if ((int)list[mid] < c) lo = mid + 1; else hi = mid;
            }
            row[c] = (uint16_t)lo;
        }
    }
    for (int64_t k = gtid; k < a.nbatch * (a.S_cap + 1); k += gsize) {
        const int64_t b = k / (a.S_cap + 1), s = k % (a.S_cap + 1);
        const int64_t i0 = b * 32 < a.M ? b * 32 : a.M, i1 = (b + 1) * 32 < a.M ? (b + 1) * 32 : a.M;
        int64_t lo = a.w_off[i0] + a.s_off[i0], hi = a.w_off[i1] + a.s_off[i1];
        if (s < S) {
            const uint32_t slot0 = (uint32_t)a.seg_chunk0[s] * 32u;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (__ldcg(a.q_of_det + a.ord_bat[mid]) < slot0) lo = mid + 1; else hi = mid;
            }
        } else {
            lo = hi;
        }
        a.bqoff[k] = (uint32_t)lo;
    }
    grid_sync(a.bar, epoch);

    // ---- C
    {
        uint32_t before = 0, all = 0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += kPostThreads) {
            const uint32_t v = __ldcg(a.table + b);
            all += v;
            before += b < (int)blockIdx.x ? v : 0u;
        }
        uint32_t carry, grand;
        block_exclusive_scan<kPostThreads>(before, ws, &carry);
        block_exclusive_scan<kPostThreads>(all, ws, &grand);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            a.meta->Ev = grand;
            a.seg_ev0[S] = grand;
        }
        uint32_t run = carry + warp_base;
        for (int64_t ch = w0; ch < w1; ++ch) {
            const uint16_t tp = __ldcg(a.slot_tp + ch * 32 + lane);
            const unsigned b = __ballot_sync(kFull, tp != 0);
            if (tp != 0) {                                 // the dense event stream, in slot order
                const uint32_t at = run + __popc(b & ((1u << lane) - 1u));
                a.ev_img[at] = __ldcg(a.slot_img + ch * 32 + lane);
                a.ev_mask[at] = tp;
            }
            if (lane == 0) {
                __stcg(a.evbase + ch, run);
                int64_t lo = 0, hi = S;                    // is this chunk the first of a segment?
                while (lo < hi) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (a.seg_chunk0[mid] < ch) lo = mid + 1; else hi = mid;
                }
                if (lo < S && a.seg_chunk0[lo] == ch) a.seg_ev0[lo] = run;
            }
            run += __popc(b);
        }
    }
    for (int64_t v = gtid; v < a.n; v += gsize) {
        const uint32_t u = a.ord_bat[v];
        a.bq[v] = make_uint2(__ldcg(a.q_of_det + u) | (u >= a.d.Dw ? 0x80000000u : 0u),
                             ((a.img_all[u] & 31u) << 27) | __ldcg(a.ownpos + u));
    }
    grid_sync(a.bar, epoch);

    // ---- D: per image the classes in which it has a detection from either detector, ordered by how deep the AP
    //      sweep of the class will go for this image as a target: the reverse sweep runs from the class's lowest
    //      confidence up to the image's most confident own detection, so its length is proportional to the number
    //      of events (true positives of the dataset) behind that detection.  The AP kernel gives neighbouring entries
    //      to the lanes of one warp, so lanes that work side by side stop at about the same time.
    for (int64_t im = gwarp; im < a.M; im += nwarps) {
        const uint16_t *wcs = a.own_w_cs + im * (a.C + 1), *scs = a.own_s_cs + im * (a.C + 1);
        const uint32_t *wq = a.own_w_q + a.w_off[im], *sq = a.own_s_q + a.s_off[im];
        uint16_t *tmp = a.act_tmp + im * a.C, *out = a.act_cls + im * a.C;
        uint32_t *key = a.act_key + im * a.C;
        uint32_t n = 0;
        for (int64_t ci0 = 0; ci0 < a.C; ci0 += 32) {
            const int64_t ci = ci0 + lane;
            int c = 0;
            bool act = false;
            uint32_t depth = 0;
            if (ci < a.C) {
                c = __ldcg(a.cls_order + ci);
                const int w0c = __ldcg(wcs + c), w1c = __ldcg(wcs + c + 1), s0c = __ldcg(scs + c), s1c = __ldcg(scs + c + 1);
                act = w1c > w0c || s1c > s0c;
                if (act) {
                    // own lists are (class, confidence desc): the first entry of the class is its most confident row
                    uint32_t q = 0xffffffffu;
                    if (w1c > w0c) q = min(q, __ldcg(wq + w0c));
                    if (s1c > s0c) q = min(q, __ldcg(sq + s0c));
                    const uint32_t last = a.pad_off[c + 1] / 32u - 1u;           // last chunk of the class (padding: no events)
                    const uint32_t ch = min(q / 32u, last);
                    depth = __ldcg(a.evbase + last) - __ldcg(a.evbase + ch);
                }
            }
            const unsigned b = __ballot_sync(kFull, act);
            if (act) {
                const uint32_t at = n + __popc(b & ((1u << lane) - 1u));
                tmp[at] = (uint16_t)c;
                key[at] = depth;
            }
            n += __popc(b);
        }
        __syncwarp();
        const bool fits = n <= (uint32_t)kRankStage;
        if (fits)
            for (uint32_t k = lane; k < n; k += 32) staged[warp][k] = key[k];
        __syncwarp();
        for (uint32_t e0 = 0; e0 < n; e0 += 32) {           // stable rank by descending depth, by counting
            const uint32_t e = e0 + lane;
            if (e < n) {
                const uint32_t me = fits ? staged[warp][e] : key[e];
                uint32_t rank = 0;
                for (uint32_t o = 0; o < n; ++o) {
                    const uint32_t ko = fits ? staged[warp][o] : key[o];
                    rank += (ko > me || (ko == me && o < e)) ? 1u : 0u;
                }
                out[rank] = tmp[e];
            }
        }
        __syncwarp();
        if (lane == 0) a.nact[im] = n;
    }
}

static int bits_for(int64_t n) {  // bits needed to represent values in [0, n)
    int b = 1;
    while (((int64_t)1 << b) < n) ++b;
    return b;
}

// The index's memory comes from a stream-ordered pool that is PRIVATE to this library (one per device, created on
// first use, never trimmed so that rebuilding an index does not go back to the driver).  The device's default pool
// — which the host application may be using — is never touched.
struct DeviceState {
    cudaMemPool_t pool = nullptr;
    int sort_blocks = 0, bucket_blocks = 0, post_blocks = 0, sms = 0;
};
static int device_state(DeviceState **out) {
    static std::mutex mu;
    static DeviceState states[64];
    int dev = 0;
    ORIE_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) {
        set_error("orie_index_build: device ordinal %d not supported", dev);
        return ORIE_ELIMIT;
    }
    std::lock_guard<std::mutex> lock(mu);
    DeviceState &s = states[dev];
    if (!s.pool) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool;
        ORIE_CUDA(cudaMemPoolCreate(&pool, &props));
        uint64_t keep_all = UINT64_MAX;
        ORIE_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep_all));
        ORIE_CUDA(cudaDeviceGetAttribute(&s.sms, cudaDevAttrMultiProcessorCount, dev));
        ORIE_TRY(sort_max_blocks(&s.sort_blocks));
        ORIE_TRY(bucket_sort_max_blocks(&s.bucket_blocks));
        ORIE_TRY(coop_max_blocks(post_kernel, kPostThreads, 0, &s.post_blocks));
        s.pool = pool;
    }
    *out = &s;
    return ORIE_OK;
}

// Stream-ordered arena: the sizes are collected first, then ONE allocation from the library's pool backs all of them.
struct Arena {
    struct Item {
        void **p;
        size_t bytes;
    };
    std::vector<Item> items;
    template <typename Tp>
    void add(Tp **p, int64_t count) {
        items.push_back({(void **)p, (size_t)round_up(std::max<int64_t>(count, 1) * (int64_t)sizeof(Tp), 256)});
    }
    size_t total() const {
        size_t t = 0;
        for (auto &it : items) t += it.bytes;
        return std::max<size_t>(t, 256);
    }
    void bind(void *base) {
        size_t at = 0;
        for (auto &it : items) {
            *it.p = (char *)base + at;
            at += it.bytes;
        }
        items.clear();
    }
};

// Frees the temporaries in stream order when the build function returns, on success and on error alike.
struct TempGuard {
    cudaStream_t st;
    void *base = nullptr;
    ~TempGuard() {
        if (base) cudaFreeAsync(base, st);
    }
};

// The whole build is enqueued on `st` without a single host synchronisation: allocation sizes, launch grids and table
// strides come from upper bounds the host can compute from (M, C, Dw, Ds, G); what depends on the data stays on the
// device (IndexMeta).
// Memory: caller-provided (orie_index_build_into: nothing is allocated, so the build can be captured in a CUDA graph) or
// taken from the library's pool (orie_index_build).  plan != nullptr: only report the two sizes.
struct BuildMemory {
    void *index_mem = nullptr, *temp_mem = nullptr;
    size_t index_bytes = 0, temp_bytes = 0;
    bool external = false;
};
static int build(orie_index *ix, const int64_t *w_off, const int32_t *w_cls, const double *w_conf, const uint16_t *w_tp,
                 const int64_t *s_off, const int32_t *s_cls, const double *s_conf, const uint16_t *s_tp,
                 const int64_t *l_off, const int32_t *l_cls, const orie_tuning_t &tune, cudaEvent_t tp_ready, cudaStream_t st,
                 const BuildMemory &mem, size_t *plan /* [2] or nullptr */) {
    const int64_t M = ix->M, C = ix->C;
    const int64_t Dw = ix->Dw, Ds = ix->Ds, G = ix->G;
    if (Dw < 0 || Ds < 0 || G < 0 || Dw >= ((int64_t)1 << 27) || Ds >= ((int64_t)1 << 27) || G >= ((int64_t)1 << 31)) {
        set_error("orie_index_build: row counts out of range (weak %lld, strong %lld, labels %lld; limit 2^27-1 detections per detector)",
                  (long long)Dw, (long long)Ds, (long long)G);
        return ORIE_ELIMIT;
    }
    DeviceState *ds = nullptr;
    ORIE_TRY(device_state(&ds));
    ix->sms = ds->sms;
    const int64_t n = Dw + Ds;
    const int64_t Nmax = std::max<int64_t>(std::max(n, G), 1);
    const Dets dets{Dw, Ds, w_cls, s_cls, w_conf, s_conf, w_tp, s_tp};
    int sort_blocks = ds->sort_blocks, bucket_blocks = ds->bucket_blocks, post_blocks = ds->post_blocks;
    if (tune.sort_max_blocks > 0) {
        sort_blocks = std::min(sort_blocks, tune.sort_max_blocks);
        bucket_blocks = std::min(bucket_blocks, tune.sort_max_blocks);
    }
    const bool use_buckets = !tune.sort_lsd && bucket_sort_applicable(n, C);

    // ---- capacities.  Detection stream: class c takes ceil((cnt_c + 1) / 32) <= cnt_c / 32 + 1 chunks; label stream:
    //      ceil(cnt_c / 32).  A class of nch chunks has ceil(nch / seg_chunks) <= nch / seg_chunks + 1 segments.
    const int64_t chunks_cap = Dw / 32 + C, chunksL_cap = G / 32 + C;
    // Segment length: about two segments per average class — a class is cut into equal-length segments no longer
    // than this (layout_kernel).  Measured on COCO- and VOC-shaped data (profiles/): shorter segments fill the walk's
    // last wave of blocks better, but the AP sweep and the per-(batch, segment) query tables of the index pay per
    // segment, and the sum is flat between two and three per class.  Enough (batch, segment) warp items for >= 4
    // waves of 148 SMs x 8 warps, and no segment so long that a single warp becomes the tail of the walk.  An event
    // record keeps its rank inside the segment in 16 bits: at most 2047 chunks = 65504 slots per segment.
    const int64_t seg_target = std::max<int64_t>(2 * C, ceil_div(4 * 148 * 8, ix->nbatch));
    const int seg_chunks = tune.seg_chunks > 0 ? std::min(tune.seg_chunks, 2047)
                                               : (int)std::min<int64_t>(512, std::max<int64_t>(16, ceil_div(chunks_cap, seg_target)));
    ix->seg_chunks = seg_chunks;
    ix->nchunks_cap = chunks_cap;
    ix->P_cap = chunks_cap * kChunk;
    ix->S_cap = chunks_cap / seg_chunks + C + 1;
    ix->PL_cap = chunksL_cap * kChunk;
    ix->SL_cap = chunksL_cap / seg_chunks + C + 1;
    ix->Ev_cap = std::min<int64_t>(Dw, G * ix->T);     // a label is claimed by at most one detection per threshold

    // ---- the index and the temporaries, one allocation each
    Arena A;
    A.add(&ix->meta, 1);
    A.add(&ix->w_off, M + 1);
    A.add(&ix->s_off, M + 1);
    A.add(&ix->own_w_q, Dw);
    A.add(&ix->own_w_m, Dw);
    A.add(&ix->own_s_q, Ds);
    A.add(&ix->own_s_m, Ds);
    A.add(&ix->own_w_cs, M * (C + 1));
    A.add(&ix->own_s_cs, M * (C + 1));
    A.add(&ix->act_cls, M * C);
    A.add(&ix->nact, M);
    A.add(&ix->bq, n);
    A.add(&ix->gtcnt, M * C);
    A.add(&ix->slot_img, ix->P_cap);
    A.add(&ix->slot_tp, ix->P_cap);
    const bool packed = M <= 65535 && !tune.walk_unpacked;     // image and mask of a slot fit one 32-bit word
    if (packed) A.add(&ix->slot_pk, ix->P_cap);
    A.add(&ix->ev_img, ix->Ev_cap);
    A.add(&ix->ev_mask, ix->Ev_cap);
    A.add(&ix->seg_chunk0, ix->S_cap);
    A.add(&ix->seg_nch, ix->S_cap);
    A.add(&ix->seg_order, ix->S_cap);
    A.add(&ix->seg_ev0, ix->S_cap + 1);
    A.add(&ix->cls_seg0, C + 1);
    A.add(&ix->cls_order, C);
    A.add(&ix->bqoff, ix->nbatch * (ix->S_cap + 1));
    A.add(&ix->lab_slot_img, ix->PL_cap);
    A.add(&ix->lseg_chunk0, ix->SL_cap);
    A.add(&ix->lseg_nch, ix->SL_cap);
    A.add(&ix->lcls_seg0, C + 1);
    const size_t index_bytes = A.total();
    Arena Tm;

    uint64_t *keys, *keys_tmp;
    uint32_t *vtmp, *img_all, *img_l, *order, *ord_bat, *pos_of_det, *lcursor, *hist, *wpre, *q_of_det, *ownpos;
    uint32_t *d_cls_off, *d_pad_off, *d_lcls_off, *d_lpad_off;
    uint16_t *own_w_c, *own_s_c, *act_tmp;
    uint32_t *evbase, *act_key;
    unsigned long long *key_and_or;
    char *scratch;
    TempGuard temp{st};
    Tm.add(&keys, n);
    Tm.add(&keys_tmp, n);
    Tm.add(&vtmp, Nmax);
    Tm.add(&img_all, n);
    Tm.add(&img_l, G);
    Tm.add(&order, n);
    Tm.add(&ord_bat, n);
    Tm.add(&pos_of_det, n);
    Tm.add(&lcursor, C);
    Tm.add(&hist, 3 * C);
    Tm.add(&wpre, n);
    Tm.add(&q_of_det, n);
    Tm.add(&ownpos, n);
    Tm.add(&own_w_c, Dw);
    Tm.add(&own_s_c, Ds);
    Tm.add(&act_tmp, M * C);
    Tm.add(&act_key, M * C);
    Tm.add(&evbase, ix->nchunks_cap);
    Tm.add(&d_cls_off, C + 1);
    Tm.add(&d_pad_off, C + 1);
    Tm.add(&d_lcls_off, C + 1);
    Tm.add(&d_lpad_off, C + 1);
    Tm.add(&key_and_or, 2);
    Tm.add(&scratch, (int64_t)std::max(std::max(sort_scratch_bytes(sort_blocks), (size_t)post_blocks * 4 + 256),
                                      use_buckets ? bucket_sort_scratch_bytes(n, C, bucket_blocks) : (size_t)0));
    const size_t temp_bytes = Tm.total();
    if (plan) {
        plan[0] = index_bytes; plan[1] = temp_bytes;
        return ORIE_OK;
    }
    if (mem.external) {
        if (mem.index_bytes < index_bytes || mem.temp_bytes < temp_bytes || !mem.index_mem || !mem.temp_mem ||
            ((uintptr_t)mem.index_mem & 255) || ((uintptr_t)mem.temp_mem & 255)) {
            set_error("orie_index_build_into: needs %zu + %zu bytes (index, temporaries), 256-byte aligned; got %zu + %zu",
                      index_bytes, temp_bytes, mem.index_bytes, mem.temp_bytes);
            return ORIE_EWORKSPACE;
        }
        A.bind(mem.index_mem);
        Tm.bind(mem.temp_mem);
    } else {
        void *base = nullptr;
        ORIE_CUDA(cudaMallocFromPoolAsync(&base, index_bytes, ds->pool, st));
        ix->allocs[ix->n_allocs++] = base;
        A.bind(base);
        ORIE_CUDA(cudaMallocFromPoolAsync(&temp.base, temp_bytes, ds->pool, st));
        Tm.bind(temp.base);
    }
    ix->device_bytes += (int64_t)index_bytes;

    ORIE_CUDA(cudaMemcpyAsync(ix->w_off, w_off, (size_t)(M + 1) * 8, cudaMemcpyDeviceToDevice, st));
    ORIE_CUDA(cudaMemcpyAsync(ix->s_off, s_off, (size_t)(M + 1) * 8, cudaMemcpyDeviceToDevice, st));
    ORIE_CUDA(cudaMemsetAsync(ix->meta, 0, sizeof(IndexMeta), st));
    ORIE_CUDA(cudaMemsetAsync(hist, 0, (size_t)(3 * C) * 4, st));
    ORIE_CUDA(cudaMemsetAsync(ix->gtcnt, 0, (size_t)(M * C) * 4, st));
    ORIE_CUDA(cudaMemsetAsync(lcursor, 0, (size_t)C * 4, st));
    ORIE_CUDA(cudaMemsetAsync(key_and_or, 0xff, 8, st));
    ORIE_CUDA(cudaMemsetAsync(key_and_or + 1, 0, 8, st));

    const int cbits = bits_for(C), bbits = bits_for(ix->nbatch);

    // ---- prep: images of rows, confidence keys, class histograms, ground-truth counts, validation
    {
        PrepArgs pa{M, C, Dw, Ds, G, w_off, s_off, l_off, w_cls, s_cls, l_cls, w_conf, s_conf,
                    keys, img_all, img_l, hist, ix->gtcnt, key_and_or, &ix->meta->status, 3 * C <= kPrepMaxSmemBins ? (int)(3 * C) : 0};
        const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(3 * M, kPrepThreads / 32), 148 * 8);
        prep_kernel<<<grid, kPrepThreads, (size_t)pa.smem_bins * 4, st>>>(pa);
        ORIE_LAUNCH_CHECK();
    }
    // ---- layout: class counts -> padded layouts and segment tables, on the device
    {
        LayoutArgs la{C, ix->S_cap, ix->SL_cap, seg_chunks, hist, d_cls_off, d_pad_off, d_lcls_off, d_lpad_off,
                      ix->cls_seg0, ix->lcls_seg0, ix->seg_chunk0, ix->seg_nch, ix->lseg_chunk0, ix->lseg_nch, ix->seg_order, ix->meta};
        layout_kernel<<<1, kLayoutThreads, 0, st>>>(la);
        ORIE_LAUNCH_CHECK();
    }

    // ---- ONE sort of all detections: confidence desc (64-bit key), then class (stable).  Weak rows precede
    //      strong rows in the id space, so on exact confidence ties weak sorts first, then image, then row.
    //      The epilogue ranks every position among the weak detections (wpre).
    if (n) {
        SortJob j;
        j.n = n;
        j.keys0 = keys; j.keys1 = keys_tmp;
        j.vals_a = order; j.vals_b = vtmp;
        j.cls_lo = w_cls; j.cls_hi = s_cls; j.split = (uint32_t)Dw;
        j.rank_out = wpre; j.rank_split = (uint32_t)Dw;
        ORIE_TRY(sort_add_passes(&j, kDigitKey, 0, 64));
        ORIE_TRY(sort_add_passes(&j, kDigitClass, 0, cbits));
        if (use_buckets) {
            // one binning pass + per-bucket bitonic sorts (sort.cu); the radix passes above are its fallback
            BucketSortJob bj;
            bj.n = n;
            bj.key_and_or = (const uint64_t *)key_and_or;
            bj.lsd = j;
            ORIE_TRY(bucket_sort_run(bj, C, bucket_blocks, scratch, st));
        } else {
            ORIE_TRY(sort_run(j, sort_blocks, scratch, st));
        }
        // the same order regrouped by 32-image batch: the per-batch query lists of the walk (weak and strong rows
        // stay interleaved, ascending by slot)
        SortJob r;
        r.n = n;
        r.vals_in = order; r.vals_a = ord_bat; r.vals_b = vtmp;
        r.img = img_all;
        ORIE_TRY(sort_add_passes(&r, kDigitBatch, 0, bbits));
        ORIE_TRY(sort_run(r, sort_blocks, scratch, st));
    }

    // ---- everything else in one cooperative launch (post_kernel)
    if (tp_ready) ORIE_CUDA(cudaStreamWaitEvent(st, tp_ready, 0));     // first reader of the true-positive masks
    {
        PostArgs pa;
        memset(&pa, 0, sizeof(pa));
        pa.d = dets;
        pa.n = n; pa.M = M; pa.C = C; pa.G = G; pa.nbatch = ix->nbatch; pa.S_cap = ix->S_cap;
        pa.order = order; pa.wpre = wpre; pa.img_all = img_all; pa.ord_bat = ord_bat; pa.img_l = img_l; pa.hist = hist;
        pa.l_cls = l_cls; pa.seg_chunk0 = ix->seg_chunk0;
        pa.cls_off = d_cls_off; pa.pad_off = d_pad_off; pa.lcls_off = d_lcls_off; pa.lpad_off = d_lpad_off;
        pa.w_off = ix->w_off; pa.s_off = ix->s_off;
        pa.lcursor = lcursor;
        pa.slot_img = ix->slot_img; pa.lab_slot_img = ix->lab_slot_img; pa.slot_tp = ix->slot_tp;
        pa.slot_pk = ix->slot_pk; pa.ev_img = ix->ev_img; pa.ev_mask = ix->ev_mask;
        pa.q_of_det = q_of_det; pa.pos_of_det = pos_of_det; pa.ownpos = ownpos;
        pa.own_w_q = ix->own_w_q; pa.own_s_q = ix->own_s_q; pa.own_w_m = ix->own_w_m; pa.own_s_m = ix->own_s_m;
        pa.own_w_c = own_w_c; pa.own_s_c = own_s_c; pa.own_w_cs = ix->own_w_cs; pa.own_s_cs = ix->own_s_cs;
        pa.cls_order = ix->cls_order; pa.act_cls = ix->act_cls; pa.nact = ix->nact;
        pa.evbase = evbase; pa.act_key = act_key; pa.act_tmp = act_tmp;
        pa.bq = ix->bq; pa.bqoff = ix->bqoff; pa.seg_ev0 = ix->seg_ev0;
        pa.meta = ix->meta;
        pa.bar = (unsigned *)scratch;
        pa.table = (uint32_t *)(scratch + 256);
        // three CTAs per SM: more only make the grid barriers dearer (measured, profiles/)
        post_blocks = std::min(post_blocks, 3 * ds->sms);
        if (tune.post_blocks > 0) post_blocks = std::max(1, std::min(post_blocks, tune.post_blocks));
        const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(post_blocks, ceil_div(std::max(n, ix->P_cap), kPostThreads)));
        ORIE_CUDA(cudaMemsetAsync(pa.bar, 0, 4, st));
        void *args[] = {&pa};
        ORIE_CUDA(cudaLaunchCooperativeKernel((const void *)post_kernel, dim3(blocks), dim3(kPostThreads), args, 0, st));
        ORIE_LAUNCH_CHECK();
    }
    return ORIE_OK;
}

int resolve(const orie_index *cix) {
    orie_index *ix = const_cast<orie_index *>(cix);
    if (ix->resolved) {
        if (ix->resolved_rc != ORIE_OK) set_error("orie_index: the build of this index failed (code %d)", ix->resolved_rc);
        return ix->resolved_rc;
    }
    IndexMeta m;
    ORIE_CUDA(cudaStreamSynchronize(ix->stream));
    ORIE_CUDA(cudaMemcpy(&m, ix->meta, sizeof(m), cudaMemcpyDeviceToHost));
    ix->P = m.P; ix->nchunks = m.nchunks; ix->S = m.S; ix->Ev = m.Ev;
    ix->PL = m.PL; ix->nchunksL = m.nchunksL; ix->SL = m.SL;
    int rc = ORIE_OK;
    if (m.status & kStatusCounts) {
        set_error("orie_index_build: row counts (%lld, %lld, %lld) do not match the offset arrays",
                  (long long)ix->Dw, (long long)ix->Ds, (long long)ix->G);
        rc = ORIE_EDATA;
    } else if (m.status & kStatusClass) {
        set_error("orie_index_build: class id outside [0, %lld)", (long long)ix->C);
        rc = ORIE_EDATA;
    } else if (m.status & kStatusRows) {
        set_error("orie_index_build: an image has more than 65535 rows in one file");
        rc = ORIE_ELIMIT;
    }
    ix->resolved = true;
    ix->resolved_rc = rc;
    return rc;
}

}  // namespace orie

using namespace orie;

static int index_create(int64_t M, int64_t C, int T, int64_t num_weak, int64_t num_strong, int64_t num_labels,
                        const int64_t *w_off, const int32_t *w_cls, const double *w_conf, const uint16_t *w_tp,
                        const int64_t *s_off, const int32_t *s_cls, const double *s_conf, const uint16_t *s_tp,
                        const int64_t *l_off, const int32_t *l_cls, const orie_tuning_t *tuning, orie_event_t tp_ready,
                        orie_stream_t stream, const BuildMemory &mem, size_t *plan, orie_index_t **out) {
    if (!out && !plan) {
        set_error("orie_index_build: out is NULL");
        return ORIE_EINVAL;
    }
    if (out) *out = nullptr;
    if (M < 1 || C < 1 || (!plan && (!w_off || !s_off || !l_off))) {
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
    orie_tuning_t tune;
    memset(&tune, 0, sizeof(tune));
    if (tuning) tune = *tuning;
    orie_index *ix = new orie_index();
    ix->M = M; ix->C = C; ix->T = T;
    ix->Dw = num_weak; ix->Ds = num_strong; ix->G = num_labels;
    ix->stream = stream;
    ix->nbatch = ceil_div(M, 32);
    ix->ens_words = ceil_div(M + 1, 32);
    ix->cls_per_warp = 32 / T;
    ix->class_groups = ceil_div(C, ix->cls_per_warp);
    ix->walk_gmem = tune.walk_gmem;
    ix->walk_single = tune.walk_single;
    ix->ap_mode = tune.ap_mode;
    ix->walk_waves = tune.walk_waves;
    const int rc = build(ix, w_off, w_cls, w_conf, w_tp, s_off, s_cls, s_conf, s_tp, l_off, l_cls, tune, tp_ready, stream, mem, plan);
    if (rc != ORIE_OK || plan) {
        orie_index_destroy(ix);
        return rc;
    }
    *out = ix;
    return ORIE_OK;
}

extern "C" int orie_index_build(int64_t M, int64_t C, int T, int64_t num_weak, int64_t num_strong, int64_t num_labels,
                                const int64_t *w_off, const int32_t *w_cls, const double *w_conf, const uint16_t *w_tp,
                                const int64_t *s_off, const int32_t *s_cls, const double *s_conf, const uint16_t *s_tp,
                                const int64_t *l_off, const int32_t *l_cls, const orie_tuning_t *tuning, orie_event_t tp_ready,
                                orie_stream_t stream, orie_index_t **out) {
    return index_create(M, C, T, num_weak, num_strong, num_labels, w_off, w_cls, w_conf, w_tp, s_off, s_cls, s_conf, s_tp,
                        l_off, l_cls, tuning, tp_ready, stream, BuildMemory(), nullptr, out);
}

extern "C" int orie_index_sizes(int64_t M, int64_t C, int T, int64_t num_weak, int64_t num_strong, int64_t num_labels,
                                const orie_tuning_t *tuning, size_t *index_bytes, size_t *temp_bytes) {
    if (!index_bytes || !temp_bytes) {
        set_error("orie_index_sizes: null output");
        return ORIE_EINVAL;
    }
    size_t plan[2] = {0, 0};
    ORIE_TRY(index_create(M, C, T, num_weak, num_strong, num_labels, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                          nullptr, nullptr, nullptr, nullptr, tuning, nullptr, nullptr, BuildMemory(), plan, nullptr));
    *index_bytes = plan[0];
    *temp_bytes = plan[1];
    return ORIE_OK;
}

extern "C" int orie_index_build_into(int64_t M, int64_t C, int T, int64_t num_weak, int64_t num_strong, int64_t num_labels,
                                     const int64_t *w_off, const int32_t *w_cls, const double *w_conf, const uint16_t *w_tp,
                                     const int64_t *s_off, const int32_t *s_cls, const double *s_conf, const uint16_t *s_tp,
                                     const int64_t *l_off, const int32_t *l_cls, const orie_tuning_t *tuning,
                                     void *index_mem, size_t index_bytes, void *temp_mem, size_t temp_bytes,
                                     orie_event_t tp_ready, orie_stream_t stream, orie_index_t **out) {
    BuildMemory mem;
    mem.index_mem = index_mem; mem.index_bytes = index_bytes; mem.temp_mem = temp_mem; mem.temp_bytes = temp_bytes;
    mem.external = true;
    return index_create(M, C, T, num_weak, num_strong, num_labels, w_off, w_cls, w_conf, w_tp, s_off, s_cls, s_conf, s_tp,
                        l_off, l_cls, tuning, tp_ready, stream, mem, nullptr, out);
}

extern "C" void orie_index_destroy(orie_index_t *ix) {
    if (!ix) return;
    for (int i = 0; i < ix->n_allocs; ++i) cudaFreeAsync(ix->allocs[i], ix->stream);
    if (ix->ev_fork) cudaEventDestroy(ix->ev_fork);
    if (ix->ev_join) cudaEventDestroy(ix->ev_join);
    delete ix;
}

extern "C" int orie_index_set_aux_stream(orie_index_t *ix, orie_stream_t aux) {
    if (!ix) {
        set_error("orie_index_set_aux_stream: index is NULL");
        return ORIE_EINVAL;
    }
    if (aux && !ix->ev_fork) {
        ORIE_CUDA(cudaEventCreateWithFlags(&ix->ev_fork, cudaEventDisableTiming));
        ORIE_CUDA(cudaEventCreateWithFlags(&ix->ev_join, cudaEventDisableTiming));
    }
    ix->aux = aux;
    return ORIE_OK;
}

extern "C" int orie_index_status(const orie_index_t *ix) {
    if (!ix) {
        set_error("orie_index_status: index is NULL");
        return ORIE_EINVAL;
    }
    ORIE_TRY(resolve(ix));
    // the reward pass may have flagged a workspace that was too small for the event lists since the build
    IndexMeta m;
    ORIE_CUDA(cudaStreamSynchronize(ix->stream));
    ORIE_CUDA(cudaMemcpy(&m, ix->meta, sizeof(m), cudaMemcpyDeviceToHost));
    if (m.status & kStatusWorkspace) {
        set_error("orie_reward: the workspace of an earlier call was too small for the event lists (%u events per target); "
                  "size it with orie_reward_workspace_bytes or orie_reward_workspace_bound", m.Ev);
        return ORIE_EWORKSPACE;
    }
    return ORIE_OK;
}

extern "C" int orie_index_info(const orie_index_t *ix, orie_index_info_t *info) {
    if (!ix || !info) {
        set_error("orie_index_info: null argument");
        return ORIE_EINVAL;
    }
    ORIE_TRY(resolve(ix));
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
