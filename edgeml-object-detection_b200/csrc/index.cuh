// Device-resident dataset index (see include/orie_b200.h: orie_index_t).
#pragma once
#include "common.cuh"

struct orie_index {
    int64_t M = 0, C = 0;
    int T = 0;
    int64_t Dw = 0, Ds = 0, G = 0;
    int seg_chunks = 0;
    int64_t nbatch = 0;          // ceil(M / 32): targets are processed 32 at a time (one per lane)
    int64_t ens_words = 0;       // ceil((M + 1) / 32): bit M is the never-a-member sentinel of padding slots

    // ---- detection stream: weak detections in (class asc, conf desc) order, each class padded
    //      to a whole number of 32-slot chunks with at least one padding slot
    int64_t P = 0, nchunks = 0, S = 0, Ev = 0;
    uint32_t *slot_img = nullptr;    // [P] image of the detection in the slot (M for padding)
    uint16_t *slot_tp = nullptr;     // [P] its true-positive mask (0 for padding); non-zero = "event"
    uint32_t *evbase = nullptr;      // [nchunks] number of events in front of the chunk
    int32_t *seg_chunk0 = nullptr;   // [S]
    int32_t *seg_nch = nullptr;      // [S]
    uint32_t *seg_ev0 = nullptr;     // [S] evbase[seg_chunk0[s]]
    int32_t *cls_seg0 = nullptr;     // [C+1]
    int32_t *cls_order = nullptr;    // [C] classes by descending weak-detection count (AP warps take neighbours)

    // ---- per-batch query list: the own detections (both detectors) of a batch's 32 images, ascending by query slot
    uint2 *bq = nullptr;             // [Dw + Ds] {q | is_strong << 31, lane << 27 | own position}
    uint32_t *bqoff = nullptr;       // [nbatch][S+1] first entry of (batch, segment)

    // ---- own lists: image-major (aligned with w_off / s_off), (class asc, conf desc)
    int64_t *w_off = nullptr, *s_off = nullptr;   // [M+1] device copies
    uint32_t *own_w_q = nullptr, *own_s_q = nullptr;   // query slot: own weak slot / strong insertion slot
    uint16_t *own_w_m = nullptr, *own_s_m = nullptr;   // TP masks
    uint16_t *own_w_cs = nullptr, *own_s_cs = nullptr; // [M][C+1] start of each class inside the image's list

    // ---- label stream: ground-truth objects sorted by class, padded to chunks
    int64_t PL = 0, nchunksL = 0, SL = 0;
    uint32_t *lab_slot_img = nullptr;  // [PL]
    int32_t *lseg_chunk0 = nullptr, *lseg_nch = nullptr;  // [SL]
    int32_t *lcls_seg0 = nullptr;      // [C+1]
    uint32_t *gtcnt = nullptr;         // [M][C] ground-truth objects per image and class

    // ---- AP work decomposition
    int cls_per_warp = 0;              // 32 / T classes are integrated side by side by one warp
    int64_t class_groups = 0;          // ceil(C / cls_per_warp)

    int64_t device_bytes = 0;
    cudaStream_t stream = nullptr;   // stream the index was built on; its memory is freed on it
    void *allocs[48] = {};
    int n_allocs = 0;
};
