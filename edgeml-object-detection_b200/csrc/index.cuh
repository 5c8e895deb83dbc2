// Device-resident dataset index (see include/orie_b200.h: orie_index_t).
#pragma once
#include "common.cuh"

namespace orie {

// Facts about the index that only the DEVICE knows until somebody asks (orie_index_info and friends synchronise
// once and cache them on the host): the padded sizes follow from the class histogram, the number of events from
// the true-positive masks.  Kernels of the reward pass read them from here, so the host never has to wait for the
// build before it enqueues the reward pass.
struct IndexMeta {
    uint32_t P, nchunks, S;        // detection stream: slots, chunks, segments
    uint32_t PL, nchunksL, SL;     // label stream
    uint32_t Ev;                   // events (slots holding a true positive)
    uint32_t status;               // sticky error bits, kStatus*
};
enum : uint32_t {
    kStatusRows = 1,        // an image has more than 65535 rows in one file
    kStatusClass = 2,       // class id outside [0, C)
    kStatusCounts = 4,      // off[M] does not match the row count given by the caller
    kStatusWorkspace = 8,   // a reward call's workspace is too small for the event lists (set by the reward pass)
};

}  // namespace orie

struct orie_index {
    int64_t M = 0, C = 0;
    int T = 0;
    int64_t Dw = 0, Ds = 0, G = 0;
    int seg_chunks = 0;
    int64_t nbatch = 0;          // ceil(M / 32): targets are processed 32 at a time (one per lane)
    int64_t ens_words = 0;       // ceil((M + 1) / 32): bit M is the never-a-member sentinel of padding slots

    // ---- capacities known on the host when the build is enqueued (upper bounds of the device-side sizes);
    //      they size the allocations, the launch grids and the strides of the per-segment tables
    int64_t P_cap = 0, nchunks_cap = 0, S_cap = 0, PL_cap = 0, SL_cap = 0, Ev_cap = 0;
    // ---- exact sizes, valid once `resolved` (orie::resolve synchronises with the build and reads IndexMeta)
    bool resolved = false;
    int resolved_rc = 0;
    int64_t P = 0, nchunks = 0, S = 0, Ev = 0, PL = 0, nchunksL = 0, SL = 0;
    orie::IndexMeta *meta = nullptr;   // device

    // ---- detection stream: weak detections in (class asc, conf desc) order, each class padded
    //      to a whole number of 32-slot chunks with at least one padding slot
    uint32_t *slot_img = nullptr;    // [P_cap] image of the detection in the slot (M for padding)
    uint16_t *slot_tp = nullptr;     // [P_cap] its true-positive mask (0 for padding); non-zero = "event"
    uint32_t *slot_pk = nullptr;     // [P_cap] image | mask << 16: one load per slot for the walk; only with M <= 65535
    uint32_t *ev_img = nullptr;      // [Ev_cap] dense event stream: image ...
    uint16_t *ev_mask = nullptr;     // [Ev_cap] ... and true-positive mask of the slots that hold an event, in slot order
    int32_t *seg_chunk0 = nullptr;   // [S_cap]
    int32_t *seg_nch = nullptr;      // [S_cap]
    int32_t *seg_order = nullptr;    // [S_cap] segments by descending length
    uint32_t *seg_ev0 = nullptr;     // [S_cap + 1] events in front of the segment; [S] = all events
    int32_t *cls_seg0 = nullptr;     // [C+1]
    int32_t *cls_order = nullptr;    // [C] classes by descending weak-detection count

    // ---- per-batch query list: the own detections (both detectors) of a batch's 32 images, ascending by query slot
    uint2 *bq = nullptr;             // [Dw + Ds] {q | is_strong << 31, lane << 27 | own position}
    uint32_t *bqoff = nullptr;       // [nbatch][S_cap+1] first entry of (batch, segment)

    // ---- own lists: image-major (aligned with w_off / s_off), (class asc, conf desc)
    int64_t *w_off = nullptr, *s_off = nullptr;   // [M+1] device copies
    uint32_t *own_w_q = nullptr, *own_s_q = nullptr;   // query slot: own weak slot / strong insertion slot
    uint16_t *own_w_m = nullptr, *own_s_m = nullptr;   // TP masks
    uint16_t *own_w_cs = nullptr, *own_s_cs = nullptr; // [M][C+1] start of each class inside the image's list
    // classes in which the image has a detection from either detector (the only ones whose AP differs between the
    // two variants of a target), in cls_order order
    uint16_t *act_cls = nullptr;     // [M][C]
    uint32_t *nact = nullptr;        // [M]

    // ---- label stream: ground-truth objects sorted by class, padded to chunks
    uint32_t *lab_slot_img = nullptr;  // [PL_cap]
    int32_t *lseg_chunk0 = nullptr, *lseg_nch = nullptr;  // [SL_cap]
    int32_t *lcls_seg0 = nullptr;      // [C+1]
    uint32_t *gtcnt = nullptr;         // [M][C] ground-truth objects per image and class

    // ---- AP work decomposition
    int cls_per_warp = 0;              // 32 / T classes are integrated side by side by one warp
    int64_t class_groups = 0;          // ceil(C / cls_per_warp)

    // ---- per-index tuning (orie_tuning_t; zeros = defaults)
    int walk_gmem = 0, ap_mode = 0, walk_single = 0;
    int sms = 0;                       // multiprocessors of the device (grid of the persistent AP kernel)
    double walk_waves = 0.0;

    int64_t device_bytes = 0;
    cudaStream_t stream = nullptr;   // stream the index was built on; its memory is freed on it
    // optional auxiliary stream (orie_index_set_aux_stream): the reward pass runs its label walk there, next to the
    // detection walk
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    void *allocs[8] = {};
    int n_allocs = 0;
};

namespace orie {
// Wait for the index's stream (once), read IndexMeta, cache the exact sizes; returns the build's error code.
int resolve(const orie_index *ix);
}  // namespace orie
