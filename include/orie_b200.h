/*
 * orie_b200.h — C ABI of the B200-native ORIE / ORI / DCSB reward engine.
 *
 * The reference (qiujiaming315/edgeml-object-detection) is pure Python and has
 * no FFI of its own; the boundary it offers is the reward.py CLI, the on-disk
 * formats and three Python call signatures.  Every entry point
 * below names the reference code it replaces (paths relative to the upstream
 * repository root).  INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add to bind them.
 *
 * Conventions
 *   - plain pointers and sizes only; every array pointer is a DEVICE pointer
 *     unless its name ends in _host;
 *   - the caller owns all buffers; the only allocations made inside are the
 *     ones owned by an orie_index_t (created by orie_index_build, released by
 *     orie_index_destroy).  Scratch for the reward pass is a caller-provided
 *     workspace (orie_reward_workspace_bytes);
 *   - work is enqueued on the given CUDA stream and is asynchronous —
 *     orie_index_build included: nothing on the path from the packed dataset
 *     to the reward vector waits for the device, so the whole job can be
 *     enqueued (or captured in a CUDA graph) in one go.  Only the functions
 *     that REPORT device-side facts synchronise: orie_index_info,
 *     orie_index_status, orie_reward_workspace_bytes, orie_reward_profile;
 *   - return value 0 = ORIE_OK, otherwise an ORIE_E* code; orie_last_error()
 *     returns a thread-local message for the last failure on this thread.
 *     Errors that only the device can detect (class id out of range, offsets
 *     that do not match the row counts, a workspace too small for the event
 *     lists) are sticky in the index: the reward pass then computes nothing
 *     and stores NaN, and orie_index_status / orie_index_info return the code;
 *   - no global mutable state visible to the caller: one host thread per GPU
 *     may drive its own index/workspace concurrently with others.  The index
 *     allocates from a stream-ordered memory pool PRIVATE to this library (one
 *     per device, created on first use); the device's default pool is never
 *     touched.
 *
 * Data layout ("packed dataset")
 *   M images, C dense class ids [0,C), T IoU thresholds (1..16).
 *   Detections of one detector: CSR by image — off int64[M+1], box f64[D,4]
 *   (x1,y1,x2,y2 = lib/metrics.py:6-18 applied on the host), conf f64[D],
 *   cls int32[D]; row order inside an image = file order.  Labels likewise
 *   without conf.
 */
#ifndef ORIE_B200_H
#define ORIE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *orie_stream_t; /* == cudaStream_t */
typedef struct CUevent_st *orie_event_t;   /* == cudaEvent_t */

enum {
    ORIE_OK = 0,
    ORIE_EINVAL = 1,     /* bad argument */
    ORIE_ECUDA = 2,      /* CUDA runtime error (message has the detail) */
    ORIE_ELIMIT = 3,     /* input exceeds a documented engine limit */
    ORIE_EWORKSPACE = 4, /* workspace too small */
    ORIE_EDATA = 5       /* inconsistent data (e.g. ensemble index out of range) */
};

#define ORIE_MAX_THRESHOLDS 16
#define ORIE_MAX_LABELS_PER_IMAGE 4096

const char *orie_last_error(void);
int orie_version(void);

/*
 * TP matching of every detection of every image at T IoU thresholds.
 * Replaces lib/metrics.py:38-64 (box_correct) + :67-86 (box_iou) as driven per
 * image by lib/data.py:63-83 (set_data).  FP64, upstream's operation order, no
 * FMA contraction.  iouv_host: host array of T thresholds (pass numpy's
 * np.linspace(0.5,0.95,10) or [0.5] verbatim).
 *   tp_mask[d]   bit t set  <=> upstream's correct[d, t]
 *   match_idx[d] label row (within the image) the detection is matched to at
 *                every threshold where it is a TP; -1 if it is a TP nowhere
 *   best_iou[d]  IoU with its best same-class label (0 if none)
 * All three outputs are required (best_iou doubles as the kernel's scratch).
 */
int orie_match(const double *det_box, const int32_t *det_cls, const int64_t *det_off,
               const double *lab_box, const int32_t *lab_cls, const int64_t *lab_off,
               const double *iouv_host, int T, int64_t M,
               uint16_t *tp_mask, int32_t *match_idx, double *best_iou, orie_stream_t stream);

/*
 * DCSB reward of every image: #(strong conf > 0.5) - #(weak conf > 0.5).
 * Replaces reward.py:55-69 (compute_dcsb).  out: int64[M].
 */
int orie_dcsb(const double *w_conf, const int64_t *w_off, const double *s_conf, const int64_t *s_off,
              int64_t M, int64_t *out, orie_stream_t stream);

/*
 * Dataset index: the weak detections sorted once by (class, confidence desc)
 * (device radix sort), laid out in padded 32-slot chunks / segments, the
 * true-positive event table, the strong detections' insertion ranks, the
 * per-image own-detection lists and the class-sorted label stream.  It is the
 * target-independent part of what reward.py:40-49 + lib/metrics.py:100-104
 * recompute for every target (gather + argsort + unique).
 * tp_ready (nullable): an event recorded after w_tp / s_tp were produced (e.g. by orie_match on another
 * stream).  The build sorts by (class, confidence) first and only waits for the event where it first reads
 * the true-positive masks, so matching — and the host-to-device copy of the boxes it needs — overlaps the sort.
 */
typedef struct orie_index orie_index_t;

typedef struct {
    int64_t num_images, num_classes;
    int32_t num_thresholds, seg_chunks;
    int64_t num_weak, num_strong, num_labels;
    int64_t slots, segments, events;          /* detection stream */
    int64_t label_slots, label_segments;      /* label stream */
    int64_t class_groups;                     /* AP work items per target */
    int64_t ens_words;                        /* uint32 words per ensemble bitmap */
    int64_t device_bytes;                     /* bytes owned by the index */
} orie_index_info_t;

/* Per-index tuning / test knobs (all zero = defaults; nothing is read from the environment). */
typedef struct {
    int32_t seg_chunks;       /* chunks per segment; 0 = auto; at most 2047 (clamped) */
    int32_t sort_max_blocks;  /* cap on the CTAs of the cooperative sort (few CTAs force its multi-tile path) */
    int32_t post_blocks;      /* cap on the CTAs of the post-layout kernel */
    int32_t walk_gmem;        /* != 0: keep the walk's membership tables in global memory even if they fit shared memory */
    int32_t ap_mode;          /* AP kernel variant (measurement knob, results identical up to summation order): 0 = default */
    int32_t sort_lsd;         /* != 0: sort the dataset with the LSD radix passes instead of the bucket sort */
    int32_t walk_unpacked;    /* != 0: the walk reads image and true-positive mask of a slot from two arrays even when
                               * they fit one 32-bit word (<= 65535 images), i.e. the path of larger datasets */
    int32_t walk_single;      /* != 0: the detection walk takes one 32-target batch per warp even where two would fit */
    double walk_waves;        /* resident-block waves the grid of the one-batch walk (labels, fallback) is sized for; 0 = default (2) */
} orie_tuning_t;

/* Asynchronous: returns as soon as the build is enqueued on `stream`; *out is usable by every call below at once
 * (they are ordered behind the build on the same stream; on another stream, order them yourself). */
int orie_index_build(int64_t M, int64_t C, int T,
                     int64_t num_weak, int64_t num_strong, int64_t num_labels, /* == off[M] of each block */
                     const int64_t *w_off, const int32_t *w_cls, const double *w_conf, const uint16_t *w_tp,
                     const int64_t *s_off, const int32_t *s_cls, const double *s_conf, const uint16_t *s_tp,
                     const int64_t *l_off, const int32_t *l_cls,
                     const orie_tuning_t *tuning /* nullable */, orie_event_t tp_ready /* nullable */,
                     orie_stream_t stream, orie_index_t **out);
/* The same build without any allocation inside — every byte comes from the caller, so the call can be recorded in
 * a CUDA graph and replayed (index_mem must stay alive as long as the index, temp_mem until the build has run on the
 * device).  orie_index_sizes reports the two sizes for a dataset shape (upper bounds computed from the row counts). */
int orie_index_sizes(int64_t M, int64_t C, int T, int64_t num_weak, int64_t num_strong, int64_t num_labels,
                     const orie_tuning_t *tuning /* nullable */, size_t *index_bytes, size_t *temp_bytes);
int orie_index_build_into(int64_t M, int64_t C, int T,
                          int64_t num_weak, int64_t num_strong, int64_t num_labels,
                          const int64_t *w_off, const int32_t *w_cls, const double *w_conf, const uint16_t *w_tp,
                          const int64_t *s_off, const int32_t *s_cls, const double *s_conf, const uint16_t *s_tp,
                          const int64_t *l_off, const int32_t *l_cls,
                          const orie_tuning_t *tuning /* nullable */,
                          void *index_mem, size_t index_bytes, void *temp_mem, size_t temp_bytes, /* 256-byte aligned */
                          orie_event_t tp_ready /* nullable */, orie_stream_t stream, orie_index_t **out);
void orie_index_destroy(orie_index_t *idx);
/* Optional: an auxiliary stream the reward pass may use for the kernels that can run side by side (the label walk
 * next to the detection walk); ordering with the main stream is kept with events inside the call.  NULL = none. */
int orie_index_set_aux_stream(orie_index_t *idx, orie_stream_t aux);
/* Waits for the build (first call only), then reports the exact sizes; returns the build's error code if the device
 * rejected the input (ORIE_EDATA: class id outside [0, C) or offsets inconsistent with the row counts; ORIE_ELIMIT:
 * more than 65535 rows in one image file). */
int orie_index_info(const orie_index_t *idx, orie_index_info_t *info);
/* Waits for the index's stream; ORIE_OK, the build's error code, or ORIE_EWORKSPACE if a reward call since the
 * build was given a workspace too small for the event lists (its outputs are NaN). */
int orie_index_status(const orie_index_t *idx);

/*
 * Ensembles.  An ensemble is a bitmap over images: uint32[ens_words] per
 * target, bit e set <=> image e is in the target's ensemble.  Targets are the
 * contiguous image range [t0, t0+nt).
 *
 * orie_ensemble_from_indices: explicit index lists ens_idx int32[nt, N]
 *   (row r = ensemble of target t0+r; the arange/shift/permutation[:N] draw of
 *   reward.py:35-38 regenerated on the host for parity runs).  status (device
 *   int32[1], zero-initialised by the caller) receives a non-zero value if an
 *   index is out of range, equals its target, or repeats.
 * orie_ensemble_sample: device-side draw of N distinct images != target per
 *   target (counter-based generator keyed by (seed, target), so the result does
 *   not depend on how targets are sharded over GPUs).  Replaces reward.py:35-38
 *   for throughput runs.
 */
int orie_ensemble_from_indices(const orie_index_t *idx, int64_t t0, int64_t nt, const int32_t *ens_idx, int64_t N,
                               uint32_t *ens_bits, int32_t *status, orie_stream_t stream);
int orie_ensemble_sample(const orie_index_t *idx, int64_t t0, int64_t nt, int64_t N, uint64_t seed,
                         uint32_t *ens_bits, orie_stream_t stream);
/* The same draw with the seed read from device memory when the kernel runs (seed_dev: device uint64[1]), so that a
 * recorded call can be replayed with another seed. */
int orie_ensemble_sample_dev(const orie_index_t *idx, int64_t t0, int64_t nt, int64_t N, const uint64_t *seed_dev,
                             uint32_t *ens_bits, orie_stream_t stream);

/*
 * ORIE rewards of targets [t0, t0+nt):
 *   reward[r] = (N+1) * ( mAP(E_weak + strong_t) - mAP(E_weak + weak_t) ), NaN -> 0.
 * Replaces reward.py:16-52 (compute_orie) + :86 (NaN -> 0) and, inside it,
 * lib/metrics.py:89-124 (ap_per_class) + :127-148 (compute_ap).  N is the
 * (already clamped) ensemble size used for the (N+1) multiplier; N = 0 is ORI.
 * Datasets whose 32-target membership table ((M+1) * 4 bytes) exceeds shared memory keep it in the workspace
 * and read it through L1 (slower walk, no limit on M below 2^27).
 * detail (nullable): f64[nt,3] = (sum of weak APs, sum of strong APs, number
 * of ground-truth classes) per target, for parity checks.
 */
/* orie_reward_workspace_bytes: exact size (waits for the build the first time).  orie_reward_workspace_bound: an
 * upper bound computed on the host without waiting (events <= min(weak rows, labels x T)); any size in between is
 * accepted as long as the event lists fit, which the reward pass checks on the device. */
size_t orie_reward_workspace_bytes(const orie_index_t *idx, int64_t nt);
size_t orie_reward_workspace_bound(const orie_index_t *idx, int64_t nt);
int orie_reward(const orie_index_t *idx, int64_t t0, int64_t nt, const uint32_t *ens_bits, int64_t N,
                void *workspace, size_t workspace_bytes, double *reward, double *detail, orie_stream_t stream);

/*
 * Per-target sums instead of rewards: sums f64[nt,3] = (sum of weak APs, sum of strong APs, number of ground-truth
 * classes).  full = 0: only the DIFFERENCE sums[.,1] - sums[.,0] is meaningful (terms common to both variants are
 * skipped); full != 0: true AP sums.  Sums are additive over a partition of the classes, which
 * is what class-sharded multi-GPU runs reduce (engine.py: compute_rewards(..., shard="classes")).
 */
int orie_reward_sums(const orie_index_t *idx, int64_t t0, int64_t nt, const uint32_t *ens_bits, int64_t N,
                     void *workspace, size_t workspace_bytes, double *sums, int full, orie_stream_t stream);

/* reward[r] = (N+1) * (sums[r,1] - sums[r,0]) / (sums[r,2] * T), 0 where sums[r,2] == 0: the last step of a class-sharded
 * run, after the per-target sums of the ranks have been added (reward.py:50,86).  sums f64[nt,3], reward f64[nt]. */
int orie_rewards_from_sums(const double *sums, int64_t nt, int T, int64_t N, double *reward, orie_stream_t stream);

/*
 * Same as orie_reward / orie_reward_sums (reward or sums may be NULL, not both), with CUDA events recorded on
 * `stream` around each kernel; synchronises and
 * writes kernel_ms_host[4] = {label walk, detection walk, AP integration, finalize} in milliseconds.
 * Measurement aid for bench.py's roofline figure; not part of the reference-facing path.
 */
int orie_reward_profile(const orie_index_t *idx, int64_t t0, int64_t nt, const uint32_t *ens_bits, int64_t N,
                        void *workspace, size_t workspace_bytes, double *reward, double *sums, int full,
                        orie_stream_t stream, float *kernel_ms_host);

/*
 * Rank normalisation of a reward vector for one cross-validation fold.  Replaces regression.py:439-441:
 *   train rows (val_mask[i] == 0):  out[i] = (argsort(argsort(train)) + 1)[i] / len(train)
 *   validation rows:                out[i] = sum(train <= reward[i]) / len(train)
 * val_mask nullable (every row is a train row).  Equal train rewards rank in row order (upstream's argsort is
 * unstable, so their order is machine-dependent there).  reward, out: f64[M]; val_mask: u8[M];
 * workspace: orie_rank_workspace_bytes(M) bytes, 256-byte aligned.
 */
size_t orie_rank_workspace_bytes(int64_t M);
int orie_rank_normalize(const double *reward, const uint8_t *val_mask, int64_t M, double *out,
                        void *workspace, size_t workspace_bytes, orie_stream_t stream);

/*
 * DCSB baseline (the "difficult-case based small-big model" estimator): threshold search on the train rows of one
 * cross-validation fold and offloading decisions for every image.  Replaces baseline.py:67-152 (fit_dcsb) with the
 * inputs baseline.py:161-206 builds for --baseline dcsb: the weak detector's outputs (packed block: xyxy boxes,
 * confidences, offsets; box areas as get_area, baseline.py:155-158), the number of ground-truth objects per image and
 * the binarised rewards (reward > 0, baseline.py:166).
 *   val_mask u8[M] nullable (1 = validation row); label_num, reward01 int64[M]; area_steps_host: host array of the
 *   area thresholds to try (pass numpy's np.arange(0.2, 0.9, 0.01) verbatim), n_count: count thresholds 1..n_count
 *   (upstream: 10).
 *   model f64[4] (device) <- confidence threshold, count threshold, area threshold, train hits of the chosen pair;
 *   est int64[M] (device) <- decision of every image (upstream's train_est / val_est, split by val_mask).
 * The confidence bisection (baseline.py:95-106) stops when upstream's tolerance is met; upstream loops forever if it
 * cannot be met (e.g. no train labels), this entry point stops after 1100 halvings.
 */
size_t orie_dcsb_fit_workspace_bytes(int64_t M, int64_t num_weak);
int orie_dcsb_fit(const double *w_box, const double *w_conf, const int64_t *w_off, int64_t M, int64_t num_weak,
                  const uint8_t *val_mask, const int64_t *label_num, const int64_t *reward01,
                  const double *area_steps_host, int n_area, int n_count,
                  double *model, int64_t *est, void *workspace, size_t workspace_bytes, orie_stream_t stream);

/*
 * Measurement aid (profiles/ap_depths.py): orie_reward with the AP kernel also recording, for every (target, class,
 * threshold) it integrates, the trip counts of the reverse sweep and of its shared-state tail loop:
 * depths uint32[nt][C][T][2], zero-initialised by the caller (entries of skipped classes stay zero).
 */
int orie_reward_depths(const orie_index_t *idx, int64_t t0, int64_t nt, const uint32_t *ens_bits, int64_t N,
                       void *workspace, size_t workspace_bytes, double *reward, uint32_t *depths, orie_stream_t stream);

/* Number of kernels this library has launched in this process (all threads). */
long long orie_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ORIE_B200_H */
