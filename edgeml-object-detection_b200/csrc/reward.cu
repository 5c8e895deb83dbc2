// ORIE reward pass (orie_reward) and ensemble construction.
//
// Replaces reward.py:16-52 (compute_orie), lib/metrics.py:89-124 (ap_per_class)
// and :127-148 (compute_ap) without re-sorting or re-matching anything per target.
//
//   K1  walk_kernel      32 targets at a time (one per lane).  A warp streams a segment of
//                        the class-sorted slot array, looks every slot's image up in a
//                        shared-memory table of 32-target membership words, transposes the
//                        32x32 bit block with 5 shuffles so lane j holds target j's membership
//                        word, and keeps per-target running ranks with popc.  It emits, per
//                        target: members per segment, the rank of every member true positive
//                        ("event"), and the rank of the target's own detections.
//   K2  ap_kernel        one warp per (target, group of 32/T classes); lane = (class, IoU
//                        threshold).  A single reverse sweep over the events integrates the
//                        101-point interpolated AP exactly as compute_ap does (precision
//                        envelope = running max in reverse, np.interp's "last knot <= x" rule,
//                        np.trapz), for the weak and the strong variant side by side.
//   K3  finalize_kernel  (N+1) * (mean strong AP - mean weak AP), NaN -> 0 (reward.py:50,86).
//
// All arithmetic that decides a branch or a reward digit is IEEE float64 in upstream's order
// (explicit _rn intrinsics; the library is built with -fmad=false).
#include <stdlib.h>

#include <algorithm>
#include <mutex>

#include "index.cuh"

namespace orie {

// ----------------------------------------------------------------------------
// 32x32 bit-matrix transpose across a warp: in  = row `lane`, bit c = column c
//                                            out = column `lane`, bit r = row r
// ----------------------------------------------------------------------------
// 13 instructions: the 16- and 8-bit stages are one byte permute each (PRMT with a per-lane selector),
// the 4/2/1-bit stages one rotate by a per-lane amount and one bit-select (LOP3) each.
struct Transposer {
    uint32_t sel16, sel8, sh4, sh2, sh1, mk4, mk2, mk1;
    __device__ __forceinline__ void init(int lane) {
        sel16 = (lane & 16) ? 0x3276u : 0x5410u;
        sel8 = (lane & 8) ? 0x3715u : 0x6240u;
        sh4 = (lane & 4) ? 28u : 4u;  mk4 = (lane & 4) ? 0xf0f0f0f0u : 0x0f0f0f0fu;
        sh2 = (lane & 2) ? 30u : 2u;  mk2 = (lane & 2) ? 0xccccccccu : 0x33333333u;
        sh1 = (lane & 1) ? 31u : 1u;  mk1 = (lane & 1) ? 0xaaaaaaaau : 0x55555555u;
    }
    __device__ __forceinline__ uint32_t operator()(uint32_t x) const {
        x = __byte_perm(x, __shfl_xor_sync(kFull, x, 16), sel16);
        x = __byte_perm(x, __shfl_xor_sync(kFull, x, 8), sel8);
        uint32_t y = __shfl_xor_sync(kFull, x, 4); y = __funnelshift_l(y, y, sh4); x = (x & mk4) | (y & ~mk4);
        y = __shfl_xor_sync(kFull, x, 2); y = __funnelshift_l(y, y, sh2); x = (x & mk2) | (y & ~mk2);
        y = __shfl_xor_sync(kFull, x, 1); y = __funnelshift_l(y, y, sh1); x = (x & mk1) | (y & ~mk1);
        return x;
    }
};

// ----------------------------------------------------------------------------
// ensembles
// ----------------------------------------------------------------------------
__global__ void ens_from_indices_kernel(const int32_t *__restrict__ ens_idx, int64_t nt, int64_t N, int64_t t0, int64_t M,
                                        int64_t words, uint32_t *__restrict__ bits, int32_t *__restrict__ status) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nt * N) return;
    const int64_t r = k / N;
    const int64_t e = ens_idx[k];
    if (e < 0 || e >= M || e == t0 + r) { atomicOr(status, 1); return; }
    const uint32_t bit = 1u << (e & 31);
    const uint32_t old = atomicOr(&bits[r * words + (e >> 5)], bit);
    if (old & bit) atomicOr(status, 2);
}

// Philox4x32-10 (Salmon et al., SC'11): counter-based, so a target's ensemble depends only on
// (seed, target) — not on the launch shape or on how targets are sharded over GPUs.
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}

// One warp per target.  Draws `want` distinct images != target into the target's bitmap, then complements the
// bitmap when the ensemble is more than half of the dataset.  Every Philox call yields four 32-bit words per lane,
// each mapped to [0, M-1) without bias (multiply-shift with rejection, Lemire 2019), i.e. 128 candidates per round.
//  * bulk rounds (the quota cannot be reached inside the round): every candidate is inserted with an atomic
//    test-and-set; the SET of accepted images and their count do not depend on the order of the atomics;
//  * last rounds: 32 candidates at a time, duplicates resolved and the quota cut in lane order, so the result is
//    a deterministic function of (seed, target).
// The bitmap is built in shared memory (SMEM = true, one row per warp) or directly in global memory.
constexpr int kSampleWarps = 4;
template <bool SMEM>
__global__ void __launch_bounds__(kSampleWarps * 32)
ens_sample_kernel(int64_t nt, int64_t N, int64_t t0, int64_t M, int64_t words, uint64_t seed_value,
                  const uint64_t *__restrict__ seed_dev, uint32_t *__restrict__ bits_all) {
    extern __shared__ uint32_t sbits[];
    const int64_t r = (int64_t)blockIdx.x * kSampleWarps + (threadIdx.x >> 5);
    if (r >= nt) return;
    const int lane = threadIdx.x & 31;
    const uint64_t seed = seed_dev ? *seed_dev : seed_value;     // from device memory: the call can be replayed in a CUDA graph
    const int64_t target = t0 + r;
    uint32_t *out = bits_all + r * words;
    uint32_t *bits = SMEM ? sbits + (int64_t)(threadIdx.x >> 5) * words : out;
    if (SMEM) {
        for (int64_t w = lane; w < words; w += 32) bits[w] = 0u;
        __syncwarp();
    }
    const uint32_t others = (uint32_t)(M - 1);                 // M < 2^27
    const bool complement = 2 * N > (int64_t)others;
    const int64_t want = complement ? (int64_t)others - N : N;
    const uint32_t reject_below = others ? (0u - others) % others : 0u;   // 2^32 mod others
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    // candidate of one random word: image index, or -1 if the word falls in the rejection zone
    auto candidate = [&](uint32_t rnd) -> int {
        const uint64_t m = (uint64_t)rnd * others;
        if ((uint32_t)m < reject_below) return -1;
        const uint32_t e = (uint32_t)(m >> 32);
        return (int)(e + (e >= (uint32_t)target ? 1u : 0u));
    };
    int64_t have = 0;
    for (uint32_t round = 0; have < want; ++round) {
        const uint4 rnd = philox4x32(make_uint4((uint32_t)target, (uint32_t)(target >> 32), round, (uint32_t)lane), key);
        const uint32_t word[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
        if (have + 128 <= want) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = candidate(word[j]);
                bool fresh = false;
                if (e >= 0) {
                    const uint32_t bit = 1u << (e & 31);
                    fresh = !(atomicOr(&bits[e >> 5], bit) & bit);
                }
                have += __popc(__ballot_sync(kFull, fresh));
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (have >= want) break;                       // uniform
                const int e = candidate(word[j]);
                const unsigned peers = __match_any_sync(kFull, e);
                bool ok = e >= 0 && lane == (__ffs(peers) - 1);
                const uint32_t bit = 1u << (e & 31);
                if (ok) ok = !(*(volatile uint32_t *)&bits[e >> 5] & bit);
                const unsigned acc = __ballot_sync(kFull, ok);
                const int before = __popc(acc & ((1u << lane) - 1u));
                if (ok && have + before < want) atomicOr(&bits[e >> 5], bit);
                __syncwarp();
                have += __popc(acc);
            }
        }
    }
    __syncwarp();
    for (int64_t w = lane; w < words; w += 32) {
        uint32_t v = *(volatile uint32_t *)&bits[w];
        if (complement) {
            v = ~v;
            const int64_t lo = w * 32;
            if (lo + 32 > M) v &= (M > lo) ? ((M - lo >= 32) ? 0xffffffffu : ((1u << (M - lo)) - 1u)) : 0u;
            if ((target >> 5) == w) v &= ~(1u << (target & 31));
        }
        if (SMEM || complement) out[w] = v;
    }
}

// ----------------------------------------------------------------------------
// K1: walk
// ----------------------------------------------------------------------------
struct WalkParams {
    int64_t M, nt, ntp, t0;
    IndexMeta *meta;            // device-side sizes of the index (segments, events) and the sticky status word
    int64_t S_cap;              // stride of bqoff rows (host-side upper bound of the segment count)
    int64_t ens_words;
    const uint32_t *ens_bits;   // [nt][ens_words]
    uint32_t *memb_global;      // [ntp / 32][ens_words * 32] membership tables in global memory (GMEM kernels)
    int memb_pairs;             // != 0: memb_global holds the tables of two batches side by side ([ntp / 64][ens_words * 32][2])
    // stream
    const uint32_t *slot_img;
    const int32_t *seg_chunk0, *seg_nch;
    int segs_per_block;
    uint32_t *tot;              // [S][ntp]
    // detection stream only
    const uint16_t *slot_tp;
    const uint32_t *slot_pk;    // image | true-positive mask << 16 (PACKED kernels: at most 65535 images)
    const uint32_t *seg_ev0;    // [S + 1]
    const uint32_t *ev_img;     // dense event stream (the slots holding a true positive, in slot order): image ...
    const uint16_t *ev_mask;    // ... and true-positive mask
    const uint2 *bq;            // per-batch query list (both detectors), ascending by slot
    const uint32_t *bqoff;      // [nbatch][S_cap+1]
    int64_t Ev;                 // capacity of one target's event list in the workspace (>= the index's event count)
    int T;
    uint32_t *evcnt;            // [S][ntp]
    uint32_t *ev;               // [ntp][Ev] event records: rank inside the segment (16 bits) | TP mask << 16
    uint16_t *kseg;             // [S][T][ntp] member true positives of the segment per IoU threshold
    uint32_t *cb_w, *cb_s;
    const int32_t *seg_order;   // detection segments, longest first
    uint32_t *pair_next;        // [ceil(ntp / 64)] next entry of seg_order to hand out, per batch pair (zero on entry)
};


// Membership table of one batch in global memory, for datasets whose table exceeds shared memory
// (more than ~58 k images): memb_global[batch][img] bit j = img in the ensemble of target 32*batch+j.
__global__ void membership_table_kernel(const WalkParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t lb = blockIdx.x, tl = lb * 32 + lane;
    Transposer transpose;
    transpose.init(lane);
    const uint32_t *row = p.ens_bits + tl * p.ens_words;
    const bool live = tl < p.nt;
    uint32_t *out = p.memb_global + lb * p.ens_words * 32;
    for (int64_t w = (int64_t)blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5); w < p.ens_words;
         w += (int64_t)gridDim.y * (blockDim.x >> 5))
        out[w * 32 + lane] = transpose(live ? row[w] : 0u);
}

// THREADS is the block size the kernel is launched with (256 / 512 / 1024, chosen from the size of the
// membership table); the register cap keeps kWalkResident threads resident per SM (at 1536 the compiler
// re-materialised the transposer's per-lane constants inside the chunk loop: 13 % more instructions).  GMEM: the membership table was
// built by membership_table_kernel and is read through L1 instead of shared memory.  PACKED (at most 65535 images):
// one 32-bit load per slot brings the image and the true-positive mask.
//
// Per segment and 32 targets a warp runs two loops:
//  * the slot loop, two chunks per trip so that two table lookups + bit transposes are in flight: running member
//    counts, an event record for every member true positive, the rank of the batch's own detections;
//  * a short dense loop over the segment's events only (32 per trip): members that are true positives at each IoU
//    threshold (one ballot per threshold) — the count the AP sweep starts from, which it would otherwise have to
//    collect by reading every event record of the class.
#ifndef ORIE_WALK_RESIDENT
#define ORIE_WALK_RESIDENT 1280
#endif
constexpr int kWalkResident = ORIE_WALK_RESIDENT;
template <bool DETS, int THREADS, bool GMEM, bool PACKED>
__global__ void __launch_bounds__(THREADS, kWalkResident / THREADS > 0 ? kWalkResident / THREADS : 1)
walk_kernel(const WalkParams p) {
    extern __shared__ uint32_t memb_s[];   // [ens_words * 32]: bit j of memb[img] = img in ensemble of target 32*batch+j
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kWarps = blockDim.x >> 5;
    // GMEM: blocks of one batch are neighbours in launch order (x = segment group), so the tables of the
    // batches in flight at any time stay L2-resident (a few dozen x 200 KB instead of one per resident block)
    const int64_t lb = GMEM ? blockIdx.y : blockIdx.x;   // local batch
    const int64_t yb = GMEM ? blockIdx.x : blockIdx.y;   // segment group
    // the grid is sized from the host's upper bound of the segment count; the exact count lives on the device
    const uint32_t status = p.meta->status;
    const int64_t S = DETS ? p.meta->S : p.meta->SL;
    if (status & (kStatusRows | kStatusClass | kStatusCounts)) return;       // the index build rejected its input
    if (DETS && (int64_t)p.meta->Ev > p.Ev) {                                // event lists would not fit the workspace
        if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) atomicOr(&p.meta->status, kStatusWorkspace);
        return;
    }
    if (yb * p.segs_per_block >= S) return;
    // the targets' event lists are packed with the index's exact event count as their stride, whatever the capacity of
    // the workspace (sized from an upper bound when the host has not waited for the build): the records of one call stay
    // within a few hundred pages instead of one page per target
    const int64_t ev_stride = DETS ? (int64_t)p.meta->Ev : 0;
    const int64_t tl = lb * 32 + lane;             // local target of this lane
    Transposer transpose;
    transpose.init(lane);
    const uint32_t *memb = !GMEM ? memb_s
                           : p.memb_pairs ? p.memb_global + (lb >> 1) * p.ens_words * 64 + (lb & 1) : p.memb_global + lb * p.ens_words * 32;
    const uint32_t memb_stride = GMEM && p.memb_pairs ? 2u : 1u;
    if (!GMEM) {
        const uint32_t *row = p.ens_bits + tl * p.ens_words;
        const bool live = tl < p.nt;
        for (int64_t w = warp; w < p.ens_words; w += kWarps) {
            const uint32_t x = live ? row[w] : 0u;
            memb_s[w * 32 + lane] = transpose(x);
        }
        __syncthreads();
    }
    // table lookup: membership word of an image (bit j: member of the ensemble of target 32*batch+j); image M, the
    // sentinel of padding slots, is a member of nothing.  The shared-memory window address is computed once.
    uint32_t memb_sa = 0u;
    if (!GMEM)        // through an opaque move: otherwise the window base is re-derived in every trip of the slot loop
        asm volatile("mov.u32 %0, %1;" : "=r"(memb_sa) : "r"((uint32_t)__cvta_generic_to_shared(memb_s)));
    auto lookup = [&](uint32_t img) -> uint32_t {
        if (GMEM) return __ldg(memb + img * memb_stride);
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(memb_sa + img * 4u));
        return v;
    };
    const uint32_t sentinel = (uint32_t)p.M;       // image M with an empty mask
    const int64_t gb = (p.t0 >> 5) + lb;           // global batch (query lists are per global batch)
    const int64_t sbeg = yb * p.segs_per_block;
    const int64_t send = min(sbeg + p.segs_per_block, S);
    for (int64_t s = sbeg + warp; s < send; s += kWarps) {
        const int64_t ch0 = p.seg_chunk0[s];
        const int nch = p.seg_nch[s];
        uint32_t cnt = 0;
        uint32_t qi = 0, q_end = 0;
        uint2 nq = make_uint2(0u, 0u);
        int nqc = 0x7fffffff;                      // chunk (relative to the segment) of the next own detection
        uint32_t *evout = nullptr;
        uint32_t ecur = 0;
        if (DETS) {
            evout = p.ev + tl * ev_stride + p.seg_ev0[s];
            asm volatile("" : "+l"(evout));        // kept in registers: otherwise re-derived for every record written
            const uint32_t *o = p.bqoff + gb * (p.S_cap + 1) + s;
            qi = o[0]; q_end = o[1];
            if (qi < q_end) {
                nq = p.bq[qi];
                nqc = (int)(((nq.x & 0x7fffffffu) >> 5) - (uint32_t)ch0);
            }
        }
        // one chunk: `x` = image of MY slot (PACKED: | TP mask << 16), `tpv` = TP mask of MY slot << 16 (PACKED: x itself,
        // the image bits below are ignored), `word` bit l: slot l holds a member of MY target
        auto chunk = [&](const uint32_t tpv, const uint32_t word, const int c) {
            if (DETS) {
                uint32_t eb = __ballot_sync(kFull, tpv >= 0x10000u);    // slots of the chunk holding an event
                while (eb) {
                    const int b = __ffs(eb) - 1;
                    eb &= eb - 1;
                    const uint32_t m = __shfl_sync(kFull, tpv, b);
                    const uint32_t sh = word << (31 - b);               // slot b on top, the slots behind it shifted out
                    if ((int)sh < 0) evout[ecur++] = (cnt + __popc(sh)) | (m & 0xffff0000u);   // 1-based rank <= 65504 (index.cu)
                }
                while (nqc <= c) {       // uniform: own detections (weak: sitting, strong: inserted) in this chunk
                    if (lane == (int)(nq.y >> 27))
                        ((nq.x >> 31) ? p.cb_s : p.cb_w)[nq.y & 0x07ffffffu] = cnt + __popc(word & ((1u << (nq.x & 31u)) - 1u));
                    ++qi;
                    nqc = 0x7fffffff;
                    if (qi < q_end) {
                        nq = p.bq[qi];
                        nqc = (int)(((nq.x & 0x7fffffffu) >> 5) - (uint32_t)ch0);
                    }
                }
            }
            cnt += __popc(word);
        };
        const uint32_t *sp = ((DETS && PACKED) ? p.slot_pk : p.slot_img) + ch0 * 32 + lane;
        const uint16_t *stp = (DETS && !PACKED) ? p.slot_tp + ch0 * 32 + lane : nullptr;
        // software pipeline: the next two chunks' slots are in flight while two are being worked on; a segment with an
        // odd number of chunks ends with a chunk of sentinels
        uint32_t x0 = sp[0], x1 = nch > 1 ? sp[32] : sentinel;
        uint32_t t0 = 0u, t1 = 0u;
        if (DETS && !PACKED) { t0 = stp[0]; t1 = nch > 1 ? (uint32_t)stp[32] : 0u; }
        for (int c = 0; c < nch; c += 2) {
            const uint32_t a = x0, b = x1, ta = t0, tb = t1;
            sp += 64;
            if (c + 2 < nch) x0 = sp[0];
            x1 = c + 3 < nch ? sp[32] : sentinel;
            if (DETS && !PACKED) {
                stp += 64;
                if (c + 2 < nch) t0 = stp[0];
                t1 = c + 3 < nch ? (uint32_t)stp[32] : 0u;
            }
            const uint32_t wa = transpose(lookup(PACKED && DETS ? a & 0xffffu : a));
            const uint32_t wb = transpose(lookup(PACKED && DETS ? b & 0xffffu : b));
            chunk(PACKED ? a : ta << 16, wa, c);
            chunk(PACKED ? b : tb << 16, wb, c + 1);
        }
        p.tot[s * p.ntp + tl] = cnt;
        if (DETS) {
            p.evcnt[s * p.ntp + tl] = ecur;
            // member true positives per threshold: 32 events per trip, counts of two thresholds share a register
            // (a segment has at most 65504 slots)
            const uint32_t e0 = p.seg_ev0[s], e1 = p.seg_ev0[s + 1];
            uint32_t kc[ORIE_MAX_THRESHOLDS / 2];
#pragma unroll
            for (int i = 0; i < ORIE_MAX_THRESHOLDS / 2; ++i) kc[i] = 0u;
            for (uint32_t w0 = e0 & ~31u; w0 < e1; w0 += 32u) {
                const uint32_t e = w0 + lane;
                const bool in = e >= e0 && e < e1;
                const uint32_t img = in ? p.ev_img[e] : sentinel;
                const uint32_t mask = in ? (uint32_t)p.ev_mask[e] : 0u;
                const uint32_t word = transpose(lookup(img));
#pragma unroll
                for (int t = 0; t < ORIE_MAX_THRESHOLDS; ++t) {
                    if (t < p.T) {          // uniform
                        const uint32_t tb = __ballot_sync(kFull, (mask >> t) & 1u);
                        kc[t >> 1] += (uint32_t)__popc(word & tb) << ((t & 1) * 16);
                    }
                }
            }
            uint16_t *ko = p.kseg + (s * p.T) * p.ntp + tl;
#pragma unroll
            for (int t = 0; t < ORIE_MAX_THRESHOLDS; ++t)
                if (t < p.T) ko[(int64_t)t * p.ntp] = (uint16_t)(kc[t >> 1] >> ((t & 1) * 16));
        }
    }
}

// ---- the detection walk for TWO batches (64 targets) per warp.  What does not depend on the target — the slot loads,
// the loop and the search for the chunk's events — is paid once for both; the membership table holds the two batches'
// words side by side (one 64-bit lookup), and the two bit transposes are independent instruction streams.  Used when two
// tables fit shared memory four blocks to the SM (or live in global memory); same outputs as walk_kernel<true, ...>.
__global__ void membership_table2_kernel(const WalkParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t lp = blockIdx.x, tlA = lp * 64 + lane, tlB = tlA + 32;
    Transposer transpose;
    transpose.init(lane);
    const uint32_t *rowA = p.ens_bits + tlA * p.ens_words, *rowB = p.ens_bits + tlB * p.ens_words;
    const bool liveA = tlA < p.nt, liveB = tlB < p.nt;
    uint2 *out = (uint2 *)p.memb_global + lp * p.ens_words * 32;
    for (int64_t w = (int64_t)blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5); w < p.ens_words;
         w += (int64_t)gridDim.y * (blockDim.x >> 5))
        out[w * 32 + lane] = make_uint2(transpose(liveA ? rowA[w] : 0u), transpose(liveB ? rowB[w] : 0u));
}

// Own detections of one batch's 32 images inside one segment (walk2_kernel): a window of 32 entries of the batch's
// query list, one per lane.  `wqc` = chunk of the lane's entry relative to the segment (INT_MAX beyond the end of the
// list range), `next` (uniform) = chunk of the first entry that has not been handled.  An entry of image i is answered
// by shuffling the running count and the membership word of lane i to the lane that holds the entry; all entries of a
// chunk are answered in one step, and the list is read 32 entries at a time instead of one dependent load per entry.
struct QueryWindow {
    uint2 wq;
    int wqc, next;
    uint32_t base, end;
    __device__ __forceinline__ void load(const uint2 *bq, uint32_t ch0, int lane) {
        const uint32_t i = base + (uint32_t)lane;
        wq = make_uint2(0u, 0u);
        wqc = 0x7fffffff;
        if (i < end) {
            wq = bq[i];
            wqc = (int)(((wq.x & 0x7fffffffu) >> 5) - ch0);
        }
    }
    __device__ __forceinline__ void open(const uint2 *bq, uint32_t first, uint32_t last, uint32_t ch0, int lane) {
        base = first; end = last;
        load(bq, ch0, lane);
        next = __reduce_min_sync(kFull, wqc);
    }
    __device__ __forceinline__ void chunk(const uint2 *bq, int c, uint32_t ch0, int lane, uint32_t word, uint32_t cnt,
                                          uint32_t *cb_w, uint32_t *cb_s) {
        while (next == c) {                   // uniform
            const bool hit = wqc == c;
            const int src = (int)(wq.y >> 27);
            const uint32_t wsrc = __shfl_sync(kFull, word, src), csrc = __shfl_sync(kFull, cnt, src);
            if (hit) ((wq.x >> 31) ? cb_s : cb_w)[wq.y & 0x07ffffffu] = csrc + __popc(wsrc & ((1u << (wq.x & 31u)) - 1u));
            int from = c + 1;                 // entries of chunk c in this window are done ...
            if (__shfl_sync(kFull, wqc, 31) <= c) {      // ... and so is the window: the next 32 entries (may continue chunk c)
                base += 32u;
                load(bq, ch0, lane);
                from = c;
            }
            next = __reduce_min_sync(kFull, wqc >= from ? wqc : 0x7fffffff);
        }
    }
};

#ifndef ORIE_WALK2_PIPE
#define ORIE_WALK2_PIPE 1
#endif
constexpr int kWalk2Threads = 256, kWalk2Blocks = 4;
template <bool GMEM, bool PACKED>
__global__ void __launch_bounds__(kWalk2Threads, kWalk2Blocks)
walk2_kernel(const WalkParams p) {
    extern __shared__ uint2 memb2_s[];     // [ens_words * 32]: .x / .y = membership words of the pair's two batches
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kWalk2Threads / 32;
    const int64_t lp = GMEM ? blockIdx.y : blockIdx.x;   // local batch pair (GMEM: the pair's blocks are neighbours in launch order)
    const uint32_t status = p.meta->status;
    const int64_t S = p.meta->S;
    if (status & (kStatusRows | kStatusClass | kStatusCounts)) return;
    if ((int64_t)p.meta->Ev > p.Ev) {
        if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) atomicOr(&p.meta->status, kStatusWorkspace);
        return;
    }
    const int64_t ev_stride = (int64_t)p.meta->Ev;        // packed event lists (see walk_kernel)
    const int64_t tlA = lp * 64 + lane, tlB = tlA + 32;   // local targets of this lane
    const bool hasB = lp * 64 + 32 < p.ntp;              // the pair's second batch exists in this call
    Transposer transpose;
    transpose.init(lane);
    const uint2 *memb = GMEM ? (const uint2 *)p.memb_global + lp * p.ens_words * 32 : memb2_s;
    if (!GMEM) {
        const uint32_t *rowA = p.ens_bits + tlA * p.ens_words, *rowB = p.ens_bits + tlB * p.ens_words;
        const bool liveA = tlA < p.nt, liveB = tlB < p.nt;
        for (int64_t w = warp; w < p.ens_words; w += kWarps)
            memb2_s[w * 32 + lane] = make_uint2(transpose(liveA ? rowA[w] : 0u), transpose(liveB ? rowB[w] : 0u));
    }
    __syncthreads();
    uint32_t memb_sa = 0u;
    if (!GMEM) asm volatile("mov.u32 %0, %1;" : "=r"(memb_sa) : "r"((uint32_t)__cvta_generic_to_shared(memb2_s)));
    auto lookup = [&](uint32_t img) -> uint2 {
        if (GMEM) return __ldg(memb + img);
        uint2 v;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(memb_sa + img * 8u));
        return v;
    };
    const uint32_t sentinel = (uint32_t)p.M;
    const int64_t gbA = (p.t0 >> 5) + lp * 2;      // global batch of the first of the two
    // Several blocks work on one batch pair (all resident at once); their warps take the pair's segments one by one,
    // longest first (seg_order), from a counter in global memory.  A fixed split left warps idle until the longest
    // share of their block was done (segments of different classes differ in length) and cut the launch into two
    // waves of blocks with a half-empty second one.
    uint32_t *next_seg = p.pair_next + lp;
    for (;;) {
        uint32_t take = 0;
        if (lane == 0) take = atomicAdd(next_seg, 1u);
        take = __shfl_sync(kFull, take, 0);
        if ((int64_t)take >= S) break;
        const int64_t s = p.seg_order[take];
        const int64_t ch0 = p.seg_chunk0[s];
        const int nch = p.seg_nch[s];
        uint32_t cntA = 0, cntB = 0, ecurA = 0, ecurB = 0;
        // own detections (weak: sitting in a slot, strong: inserted in front of one) of each batch's 32 images whose slot
        // lies in this segment, ascending by slot: a window of 32 list entries per batch, one per lane, with the chunk
        // (relative to the segment) of the lane's entry and, uniform, the chunk of the first entry not yet handled
        QueryWindow qA, qB;
        const uint32_t ev0 = p.seg_ev0[s];
        uint32_t *evoutA = p.ev + tlA * ev_stride + ev0, *evoutB = p.ev + tlB * ev_stride + ev0;
        asm volatile("" : "+l"(evoutA), "+l"(evoutB));
        {
            const uint32_t *o = p.bqoff + gbA * (p.S_cap + 1) + s;
            qA.open(p.bq, o[0], o[1], (uint32_t)ch0, lane);
            if (hasB) qB.open(p.bq, o[p.S_cap + 1], o[p.S_cap + 2], (uint32_t)ch0, lane);
            else qB.open(p.bq, 0u, 0u, (uint32_t)ch0, lane);
        }
        const uint32_t *sp = (PACKED ? p.slot_pk : p.slot_img) + ch0 * 32 + lane;
        const uint16_t *stp = PACKED ? nullptr : p.slot_tp + ch0 * 32 + lane;
#if ORIE_WALK2_PIPE
        // software pipeline, two deep: the slot words of chunk c + 2 are in flight, the table lookup and the two bit
        // transposes of chunk c + 1 are issued before the events and own detections of chunk c are worked on (their
        // shuffle latency hides behind that work); past the end of the segment the pipeline is fed sentinels
        uint32_t xc = sp[0], xn = nch > 1 ? sp[32] : sentinel;
        uint32_t tc = 0u, tn = 0u;
        if (!PACKED) { tc = stp[0]; tn = nch > 1 ? (uint32_t)stp[32] : 0u; }
        uint32_t wordA, wordB;                     // bit l: slot l holds a member of MY target
        {
            const uint2 w2 = lookup(PACKED ? xc & 0xffffu : xc);
            wordA = transpose(w2.x); wordB = transpose(w2.y);
        }
        for (int c = 0; c < nch; ++c) {
            const uint32_t tpv = PACKED ? xc : tc << 16;   // TP mask of MY slot in bits 16..31
            xc = xn; tc = tn;
            sp += 32;
            xn = c + 2 < nch ? sp[32] : sentinel;
            if (!PACKED) {
                stp += 32;
                tn = c + 2 < nch ? (uint32_t)stp[32] : 0u;
            }
            const uint2 n2 = lookup(PACKED ? xc & 0xffffu : xc);
            const uint32_t nextA = transpose(n2.x), nextB = transpose(n2.y);
#else
        uint32_t xn = sp[0], tn = PACKED ? 0u : (uint32_t)stp[0];
        for (int c = 0; c < nch; ++c) {
            const uint32_t x = xn;
            const uint32_t tpv = PACKED ? x : tn << 16;    // TP mask of MY slot in bits 16..31
            sp += 32;
            if (c + 1 < nch) xn = sp[0];
            if (!PACKED) {
                stp += 32;
                if (c + 1 < nch) tn = stp[0];
            }
            const uint2 w2 = lookup(PACKED ? x & 0xffffu : x);
            const uint32_t wordA = transpose(w2.x), wordB = transpose(w2.y);   // bit l: slot l holds a member of MY target
#endif
            uint32_t eb = __ballot_sync(kFull, tpv >= 0x10000u);               // slots of the chunk holding an event
            while (eb) {
                const int b = __ffs(eb) - 1;
                eb &= eb - 1;
                const uint32_t m = __shfl_sync(kFull, tpv, b) & 0xffff0000u;
                const uint32_t sa = wordA << (31 - b), sb = wordB << (31 - b);  // slot b on top, the slots behind it shifted out
                if ((int)sa < 0) evoutA[ecurA++] = (cntA + __popc(sa)) | m;     // 1-based rank <= 65504 (index.cu)
                if ((int)sb < 0) evoutB[ecurB++] = (cntB + __popc(sb)) | m;
            }
            qA.chunk(p.bq, c, (uint32_t)ch0, lane, wordA, cntA, p.cb_w, p.cb_s);
            qB.chunk(p.bq, c, (uint32_t)ch0, lane, wordB, cntB, p.cb_w, p.cb_s);
            cntA += __popc(wordA);
            cntB += __popc(wordB);
#if ORIE_WALK2_PIPE
            wordA = nextA; wordB = nextB;
#endif
        }
        p.tot[s * p.ntp + tlA] = cntA;
        p.evcnt[s * p.ntp + tlA] = ecurA;
        if (hasB) {
            p.tot[s * p.ntp + tlB] = cntB;
            p.evcnt[s * p.ntp + tlB] = ecurB;
        }
        // member true positives per threshold (see walk_kernel)
        const uint32_t e0 = ev0, e1 = p.seg_ev0[s + 1];
        uint32_t kcA[ORIE_MAX_THRESHOLDS / 2], kcB[ORIE_MAX_THRESHOLDS / 2];
#pragma unroll
        for (int i = 0; i < ORIE_MAX_THRESHOLDS / 2; ++i) kcA[i] = kcB[i] = 0u;
        for (uint32_t w0 = e0 & ~31u; w0 < e1; w0 += 32u) {
            const uint32_t e = w0 + lane;
            const bool in = e >= e0 && e < e1;
            const uint32_t img = in ? p.ev_img[e] : sentinel;
            const uint32_t mask = in ? (uint32_t)p.ev_mask[e] : 0u;
            const uint2 w2 = lookup(img);
            const uint32_t wordA = transpose(w2.x), wordB = transpose(w2.y);
#pragma unroll
            for (int t = 0; t < ORIE_MAX_THRESHOLDS; ++t) {
                if (t < p.T) {          // uniform
                    const uint32_t tb = __ballot_sync(kFull, (mask >> t) & 1u);
                    kcA[t >> 1] += (uint32_t)__popc(wordA & tb) << ((t & 1) * 16);
                    kcB[t >> 1] += (uint32_t)__popc(wordB & tb) << ((t & 1) * 16);
                }
            }
        }
        uint16_t *ko = p.kseg + (s * p.T) * p.ntp + tlA;
#pragma unroll
        for (int t = 0; t < ORIE_MAX_THRESHOLDS; ++t) {
            if (t < p.T) {
                ko[(int64_t)t * p.ntp] = (uint16_t)(kcA[t >> 1] >> ((t & 1) * 16));
                if (hasB) ko[(int64_t)t * p.ntp + 32] = (uint16_t)(kcB[t >> 1] >> ((t & 1) * 16));
            }
        }
    }
}

// ----------------------------------------------------------------------------
// K2: AP integration
// ----------------------------------------------------------------------------
struct Grid101 {
    double cw[102];   // cw[i]  = sum_{g<i} W_g,      W = trapezoid weight of grid point g of np.linspace(0,1,101)
    double cwx[102];  // cwx[i] = sum_{g<i} W_g x_g
    uint32_t ge[5];   // bit g: x[g] >= correctly rounded g/100 (decides exact rational ties, see ApVar); ge[4]: all set
};

struct ApParams {
    int64_t M, C, nt, ntp, t0;
    int T, cls_per_warp;
    int64_t class_groups;
    int64_t Ev;                 // stride of the per-target event lists in the workspace
    const IndexMeta *meta;
    const int32_t *cls_seg0, *seg_chunk0, *lcls_seg0, *cls_order;
    const uint16_t *act_cls;    // [M][C] classes with an own detection, per image
    const uint32_t *nact;       // [M]
    const uint32_t *seg_ev0;
    const uint32_t *tot, *evcnt, *totL;
    const uint16_t *kseg;       // [S][T][ntp] member true positives per segment and threshold (walk_kernel)
    const uint32_t *ev;
    const uint32_t *gtcnt;
    const int64_t *w_off, *s_off;
    const uint16_t *own_w_cs, *own_s_cs, *own_w_m, *own_s_m;
    const uint32_t *own_w_q, *own_s_q, *cb_w, *cb_s;
    double *partial;    // [ntp][class_groups][3]
    unsigned long long *next_item;   // work counter of the persistent warps (zero on entry)
    uint32_t *depths;   // measurement only (orie_reward_depths): [nt][C][T][2] loop trips of the sweep / its tail
};

// Reverse-sweep state of one (class, threshold, variant) AP integral; oracle/event_model.py:_Var
// is the CPU statement of the same thing and carries the argument for the three shortcuts:
//  * below the last true positive np.interp's two knots carry the same envelope value, so the
//    curve is y(x) = E = max_{k' > k} k'/p_{k'};
//  * "x_g >= fl(k/n_l)" is decided in integers: 100 k = q n_l + r, first grid point = q + (r > 0)
//    (exact ties by the constant ge table); q, r are updated incrementally as k decreases;
//  * a run of grid points sharing one E contributes E * (cw[hi+1] - cw[lo]) to np.trapz; only the
//    tail beyond the last true positive is a genuine linear ramp (lib/metrics.py:137-144), integrated
//    in closed form with cw and cwx.
// 1/p to ~1 ulp without the IEEE division sequence: float seed + two Newton steps.  Only used for
// envelope values that are multiplied into the integral (tolerance 1e-9 on mAP), never for a branch
// that changes which grid points are sampled.
__device__ __forceinline__ double fast_ratio(uint32_t num, uint32_t den) {
    const double d = (double)den;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));     // MUFU.RCP64H: ~20 good bits
    r = __fma_rn(r, __fma_rn(-d, r, 1.0), r);
    r = __fma_rn(r, __fma_rn(-d, r, 1.0), r);
    return __dmul_rn((double)num, r);
}

struct ApVar {
    double ap, E;
    uint32_t k, n_l;
    int q, r;             // 100 * k == q * n_l + r, 0 <= r < n_l   (k = true positives still in front)
    int dq, dr;           // 100 == dq * n_l + dr
    int g;                // highest grid point not yet integrated
    bool dead;

    // ge == nullptr (the kernel passes that when every bit of the tie table is set — Grid101::ge[4], which holds for
    // np.linspace's grid): an exact tie always counts as "reached" and the table lookup is skipped
    __device__ __forceinline__ int first_grid(const uint32_t *ge) const {
        int gl = q + (r > 0);
        if (ge != nullptr && r == 0 && gl <= 100 && !((ge[gl >> 5] >> (gl & 31)) & 1u)) ++gl;
        return gl;
    }
    __device__ __forceinline__ void init(const double *cw, const double *cwx, const uint32_t *ge, uint32_t K, uint32_t n_p,
                                         uint32_t nl) {
        ap = 0.0; E = 0.0; g = 99; k = K; n_l = nl; q = 0; r = 0;
        dead = (K == 0 || n_p == 0);
        if (dead) return;
        if (K <= 42000000u) {               // 100 K fits 32 bits: skip the 64-bit division routine
            const uint32_t a = K * 100u;
            q = (int)(a / nl);
            r = (int)(a - (uint32_t)q * nl);
        } else {
            const uint64_t a = (uint64_t)K * 100ull;
            q = (int)(a / nl);
            r = (int)(a - (uint64_t)q * nl);
        }
        dq = (int)(100u / nl);
        dr = (int)(100u - (uint32_t)dq * nl);
        const int gl = first_grid(ge);
        if (gl <= g) {
            const double r_k = __ddiv_rn((double)K, (double)nl);
            const double env = __ddiv_rn((double)K, (double)n_p);
            const double slope = __ddiv_rn(__dsub_rn(0.0, env), __dsub_rn(1.0, r_k));
            const double sw = __dsub_rn(cw[g + 1], cw[gl]);
            const double swx = __dsub_rn(cwx[g + 1], cwx[gl]);
            ap = __dadd_rn(__dmul_rn(slope, __dsub_rn(swx, __dmul_rn(r_k, sw))), __dmul_rn(env, sw));
            g = gl - 1;
        }
    }
    // the k-th true positive (k = current k) sits at 1-based rank pos
    __device__ __forceinline__ void step(const double *cw, const uint32_t *ge, uint32_t pos) {
        if (dead) return;
        E = fmax(E, fast_ratio(k, pos));
        --k;
        q -= dq;
        r -= dr;
        if (r < 0) { r += (int)n_l; --q; }
        const int gl = first_grid(ge);
        if (gl <= g) {
            ap = __dadd_rn(ap, __dmul_rn(E, __dsub_rn(cw[g + 1], cw[gl])));
            g = gl - 1;
        }
    }
};

// Cursor over the target's own detections of one class in one variant, walked from the back;
// the current entry is cached in registers.
struct OwnCursor {
    const uint32_t *q, *cb;
    const uint16_t *m;
    int ib, lo;           // indices relative to the image's own list
    uint32_t cq, ccb;
    bool ctp, valid;
    __device__ __forceinline__ void load(int t) {
        valid = ib >= lo;
        if (valid) { cq = q[ib]; ccb = cb[ib]; ctp = (m[ib] >> t) & 1; }
    }
    __device__ __forceinline__ void init(const uint32_t *q_, const uint32_t *cb_, const uint16_t *m_, int lo_, int hi_, int t) {
        q = q_; cb = cb_; m = m_; lo = lo_; ib = hi_ - 1; cq = 0; ccb = 0; ctp = false;
        load(t);
    }
    __device__ __forceinline__ uint32_t before() const { return (uint32_t)(ib - lo + 1); }   // own entries still in front
    // own detections of the current segment that rank behind position `pos` (pos == 0: all of the segment's)
    __device__ __forceinline__ void drain(ApVar &v, const double *cw, const uint32_t *ge, uint32_t base, uint32_t slot0,
                                          uint32_t pos, int t) {
        while (valid && cq >= slot0 && base + ccb >= pos) {
            if (ctp) v.step(cw, ge, base + ccb + before());
            --ib;
            load(t);
        }
    }
};

#ifndef ORIE_AP_THREADS
#define ORIE_AP_THREADS 128
#endif
#ifndef ORIE_AP_BLOCKS
#define ORIE_AP_BLOCKS (1024 / ORIE_AP_THREADS)
#endif
constexpr int kApThreads = ORIE_AP_THREADS;

// One warp per (target, group of 32/T classes); lane = (class slot, IoU threshold).
//
// FULL = false (normal operation): only the DIFFERENCE of the two variants' integrals is needed, so
//  * a class in which the target has no detection from either detector is skipped (identical integrals): the warp's
//    classes come from the image's list of classes WITH an own detection (act_cls, built with the index), so the
//    lanes of the warps that do run are all busy and the other warps of the target leave at once;
//  * the sweep of a (class, threshold) stops as soon as the variants can no longer differ: once every own
//    detection lies behind (lower confidence), both variants see the same ranks, the same count of remaining true
//    positives and the same grid pointer; from then on they take the max with the same ratios, so when their
//    envelope values meet they stay equal and every later (higher-confidence) term cancels exactly.
//    The sums written are then partial (common part omitted in both).
// FULL = true (detail requested): every class, both integrals completed, the sums are the true AP sums.
//
// The number of ground-truth classes (the mean's denominator, lib/metrics.py:104-107) is counted separately, lane =
// class, by the first ceil(C / 32) warps of the target.  The number of member true positives of a class at every
// threshold (K, where the reverse sweep starts) comes from the walk (kseg).
// MODE (orie_tuning_t::ap_mode selects among the instantiations): 2 = the warp's classes come from the image's
// active-class list (default), 0 = fixed cls_order groups.
template <bool FULL, int MODE, bool DEPTHS = false>
__global__ void __launch_bounds__(kApThreads, ORIE_AP_BLOCKS)
ap_kernel(const ApParams p, const Grid101 grid) {
    __shared__ double cw[102];
    __shared__ double cwx[102];
    __shared__ uint32_t ge_s[5];
    if (p.meta->status) return;                 // uniform: rejected input or a workspace flagged too small by the walk
    for (int i = threadIdx.x; i < 102; i += kApThreads) { cw[i] = grid.cw[i]; cwx[i] = grid.cwx[i]; }
    if (threadIdx.x < 5) ge_s[threadIdx.x] = grid.ge[threadIdx.x];
    __syncthreads();
    const uint32_t *ge = grid.ge[4] ? nullptr : ge_s;      // uniform
    const int lane = threadIdx.x & 31;
    // Persistent warps: every warp takes the next (class group, target) item from a counter in global memory (zeroed by
    // the walk) until none is left.  A target's first group holds its deepest sweeps (act_cls), and the items are
    // numbered group-major, so the long items are handed out first and the short ones fill the end; and a warp that
    // finishes a short item does not wait for the longest item of its block to free the slot.
    const int64_t items = p.nt * p.class_groups;
    for (;;) {
    unsigned long long took = 0;
    if (lane == 0) took = atomicAdd(p.next_item, 1ull);
    const int64_t item = (int64_t)__shfl_sync(kFull, took, 0);
    if (item >= items) break;
#ifdef ORIE_AP_TARGET_MAJOR
    const int64_t tl = item / p.class_groups, grp = item - tl * p.class_groups;
#else
    const int64_t grp = item / p.nt, tl = item - grp * p.nt;
#endif
    const int64_t j = p.t0 + tl;
    const uint32_t *tot = p.tot + tl, *evcnt = p.evcnt + tl;
    // the lists are laid out with the exact event count of the index as their stride (walk_kernel)
    const uint32_t *ev = p.ev + tl * (int64_t)p.meta->Ev;

    // ---- classes with ground truth in E + {target}: lane = class
    double has_gt = 0.0;
    if (grp * 32 < p.C) {
        const int64_t c = grp * 32 + lane;
        bool has = false;
        if (c < p.C) {
            uint32_t n_l = p.gtcnt[j * p.C + c];
            for (int ls = p.lcls_seg0[c]; ls < p.lcls_seg0[c + 1] && n_l == 0; ++ls) n_l += p.totL[(int64_t)ls * p.ntp + tl];
            has = n_l > 0;
        }
        has_gt = (double)__popc(__ballot_sync(kFull, has));
    }

    // ---- AP integrals of this warp's classes
    constexpr bool kActList = !FULL && (MODE & 2);
    const int64_t nact = kActList ? (int64_t)p.nact[j] : p.C;
    const int64_t first = grp * p.cls_per_warp;
    double ap_w = 0.0, ap_s = 0.0;
    if (first < nact) {
        const int slot = lane / p.T, t = lane % p.T;
        const int64_t ci = first + slot;
        const bool active = slot < p.cls_per_warp && ci < nact;
        int c = 0, wa = 0, wb = 0, sa = 0, sb = 0;
        uint32_t n_l = 0;
        if (active) {
            c = kActList ? (int)p.act_cls[j * p.C + ci] : p.cls_order[ci];     // classes of similar size share a warp
            n_l = p.gtcnt[j * p.C + c];
            for (int ls = p.lcls_seg0[c]; ls < p.lcls_seg0[c + 1]; ++ls) n_l += p.totL[(int64_t)ls * p.ntp + tl];
            const uint16_t *wcs = p.own_w_cs + j * (p.C + 1) + c, *scs = p.own_s_cs + j * (p.C + 1) + c;
            wa = wcs[0]; wb = wcs[1]; sa = scs[0]; sb = scs[1];
        }
        // the target has no detection of this class from either detector: both variants are the same integral
        const bool same = (wb == wa) && (sb == sa);
        const bool need = active && n_l > 0 && (FULL || !same);
        // members of the class (n_ens) and member true positives at this lane's threshold (K_ens): sums over the
        // class's segments of what the walk counted
        uint32_t n_ens = 0, K_ens = 0;
        if (need) {
            const int s0 = p.cls_seg0[c], s1 = p.cls_seg0[c + 1];
            const uint16_t *ks = p.kseg + tl + (int64_t)t * p.ntp;
            for (int s = s0; s < s1; ++s) {
                n_ens += tot[(int64_t)s * p.ntp];
                K_ens += ks[(int64_t)s * p.T * p.ntp];
            }
            const uint32_t *wq = p.own_w_q + p.w_off[j], *wcb = p.cb_w + p.w_off[j];
            const uint32_t *sq = p.own_s_q + p.s_off[j], *scb = p.cb_s + p.s_off[j];
            const uint16_t *wm = p.own_w_m + p.w_off[j], *sm = p.own_s_m + p.s_off[j];
            uint32_t K_w = K_ens, K_s = K_ens;
            for (int i = wa; i < wb; ++i) K_w += (wm[i] >> t) & 1;
            for (int i = sa; i < sb; ++i) K_s += (sm[i] >> t) & 1;
            ApVar vw, vs;
            vw.init(cw, cwx, ge, K_w, n_ens + (uint32_t)(wb - wa), n_l);
            vs.init(cw, cwx, ge, K_s, n_ens + (uint32_t)(sb - sa), n_l);
            if (same) vs.dead = true;             // FULL only: one integral serves both variants
            if (!(vw.dead && vs.dead)) {
                OwnCursor ow, os;
                ow.init(wq, wcb, wm, wa, wb, t);
                os.init(sq, scb, sm, sa, sb, t);
                // one flat loop over (segment, event) pairs, last to first, so that lanes working on classes
                // with different segment structure do not wait for each other at segment boundaries
                int s = s1, i = -1;
                uint32_t base = n_ens, slot0 = 0;
                const uint32_t *e = ev;
                bool tail = false;
                uint32_t trips_main = 0, trips_tail = 0;
                for (;;) {
                    if (DEPTHS) ++trips_main;
                    if (!FULL && !ow.valid && !os.valid) {
                        // every own detection lies behind: no true positive left in front of either variant -> nothing
                        // can change any more; both alive with the same count and grid pointer -> continue in the tail loop
                        const uint32_t kw = vw.dead ? 0u : vw.k, ks = vs.dead ? 0u : vs.k;
                        if (kw == 0u && ks == 0u) break;
                        if (!vw.dead && !vs.dead && kw == ks && vw.g == vs.g) { tail = true; break; }
                    }
                    if (i < 0) {
                        if (s < s1) {                       // leaving a segment: its remaining own detections
                            ow.drain(vw, cw, ge, base, slot0, 0u, t);
                            os.drain(vs, cw, ge, base, slot0, 0u, t);
                        }
                        if (s == s0) break;
                        --s;
                        base -= tot[(int64_t)s * p.ntp];
                        slot0 = (uint32_t)p.seg_chunk0[s] * 32u;
                        e = ev + p.seg_ev0[s];
                        i = (int)evcnt[(int64_t)s * p.ntp] - 1;
                        continue;
                    }
                    const uint32_t rec = e[i];
                    const uint32_t pos = base + (rec & 0xffffu);
                    // own detections that rank behind this event come first in the reverse sweep
                    ow.drain(vw, cw, ge, base, slot0, pos, t);
                    os.drain(vs, cw, ge, base, slot0, pos, t);
                    if ((rec >> (16 + t)) & 1u) {
                        vw.step(cw, ge, pos + ow.before());
                        vs.step(cw, ge, pos + os.before());
                    }
                    --i;
                }
                // Tail: the variants now see the same ranks and share k, q, r, g (kept in vw); only the envelope values
                // and the integrals differ, and they stop differing when the envelopes meet.  One ratio per event
                // serves both variants and events that are not true positives at this threshold cost a load and a test.
                while (tail && vw.E != vs.E) {
                    if (DEPTHS) ++trips_tail;
                    if (i < 0) {
                        if (s == s0) break;
                        --s;
                        base -= tot[(int64_t)s * p.ntp];
                        e = ev + p.seg_ev0[s];
                        i = (int)evcnt[(int64_t)s * p.ntp] - 1;
                        continue;
                    }
                    const uint32_t rec = e[i--];
                    if (!((rec >> (16 + t)) & 1u)) continue;
                    const double ratio = fast_ratio(vw.k, base + (rec & 0xffffu));
                    vw.E = fmax(vw.E, ratio);
                    vs.E = fmax(vs.E, ratio);
                    --vw.k;
                    vw.q -= vw.dq;
                    vw.r -= vw.dr;
                    if (vw.r < 0) { vw.r += (int)vw.n_l; --vw.q; }
                    const int gl = vw.first_grid(ge);
                    if (gl <= vw.g) {
                        const double d = __dsub_rn(cw[vw.g + 1], cw[gl]);
                        vw.ap = __dadd_rn(vw.ap, __dmul_rn(vw.E, d));
                        vs.ap = __dadd_rn(vs.ap, __dmul_rn(vs.E, d));
                        vw.g = gl - 1;
                    }
                }
                if (DEPTHS) {
                    uint32_t *d = p.depths + ((tl * p.C + c) * p.T + t) * 2;
                    d[0] = trips_main; d[1] = trips_tail;
                }
            }
            ap_w = vw.dead ? 0.0 : vw.ap;
            ap_s = same ? ap_w : (vs.dead ? 0.0 : vs.ap);
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            ap_w = __dadd_rn(ap_w, __shfl_xor_sync(kFull, ap_w, d));
            ap_s = __dadd_rn(ap_s, __shfl_xor_sync(kFull, ap_s, d));
        }
    }
    if (lane == 0) {
        double *out = p.partial + (tl * p.class_groups + grp) * 3;
        out[0] = ap_w; out[1] = ap_s; out[2] = has_gt;
    }
    }
}

// ----------------------------------------------------------------------------
// K3
// ----------------------------------------------------------------------------
// (N + 1) * (mean strong AP - mean weak AP) from the sums; no ground-truth class: upstream's mean over an empty AP
// table is NaN, stored as 0 (reward.py:50,86)
__device__ __forceinline__ double reward_from_sums(double sw, double ss, double nc, int T, int64_t N) {
    if (!(nc > 0.0)) return 0.0;
    const double cnt = nc * (double)T;
    return __dmul_rn(__dsub_rn(__ddiv_rn(ss, cnt), __ddiv_rn(sw, cnt)), (double)(N + 1));
}

// One warp per target: lane g adds the partial sums of class groups g, g + 32, ...; a fixed shuffle tree adds the lanes
// (the order of the additions depends on the number of groups only, never on how targets are cut into calls).
constexpr int kFinalizeThreads = 128;
__global__ void __launch_bounds__(kFinalizeThreads)
finalize_kernel(const double *__restrict__ partial, int64_t nt, int64_t groups, int T, int64_t N,
                const IndexMeta *__restrict__ meta, double *__restrict__ reward, double *__restrict__ detail) {
    const int lane = threadIdx.x & 31;
    const int64_t tl = (int64_t)blockIdx.x * (kFinalizeThreads / 32) + (threadIdx.x >> 5);
    if (tl >= nt) return;
    double sw = 0.0, ss = 0.0, nc = 0.0;
    if (meta->status) {                         // nothing was computed: make that impossible to miss
        sw = ss = nc = __longlong_as_double(0x7ff8000000000000ll);
    } else {
        for (int64_t g = lane; g < groups; g += 32) {
            const double *q = partial + (tl * groups + g) * 3;
            sw = __dadd_rn(sw, q[0]);
            ss = __dadd_rn(ss, q[1]);
            nc += q[2];
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            sw = __dadd_rn(sw, __shfl_xor_sync(kFull, sw, d));
            ss = __dadd_rn(ss, __shfl_xor_sync(kFull, ss, d));
            nc += __shfl_xor_sync(kFull, nc, d);
        }
    }
    if (lane != 0) return;
    if (reward) reward[tl] = meta->status ? sw : reward_from_sums(sw, ss, nc, T, N);
    if (detail) { detail[tl * 3] = sw; detail[tl * 3 + 1] = ss; detail[tl * 3 + 2] = nc; }
}

// rewards from (reduced) per-target sums: the last step of a class-sharded multi-GPU run, after the all-reduce
__global__ void rewards_from_sums_kernel(const double *__restrict__ sums, int64_t nt, int T, int64_t N, double *__restrict__ reward) {
    const int64_t tl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tl >= nt) return;
    reward[tl] = reward_from_sums(sums[tl * 3], sums[tl * 3 + 1], sums[tl * 3 + 2], T, N);
}

// ----------------------------------------------------------------------------
// workspace
// ----------------------------------------------------------------------------
// The per-segment tables are sized with the host-side segment capacities; the per-target event lists come last and take
// whatever the caller's workspace has left (at least the index's event count, checked on the host once the exact
// count is known, on the device otherwise).
struct WsLayout {
    size_t tot, evcnt, totL, kseg, cb_w, cb_s, partial, counter, memb, ev, fixed;     // counter: AP work counter + per-pair segment counters
};

static bool walk_in_gmem(const orie_index *ix) {
    // beyond 112 KB a shared-memory table would allow only one 1024-thread block per SM; the L1/L2-served global
    // table with six 256-thread blocks per SM is measurably faster there (profiles/: 157 vs 168 ms at M = 50 000)
    return (size_t)ix->ens_words * 32 * 4 > 112 * 1024 || ix->walk_gmem != 0;
}

static WsLayout ws_layout(const orie_index *ix, int64_t nt) {
    const int64_t ntp = round_up(nt, 32);
    WsLayout L;
    size_t o = 0;
    auto take = [&](int64_t bytes) { size_t at = o; o += (size_t)round_up(bytes > 0 ? bytes : 1, 256); return at; };
    L.tot = take(ix->S_cap * ntp * 4);
    L.evcnt = take(ix->S_cap * ntp * 4);
    L.totL = take(ix->SL_cap * ntp * 4);
    L.kseg = take(ix->S_cap * ix->T * ntp * 2);
    L.cb_w = take(ix->Dw * 4);
    L.cb_s = take(ix->Ds * 4);
    L.partial = take(ntp * ix->class_groups * 3 * 8);
    L.counter = take(8 + (ntp / 32 + 1) / 2 * 4);
    // membership tables in global memory, only when they exceed shared memory (or when forced for tests)
    L.memb = take(walk_in_gmem(ix) ? round_up(ntp / 32, 2) * ix->ens_words * 32 * 4 : 0);     // whole batch pairs
    L.ev = o;
    L.fixed = o;
    return L;
}

static size_t ws_total(const orie_index *ix, int64_t nt, int64_t events) {
    return ws_layout(ix, nt).fixed + (size_t)round_up(round_up(nt, 32) * std::max<int64_t>(events, 1) * 4, 256);
}

// one-time (per device) opt-in of the walk kernels to the shared memory their membership table may need
static int walk_attributes() {
    static std::mutex mu;
    static bool done[64] = {};
    int dev = 0;
    ORIE_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev < 0 || dev >= 64 || done[dev]) return ORIE_OK;
    const int smem = 112 * 1024;
    ORIE_CUDA(cudaFuncSetAttribute(walk_kernel<true, 256, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ORIE_CUDA(cudaFuncSetAttribute(walk_kernel<true, 512, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ORIE_CUDA(cudaFuncSetAttribute(walk_kernel<true, 1024, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ORIE_CUDA(cudaFuncSetAttribute(walk_kernel<true, 256, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ORIE_CUDA(cudaFuncSetAttribute(walk_kernel<true, 512, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ORIE_CUDA(cudaFuncSetAttribute(walk_kernel<true, 1024, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ORIE_CUDA(cudaFuncSetAttribute(walk_kernel<false, 256, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ORIE_CUDA(cudaFuncSetAttribute(walk_kernel<false, 512, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ORIE_CUDA(cudaFuncSetAttribute(walk_kernel<false, 1024, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ORIE_CUDA(cudaFuncSetAttribute(walk2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ORIE_CUDA(cudaFuncSetAttribute(walk2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    done[dev] = true;
    return ORIE_OK;
}

}  // namespace orie

using namespace orie;

extern "C" size_t orie_reward_workspace_bytes(const orie_index_t *ix, int64_t nt) {
    if (!ix || nt < 0) return 0;
    if (resolve(ix) != ORIE_OK) return 0;
    return ws_total(ix, nt, ix->Ev);
}

extern "C" size_t orie_reward_workspace_bound(const orie_index_t *ix, int64_t nt) {
    if (!ix || nt < 0) return 0;
    return ws_total(ix, nt, ix->resolved ? ix->Ev : ix->Ev_cap);
}

static int check_range(const orie_index *ix, int64_t t0, int64_t nt, const char *who) {
    if (!ix) { set_error("%s: index is NULL", who); return ORIE_EINVAL; }
    if (t0 < 0 || nt < 0 || t0 + nt > ix->M || (t0 & 31)) {
        set_error("%s: target range [%lld, %lld) must lie inside [0, %lld) and start at a multiple of 32",
                  who, (long long)t0, (long long)(t0 + nt), (long long)ix->M);
        return ORIE_EINVAL;
    }
    return ORIE_OK;
}

extern "C" int orie_ensemble_from_indices(const orie_index_t *ix, int64_t t0, int64_t nt, const int32_t *ens_idx, int64_t N,
                                          uint32_t *ens_bits, int32_t *status, orie_stream_t stream) {
    ORIE_TRY(check_range(ix, t0, nt, "orie_ensemble_from_indices"));
    if (N < 0 || N > ix->M - 1 || !ens_bits || !status || (N > 0 && !ens_idx)) {
        set_error("orie_ensemble_from_indices: need 0 <= N <= M-1 and non-null buffers");
        return ORIE_EINVAL;
    }
    if (nt == 0) return ORIE_OK;
    ORIE_CUDA(cudaMemsetAsync(ens_bits, 0, (size_t)(nt * ix->ens_words) * 4, stream));
    if (N > 0) {
        const int64_t n = nt * N;
        ens_from_indices_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(ens_idx, nt, N, t0, ix->M, ix->ens_words,
                                                                            ens_bits, status);
        ORIE_LAUNCH_CHECK();
    }
    return ORIE_OK;
}

static int ensemble_sample(const orie_index_t *ix, int64_t t0, int64_t nt, int64_t N, uint64_t seed, const uint64_t *seed_dev,
                           uint32_t *ens_bits, orie_stream_t stream) {
    ORIE_TRY(check_range(ix, t0, nt, "orie_ensemble_sample"));
    if (N < 0 || N > ix->M - 1 || !ens_bits) {
        set_error("orie_ensemble_sample: need 0 <= N <= M-1 and a bitmap buffer");
        return ORIE_EINVAL;
    }
    if (nt == 0) return ORIE_OK;
    const size_t smem = (size_t)kSampleWarps * ix->ens_words * 4;
    const unsigned grid = (unsigned)ceil_div(nt, kSampleWarps);
    if (smem <= 48 * 1024) {
        ens_sample_kernel<true><<<grid, kSampleWarps * 32, smem, stream>>>(nt, N, t0, ix->M, ix->ens_words, seed, seed_dev, ens_bits);
    } else {
        ORIE_CUDA(cudaMemsetAsync(ens_bits, 0, (size_t)(nt * ix->ens_words) * 4, stream));
        ens_sample_kernel<false><<<grid, kSampleWarps * 32, 0, stream>>>(nt, N, t0, ix->M, ix->ens_words, seed, seed_dev, ens_bits);
    }
    ORIE_LAUNCH_CHECK();
    return ORIE_OK;
}

extern "C" int orie_ensemble_sample(const orie_index_t *ix, int64_t t0, int64_t nt, int64_t N, uint64_t seed,
                                    uint32_t *ens_bits, orie_stream_t stream) {
    return ensemble_sample(ix, t0, nt, N, seed, nullptr, ens_bits, stream);
}

extern "C" int orie_ensemble_sample_dev(const orie_index_t *ix, int64_t t0, int64_t nt, int64_t N, const uint64_t *seed_dev,
                                        uint32_t *ens_bits, orie_stream_t stream) {
    if (!seed_dev) {
        set_error("orie_ensemble_sample_dev: seed_dev is NULL");
        return ORIE_EINVAL;
    }
    return ensemble_sample(ix, t0, nt, N, 0, seed_dev, ens_bits, stream);
}

template <bool DETS, int THREADS, bool GMEM, bool PACKED>
static int launch_walk_t(dim3 grid, size_t smem, cudaStream_t stream, const WalkParams &p) {
    walk_kernel<DETS, THREADS, GMEM, PACKED><<<grid, THREADS, GMEM ? 0 : smem, stream>>>(p);
    return ORIE_OK;
}
template <bool DETS, bool PACKED>
static int launch_walk_p(dim3 grid, int threads, size_t smem, bool gmem, cudaStream_t stream, const WalkParams &p) {
    if (gmem) return launch_walk_t<DETS, 256, true, PACKED>(grid, smem, stream, p);
    if (threads == 256) return launch_walk_t<DETS, 256, false, PACKED>(grid, smem, stream, p);
    if (threads == 512) return launch_walk_t<DETS, 512, false, PACKED>(grid, smem, stream, p);
    return launch_walk_t<DETS, 1024, false, PACKED>(grid, smem, stream, p);
}
// the label stream has no true-positive masks: never packed
template <bool DETS>
static int launch_walk(dim3 grid, int threads, size_t smem, bool gmem, cudaStream_t stream, const WalkParams &p) {
    if (DETS && p.slot_pk) return launch_walk_p<DETS, DETS>(grid, threads, smem, gmem, stream, p);
    return launch_walk_p<DETS, false>(grid, threads, smem, gmem, stream, p);
}

// np.linspace(0, 1, 101) and the cumulative trapezoid weights of np.trapz over it, bit for bit
static const Grid101 &grid101() {
    static const Grid101 g = [] {
        Grid101 grid;
        double x[101], w[101];
        for (int i = 0; i <= 100; ++i) { x[i] = (double)i * 0.01; w[i] = 0.0; }   // == np.linspace(0, 1, 101) bit for bit
        x[100] = 1.0;
        for (int i = 0; i < 100; ++i) {
            const double d = x[i + 1] - x[i];          // np.diff
            w[i] += d / 2; w[i + 1] += d / 2;
        }
        grid.cw[0] = grid.cwx[0] = 0.0;
        for (int i = 0; i <= 100; ++i) {
            grid.cw[i + 1] = grid.cw[i] + w[i];
            grid.cwx[i + 1] = grid.cwx[i] + w[i] * x[i];
        }
        for (int k = 0; k < 4; ++k) grid.ge[k] = 0;
        for (int i = 0; i <= 100; ++i)
            if (x[i] >= (double)i / 100.0) grid.ge[i >> 5] |= 1u << (i & 31);
        grid.ge[4] = (grid.ge[0] == 0xffffffffu && grid.ge[1] == 0xffffffffu && grid.ge[2] == 0xffffffffu &&
                      grid.ge[3] == 0x1fu) ? 1u : 0u;
        return grid;
    }();
    return g;
}

static int run_reward(const orie_index_t *ix, int64_t t0, int64_t nt, const uint32_t *ens_bits, int64_t N,
                      void *workspace, size_t workspace_bytes, double *reward, double *detail, bool full,
                      cudaStream_t stream, cudaEvent_t *marks /* 5 events or NULL */, uint32_t *depths = nullptr) {
    ORIE_TRY(check_range(ix, t0, nt, "orie_reward"));
    if (!ens_bits || (!reward && !detail) || !workspace || N < 0) {
        set_error("orie_reward: null buffer or negative N");
        return ORIE_EINVAL;
    }
    if (nt == 0) return ORIE_OK;
    if (ix->resolved && ix->resolved_rc != ORIE_OK) return resolve(ix);
    const WsLayout L = ws_layout(ix, nt);
    const int64_t ntp = round_up(nt, 32);
    // capacity of one target's event list: what the workspace has left after the fixed tables
    const int64_t ev_stride = workspace_bytes > L.fixed ? (int64_t)((workspace_bytes - L.fixed) / ((size_t)ntp * 4)) : 0;
    if (ev_stride < 1 || (ix->resolved && ev_stride < ix->Ev)) {
        set_error("orie_reward: workspace has %zu bytes, %zu needed for %lld targets", workspace_bytes,
                  ws_total(ix, nt, ix->resolved ? ix->Ev : 1), (long long)nt);
        return ORIE_EWORKSPACE;
    }
    if ((uintptr_t)workspace & 255) {
        set_error("orie_reward: workspace must be 256-byte aligned");
        return ORIE_EINVAL;
    }
    ORIE_TRY(walk_attributes());
    char *ws = (char *)workspace;
    const size_t smem = (size_t)ix->ens_words * 32 * 4;
    const bool gmem = walk_in_gmem(ix);
    static_assert(sizeof(WalkParams) < 4000 && sizeof(ApParams) + sizeof(Grid101) < 4000, "kernel parameter space");

    WalkParams wp;
    memset(&wp, 0, sizeof(wp));
    wp.M = ix->M; wp.nt = nt; wp.ntp = ntp; wp.t0 = t0;
    wp.meta = ix->meta; wp.S_cap = ix->S_cap;
    wp.ens_words = ix->ens_words; wp.ens_bits = ens_bits;
    const int64_t nb = ntp / 32;
    const int walk_threads = gmem ? 256 : smem <= 56 * 1024 ? 256 : 512;   // keep the SM full of warps
    // the detection walk takes two batches per warp when two tables fit shared memory four blocks to the SM
    const bool pairs = !ix->walk_single && (gmem || 2 * smem <= 56 * 1024);
    auto segs_per_block = [&](int64_t S, bool two) {
        // enough blocks for two full waves of resident blocks when the data allows, at least one segment per warp
        const int64_t resident = two ? 148 * kWalk2Blocks : 148 * (kWalkResident / walk_threads > 0 ? kWalkResident / walk_threads : 1);
        const int64_t nb = two ? ceil_div(ntp / 32, 2) : ntp / 32;
        const double waves = ix->walk_waves > 0.0 ? ix->walk_waves : 2.0;
        int64_t want_y = (int64_t)((waves * (double)resident + (double)nb - 1.0) / (double)nb);
        if (gmem) want_y = std::max<int64_t>(want_y, 32);     // many blocks per batch: few tables in flight
        int64_t spb = ceil_div(S, want_y > 0 ? want_y : 1);
        // the single-batch kernel splits a block's segments evenly over its warps; the pair kernel hands them out one by one
        spb = two ? std::max<int64_t>(spb, kWalk2Threads / 32) : round_up(spb > 0 ? spb : 1, walk_threads / 32);
        return (int)spb;
    };
    if (gmem && nb > 65535) {
        set_error("orie_reward: at most %d targets per call for datasets of this size (%lld requested); run them in waves",
                  65535 * 32, (long long)nt);
        return ORIE_ELIMIT;
    }
    // launch grids come from the exact segment counts when the host already knows them, else from their upper bounds
    const int64_t S_grid = ix->resolved ? ix->S : ix->S_cap, SL_grid = ix->resolved ? ix->SL : ix->SL_cap;
    if (marks) ORIE_CUDA(cudaEventRecord(marks[0], stream));
    const int64_t nb2 = ceil_div(nb, 2);
    ORIE_CUDA(cudaMemsetAsync(ws + L.counter, 0, 8 + (size_t)nb2 * 4, stream));      // work counters of the walk and the AP kernel
    if (gmem) {
        wp.memb_global = (uint32_t *)(ws + L.memb);
        wp.memb_pairs = pairs ? 1 : 0;
        dim3 grid((unsigned)(pairs ? nb2 : nb), (unsigned)std::min<int64_t>(ceil_div(ix->ens_words, 8), 64));
        if (pairs) membership_table2_kernel<<<grid, 256, 0, stream>>>(wp);
        else membership_table_kernel<<<grid, 256, 0, stream>>>(wp);
        ORIE_LAUNCH_CHECK();
    }
    // labels: a small grid — on the auxiliary stream, if the caller gave one, it runs next to the detection walk
    // (and fills its last, partly empty wave of blocks); not when per-kernel durations are being measured, and not
    // with the membership tables in global memory: the detection walk's block order keeps the tables in flight
    // L2-resident, and a second kernel walking them in another order evicts them (50k sweep: +17 ms)
    const bool side = ix->aux != nullptr && ix->aux != stream && !marks && !gmem;
    if (SL_grid > 0) {
        WalkParams lp = wp;
        lp.slot_img = ix->lab_slot_img; lp.seg_chunk0 = ix->lseg_chunk0; lp.seg_nch = ix->lseg_nch;
        lp.segs_per_block = segs_per_block(SL_grid, false);
        lp.tot = (uint32_t *)(ws + L.totL);
        const unsigned ny = (unsigned)ceil_div(SL_grid, lp.segs_per_block);
        dim3 grid = gmem ? dim3(ny, (unsigned)nb) : dim3((unsigned)nb, ny);
        cudaStream_t ls = stream;
        if (side) {
            ORIE_CUDA(cudaEventRecord(ix->ev_fork, stream));
            ORIE_CUDA(cudaStreamWaitEvent(ix->aux, ix->ev_fork, 0));
            ls = ix->aux;
        }
        ORIE_TRY(launch_walk<false>(grid, walk_threads, smem, gmem, ls, lp));
        ORIE_LAUNCH_CHECK();
        if (side) ORIE_CUDA(cudaEventRecord(ix->ev_join, ix->aux));
    }
    if (marks) ORIE_CUDA(cudaEventRecord(marks[1], stream));
    // detections
    {
        wp.slot_img = ix->slot_img; wp.seg_chunk0 = ix->seg_chunk0; wp.seg_nch = ix->seg_nch;
        wp.segs_per_block = segs_per_block(S_grid, pairs);
        wp.tot = (uint32_t *)(ws + L.tot);
        wp.slot_tp = ix->slot_tp; wp.slot_pk = ix->slot_pk; wp.seg_ev0 = ix->seg_ev0;
        wp.ev_img = ix->ev_img; wp.ev_mask = ix->ev_mask; wp.T = ix->T;
        wp.kseg = (uint16_t *)(ws + L.kseg);
        wp.bq = ix->bq; wp.bqoff = ix->bqoff;
        wp.Ev = ev_stride;
        wp.evcnt = (uint32_t *)(ws + L.evcnt);
        wp.ev = (uint32_t *)(ws + L.ev);
        wp.cb_w = (uint32_t *)(ws + L.cb_w);
        wp.cb_s = (uint32_t *)(ws + L.cb_s);
        wp.seg_order = ix->seg_order;
        wp.pair_next = (uint32_t *)(ws + L.counter + 8);
        unsigned ny = (unsigned)ceil_div(S_grid, wp.segs_per_block);
        if (pairs) {
            // blocks per pair: what is resident at once shared out over the pairs, at least one segment per warp
            const int64_t resident = (int64_t)std::max(ix->sms, 1) * kWalk2Blocks;
            // (tables in global memory: at least 32 blocks per pair, launched pair by pair, so that only the tables of
            // the few pairs in flight compete for L2 — with every pair resident at once the 50k sweep's walk took
            // 74 ms instead of 59)
            ny = (unsigned)std::max<int64_t>(1, std::min<int64_t>(std::max<int64_t>(resident / nb2, gmem ? 32 : 1),
                                                                  ceil_div(S_grid, kWalk2Threads / 32)));
            dim3 grid = gmem ? dim3(ny, (unsigned)nb2) : dim3((unsigned)nb2, ny);
            if (gmem && wp.slot_pk) walk2_kernel<true, true><<<grid, kWalk2Threads, 0, stream>>>(wp);
            else if (gmem) walk2_kernel<true, false><<<grid, kWalk2Threads, 0, stream>>>(wp);
            else if (wp.slot_pk) walk2_kernel<false, true><<<grid, kWalk2Threads, 2 * smem, stream>>>(wp);
            else walk2_kernel<false, false><<<grid, kWalk2Threads, 2 * smem, stream>>>(wp);
        } else {
            dim3 grid = gmem ? dim3(ny, (unsigned)nb) : dim3((unsigned)nb, ny);
            ORIE_TRY(launch_walk<true>(grid, walk_threads, smem, gmem, stream, wp));
        }
        ORIE_LAUNCH_CHECK();
    }
    if (side && SL_grid > 0) ORIE_CUDA(cudaStreamWaitEvent(stream, ix->ev_join, 0));
    if (marks) ORIE_CUDA(cudaEventRecord(marks[2], stream));
    // AP
    ApParams ap;
    memset(&ap, 0, sizeof(ap));
    ap.M = ix->M; ap.C = ix->C; ap.nt = nt; ap.ntp = ntp; ap.t0 = t0;
    ap.T = ix->T; ap.cls_per_warp = ix->cls_per_warp; ap.class_groups = ix->class_groups;
    ap.Ev = ev_stride; ap.meta = ix->meta;
    ap.cls_order = ix->cls_order; ap.cls_seg0 = ix->cls_seg0; ap.seg_chunk0 = ix->seg_chunk0; ap.lcls_seg0 = ix->lcls_seg0; ap.seg_ev0 = ix->seg_ev0;
    ap.act_cls = ix->act_cls; ap.nact = ix->nact;
    ap.tot = (const uint32_t *)(ws + L.tot); ap.evcnt = (const uint32_t *)(ws + L.evcnt);
    ap.totL = (const uint32_t *)(ws + L.totL); ap.ev = (const uint32_t *)(ws + L.ev);
    ap.kseg = (const uint16_t *)(ws + L.kseg);
    ap.gtcnt = ix->gtcnt; ap.w_off = ix->w_off; ap.s_off = ix->s_off;
    ap.own_w_cs = ix->own_w_cs; ap.own_s_cs = ix->own_s_cs; ap.own_w_m = ix->own_w_m; ap.own_s_m = ix->own_s_m;
    ap.own_w_q = ix->own_w_q; ap.own_s_q = ix->own_s_q;
    ap.cb_w = (const uint32_t *)(ws + L.cb_w); ap.cb_s = (const uint32_t *)(ws + L.cb_s);
    ap.partial = (double *)(ws + L.partial);
    ap.next_item = (unsigned long long *)(ws + L.counter);
    ap.depths = depths;
    const int64_t items = nt * ix->class_groups;
    // persistent warps: as many blocks as are resident at once
    const unsigned ap_grid = (unsigned)std::min<int64_t>(ceil_div(items, kApThreads / 32), (int64_t)std::max(ix->sms, 1) * ORIE_AP_BLOCKS);
    if (depths) ap_kernel<false, 0, true><<<ap_grid, kApThreads, 0, stream>>>(ap, grid101());
    else if (full) ap_kernel<true, 0><<<ap_grid, kApThreads, 0, stream>>>(ap, grid101());
    else if (ix->ap_mode == 1) ap_kernel<false, 0><<<ap_grid, kApThreads, 0, stream>>>(ap, grid101());   // fixed cls_order groups
    else ap_kernel<false, 2><<<ap_grid, kApThreads, 0, stream>>>(ap, grid101());      // default: depth-ordered active classes
    ORIE_LAUNCH_CHECK();
    if (marks) ORIE_CUDA(cudaEventRecord(marks[3], stream));
    finalize_kernel<<<(unsigned)ceil_div(nt, kFinalizeThreads / 32), kFinalizeThreads, 0, stream>>>(ap.partial, nt, ix->class_groups, ix->T, N, ix->meta,
                                                                                                   reward, detail);
    ORIE_LAUNCH_CHECK();
    if (marks) ORIE_CUDA(cudaEventRecord(marks[4], stream));
    return ORIE_OK;
}

extern "C" int orie_reward(const orie_index_t *ix, int64_t t0, int64_t nt, const uint32_t *ens_bits, int64_t N,
                           void *workspace, size_t workspace_bytes, double *reward, double *detail, orie_stream_t stream) {
    return run_reward(ix, t0, nt, ens_bits, N, workspace, workspace_bytes, reward, detail, detail != nullptr, stream, nullptr);
}

extern "C" int orie_reward_sums(const orie_index_t *ix, int64_t t0, int64_t nt, const uint32_t *ens_bits, int64_t N,
                                void *workspace, size_t workspace_bytes, double *sums, int full, orie_stream_t stream) {
    if (!sums) {
        set_error("orie_reward_sums: sums is NULL");
        return ORIE_EINVAL;
    }
    return run_reward(ix, t0, nt, ens_bits, N, workspace, workspace_bytes, nullptr, sums, full != 0, stream, nullptr);
}

extern "C" int orie_reward_depths(const orie_index_t *ix, int64_t t0, int64_t nt, const uint32_t *ens_bits, int64_t N,
                                  void *workspace, size_t workspace_bytes, double *reward, uint32_t *depths, orie_stream_t stream) {
    if (!depths || !reward) {
        set_error("orie_reward_depths: null buffer");
        return ORIE_EINVAL;
    }
    return run_reward(ix, t0, nt, ens_bits, N, workspace, workspace_bytes, reward, nullptr, false, stream, nullptr, depths);
}

extern "C" int orie_rewards_from_sums(const double *sums, int64_t nt, int T, int64_t N, double *reward, orie_stream_t stream) {
    if (!sums || !reward || nt < 0 || T < 1 || N < 0) {
        set_error("orie_rewards_from_sums: null buffer or bad size");
        return ORIE_EINVAL;
    }
    if (nt == 0) return ORIE_OK;
    rewards_from_sums_kernel<<<(unsigned)ceil_div(nt, 256), 256, 0, stream>>>(sums, nt, T, N, reward);
    ORIE_LAUNCH_CHECK();
    return ORIE_OK;
}

extern "C" int orie_reward_profile(const orie_index_t *ix, int64_t t0, int64_t nt, const uint32_t *ens_bits, int64_t N,
                                   void *workspace, size_t workspace_bytes, double *reward, double *detail, int full,
                                   orie_stream_t stream, float *kernel_ms_host) {
    if (!kernel_ms_host) {
        set_error("orie_reward_profile: kernel_ms_host is NULL");
        return ORIE_EINVAL;
    }
    cudaEvent_t marks[5];
    for (int i = 0; i < 5; ++i) ORIE_CUDA(cudaEventCreate(&marks[i]));
    int rc = run_reward(ix, t0, nt, ens_bits, N, workspace, workspace_bytes, reward, detail, full != 0, stream, marks);
    if (rc == ORIE_OK && nt > 0) {
        cudaError_t e = cudaEventSynchronize(marks[4]);
        if (e != cudaSuccess) {
            set_error("orie_reward_profile: %s", cudaGetErrorString(e));
            rc = ORIE_ECUDA;
        }
        for (int i = 0; i < 4 && rc == ORIE_OK; ++i) cudaEventElapsedTime(&kernel_ms_host[i], marks[i], marks[i + 1]);
    }
    for (int i = 0; i < 5; ++i) cudaEventDestroy(marks[i]);
    return rc;
}
