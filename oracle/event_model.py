"""CPU model of the ENGINE's formulation of the ORIE reward (not of upstream's).

TEST INFRASTRUCTURE ONLY (same rules as ``orie_oracle.py``).  The CUDA engine
does not re-sort every ensemble; it sorts the whole dataset once by
(class, confidence desc) into padded *slots*, turns an ensemble into a
membership test per slot, keeps only the ensemble's true-positive *events*
with their rank among ensemble members, and integrates the 101-point AP by a
single reverse sweep over those events.  This module restates exactly that
data flow in numpy / plain Python so the formulation itself can be checked
against the straightforward oracle on the CPU before any GPU time is spent:

  build_index   <-> csrc/index.cu      (slots, segments, events, own lists)
  walk_target   <-> csrc/reward.cu K1  (segment totals, event ranks, member true positives per segment and
                                        threshold from the dense event stream, own-detection ranks)
  ap_reverse    <-> csrc/reward.cu K2  (reverse sweep == lib/metrics.py:118-123,137-144)
  reward_target <-> K2 + K3            (reward.py:43-50)

It is slow (Python loops) and meant for small cases.
"""
from __future__ import annotations

import numpy as np

GRID = np.linspace(0, 1, 101)
GRID_D = np.diff(GRID)
CHUNK = 32


def _roundup(x, m):
    return (x + m - 1) // m * m


class Index:
    pass


def build_index(M, C, off_w, cls_w, conf_w, tpm_w, off_s, cls_s, conf_s, tpm_s, lab_off, lab_cls, seg_chunks=4):
    ix = Index()
    ix.M, ix.C, ix.seg_chunks = M, C, seg_chunks
    Dw, Ds = len(cls_w), len(cls_s)
    img_w = np.repeat(np.arange(M), np.diff(off_w))
    img_s = np.repeat(np.arange(M), np.diff(off_s))
    # --- global weak order: class asc, conf desc, stable by original index
    order_w = np.lexsort((np.arange(Dw), -conf_w, cls_w))
    cnt_w = np.bincount(cls_w, minlength=C)
    cls_off = np.concatenate([[0], np.cumsum(cnt_w)])
    padlen = _roundup(cnt_w + 1, CHUNK)
    pad_off = np.concatenate([[0], np.cumsum(padlen)])
    P = int(pad_off[-1])
    ix.P, ix.pad_off, ix.cnt_w = P, pad_off, cnt_w
    slot_img = np.full(P, M, dtype=np.int64)
    slot_tpm = np.zeros(P, dtype=np.int64)
    prank = np.empty(Dw, dtype=np.int64)
    r = np.arange(Dw)
    c_sorted = cls_w[order_w]
    slot = pad_off[c_sorted] + (r - cls_off[c_sorted])
    slot_img[slot] = img_w[order_w]
    slot_tpm[slot] = tpm_w[order_w]
    prank[order_w] = slot
    ix.slot_img, ix.slot_tpm = slot_img, slot_tpm
    # --- segments (never cross a class boundary)
    seg_cls, seg_chunk0, seg_nch = [], [], []
    cls_seg0 = np.zeros(C + 1, dtype=np.int64)
    for c in range(C):
        cls_seg0[c] = len(seg_cls)
        nch = padlen[c] // CHUNK
        c0 = pad_off[c] // CHUNK
        for k in range(0, nch, seg_chunks):
            seg_cls.append(c)
            seg_chunk0.append(c0 + k)
            seg_nch.append(min(seg_chunks, nch - k))
    cls_seg0[C] = len(seg_cls)
    ix.seg_cls, ix.seg_chunk0, ix.seg_nch, ix.cls_seg0 = map(np.asarray, (seg_cls, seg_chunk0, seg_nch, cls_seg0))
    ix.S = len(seg_cls)
    seg_of_chunk = np.empty(P // CHUNK, dtype=np.int64)
    for s in range(ix.S):
        seg_of_chunk[ix.seg_chunk0[s]:ix.seg_chunk0[s] + ix.seg_nch[s]] = s
    ix.seg_of_chunk = seg_of_chunk
    # --- strong insertion slot: after every weak detection of the class with conf >= its conf
    q_s = np.empty(Ds, dtype=np.int64)
    for c in range(C):
        sel = np.nonzero(cls_s == c)[0]
        if len(sel) == 0:
            continue
        seg_conf = conf_w[order_w[cls_off[c]:cls_off[c + 1]]]          # descending
        k = np.searchsorted(-seg_conf, -conf_s[sel], side="right")
        q_s[sel] = pad_off[c] + k
    # --- own lists: image-major, (class asc, conf desc, row)
    own_w = np.lexsort((np.arange(Dw), -conf_w, cls_w, img_w))
    own_s = np.lexsort((np.arange(Ds), -conf_s, cls_s, img_s))
    ix.off_w, ix.off_s = off_w, off_s
    ix.own_w_q, ix.own_w_m, ix.own_w_c = prank[own_w], tpm_w[own_w], cls_w[own_w]
    ix.own_s_q, ix.own_s_m, ix.own_s_c = q_s[own_s], tpm_s[own_s], cls_s[own_s]
    # --- labels: per image per class counts, and class-sorted stream (here: dense counts do both jobs)
    gt = np.zeros((M, C), dtype=np.int64)
    img_l = np.repeat(np.arange(M), np.diff(lab_off))
    np.add.at(gt, (img_l, lab_cls), 1)
    ix.gtcnt = gt
    # --- dense event stream (index.cu post_kernel phase C): the slots holding a true positive, in slot order, with
    #     the number of events in front of every segment (seg_ev0[S] = all events)
    ev_slot = np.nonzero(ix.slot_tpm != 0)[0]
    ix.ev_img, ix.ev_mask = ix.slot_img[ev_slot], ix.slot_tpm[ev_slot]
    seg_first = ix.seg_chunk0 * CHUNK
    ix.seg_ev0 = np.concatenate([np.searchsorted(ev_slot, seg_first, side="left"), [len(ev_slot)]]).astype(np.int64)
    return ix


def walk_target(ix: Index, member: np.ndarray, j: int):
    """K1 for one target: member[M+1] bool (member[M] = False sentinel).

    Returns tot[S], events (list per seg of (p_rel, mask)), cb_w, cb_s (rank of
    each own detection among ensemble members of its segment, exclusive).  ``ix.kseg`` (set as a side effect, like the
    kernel's workspace array) holds the member true positives of every segment per threshold bit, counted the way
    walk_kernel / walk2_kernel count them: from the dense event stream, not from the records."""
    mem = member[ix.slot_img]                      # per slot
    tot = np.zeros(ix.S, dtype=np.int64)
    events = []
    cum = np.zeros(ix.P + 1, dtype=np.int64)       # exclusive prefix within segment
    for s in range(ix.S):
        a = ix.seg_chunk0[s] * CHUNK
        b = a + ix.seg_nch[s] * CHUNK
        m = mem[a:b].astype(np.int64)
        inc = np.cumsum(m)
        cum[a:b] = inc - m
        tot[s] = inc[-1]
        hit = np.nonzero((m == 1) & (ix.slot_tpm[a:b] != 0))[0]
        events.append([(int(inc[h]), int(ix.slot_tpm[a + h])) for h in hit])
    memb_ev = member[ix.ev_img]
    kseg = np.zeros((ix.S, 16), dtype=np.int64)
    for s in range(ix.S):
        e0, e1 = ix.seg_ev0[s], ix.seg_ev0[s + 1]
        masks = ix.ev_mask[e0:e1][memb_ev[e0:e1]].astype(np.int64)
        for t in range(16):
            kseg[s, t] = int(((masks >> t) & 1).sum())
        assert e1 - e0 >= len(events[s]) and int(memb_ev[e0:e1].sum()) == len(events[s])
    ix.kseg = kseg
    wq = ix.own_w_q[ix.off_w[j]:ix.off_w[j + 1]]
    sq = ix.own_s_q[ix.off_s[j]:ix.off_s[j + 1]]
    return tot, events, cum[wq], cum[sq]


GRID_C = np.array([g / 100.0 for g in range(101)])       # correctly rounded g/100
GRID_GE = GRID >= GRID_C                                  # np.linspace's g*0.01 against it (exact-tie rule)
# trapezoid weights per grid point: sum_g d_g (y_g + y_{g+1}) / 2 == sum_g W_g y_g
GRID_W = np.zeros(101)
GRID_W[:100] += GRID_D / 2
GRID_W[1:] += GRID_D / 2
GRID_CW = np.concatenate([[0.0], np.cumsum(GRID_W)])             # CW[i]  = sum_{g<i} W_g
GRID_CWX = np.concatenate([[0.0], np.cumsum(GRID_W * GRID)])     # CWX[i] = sum_{g<i} W_g x_g


def grid_lo(a, n_l):
    """Smallest grid index g with x_g >= (a/100)/n_l in numpy's float comparison, decided in
    integers: g*n_l > a, or == a and GRID_GE[g]."""
    q, r = divmod(a, n_l)
    g = q + (1 if r > 0 else 0)
    if r == 0 and g <= 100 and not GRID_GE[g]:
        g += 1
    return g


class _Var:
    """Reverse-sweep state of one (class, threshold, variant) AP integral.

    Three facts make the sweep cheap (csrc/reward.cu uses the same):
    * below the last true positive the interpolated curve is the plain precision
      envelope: with j = p_{k+1}-1, precision k/j never exceeds (k+1)/p_{k+1}
      (because p_{k+1} >= k+1), so np.interp's two knots carry the same value and
      y(x) = E_{k+1} = max_{k'>k} k'/p_{k'};
    * "x_g >= fl(k/n_l)" is decided in integers: 100*k vs g*n_l, and on exact
      rational ties by the constant table GRID_GE.  Different rationals differ by
      >= 1/(100 n_l), far above float64 rounding, so this equals numpy's comparison;
    * a run of grid points that share one envelope value E contributes
      E * (CW[hi+1] - CW[lo]) to np.trapz; only the tail beyond the last true
      positive is a genuine linear ramp (closed form with CW and CWX).
    Summation order differs from np.trapz by O(1e-16) relative."""

    def __init__(self, K, n_p, n_l):
        self.n_l = n_l
        self.ap = 0.0
        self.g = 99                  # highest grid point not yet integrated (y at 100 is always 0)
        self.k = K
        self.E = 0.0                 # running envelope max
        self.dead = (K == 0 or n_p == 0)
        if self.dead:
            return
        self.q, self.r = divmod(K * 100, n_l)     # 100 k == q n_l + r, updated incrementally below
        gl = self._first_grid()
        assert gl == grid_lo(K * 100, n_l)
        if gl <= self.g:
            r_k = K / n_l
            env = K / n_p
            slope = (0.0 - env) / (1.0 - r_k)
            sw = GRID_CW[self.g + 1] - GRID_CW[gl]
            swx = GRID_CWX[self.g + 1] - GRID_CWX[gl]
            self.ap = slope * (swx - r_k * sw) + env * sw
            self.g = gl - 1

    def _first_grid(self):
        gl = self.q + (1 if self.r > 0 else 0)
        if self.r == 0 and gl <= 100 and not GRID_GE[gl]:
            gl += 1
        return gl

    def step(self, pos):
        """The k-th true positive (k = self.k) sits at 1-based rank ``pos``."""
        if self.dead:
            return
        self.E = max(self.E, self.k / pos)
        self.k -= 1
        self.r -= 100
        while self.r < 0:
            self.r += self.n_l
            self.q -= 1
        gl = self._first_grid()
        assert gl == grid_lo(self.k * 100, self.n_l)
        if gl <= self.g:
            self.ap += self.E * (GRID_CW[self.g + 1] - GRID_CW[gl])
            self.g = gl - 1


def ap_reverse(ix, c, t, tot, events, own, n_l):
    """AP of class c at threshold bit t for one variant.

    ``own`` = (q[], mask[], cb_rel[]) of the target's own detections of class c
    in this variant, ascending by q."""
    q, mk, cb = own
    s0, s1 = ix.cls_seg0[c], ix.cls_seg0[c + 1]
    n_ens = int(tot[s0:s1].sum())
    K_ens = int(ix.kseg[s0:s1, t].sum())                 # what K1 hands over; K2 no longer reads every record for it
    assert K_ens == sum(1 for s in range(s0, s1) for (_, m) in events[s] if (m >> t) & 1)
    K = K_ens + sum(1 for m in mk if (m >> t) & 1)
    n_p = n_ens + len(q)
    v = _Var(K, n_p, n_l)
    if v.dead:
        return 0.0
    ib = len(q) - 1
    rem = n_ens
    for s in range(s1 - 1, s0 - 1, -1):
        rem -= int(tot[s])
        base = rem
        slot0 = ix.seg_chunk0[s] * CHUNK
        for (p_rel, m) in reversed(events[s]):
            p = base + p_rel
            while ib >= 0 and q[ib] >= slot0 and base + cb[ib] >= p:
                if (mk[ib] >> t) & 1:
                    v.step(base + int(cb[ib]) + 1 + ib)
                ib -= 1
            if (m >> t) & 1:
                v.step(p + ib + 1)
        while ib >= 0 and q[ib] >= slot0:
            if (mk[ib] >> t) & 1:
                v.step(base + int(cb[ib]) + 1 + ib)
            ib -= 1
    assert ib == -1 and v.k == 0 and v.g == -1, (ib, v.k, v.g)
    return v.ap


def delta_reverse(ix, c, t, tot, events, own_w, own_s, n_l):
    """AP_strong - AP_weak of class c at threshold bit t, the way ap_kernel<false> computes it: both variants are
    swept together and the sweep stops once they can no longer differ (all own detections behind and either no true
    positive left in front, or equal envelopes at an equal remaining-TP count and grid pointer) — every later term is
    identical in both and cancels."""
    if len(own_w[0]) == 0 and len(own_s[0]) == 0:
        return 0.0, 0
    s0, s1 = ix.cls_seg0[c], ix.cls_seg0[c + 1]
    n_ens = int(tot[s0:s1].sum())
    K_ens = int(ix.kseg[s0:s1, t].sum())
    assert K_ens == sum(1 for s in range(s0, s1) for (_, m) in events[s] if (m >> t) & 1)
    var, cur = [], []
    for q, mk, cb in (own_w, own_s):
        var.append(_Var(K_ens + sum(1 for m in mk if (m >> t) & 1), n_ens + len(q), n_l))
        cur.append(len(q) - 1)
    owns = (own_w, own_s)
    steps = 0

    def drain(v, base, slot0, limit):
        q, mk, cb = owns[v]
        while cur[v] >= 0 and q[cur[v]] >= slot0 and base + cb[cur[v]] >= limit:
            if (mk[cur[v]] >> t) & 1:
                var[v].step(base + int(cb[cur[v]]) + 1 + cur[v])
            cur[v] -= 1

    def converged():
        """ap_kernel's two exits once every own detection lies behind: (1) no true positive is left in front of either
        variant, nothing can change any more; (2) both alive — they then share the remaining-TP count and the grid
        pointer (asserted here: it is what lets the kernel's tail loop keep ONE copy of that state and compute one
        ratio per event for both variants) — and the envelopes have met."""
        a, b = var
        if cur[0] >= 0 or cur[1] >= 0:
            return False
        ka, kb = (0 if a.dead else a.k), (0 if b.dead else b.k)
        if ka == 0 and kb == 0:
            return True
        if not a.dead and not b.dead:
            assert a.k == b.k and a.g == b.g, (a.k, b.k, a.g, b.g)
            return a.E == b.E
        return False

    if not (var[0].dead and var[1].dead):
        rem, done = n_ens, False
        for s in range(s1 - 1, s0 - 1, -1):
            rem -= int(tot[s])
            base, slot0 = rem, ix.seg_chunk0[s] * CHUNK
            for (p_rel, m) in reversed(events[s]):
                if converged():
                    done = True
                    break
                p = base + p_rel
                drain(0, base, slot0, p)
                drain(1, base, slot0, p)
                if (m >> t) & 1:
                    var[0].step(p + cur[0] + 1)
                    var[1].step(p + cur[1] + 1)
                    steps += 1
            if done:
                break
            drain(0, base, slot0, 0)
            drain(1, base, slot0, 0)
    aw = 0.0 if var[0].dead else var[0].ap
    as_ = 0.0 if var[1].dead else var[1].ap
    return as_ - aw, steps


def reward_target_delta(ix: Index, j: int, ens_idx, T: int):
    """reward_target via delta_reverse (difference-only sweep); also returns the number of TP steps taken."""
    M, C = ix.M, ix.C
    member = np.zeros(M + 1, dtype=bool)
    member[np.asarray(ens_idx, dtype=np.int64)] = True
    tot, events, cb_w, cb_s = walk_target(ix, member, j)
    n_l_all = ix.gtcnt[member[:M]].sum(axis=0) + ix.gtcnt[j]
    sl = lambda a, off: a[off[j]:off[j + 1]]
    wc, sc = sl(ix.own_w_c, ix.off_w), sl(ix.own_s_c, ix.off_s)
    wq, sq = sl(ix.own_w_q, ix.off_w), sl(ix.own_s_q, ix.off_s)
    wm, sm = sl(ix.own_w_m, ix.off_w), sl(ix.own_s_m, ix.off_s)
    delta, nc, steps = 0.0, 0, 0
    for c in range(C):
        n_l = int(n_l_all[c])
        if n_l == 0:
            continue
        nc += 1
        a, b = wc == c, sc == c
        for t in range(T):
            d, st = delta_reverse(ix, c, t, tot, events, (wq[a], wm[a], cb_w[a]), (sq[b], sm[b], cb_s[b]), n_l)
            delta += d
            steps += st
    if nc == 0:
        return 0.0, 0
    return delta / (nc * T) * (len(ens_idx) + 1), steps


def reward_target(ix: Index, j: int, ens_idx, T: int):
    M, C = ix.M, ix.C
    member = np.zeros(M + 1, dtype=bool)
    member[np.asarray(ens_idx, dtype=np.int64)] = True
    assert not member[j]
    tot, events, cb_w, cb_s = walk_target(ix, member, j)
    n_l_all = ix.gtcnt[member[:M]].sum(axis=0) + ix.gtcnt[j]
    sw = ss = 0.0
    nc = 0
    wc = ix.own_w_c[ix.off_w[j]:ix.off_w[j + 1]]
    sc = ix.own_s_c[ix.off_s[j]:ix.off_s[j + 1]]
    wq = ix.own_w_q[ix.off_w[j]:ix.off_w[j + 1]]
    sq = ix.own_s_q[ix.off_s[j]:ix.off_s[j + 1]]
    wm = ix.own_w_m[ix.off_w[j]:ix.off_w[j + 1]]
    sm = ix.own_s_m[ix.off_s[j]:ix.off_s[j + 1]]
    for c in range(C):
        n_l = int(n_l_all[c])
        if n_l == 0:
            continue
        nc += 1
        a = wc == c
        b = sc == c
        for t in range(T):
            sw += ap_reverse(ix, c, t, tot, events, (wq[a], wm[a], cb_w[a]), n_l)
            ss += ap_reverse(ix, c, t, tot, events, (sq[b], sm[b], cb_s[b]), n_l)
    if nc == 0:
        return 0.0
    return (ss / (nc * T) - sw / (nc * T)) * (len(ens_idx) + 1)
