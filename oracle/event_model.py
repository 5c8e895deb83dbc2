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
  walk_target   <-> csrc/reward.cu K1  (segment totals, event ranks, own-detection ranks)
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
    return ix


def walk_target(ix: Index, member: np.ndarray, j: int):
    """K1 for one target: member[M+1] bool (member[M] = False sentinel).

    Returns tot[S], events (list per seg of (p_rel, mask)), cb_w, cb_s (rank of
    each own detection among ensemble members of its segment, exclusive)."""
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
    wq = ix.own_w_q[ix.off_w[j]:ix.off_w[j + 1]]
    sq = ix.own_s_q[ix.off_s[j]:ix.off_s[j + 1]]
    return tot, events, cum[wq], cum[sq]


class _Var:
    """Reverse-sweep state of one (class, threshold, variant) AP integral."""

    def __init__(self, K, n_p, n_l):
        self.n_l = n_l
        self.ap = 0.0
        self.g = 99
        self.y_next = 0.0            # y at grid point 100 is always 0
        self.k = K
        self.E = -1.0
        self.dead = (K == 0 or n_p == 0)
        if self.dead:
            return
        r_k = K / n_l
        self._consume(r_k, K / n_p, 1.0, 0.0)
        self.r_cur = r_k

    def _consume(self, r_lo, env_lo, r_hi, env_hi):
        if self.g >= 0 and GRID[self.g] >= r_lo:
            slope = (env_hi - env_lo) / (r_hi - r_lo)
            while self.g >= 0 and GRID[self.g] >= r_lo:
                x = GRID[self.g]
                y = env_lo if x == r_lo else slope * (x - r_lo) + env_lo
                self.ap += GRID_D[self.g] * (self.y_next + y) / 2.0
                self.y_next = y
                self.g -= 1

    def step(self, pos):
        """The k-th true positive (k = self.k) sits at 1-based rank ``pos``."""
        if self.dead:
            return
        k = self.k
        self.E = max(self.E, k / pos)
        r_lo = (k - 1) / self.n_l
        j = pos - 1
        prec_j = 1.0 if j == 0 else (k - 1) / j
        self._consume(r_lo, max(prec_j, self.E), self.r_cur, self.E)
        self.r_cur = r_lo
        self.k = k - 1


def ap_reverse(ix, c, t, tot, events, own, n_l):
    """AP of class c at threshold bit t for one variant.

    ``own`` = (q[], mask[], cb_rel[]) of the target's own detections of class c
    in this variant, ascending by q."""
    q, mk, cb = own
    s0, s1 = ix.cls_seg0[c], ix.cls_seg0[c + 1]
    n_ens = int(tot[s0:s1].sum())
    K = sum(1 for s in range(s0, s1) for (_, m) in events[s] if (m >> t) & 1) + sum(1 for m in mk if (m >> t) & 1)
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
