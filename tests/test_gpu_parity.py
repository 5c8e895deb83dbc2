"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle."""
import numpy as np
import pytest

from helpers import O, flat_tp, make_packed, oracle_cache

pytestmark = pytest.mark.gpu


def _engine(pk, iouv, **kw):
    from orie_b200.engine import Engine
    return Engine(pk, iouv=iouv, **kw)


@pytest.mark.parametrize("T", [1, 10])
def test_matching_bit_exact(T):
    iouv = O.IOU_05 if T == 1 else O.IOU_05_095
    _, pk = make_packed(M=300, seed=11, empty_det_frac=0.03)
    eng = _engine(pk, iouv)
    wtp, stp, wm, sm = eng.tp_flags()
    wd, sd, _ = oracle_cache(pk, iouv)
    assert np.array_equal(wtp, flat_tp(wd, len(pk.w_cls), T))
    assert np.array_equal(stp, flat_tp(sd, len(pk.s_cls), T))
    # match indices: the label a TP is matched to
    for off, box, cls, got, tp in ((pk.w_off, pk.w_box, pk.w_cls, wm, wtp), (pk.s_off, pk.s_box, pk.s_cls, sm, stp)):
        for i in range(0, pk.num_images, 7):
            a, b = off[i], off[i + 1]
            la, lb = pk.l_off[i], pk.l_off[i + 1]
            _, best, _ = O.match_detections(box[a:b], cls[a:b], pk.l_box[la:lb], pk.l_cls[la:lb], iouv)
            want = np.where(tp[a:b].any(axis=1), best, -1)
            assert np.array_equal(got[a:b], want)
    eng.close()


def test_dcsb_exact():
    _, pk = make_packed(M=300, seed=12, empty_det_frac=0.05)
    eng = _engine(pk, O.IOU_05)
    wd, sd, _ = oracle_cache(pk, O.IOU_05)
    assert np.array_equal(eng.dcsb(), O.dcsb_all(wd, sd))
    eng.close()


@pytest.mark.parametrize("M,N,T,segch", [(200, 50, 10, 0), (200, 50, 1, 2), (96, 0, 10, 1), (70, 69, 10, 0), (257, 31, 1, 3)])
def test_orie_explicit_ensembles(M, N, T, segch):
    iouv = O.IOU_05 if T == 1 else O.IOU_05_095
    _, pk = make_packed(M=M, seed=100 + M, empty_det_frac=0.04)
    eng = _engine(pk, iouv, seg_chunks=segch)
    em = O.ensemble_matrix(M, N, 77)
    got, det = eng.orie(N, ens_matrix=em, detail=True)
    wd, sd, lc = oracle_cache(pk, iouv)
    want = O.orie_all(wd, sd, lc, em)
    err = np.abs(got - want)
    assert err.max() < 1e-9, (err.max(), int(err.argmax()), got[err.argmax()], want[err.argmax()])
    # per-target AP sums (detail) for a few targets
    for i in range(0, M, 37):
        _, wap, sap = O.orie_one(i, wd, sd, lc, em[i])
        assert abs(det[i, 0] - wap.sum()) < 1e-9 and abs(det[i, 1] - sap.sum()) < 1e-9
        assert det[i, 2] == wap.shape[0]
    eng.close()


def test_orie_waves_and_ranges_agree():
    M, N = 200, 40
    _, pk = make_packed(M=M, seed=5)
    eng = _engine(pk, O.IOU_05_095)
    em = O.ensemble_matrix(M, N, 3)
    full = eng.orie(N, ens_matrix=em)
    small = eng.orie(N, ens_matrix=em, workspace_budget=eng.workspace_bytes(64))
    assert np.array_equal(full, small)
    part = eng.orie(N, ens_matrix=em[64:160], t0=64, nt=96)
    assert np.array_equal(full[64:160], part)
    eng.close()


def test_orie_sampled_ensembles():
    import torch
    M, N = 150, 60
    _, pk = make_packed(M=M, seed=9)
    eng = _engine(pk, O.IOU_05_095)
    bits = eng.sample_bits(N, seed=1234)
    assert bits.shape[0] == M
    member = ((bits[:, :, None] >> np.arange(32)[None, None, :]) & 1).reshape(M, -1)[:, :M].astype(bool)
    assert (member.sum(axis=1) == N).all()
    assert not member[np.arange(M), np.arange(M)].any()
    got = eng.orie(N, seed=1234)
    wd, sd, lc = oracle_cache(pk, O.IOU_05_095)
    em = np.stack([np.nonzero(member[i])[0] for i in range(M)])
    want = O.orie_all(wd, sd, lc, em)
    assert np.abs(got - want).max() < 1e-9
    # more than half of the dataset -> complement path
    N2 = 120
    bits = eng.sample_bits(N2, seed=5)
    member = ((bits[:, :, None] >> np.arange(32)[None, None, :]) & 1).reshape(M, -1)[:, :M].astype(bool)
    assert (member.sum(axis=1) == N2).all() and not member[np.arange(M), np.arange(M)].any()
    # different seeds differ, same seed repeats
    assert np.array_equal(eng.sample_bits(N, seed=1234), eng.sample_bits(N, seed=1234))
    assert not np.array_equal(eng.sample_bits(N, seed=1), eng.sample_bits(N, seed=2))
    eng.close()


def test_bad_ensemble_is_reported():
    from orie_b200._lib import OrieError
    M, N = 64, 5
    _, pk = make_packed(M=M, seed=2)
    eng = _engine(pk, O.IOU_05)
    em = O.ensemble_matrix(M, N, 1)
    em[3, 0] = 3      # a target inside its own ensemble
    with pytest.raises(OrieError):
        eng.orie(N, ens_matrix=em)
    eng.close()


def test_global_memory_membership_table_matches():
    """Datasets beyond ~58 k images keep the 32-target membership table in global memory instead of shared
    memory; force that path on a small dataset and demand identical bits."""
    M, N = 130, 50
    _, pk = make_packed(M=M, seed=44)
    eng = _engine(pk, O.IOU_05_095)
    em = O.ensemble_matrix(M, N, 9)
    ref = eng.orie(N, ens_matrix=em)
    eng2 = _engine(pk, O.IOU_05_095, tuning=dict(walk_gmem=1))
    got = eng2.orie(N, ens_matrix=em)
    assert np.array_equal(ref, got)
    eng.close(); eng2.close()


def test_walk_variants_match():
    """The detection walk has four code paths: one or two 32-target batches per warp (two when two membership tables
    fit shared memory), slot image + TP mask in one packed word (up to 65535 images) or in two arrays, and for each the
    membership table in shared or in global memory.  Force every combination on a small dataset (an odd number of
    batches, so the last pair is half empty): identical bits, and right against the oracle."""
    M, N = 140, 60
    _, pk = make_packed(M=M, seed=45)
    eng = _engine(pk, O.IOU_05_095)
    em = O.ensemble_matrix(M, N, 10)
    ref = eng.orie(N, ens_matrix=em)
    wd, sd, lc = oracle_cache(pk, O.IOU_05_095)
    assert np.abs(ref - O.orie_all(wd, sd, lc, em)).max() < 1e-9
    for single in (0, 1):
        for unpacked in (0, 1):
            for gmem in (0, 1):
                tv = dict(walk_single=single, walk_unpacked=unpacked, walk_gmem=gmem)
                eng2 = _engine(pk, O.IOU_05_095, tuning=tv)
                assert np.array_equal(ref, eng2.orie(N, ens_matrix=em)), tv
                eng2.close()
    eng.close()


def test_class_sharded_sums_add_up():
    """Multi-GPU decomposition by class, emulated on one GPU: every shard keeps all images and its share of the
    classes; per-target AP sums of the shards add up to the un-sharded sums, so one all-reduce recovers the rewards."""
    import torch
    from orie_b200.engine import Engine, class_shard, clamp_ensemble, rewards_from_sums
    M, N, world = 150, 60, 3
    _, pk = make_packed(M=M, seed=55, zipf=0.8)
    eng = Engine(pk, iouv=O.IOU_05_095)
    em = O.ensemble_matrix(M, N, 21)
    ref, ref_detail = eng.orie(N, ens_matrix=em, detail=True)
    eng.close()
    for full in (True, False):
        total = torch.zeros((M, 3), dtype=torch.float64, device="cuda")
        for r in range(world):
            e = Engine(class_shard(pk, r, world), iouv=O.IOU_05_095)
            total += e.orie_sums_device(N, ens_matrix=em, full=full, total_images=M)
            e.check_status()
            e.close()
        got = rewards_from_sums(total, 10, clamp_ensemble(M, N)).cpu().numpy()
        assert np.abs(got - ref).max() < 1e-11
        if full:
            assert np.abs(total.cpu().numpy() - ref_detail).max() < 1e-9
    # device-drawn ensembles depend on (seed, target, M) only, so every shard draws the same ones
    a = Engine(class_shard(pk, 0, world), iouv=O.IOU_05).sample_bits(N, seed=3)
    b = Engine(class_shard(pk, 2, world), iouv=O.IOU_05).sample_bits(N, seed=3)
    assert np.array_equal(a, b)


def test_multi_tile_sort_path_matches():
    """The cooperative radix sort keeps a CTA's items in registers when they fit one tile; larger ranges (the 50k
    sweep) stream tile by tile.  Force that path on a small dataset (2 CTAs for ~35 k detections) and demand
    identical bits, TP flags included."""
    M, N = 300, 120
    _, pk = make_packed(M=M, seed=61)
    em = O.ensemble_matrix(M, N, 4)
    eng = _engine(pk, O.IOU_05_095)
    ref, ref_detail = eng.orie(N, ens_matrix=em, detail=True)
    ref_dev = eng.orie(N, seed=17)
    eng2 = _engine(pk, O.IOU_05_095, tuning=dict(sort_max_blocks=2, sort_lsd=1))
    got, got_detail = eng2.orie(N, ens_matrix=em, detail=True)
    got_dev = eng2.orie(N, seed=17)
    assert eng.info == eng2.info
    assert np.array_equal(ref, got) and np.array_equal(ref_detail, got_detail) and np.array_equal(ref_dev, got_dev)
    eng.close(); eng2.close()


def _orie_with_the_engine_tie_rule(i, wd, sd, lc, ens):
    """The oracle evaluated in the engine's documented tie order: equal confidences sort weak before strong, then by
    image index, then by row — i.e. a stable sort of the records concatenated in image order, the target's strong
    record last."""
    ens = np.sort(np.asarray(ens, dtype=np.int64))
    gt = np.concatenate([lc[s] for s in list(ens) + [i]]).astype(int)
    cat = lambda recs: [np.concatenate(col, axis=0) for col in zip(*recs)]
    weak_ap = O.ap_by_class(*cat([wd[s] for s in np.sort(np.append(ens, i))]), gt)
    strong_ap = O.ap_by_class(*cat([wd[s] for s in ens] + [sd[i]]), gt)
    r = (np.mean(strong_ap) - np.mean(weak_ap)) * (len(ens) + 1) if len(gt) else 0.0
    return 0.0 if np.isnan(r) else float(r)


def test_confidence_ties_and_many_classes_follow_the_tie_rule():
    """Exact confidence ties (as in %g-rounded files) are machine-dependent upstream (unstable argsort); the engine
    sorts them weak before strong, then by image, then by row (tie rule).  300 class ids exercise the
    two-pass class digit of the sort."""
    M, N = 90, 30
    ds, pk = make_packed(M=M, seed=8)
    pk.w_conf[:] = np.round(pk.w_conf, 2)                    # heavy ties inside and across images and detectors
    pk.s_conf[:] = np.round(pk.s_conf, 2)
    remap = np.random.default_rng(0).permutation(300)[:pk.num_classes].astype(np.int32)   # spread over [0, 300)
    pk.w_cls[:] = remap[pk.w_cls]; pk.s_cls[:] = remap[pk.s_cls]; pk.l_cls[:] = remap[pk.l_cls]
    pk.num_classes = 300
    pk.class_values = np.arange(300, dtype=np.int64)
    eng = _engine(pk, O.IOU_05_095)
    em = O.ensemble_matrix(M, N, 13)
    got = eng.orie(N, ens_matrix=em)
    wd, sd, lc = oracle_cache(pk, O.IOU_05_095)
    want = np.array([_orie_with_the_engine_tie_rule(i, wd, sd, lc, em[i]) for i in range(M)])
    assert np.abs(got - want).max() < 1e-9
    eng.close()


def test_images_with_more_rows_than_the_ranking_stage():
    """Own lists are ranked per image in a 512-row shared-memory stage; images with more rows go through it tile by
    tile."""
    from orie_b200 import data, synth
    from orie_b200.synth import DetectorShape
    M, N = 40, 15
    ds = synth.generate(M, 12, 6.0, 0.0, DetectorShape(.6, .08, 700, 1500), DetectorShape(.8, .04, 600, 1200), 91)
    pk = data.pack(ds.labels, ds.weak, ds.strong)
    assert np.diff(pk.w_off).max() > 512 and np.diff(pk.s_off).max() > 512
    eng = _engine(pk, O.IOU_05_095)
    em = O.ensemble_matrix(M, N, 2)
    got = eng.orie(N, ens_matrix=em)
    wd, sd, lc = oracle_cache(pk, O.IOU_05_095)
    assert np.abs(got - O.orie_all(wd, sd, lc, em)).max() < 1e-9
    eng.close()


def test_device_detected_errors_are_sticky_and_reported():
    """The index build is asynchronous: input only the device can reject (a class id outside [0, C)) and a workspace
    too small for the event lists surface through the index status — the rewards are NaN, never silently wrong."""
    import ctypes as C
    import torch
    from orie_b200 import _lib
    from orie_b200._lib import OrieError
    M, N = 64, 10
    _, pk = make_packed(M=M, seed=3)
    good = _engine(pk, O.IOU_05)
    ref = good.orie(N, seed=5)
    bad_cls = pk.w_cls.copy()
    bad_cls[7] = pk.num_classes + 3
    import dataclasses
    bad = _engine(dataclasses.replace(pk, w_cls=bad_cls), O.IOU_05)
    out = bad.orie_device(N, seed=5)
    assert torch.isnan(out).all()
    with pytest.raises(OrieError):
        bad.check_status()
    with pytest.raises(OrieError):
        bad.info
    bad.close()
    # a workspace below the exact size: rejected on the host once the sizes are known ...
    lib = good.lib
    exact = good.workspace_bytes(M)
    assert good.workspace_bound(M) >= exact
    ws = torch.empty(exact, dtype=torch.uint8, device="cuda")
    bits = torch.from_numpy(good.sample_bits(N, seed=5).view(np.int32)).cuda()
    rw = torch.empty(M, dtype=torch.float64, device="cuda")
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    small = exact - 4096 * 64
    if good.info["events"] > 0 and small > 0:
        assert lib.orie_reward(good._handle, 0, M, C.c_void_p(bits.data_ptr()), N, C.c_void_p(ws.data_ptr()), small,
                               C.c_void_p(rw.data_ptr()), C.c_void_p(0), s) == 4       # ORIE_EWORKSPACE
    _lib.check(lib.orie_reward(good._handle, 0, M, C.c_void_p(bits.data_ptr()), N, C.c_void_p(ws.data_ptr()), exact,
                               C.c_void_p(rw.data_ptr()), C.c_void_p(0), s))
    assert np.array_equal(rw.cpu().numpy(), ref)
    # ... and on the device when the call is enqueued before the host knows them
    fresh = _engine(pk, O.IOU_05)
    if good.info["events"] > 0 and small > 0:
        rc = lib.orie_reward(fresh._handle, 0, M, C.c_void_p(bits.data_ptr()), N, C.c_void_p(ws.data_ptr()), small,
                             C.c_void_p(rw.data_ptr()), C.c_void_p(0), s)
        if rc == 0:        # the build had not been waited for: the device flags it
            assert torch.isnan(rw).all()
            assert lib.orie_index_status(fresh._handle) == 4
    fresh.close(); good.close()


def test_bucket_sort_agrees_with_the_radix_passes():
    """The dataset sort of the index build is a bucket sort (one binning pass, per-bucket bitonic sorts) with the LSD
    radix passes as its fallback.  Same order bit for bit on ordinary data, on data with heavy exact confidence ties
    (buckets beyond a warp's 256 items are sorted by a whole CTA) and on data that defeats the binning altogether
    (one confidence value for tens of thousands of rows: the kernel falls back to the radix passes)."""
    from orie_b200 import data, synth
    from orie_b200.synth import DetectorShape
    M, N = 120, 40
    ds = synth.generate(M, 6, 5.0, 0.0, DetectorShape(.6, .08, 250, 300), DetectorShape(.8, .04, 200, 300), 17)
    base = data.pack(ds.labels, ds.weak, ds.strong)
    em = O.ensemble_matrix(M, N, 4)
    cases = {"plain": lambda c: c, "ties": lambda c: np.round(c, 1), "one value": lambda c: np.full_like(c, 0.25)}
    for name, f in cases.items():
        import dataclasses
        pk = dataclasses.replace(base, w_conf=f(base.w_conf), s_conf=f(base.s_conf))
        a = _engine(pk, O.IOU_05_095)
        b = _engine(pk, O.IOU_05_095, tuning=dict(sort_lsd=1))
        ra, da = a.orie(N, ens_matrix=em, detail=True)
        rb, db = b.orie(N, ens_matrix=em, detail=True)
        assert a.info == b.info, name
        assert np.array_equal(ra, rb) and np.array_equal(da, db), name
        a.close(); b.close()
    # and against the oracle in the documented tie order, on the tied data
    pk = dataclasses.replace(base, w_conf=np.round(base.w_conf, 1), s_conf=np.round(base.s_conf, 1))
    eng = _engine(pk, O.IOU_05_095)
    got = eng.orie(N, ens_matrix=em)
    wd, sd, lc = oracle_cache(pk, O.IOU_05_095)
    want = np.array([_orie_with_the_engine_tie_rule(i, wd, sd, lc, em[i]) for i in range(0, M, 7)])
    assert np.abs(got[::7] - want).max() < 1e-9
    eng.close()


def test_replayed_job_matches_the_plain_calls():
    """ReplayJob records a whole job (matching, index build, draw, walk, AP) in a CUDA graph; replays with different
    seeds reproduce what the plain calls compute, from HBM-resident inputs and from pinned host arrays that are
    overwritten in place between replays."""
    import torch
    from orie_b200.engine import DevicePacked, Engine, HostPacked, ReplayJob
    M, N = 200, 60
    _, pk = make_packed(M=M, seed=23)
    eng = _engine(pk, O.IOU_05_095)
    want = {s: eng.orie(N, seed=s) for s in (1, 2, 77)}
    want_sums = eng.orie_sums_device(N, seed=2, t0=64, nt=96).cpu().numpy()
    eng.close()
    dp = DevicePacked(HostPacked(pk), "cuda")
    job = ReplayJob(dp, iouv=O.IOU_05_095, num_ensemble=N)
    assert job.launches_per_replay >= 8
    for s in (1, 2, 77, 1):
        got = job.run(s).cpu().numpy()
        assert np.array_equal(got, want[s])
    job.check_status()
    job.close()
    part = ReplayJob(dp, iouv=O.IOU_05_095, num_ensemble=N, t0=64, nt=96, sums=True)
    assert np.array_equal(part.run(2).cpu().numpy(), want_sums)
    part.close()
    # from pinned host memory: the upload is part of every replay
    hp = HostPacked(pk)
    job = ReplayJob(hp, iouv=O.IOU_05_095, num_ensemble=N)
    assert np.array_equal(job.run(2).cpu().numpy(), want[2])
    hp.w_conf.mul_(0.5)                       # new data in the same buffers (halving keeps the order, changes DCSB-like counts only)
    hp.s_conf.copy_(torch.from_numpy(pk.s_conf[::-1].copy()))       # and a really different strong detector
    import dataclasses
    pk2 = dataclasses.replace(pk, w_conf=pk.w_conf * 0.5, s_conf=pk.s_conf[::-1].copy())
    eng = _engine(pk2, O.IOU_05_095)
    want2 = eng.orie(N, seed=5)
    eng.close()
    assert np.array_equal(job.run(5).cpu().numpy(), want2)
    job.close()
