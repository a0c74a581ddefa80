"""Parity of the CUDA ranking path (junk filter, exact positive ranks by counting, AP, CMC, top-k)
with the reference's mean_ap / cmc (reid_dataset_evaluator.py:283-439) through the golden fixtures
written by the unmodified reference, and with the oracle on seeded inputs.

On the reference's OWN distance matrix the integer outputs (valid flags, first-match ranks, CMC rows,
<=-counts) must be bit-exact and AP equal to the last float64 bit (tolerance 1e-12); from features the
distances differ in the last fp32 bits, so mAP is compared at 1e-6 absolute (BASELINE.json)."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES
from oracle import pps_oracle as O

pytestmark = pytest.mark.gpu


def _ids(d):
    return dict(query_ids=d["qid"], gallery_ids=d["gid"], query_cams=d["qcam"], gallery_cams=d["gcam"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_mean_ap_on_reference_matrix(golden, name):
    import pps_b200
    d = golden(name)
    assert abs(pps_b200.mean_ap(d["dist"], **_ids(d)) - float(d["mAP"])) < 1e-12
    aps, valid = pps_b200.mean_ap(d["dist"], average=False, **_ids(d))
    np.testing.assert_array_equal(valid, d["valid"])
    np.testing.assert_allclose(aps, d["aps"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_cmc_on_reference_matrix(golden, name):
    import pps_b200
    d = golden(name)
    if name == "dup_ties":
        # exact ties: the reference's np.argsort is unstable; compare with the oracle's stable order
        want = O.cmc(d["dist"], topk=10, first_match_break=True, stable=True, **_ids(d))
        got = pps_b200.cmc(d["dist"], topk=10, first_match_break=True, **_ids(d))
        np.testing.assert_allclose(got, want, atol=1e-12)
        return
    np.testing.assert_allclose(pps_b200.cmc(d["dist"], topk=10, first_match_break=True, **_ids(d)), d["cmc_fmb"], atol=1e-12)
    np.testing.assert_allclose(pps_b200.cmc(d["dist"], topk=20, first_match_break=False, **_ids(d)), d["cmc_all"], atol=1e-12)
    rows, v = pps_b200.cmc(d["dist"], topk=10, first_match_break=True, average=False, **_ids(d))
    np.testing.assert_array_equal(rows, d["cmc_rows"])
    np.testing.assert_array_equal(v, d["cmc_valid"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_integer_outputs_equal_count_oracle(golden, name):
    import pps_b200
    d = golden(name)
    res = pps_b200.rank_distmat(d["dist"], d["qid"], d["gid"], d["qcam"], d["gcam"], want_neg_before=True, topk=16)
    ap, valid, first, neg_before = O.rank_counts(d["dist"], d["qid"], d["gid"], d["qcam"], d["gcam"])
    np.testing.assert_array_equal(res.is_valid, valid)
    np.testing.assert_array_equal(res.first_rank, first)
    p = res.pairs
    for i in range(p.nq):
        e = np.arange(p.off[i], p.off[i + 1])
        np.testing.assert_array_equal(res.neg_before[e[p.pos[e] == 1]], neg_before[i])
    ti, td = O.topk_filtered(d["dist"], d["qid"], d["gid"], d["qcam"], d["gcam"], 16)
    np.testing.assert_array_equal(res.topk_index, ti)
    np.testing.assert_array_equal(res.topk_dist, td)


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("precision", ["f16x3", "bf16x3", "bf16x6"])
def test_rank_eval_from_features(golden, name, precision):
    import torch
    import pps_b200
    d = golden(name)
    res = pps_b200.rank_eval(torch.from_numpy(d["q"]).cuda(), torch.from_numpy(d["g"]).cuda(),
                             d["qid"], d["gid"], d["qcam"], d["gcam"], precision=precision)
    np.testing.assert_array_equal(res.is_valid, d["valid"].astype(np.uint8))
    assert abs(res.mean_ap() - float(d["mAP"])) < 1e-6 or name == "dup_ties"
    if name == "dup_ties":
        # duplicated gallery rows: the reference's own fp32 distances differ between duplicates by
        # rounding noise, ours are bit-identical -> AP may legitimately differ inside the tie tolerance
        assert abs(res.mean_ap() - float(d["mAP"])) < 5e-3
    else:
        np.testing.assert_allclose(res.cmc(10, True), d["cmc_fmb"], atol=1e-12)


def test_chunked_gallery_equals_single_block(golden):
    """Counters are integers: sweeping the gallery in chunks gives the same bits as one block."""
    import torch
    import pps_b200
    d = golden("small_mid")
    q, g = torch.from_numpy(d["q"]).cuda(), torch.from_numpy(d["g"]).cuda()
    one = pps_b200.rank_eval(q, g, d["qid"], d["gid"], d["qcam"], d["gcam"], topk=20, want_neg_before=True)
    many = pps_b200.rank_eval(q, g, d["qid"], d["gid"], d["qcam"], d["gcam"], topk=20, want_neg_before=True,
                              max_block_bytes=96 * 256 * 4)
    np.testing.assert_array_equal(one.ap, many.ap)
    np.testing.assert_array_equal(one.first_rank, many.first_rank)
    np.testing.assert_array_equal(one.neg_before, many.neg_before)
    np.testing.assert_array_equal(one.topk_index, many.topk_index)
    np.testing.assert_array_equal(one.topk_dist, many.topk_dist)


def test_compacted_threshold_pass_equals_full_sweep(golden):
    """Multi-chunk galleries take their thresholds from the product queries x (rows that appear in a pair);
    it must give the bits of the two-sweep form, including when those rows need several blocks themselves."""
    import torch
    from pps_b200 import evaluator
    for name in ("small_mid", "many_pos", "ragged_dim", "dup_ties"):
        d = golden(name)
        q, g = torch.from_numpy(d["q"]).cuda(), torch.from_numpy(d["g"]).cuda()
        out = []
        for compact in (False, True):
            eng = evaluator.RankEngine(d["qid"], d["gid"], d["qcam"], d["gcam"], nq=q.shape[0], ng_local=g.shape[0],
                                       dim=q.shape[1], topk=7, want_neg_before=True, max_block_bytes=q.shape[0] * 256 * 4)
            eng.compact_thresholds = compact
            eng.use_c_pass = False                      # (the launch-by-launch Python path has both forms)
            assert eng.n_chunks > 1
            out.append(eng.run(q, g))
            if compact:
                same_id = np.isin(d["gid"], np.unique(d["qid"]))
                assert eng.threshold_rows == int(same_id.sum())
        a, b = out
        np.testing.assert_array_equal(a.ap, b.ap)
        np.testing.assert_array_equal(a.first_rank, b.first_rank)
        np.testing.assert_array_equal(a.neg_before, b.neg_before)
        np.testing.assert_array_equal(a.topk_index, b.topk_index)


@pytest.mark.parametrize("name", ["small_mid", "ragged_dim", "dup_ties"])
@pytest.mark.parametrize("precision", ["bf16x3", "bf16x1"])
def test_fused_rank_epilogue_equals_count_kernel(golden, name, precision):
    """pps_dist_rank_tc (counters in the GEMM epilogue, no distance block) == pps_dist_tc + pps_rank_count, bit for
    bit: AP, first-match ranks (exact duplicate ties included) and per-positive negatives-before counts."""
    import torch
    from pps_b200 import evaluator
    d = golden(name)
    q, g = torch.from_numpy(d["q"]).cuda(), torch.from_numpy(d["g"]).cuda()
    out = []
    for fused in (False, True):
        eng = evaluator.RankEngine(d["qid"], d["gid"], d["qcam"], d["gcam"], nq=q.shape[0], ng_local=g.shape[0],
                                   dim=q.shape[1], want_neg_before=True, precision=precision,
                                   max_block_bytes=q.shape[0] * 256 * 4)
        eng.fused_rank = fused
        assert eng.n_chunks > 1
        out.append(eng.run(q, g))
        assert eng.used_fused_rank == fused
    a, b = out
    np.testing.assert_array_equal(a.is_valid, b.is_valid)
    np.testing.assert_array_equal(a.ap, b.ap)
    np.testing.assert_array_equal(a.first_rank, b.first_rank)
    np.testing.assert_array_equal(a.neg_before, b.neg_before)


def test_fused_rank_epilogue_shapes():
    """Row / column edges of the fused kernel: queries not a multiple of 256 (and > 256 so that several m tiles and
    the superblock schedule are exercised), gallery blocks not a multiple of 256, fp16 inputs."""
    import torch
    import pps_b200
    from pps_b200 import evaluator, synthetic
    for nq, ng, dim, dt in ((300, 3001, 192, torch.float32), (1, 700, 64, torch.float32), (2600, 1500, 128, torch.float16)):
        d = synthetic.make_reid_set(nq=nq, ng=ng, dim=dim, n_ids=max(nq // 6, 1), n_cams=3, n_distractors=ng // 5,
                                    sigma=3.0, seed=nq)
        q, g = torch.from_numpy(d["q"]).cuda().to(dt), torch.from_numpy(d["g"]).cuda().to(dt)
        out = []
        for fused in (False, True):
            eng = evaluator.RankEngine(d["qid"], d["gid"], d["qcam"], d["gcam"], nq=nq, ng_local=ng, dim=dim,
                                       max_block_bytes=max(nq, 1) * 700 * 4, in_dtype=dt, precision="bf16x3")
            eng.fused_rank = fused             # (the counting epilogue has no row-scale path: bf16 planes or fp16 inputs)
            eng.use_c_pass = False
            out.append(eng.run(q, g))
            assert eng.used_fused_rank == (fused and eng.n_chunks > 1)
        np.testing.assert_array_equal(out[0].ap, out[1].ap)
        np.testing.assert_array_equal(out[0].first_rank, out[1].first_rank)


def test_prefiltered_pair_lists_equal_plain(monkeypatch):
    """pps_pairs_prefilter + sweeps on the candidates + pps_pairs_remap == the plain sweeps over the whole gallery
    (ids far apart, negative ids, an id shared by many queries, block edges of the ordered compaction)."""
    import torch
    from pps_b200 import evaluator
    rs = np.random.RandomState(11)
    nq, ng = 500, 20000
    qid = rs.randint(-50, 200, size=nq).astype(np.int64) * 1000003
    gid = np.where(rs.rand(ng) < 0.1, rs.randint(-50, 200, size=ng) * 1000003, rs.randint(10 ** 9, 2 * 10 ** 9, size=ng)).astype(np.int64)
    gid[2047] = gid[2048] = gid[4095] = qid[0]
    gid[ng - 1] = qid[1]
    qcam, gcam = rs.randint(0, 4, size=nq), rs.randint(0, 4, size=ng)
    dev = torch.device("cuda")
    plain = evaluator.DevicePairs(qid, qcam, gid, gcam, dev)
    assert not plain.prefilter
    plain.begin().finish()
    monkeypatch.setattr(evaluator, "PREFILTER_MIN_ROWS", 1000)
    filt = evaluator.DevicePairs(qid, qcam, gid, gcam, dev)
    assert filt.prefilter
    filt.begin().finish()
    assert filt.n_cand == int(np.isin(gid, np.unique(qid)).sum())
    assert filt.n_pairs == plain.n_pairs and filt.max_pairs == plain.max_pairs
    n = plain.n_pairs
    np.testing.assert_array_equal(filt.off, plain.off)
    np.testing.assert_array_equal(filt.q[:n], plain.q[:n])
    np.testing.assert_array_equal(filt.g[:n], plain.g[:n])
    np.testing.assert_array_equal(filt.pos[:n], plain.pos[:n])
    # no query id at all in the gallery
    none = evaluator.DevicePairs(qid, qcam, np.arange(ng, dtype=np.int64) + 10 ** 12, gcam, dev)
    none.begin().finish()
    assert none.n_cand == 0 and none.n_pairs == 0


def test_compact_rows_layout():
    """pps_pairs_compact_rows: gp_rows = the distinct same-id gallery rows of the window grouped by id (one
    representative query per id), pair_col = where each pair's row sits in it (or -1 outside the window)."""
    import torch
    from pps_b200 import _lib, evaluator
    rs = np.random.RandomState(3)
    nq, ng = 300, 4000
    qid, gid = rs.randint(1, 40, size=nq), rs.randint(0, 60, size=ng)
    qcam, gcam = rs.randint(0, 3, size=nq), rs.randint(0, 3, size=ng)
    lib = _lib.load()
    dev = torch.device("cuda")
    p = evaluator.DevicePairs(qid, qcam, gid, gcam, dev)
    p.begin().finish()
    n = p.n_pairs
    for lo, hi in ((0, ng), (1000, 2500), (3999, 4000), (10, 10)):
        ws = torch.empty(int(lib.pps_pairs_compact_workspace_bytes(nq)), dtype=torch.uint8, device=dev)
        rows = torch.full((n,), -7, dtype=torch.int32, device=dev)
        col = torch.full((n,), -7, dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(lib.pps_pairs_compact_rows(_lib.ptr(p.qid), nq, _lib.ptr(p.off_d), _lib.ptr(p.q_d), _lib.ptr(p.g_d), n,
                                              lo, hi, _lib.ptr(ws), _lib.ptr(rows), _lib.ptr(col), _lib.ptr(cnt),
                                              _lib.stream_ptr()), "pps_pairs_compact_rows")
        torch.cuda.synchronize()
        n_rows = int(cnt[0])
        rows, col = rows.cpu().numpy()[:n_rows], col.cpu().numpy()
        inside = np.isin(gid, np.unique(qid)) & (np.arange(ng) >= lo) & (np.arange(ng) < hi)
        assert n_rows == int(inside.sum())
        np.testing.assert_array_equal(np.sort(rows), np.nonzero(inside)[0])       # every such row exactly once
        pg, pq = p.g[:n], p.q[:n]
        win = (pg >= lo) & (pg < hi)
        assert (col[~win] == -1).all()
        np.testing.assert_array_equal(rows[col[win]], pg[win])                    # the column holds the pair's row
        # grouped by id in the order of each id's first query
        first_q = [np.nonzero(qid == i)[0][0] for i in gid[rows]]
        assert (np.diff(first_q) >= 0).all()


def test_more_positives_than_one_window():
    """> 63 positives per query: the count kernel walks several threshold windows."""
    import pps_b200
    rs = np.random.RandomState(7)
    nq, ng = 9, 5000
    dist = rs.rand(nq, ng).astype(np.float32)
    dist[:, ::7] = dist[:, 1::7][:, :dist[:, ::7].shape[1]]       # plenty of exact ties
    qid = np.arange(1, nq + 1)
    gid = rs.randint(1, 4, size=ng)                               # ~1 666 same-id items for ids 1..3
    gid[:50] = np.arange(50) % 9 + 1
    qcam = rs.randint(0, 3, size=nq)
    gcam = rs.randint(0, 3, size=ng)
    res = pps_b200.rank_distmat(dist, qid, gid, qcam, gcam, want_neg_before=True)
    ap, valid, first, neg_before = O.rank_counts(dist, qid, gid, qcam, gcam)
    np.testing.assert_array_equal(res.is_valid, valid)
    np.testing.assert_array_equal(res.first_rank, first)
    np.testing.assert_allclose(res.ap, ap, rtol=0, atol=1e-12)
    assert abs(pps_b200.mean_ap(dist, qid, gid, qcam, gcam) - O.mean_ap(dist, qid, gid, qcam, gcam)) < 1e-12


def test_no_valid_query_raises():
    import pps_b200
    dist = np.random.RandomState(0).rand(4, 10).astype(np.float32)
    qid, gid = np.arange(4) + 100, np.arange(10)
    cams_q, cams_g = np.zeros(4, np.int64), np.ones(10, np.int64)
    with pytest.raises(RuntimeError, match="No valid query"):
        pps_b200.cmc(dist, qid, gid, cams_q, cams_g, topk=5, first_match_break=True)
    # junk-only matches (same id, same camera) do not make a query valid either (:327-332)
    gid2 = gid.copy(); gid2[:4] = qid
    with pytest.raises(RuntimeError, match="No valid query"):
        pps_b200.cmc(dist, qid, gid2, cams_q, np.zeros(10, np.int64), topk=5, first_match_break=True)
    with pytest.raises(AssertionError):
        pps_b200.mean_ap(dist.tolist(), qid, gid, cams_q, cams_g)   # :411-415 isinstance asserts


@pytest.mark.parametrize("name", ["small_mid", "some_invalid"])
def test_evaluate_host_entry_point(golden, name):
    """pps_evaluate_host: host buffers in, metrics out (what bench.py times as e2e)."""
    import pps_b200
    d = golden(name)
    out = pps_b200.evaluate_host(d["q"], d["g"], d["qid"], d["gid"], d["qcam"], d["gcam"], cmc_topk=10, topk=8)
    assert abs(out["mAP"] - float(d["mAP"])) < 1e-6
    np.testing.assert_allclose(out["cmc"], d["cmc_fmb"], atol=1e-12)
    np.testing.assert_array_equal(out["valid"], d["valid"].astype(np.uint8))
    ti, td = O.topk_filtered(d["dist"], d["qid"], d["gid"], d["qcam"], d["gcam"], 8)
    assert np.mean(out["topk_index"] == ti) > 0.99           # modulo last-bit distance differences
    np.testing.assert_allclose(out["topk_dist"], td, rtol=1e-4)


def test_evaluate_arrays_single_and_multi_query(golden):
    """evaluate(): marks 0/1/2 split, single-query + multi-query branches (:57-159)."""
    import pps_b200
    d = golden("small_mid")
    rs = np.random.RandomState(1)
    n_mq = 60
    mq_src = rs.randint(0, len(d["qid"]), size=n_mq)
    mq_feat = d["q"][mq_src] + 0.01 * rs.randn(n_mq, d["q"].shape[1]).astype(np.float32)
    feats = np.concatenate([d["q"], d["g"], mq_feat]).astype(np.float32)
    ids = np.concatenate([d["qid"], d["gid"], d["qid"][mq_src]])
    cams = np.concatenate([d["qcam"], d["gcam"], d["qcam"][mq_src]])
    marks = np.concatenate([np.zeros(len(d["qid"])), np.ones(len(d["gid"])), np.full(n_mq, 2)]).astype(np.int64)
    mAP, cmc_scores, mq_mAP, mq_cmc = pps_b200.evaluate_arrays(feats, ids, cams, marks)
    assert abs(mAP - float(d["mAP"])) < 1e-6
    np.testing.assert_allclose(cmc_scores, d["cmc_fmb"], atol=1e-12)
    # multi-query oracle: mean of the mark==2 features per (id, cam), then the same scoring (:131-159)
    keys, pooled = [], []
    seen = {}
    for k, (i, c) in enumerate(zip(ids[marks == 2], cams[marks == 2])):
        seen.setdefault((i, c), []).append(k)
    for key, idx in seen.items():
        keys.append(key)
        pooled.append(mq_feat[idx].mean(axis=0))
    pooled = np.stack(pooled).astype(np.float32)
    dist = O.compute_dist(pooled, d["g"])
    kid, kcam = np.array([k[0] for k in keys]), np.array([k[1] for k in keys])
    assert abs(mq_mAP - O.mean_ap(dist, kid, d["gid"], kcam, d["gcam"])) < 1e-6
    np.testing.assert_allclose(mq_cmc, O.cmc(dist, kid, d["gid"], kcam, d["gcam"], topk=10, first_match_break=True), atol=1e-12)

    class FakeDataset:                      # evaluate(json_dataset, all_feats, output_dir)
        def get_roidb(self, gt=True):
            return [dict(image="/x/%08d_%04d_%08d.jpg" % (i, c, k), mark=int(m))
                    for k, (i, c, m) in enumerate(zip(ids, cams, marks))]
    r = pps_b200.evaluate(FakeDataset(), feats, None, verbose=False)
    assert r[0] == mAP and r[2] == mq_mAP
    res = pps_b200.reid_results(r, "market1501_test")
    assert res["market1501_test"]["ReID"]["CMC1"] == cmc_scores[0]


def test_cuhk03_shape_full_parity_with_oracle():
    """BASELINE configs[0]: 1 400 x 5 332 x 2048 — whole evaluation against the oracle run in full."""
    import torch
    import pps_b200
    from pps_b200 import synthetic
    d = synthetic.make_config("cuhk03")
    res = pps_b200.rank_eval(torch.from_numpy(d["q"]).cuda(), torch.from_numpy(d["g"]).cuda(),
                             d["qid"], d["gid"], d["qcam"], d["gcam"], topk=10)
    dist = O.compute_dist(d["q"], d["g"])
    ids = _ids(d)
    want_map = O.mean_ap(dist, ap_fn=O.average_precision_step, **ids)
    want_cmc = O.cmc(dist, topk=10, first_match_break=True, **ids)
    assert abs(res.mean_ap() - want_map) < 1e-6
    # CMC is a count of integer ranks; a rank can move only where two distances agree to 1e-4 relative
    ap, valid, first, _ = O.rank_counts(dist, d["qid"], d["gid"], d["qcam"], d["gcam"])
    np.testing.assert_array_equal(res.is_valid, valid)
    moved = np.nonzero(res.first_rank != first)[0]
    assert len(moved) <= 2, "first-match ranks differ for %d queries" % len(moved)
    assert np.max(np.abs(res.cmc(10, True) - want_cmc)) <= 2.0 / max(valid.sum(), 1)


def _full_shape_parity(name, topk, small_block_bytes):
    """Whole BASELINE config on the GPU against the oracle run IN FULL (every query x the whole gallery):
    (i) gallery-chunked == single block, bit for bit; (ii) mean AP within 1e-6 of the oracle (north_star); (iii) valid
    flags identical; (iv) every first-match rank / top-k entry that differs from the oracle's is proven to be a distance
    tie within 1e-4 relative (tests/parity_util.py) - no "most of them agree"; (v) on the GPU's own distance rows the
    integer outputs are bit-exact and AP agrees to 1e-12."""
    import torch
    import pps_b200
    from pps_b200 import synthetic
    import parity_util as P
    d = synthetic.make_config(name)
    q, g = torch.from_numpy(d["q"]).cuda(), torch.from_numpy(d["g"]).cuda()
    ids = (d["qid"], d["gid"], d["qcam"], d["gcam"])
    one = pps_b200.rank_eval(q, g, *ids, topk=topk)
    many = pps_b200.rank_eval(q, g, *ids, topk=topk, max_block_bytes=small_block_bytes)
    np.testing.assert_array_equal(one.ap, many.ap)
    np.testing.assert_array_equal(one.first_rank, many.first_rank)
    if topk:
        np.testing.assert_array_equal(one.topk_index, many.topk_index)
        np.testing.assert_array_equal(one.topk_dist, many.topk_dist)
    dist = O.compute_dist(d["q"], d["g"])
    ap, valid, first, _ = O.rank_counts(dist, *ids)
    np.testing.assert_array_equal(one.is_valid, valid)
    assert abs(one.mean_ap() - float(ap.sum() / valid.sum())) < 1e-6
    moved = P.assert_first_rank_parity(one.first_rank, dist, d["qid"], d["qcam"], d["gid"], d["gcam"], what=name)
    assert moved <= 0.02 * len(first), "%d first-match ranks sit on a tie: implausibly many" % moved
    want_cmc = O.cmc(dist, topk=10, first_match_break=True, query_ids=d["qid"], gallery_ids=d["gid"], query_cams=d["qcam"],
                     gallery_cams=d["gcam"])
    assert np.max(np.abs(one.cmc(10, True) - want_cmc)) <= moved / max(valid.sum(), 1) + 1e-12
    sel = np.arange(0, len(first), max(len(first) // 96, 1))[:96]
    if topk:
        P.assert_topk_parity(one.topk_index[sel], one.topk_dist[sel], dist[sel], d["qid"][sel], d["qcam"][sel], d["gid"],
                             d["gcam"], what=name)
    own = pps_b200.compute_dist(q[torch.from_numpy(sel).cuda()], g).cpu().numpy()
    P.assert_counts_exact_on_own_distances(one, own, d["qid"], d["qcam"], d["gid"], d["gcam"], O, sel=sel, topk=topk)
    np.testing.assert_allclose(own, dist[sel], rtol=1e-4, atol=1e-6)


def test_market_shape_full_parity_with_oracle():
    """BASELINE configs[1]: 3 368 x 19 732 x 2048, top-100."""
    _full_shape_parity("market1501", 100, 64 << 20)


@pytest.mark.parametrize("nq,ng,n_ids", [(70, 600, 20), (33, 5000, 7), (1, 1, 1), (257, 2049, 300), (5, 70000, 3)])
def test_device_pair_lists_equal_host_builder(nq, ng, n_ids):
    """pairs.cu (count / scan / fill on the device) reproduces pps_pairs_fill element for element."""
    from pps_b200 import evaluator
    rs = np.random.RandomState(nq + ng)
    qid, gid = rs.randint(1, n_ids + 1, size=nq), rs.randint(0, n_ids + 1, size=ng)
    qcam, gcam = rs.randint(0, 3, size=nq), rs.randint(0, 3, size=ng)
    host = evaluator.PairLists(qid, qcam, gid, gcam)
    dev = evaluator.DevicePairs(qid, qcam, gid, gcam, "cuda").begin().finish()
    assert dev.n_pairs == host.n_pairs and dev.max_pairs == host.max_pairs
    np.testing.assert_array_equal(dev.off, host.off)
    n = host.n_pairs
    np.testing.assert_array_equal(dev.q[:n], host.q[:n])
    np.testing.assert_array_equal(dev.g[:n], host.g[:n])
    np.testing.assert_array_equal(dev.pos[:n], host.pos[:n])
    np.testing.assert_array_equal(dev.n_pos_per_q, host.n_pos_per_q)
    # a second build on the same object (what RankEngine does every step) gives the same lists
    dev.begin().finish()
    np.testing.assert_array_equal(dev.g[:n], host.g[:n])


def test_device_pair_lists_no_matches():
    from pps_b200 import evaluator
    dev = evaluator.DevicePairs(np.array([5, 6]), np.array([0, 1]), np.array([1, 2, 3]), np.array([0, 0, 0]), "cuda")
    dev.begin().finish()
    assert dev.n_pairs == 0 and dev.max_pairs == 0 and list(dev.off) == [0, 0, 0]


def test_host_context_reuse_gives_identical_results(golden):
    import pps_b200
    d = golden("small_mid")
    a = pps_b200.evaluate_host(d["q"], d["g"], d["qid"], d["gid"], d["qcam"], d["gcam"], cmc_topk=10, topk=5)
    d2 = golden("many_pos")
    pps_b200.evaluate_host(d2["q"], d2["g"], d2["qid"], d2["gid"], d2["qcam"], d2["gcam"], cmc_topk=10)
    b = pps_b200.evaluate_host(d["q"], d["g"], d["qid"], d["gid"], d["qcam"], d["gcam"], cmc_topk=10, topk=5)
    assert a["mAP"] == b["mAP"]
    np.testing.assert_array_equal(a["ap"], b["ap"])
    np.testing.assert_array_equal(a["topk_index"], b["topk_index"])


def test_rank_on_matrix_with_negative_and_tied_values():
    """mean_ap / cmc accept any real-valued 'distance' (e.g. a negated similarity): the count kernel's integer
    keys must stay monotone across the sign, over exact ties and over a very wide dynamic range."""
    import pps_b200
    rs = np.random.RandomState(21)
    nq, ng = 24, 3000
    dist = (rs.randn(nq, ng) * np.exp(rs.uniform(-12, 12, size=(nq, 1)))).astype(np.float32)
    dist[:, 100:200] = dist[:, 300:400]                       # exact ties
    dist[3] = np.round(dist[3] / np.abs(dist[3]).max() * 4)   # a row with only 9 distinct values
    dist[4] = 0.0
    dist[5, ::2] = -0.0
    qid = rs.randint(1, 9, size=nq)
    gid = rs.randint(0, 9, size=ng)
    qcam, gcam = rs.randint(0, 3, size=nq), rs.randint(0, 3, size=ng)
    res = pps_b200.rank_distmat(dist, qid, gid, qcam, gcam, want_neg_before=True)
    ap, valid, first, neg_before = O.rank_counts(dist, qid, gid, qcam, gcam)
    np.testing.assert_array_equal(res.is_valid, valid)
    np.testing.assert_array_equal(res.first_rank, first)
    np.testing.assert_allclose(res.ap, ap, rtol=0, atol=1e-12)
    p = res.pairs
    for i in range(p.nq):
        e = np.arange(p.off[i], p.off[i + 1])
        np.testing.assert_array_equal(res.neg_before[e[p.pos[e] == 1]], neg_before[i])


def test_duke_shape_full_parity_with_oracle():
    """BASELINE configs[2] (DukeMTMC-reID-shaped, 2 228 x 17 661 x 2048): whole evaluation against the oracle in full."""
    _full_shape_parity("duke", 0, 48 << 20)


def test_features_npy_cli_roundtrip(tmp_path, golden):
    """features.npy + COCO-style json on disk -> the CLI -> the reference's result dict (task_evaluation.py:437-452)."""
    import json
    from pps_b200 import dataset_io
    d = golden("small_mid")
    feats = np.concatenate([d["q"], d["g"]]).astype(np.float32)
    ids = np.concatenate([d["qid"], d["gid"]]); cams = np.concatenate([d["qcam"], d["gcam"]])
    marks = np.concatenate([np.zeros(len(d["qid"])), np.ones(len(d["gid"]))]).astype(int)
    order = np.random.RandomState(0).permutation(len(ids))            # image ids in a shuffled order: the reader sorts them
    images = [dict(id=int(i), file_name="%08d_%04d_%08d.jpg" % (ids[i], cams[i], i)) for i in order]
    anns = [dict(id=1000 + int(i), image_id=int(i), mark=int(marks[i])) for i in order]
    (tmp_path / "split.json").write_text(json.dumps(dict(images=images, annotations=anns)))
    np.save(tmp_path / "features.npy", feats)
    res = dataset_io.main(["--features", str(tmp_path / "features.npy"), "--annotations", str(tmp_path / "split.json"),
                           "--output", str(tmp_path / "out.json")])
    r = res["split"]["ReID"]
    assert abs(r["mAP"] - float(d["mAP"])) < 1e-6 and abs(r["CMC1"] - d["cmc_fmb"][0]) < 1e-12 and r["mq_mAP"] == -1
    assert json.loads((tmp_path / "out.json").read_text())["split"]["ReID"]["CMC10"] == r["CMC10"]


def test_new_entry_points_degenerate_inputs():
    """Empty / tiny inputs of the large-gallery entry points: nothing launched, sane outputs, argument errors."""
    import torch
    from pps_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda")
    s = _lib.stream_ptr()
    n_out = torch.full((1,), 7, dtype=torch.int32, device=dev)
    # pre-filter with no queries / no gallery rows -> zero candidates
    assert lib.pps_pairs_prefilter(None, 0, None, None, 0, None, None, None, None, _lib.ptr(n_out), s) == 0
    torch.cuda.synchronize()
    assert int(n_out[0]) == 0
    assert lib.pps_pairs_prefilter(None, 0, None, None, 0, None, None, None, None, None, s) == _lib.PPS_ERR_INVALID_ARG
    # compaction with no pairs
    n_out.fill_(7)
    assert lib.pps_pairs_compact_rows(None, 5, None, None, None, 0, 0, 10, None, None, None, _lib.ptr(n_out), s) == 0
    torch.cuda.synchronize()
    assert int(n_out[0]) == 0
    assert lib.pps_pairs_remap(None, 0, None, 0, s) == 0
    # gather split / fused distance with zero rows are no-ops; bad table widths are rejected
    assert lib.pps_split_rows_gather(None, _lib.DTYPE_F32, None, 0, 0, 64, 64, 2, None, None, s) == 0
    assert lib.pps_rank_tab_elems(300, 12) == _lib.PPS_ERR_INVALID_ARG and lib.pps_rank_tab_elems(300, 72) == _lib.PPS_ERR_INVALID_ARG
    assert lib.pps_rank_tab_elems(300, 16) == 512 * 16
    assert lib.pps_dist_rank_tc(None, None, 0, 1, 0, None, None, 0, 1, 0, 64, _lib.PREC_BF16X3, 0, 0, 16, None, None, None, None,
                                None, s) == 0
    # sweep with k out of range
    assert lib.pps_rank_sweep(None, 0, 1, 1, 0, None, None, None, None, 0, None, None, None, 0, 1, s) == _lib.PPS_ERR_INVALID_ARG
    assert lib.pps_rank_sweep(None, 0, 0, 0, 0, None, None, None, None, 0, None, None, None, 10, 1, s) == 0


def test_rank_eval_topk_without_any_pair():
    """No query id occurs in the gallery: nothing to count, but the top-k sweep still has to run."""
    import torch
    import pps_b200
    rs = np.random.RandomState(2)
    q, g = rs.randn(5, 64).astype(np.float32), rs.randn(900, 64).astype(np.float32)
    res = pps_b200.rank_eval(torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda(), np.arange(5) + 1000, np.arange(900),
                             np.zeros(5, np.int64), np.zeros(900, np.int64), topk=9)
    assert res.is_valid.sum() == 0
    d = O.compute_dist(q, g)
    np.testing.assert_array_equal(res.topk_index, np.argsort(d, axis=1, kind="stable")[:, :9])


def test_topk_admission_in_distance_epilogue(golden):
    """Blocks after the first take their top-k candidates in the epilogue of the distance kernel; the result must be
    the bits of the one-read sweep, with and without the junk filter, and an overflowing candidate buffer must fall
    back to the sweep (forced here by a 4-entry buffer and a gallery whose nearest rows come last)."""
    import torch
    from pps_b200 import evaluator
    for name in ("small_mid", "dup_ties", "many_pos"):
        d = golden(name)
        q, g = torch.from_numpy(d["q"]).cuda(), torch.from_numpy(d["g"]).cuda()
        for filtered in (True, False):
            out = []
            for fused in (False, True):
                eng = evaluator.RankEngine(d["qid"], d["gid"], d["qcam"], d["gcam"], nq=q.shape[0], ng_local=g.shape[0],
                                           dim=q.shape[1], topk=13, topk_filtered=filtered, max_block_bytes=q.shape[0] * 256 * 4)
                eng.fused_topk = fused
                eng.use_c_pass = False                  # (tests/test_gpu_pass.py covers the C-driven pass)
                assert eng.n_chunks > 1
                out.append(eng.run(q, g))
                assert eng.used_fused_topk == fused
            np.testing.assert_array_equal(out[0].topk_index, out[1].topk_index)
            np.testing.assert_array_equal(out[0].topk_dist, out[1].topk_dist)
            np.testing.assert_array_equal(out[0].ap, out[1].ap)
    # overflow -> fallback
    d = golden("small_mid")
    order = np.argsort(-O.compute_dist(d["q"][:1], d["g"])[0])            # farthest rows first
    g = torch.from_numpy(np.ascontiguousarray(d["g"][order])).cuda()
    q = torch.from_numpy(d["q"]).cuda()
    gid, gcam = d["gid"][order], d["gcam"][order]
    ref = evaluator.RankEngine(d["qid"], gid, d["qcam"], gcam, nq=q.shape[0], ng_local=g.shape[0], dim=q.shape[1], topk=13)
    want = ref.run(q, g)
    eng = evaluator.RankEngine(d["qid"], gid, d["qcam"], gcam, nq=q.shape[0], ng_local=g.shape[0], dim=q.shape[1], topk=13,
                               max_block_bytes=q.shape[0] * 256 * 4)
    eng.TOPK_CAND_CAP = 4
    eng.use_c_pass = False
    got = eng.run(q, g)
    assert eng.fused_topk and not eng.used_fused_topk                    # the repeat ran without the epilogue path
    np.testing.assert_array_equal(got.topk_index, want.topk_index)
    np.testing.assert_array_equal(got.ap, want.ap)


def test_group_mean_rows_is_numpy_mean():
    """pps_group_mean_rows == np.stack([np.mean(feats[rows], axis=0) ...]) bit for bit (float32 sum in list order, then
    one division), the multi-query pooling of reid_dataset_evaluator.py:131-143."""
    import torch
    from pps_b200 import evaluator
    rs = np.random.RandomState(4)
    f = rs.randn(50, 77).astype(np.float32)
    groups = [[3], [0, 49, 7], list(range(10, 40)), [5, 5, 6]]
    got = evaluator.group_mean_rows(torch.from_numpy(f).cuda(), groups).cpu().numpy()
    want = np.stack([np.mean(f[g], axis=0) for g in groups])
    np.testing.assert_array_equal(got, want)
    assert evaluator.group_mean_rows(torch.from_numpy(f).cuda(), []).shape == (0, 77)


@pytest.mark.parametrize("name", ["small_mid", "ragged_dim", "many_pos", "dup_ties", "some_invalid"])
def test_trapezoid_ap_definition(golden, name):
    """mean_ap(ap_definition='trapezoid') == the scikit-learn 0.18.1 definition the reference asks for (:398-407),
    restated in the oracle; exact duplicate distances (dup_ties) go through the exact-tie counts."""
    import pps_b200
    d = golden(name)
    ids = dict(query_ids=d["qid"], gallery_ids=d["gid"], query_cams=d["qcam"], gallery_cams=d["gcam"])
    want_aps, want_valid = O.mean_ap(d["dist"], average=False, ap_fn=O.average_precision_trapezoid, **ids)
    aps, valid = pps_b200.mean_ap(d["dist"], average=False, ap_definition="trapezoid", **ids)
    np.testing.assert_array_equal(valid, want_valid)
    np.testing.assert_allclose(aps, want_aps, rtol=0, atol=1e-12)
    step = pps_b200.mean_ap(d["dist"], **ids)
    trap = pps_b200.mean_ap(d["dist"], ap_definition="trapezoid", **ids)
    assert abs(step - float(d["mAP"])) < 1e-12 and trap != step


@pytest.mark.parametrize("name", ["small_mid", "ragged_dim", "many_pos", "some_invalid"])
def test_cmc_separate_camera_set(golden, name):
    """The cmc branch the reference's evaluate() never selects but its signature offers (:329-331): every gallery item of
    the query's camera is removed.  Pinned to the unmodified reference; single_gallery_shot stays NotImplemented."""
    import pps_b200
    d, s = golden(name), golden("cmc_sep_" + name)
    ids = _ids(d)
    np.testing.assert_allclose(pps_b200.cmc(d["dist"], topk=10, separate_camera_set=True, first_match_break=True, **ids), s["cmc_fmb"], atol=1e-12)
    np.testing.assert_allclose(pps_b200.cmc(d["dist"], topk=20, separate_camera_set=True, first_match_break=False, **ids), s["cmc_all"], atol=1e-12)
    rows, valid = pps_b200.cmc(d["dist"], topk=10, separate_camera_set=True, first_match_break=True, average=False, **ids)
    np.testing.assert_array_equal(rows, s["cmc_rows"])
    np.testing.assert_array_equal(valid, s["cmc_valid"])
    with pytest.raises(NotImplementedError):
        pps_b200.cmc(d["dist"], single_gallery_shot=True, **ids)
