"""The oracle (oracle/pps_oracle.py) against the fixtures produced by the UNMODIFIED reference
functions (oracle/make_golden.py), and against the live reference when /root/reference exists."""
import contextlib
import io

import numpy as np
import pytest

from oracle import pps_oracle as O
from oracle import ref_loader
from conftest import GOLDEN_CASES, POOL_GOLDEN_CASES, pool_fixture_expected


def _ids(d):
    return dict(query_ids=d["qid"], gallery_ids=d["gid"], query_cams=d["qcam"], gallery_cams=d["gcam"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_compute_dist_matches_reference_fixture(golden, name):
    d = golden(name)
    dist = O.compute_dist(d["q"], d["g"])
    assert dist.dtype == np.float32 and dist.shape == d["dist"].shape
    # same formula, but BLAS / SIMD sgemm order may differ between hosts
    np.testing.assert_allclose(dist, d["dist"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_mean_ap_and_cmc_match_reference_fixture(golden, name):
    d = golden(name)
    dist = d["dist"]          # rank on the reference's own matrix: outputs must agree to the last bit
    assert abs(O.mean_ap(dist, **_ids(d)) - float(d["mAP"])) < 1e-12
    aps, valid = O.mean_ap(dist, average=False, **_ids(d))
    np.testing.assert_allclose(aps, d["aps"], rtol=0, atol=1e-12)
    np.testing.assert_array_equal(valid, d["valid"])
    if name != "dup_ties":    # CMC under exact ties depends on the (unstable) sort
        np.testing.assert_allclose(O.cmc(dist, topk=10, first_match_break=True, **_ids(d)), d["cmc_fmb"], atol=1e-12)
        np.testing.assert_allclose(O.cmc(dist, topk=20, first_match_break=False, **_ids(d)), d["cmc_all"], atol=1e-12)
        rows, v = O.cmc(dist, topk=10, first_match_break=True, average=False, **_ids(d))
        np.testing.assert_array_equal(rows, d["cmc_rows"])
        np.testing.assert_array_equal(v, d["cmc_valid"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_restated_ap_definitions(golden, name):
    d = golden(name)
    dist = d["dist"]
    step = O.mean_ap(dist, ap_fn=O.average_precision_step, **_ids(d))
    assert abs(step - float(d["mAP"])) < 1e-12          # == installed scikit-learn (>= 0.19)
    trap = O.mean_ap(dist, ap_fn=O.average_precision_trapezoid, **_ids(d))
    assert 0.0 < trap <= 1.0 and trap != step            # the 0.18.1 definition is a different number


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_count_based_restatement_equals_sort_based(golden, name):
    """What the GPU kernels compute (<=-counts) equals the reference's sort + sklearn AP."""
    d = golden(name)
    dist = d["dist"]
    ap, valid, first, neg_before = O.rank_counts(dist, d["qid"], d["gid"], d["qcam"], d["gcam"])
    np.testing.assert_allclose(ap, d["aps"], rtol=0, atol=1e-12)
    np.testing.assert_array_equal(valid, d["valid"].astype(np.uint8))
    rows, v = O.cmc(dist, topk=10, first_match_break=True, average=False, stable=True, **_ids(d))
    first_from_rows = np.where(rows[:, -1] > 0, (rows == 0).sum(axis=1), -1)
    ok = (first >= 0) & (first < 10)
    np.testing.assert_array_equal(first[ok], first_from_rows[ok])
    assert np.all(first_from_rows[(valid > 0) & ~ok] == -1)


def test_uniform_partition_split_tables():
    assert O.uniform_partition_split(6) == [4] * 6
    assert O.uniform_partition_split(5) == [5, 5, 4, 5, 5]
    assert O.uniform_partition_split(7) == [3, 3, 4, 4, 4, 3, 3]
    assert O.uniform_partition_split(6, 384, 1.0 / 8) == [8] * 6
    assert O.uniform_partition_split(5, 384, 1.0 / 8) == [10, 10, 8, 10, 10]
    assert sum(O.uniform_partition_split(10)) == 24 and sum(O.uniform_partition_split(9)) == 24


def test_pooling_restatement_float64_crosscheck():
    rs = np.random.RandomState(3)
    x = np.maximum(rs.randn(3, 40, 24, 8), 0).astype(np.float32)
    for mode in ("max_ave", "avg_max"):
        y32 = O.pps_pool(x, 6, mode=mode)
        y64 = O.pps_pool(x.astype(np.float64), 6, mode=mode, dtype=np.float64)
        assert y32.shape == (3, 63, 40)
        np.testing.assert_allclose(y32, y64, rtol=2e-6, atol=1e-7)
    # equal strips: mean of strip averages == region average
    avg, _ = O.strip_pools(x.astype(np.float64), [4] * 6, np.float64)
    y = O.pps_pool(x.astype(np.float64), 6, mode="avg_max", dtype=np.float64)
    np.testing.assert_allclose(y[:, 0, :], avg[0])
    np.testing.assert_allclose(y[:, 2, :], np.maximum(avg[0], avg[1]))   # mask 3 = parts {0,1}


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_oracle_against_live_reference():
    from pps_b200 import synthetic
    ref = ref_loader.load()
    d = synthetic.make_reid_set(nq=60, ng=500, dim=80, n_ids=15, n_cams=3, n_distractors=30, sigma=2.5, seed=21)
    ids = _ids(d)
    with contextlib.redirect_stdout(io.StringIO()):
        dist_ref = ref.compute_dist(d["q"], d["g"], type="euclidean")
        map_ref = ref.mean_ap(distmat=dist_ref, **ids)
        cmc_ref = ref.cmc(distmat=dist_ref, topk=10, first_match_break=True, **ids)
    np.testing.assert_array_equal(O.compute_dist(d["q"], d["g"]), dist_ref)
    assert O.mean_ap(dist_ref, **ids) == map_ref
    np.testing.assert_array_equal(O.cmc(dist_ref, topk=10, first_match_break=True, **ids), cmc_ref)


RERANK_CASES = ["rerank_small", "rerank_tiny", "rerank_wide"]


@pytest.mark.parametrize("name", RERANK_CASES)
def test_re_ranking_restatement_matches_reference_fixture(golden, name):
    """oracle.re_ranking == the reference's re_ranking (reid_dataset_evaluator.py:442-519) bit for bit, for the
    default parameters and for k1=7, k2=1 (no query expansion), lambda=0.5."""
    d = golden(name)
    np.testing.assert_array_equal(O.re_ranking(d["q_g"], d["q_q"], d["g_g"]), d["rerank"])
    np.testing.assert_array_equal(O.re_ranking(d["q_g"], d["q_q"], d["g_g"], k1=7, k2=1, lambda_value=0.5),
                                  d["rerank_k7_k2_1"])
    ids = dict(query_ids=d["qid"], gallery_ids=d["gid"], query_cams=d["qcam"], gallery_cams=d["gcam"])
    assert abs(O.mean_ap(d["rerank"], **ids) - float(d["mAP"])) < 1e-12


@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference (authoring container)")
def test_re_ranking_restatement_matches_live_reference():
    from pps_b200 import synthetic
    ref = ref_loader.load()
    d = synthetic.make_reid_set(nq=25, ng=150, dim=48, n_ids=8, n_cams=3, n_distractors=10, sigma=2.5, seed=31)
    with contextlib.redirect_stdout(io.StringIO()):
        q_g, q_q, g_g = (ref.compute_dist(a, b) for a, b in ((d["q"], d["g"]), (d["q"], d["q"]), (d["g"], d["g"])))
        want = ref.re_ranking(q_g, q_q, g_g, k1=12, k2=4, lambda_value=0.2)
    np.testing.assert_array_equal(O.re_ranking(q_g, q_q, g_g, k1=12, k2=4, lambda_value=0.2), want)


@pytest.mark.parametrize("name", POOL_GOLDEN_CASES)
def test_pooling_restatement_matches_reference_graph_fixture(golden, name):
    """oracle.pps_pool == the blobs the reference's OWN graph builders (bpm_heads.py + pps_heads.py, executed eagerly by
    oracle/ref_pool_loader.py) return: same count, same order, same bits; names follow the reference's scheme."""
    from pps_b200 import pooling
    d = golden(name)
    want = [d["y%03d" % k] for k in range(int(d["n_out"]))]
    got = pool_fixture_expected(d, lambda x, n, split, mode: O.pps_pool(x, n, split=split, mode=mode), O.uniform_partition_split)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)
    names = [str(s) for s in d["names"]]
    n = int(d["strip_num"])
    if not int(d["fpn_on"]) or not int(d["train"]):
        assert names == pooling.blob_names(n, "pps")
    elif not int(d["fpn_shared"]):
        assert names == [nm for i in range(int(d["n_levels"])) for nm in pooling.blob_names(n, "pps_%d_" % i)]


@pytest.mark.skipif(not __import__("oracle.ref_pool_loader", fromlist=["x"]).available(), reason="needs /root/reference")
def test_pooling_restatement_matches_live_reference_graph():
    from oracle import ref_pool_loader as R
    x = np.maximum(np.random.RandomState(77).randn(2, 24, 24, 8), 0).astype(np.float32)
    for n, max_ave in ((6, True), (9, False), (10, True), (3, True)):
        names, arrs, dims, ops = R.run_pps_head(x, n, max_ave)
        want = O.pps_pool(x, n, split=O.uniform_partition_split(n), mode="max_ave" if max_ave else "avg_max")
        assert len(arrs) == (1 << n) - 1 and dims == [24] * len(arrs)
        np.testing.assert_array_equal(np.stack([a.reshape(2, 24) for a in arrs], 1), want)


EMBED_CASES = ["embed_n3_c48", "embed_n6_c64", "embed_n4_raw"]


@pytest.mark.parametrize("name", EMBED_CASES)
def test_embedding_restatement_matches_reference_graph_fixture(golden, name):
    """oracle.reid_embed == what the reference's own add_reid_outputs builder (reid_heads.py:34-127, executed eagerly
    one image at a time by oracle/ref_pool_loader.run_reid_outputs) produces."""
    d = golden(name)
    n = int(d["n_parts"])
    pooled = np.transpose(O.pps_pool(d["x"], n, split=[24 // n] * n, mode="max_ave"), (1, 0, 2))
    p = {k: d[k] for k in ("conv_bias", "bn_scale", "bn_bias", "bn_mean", "bn_var")}
    got = O.reid_embed(pooled, d["weight"], normalize=bool(int(d["normalize"])), **p)
    np.testing.assert_allclose(got, d["feature"], rtol=1e-12, atol=1e-14)


def test_pool_gradient_restatement_equals_finite_differences():
    """oracle.pps_pool_grad (the Caffe2 operator gradients restated) against central differences of the float64 forward:
    the CPU half of the gradient check of tests/test_gpu_pool.py (test_batch_permutation_op.py:43-50 style)."""
    rs = np.random.RandomState(0)
    for mode in ("max_ave", "avg_max"):
        for n, split in ((6, None), (5, [5, 5, 4, 5, 5])):
            x = rs.randn(1, 2, 24, 8)
            dy = rs.randn(1, (1 << n) - 1, 2)
            g = O.pps_pool_grad(x, dy, n, split, mode)
            f = lambda z: float((O.pps_pool(z, n, split, mode, dtype=np.float64) * dy).sum())
            num = np.zeros_like(x)
            eps = 1e-6
            for idx in np.ndindex(*x.shape):
                xp, xm = x.copy(), x.copy()
                xp[idx] += eps
                xm[idx] -= eps
                num[idx] = (f(xp) - f(xm)) / (2 * eps)
            np.testing.assert_allclose(g, num, rtol=0, atol=1e-6 * max(1.0, np.abs(num).max()))


@pytest.mark.parametrize("name", ["small_mid", "ragged_dim", "many_pos", "some_invalid"])
def test_cmc_separate_camera_set_matches_reference_fixture(golden, name):
    """cmc(separate_camera_set=True) (reid_dataset_evaluator.py:329-331) against the unmodified reference
    (tests/golden/cmc_sep_*.npz, oracle/make_golden_cmc_sep.py)."""
    d, s = golden(name), golden("cmc_sep_" + name)
    ids = dict(query_ids=d["qid"], gallery_ids=d["gid"], query_cams=d["qcam"], gallery_cams=d["gcam"])
    np.testing.assert_allclose(O.cmc(d["dist"], topk=10, first_match_break=True, separate_camera_set=True, **ids), s["cmc_fmb"], atol=1e-12)
    np.testing.assert_allclose(O.cmc(d["dist"], topk=20, first_match_break=False, separate_camera_set=True, **ids), s["cmc_all"], atol=1e-12)
    rows, valid = O.cmc(d["dist"], topk=10, first_match_break=True, separate_camera_set=True, average=False, **ids)
    np.testing.assert_array_equal(rows, s["cmc_rows"])
    np.testing.assert_array_equal(valid, s["cmc_valid"])


# ---- triplet-mining ops: the restatement against the reference's own operators ----
def _batch_hard_cases(golden):
    d = golden("triplet_ref_batch_hard")
    i = 0
    while "xd%d" % i in d:
        yield {k: d["%s%d" % (k, i)] for k in ("xd", "labels", "dap", "dan", "ap", "an", "dx")}
        i += 1


def test_batch_hard_restatement_matches_reference_operator_fixture(golden):
    """tests/golden/triplet_ref_batch_hard.npz holds what BatchHardOp / BatchHardGradientOp of
    /root/reference/detectron/ops/batch_hard_op.cc - compiled unmodified (oracle/build_ref_ops.py) - computed; the
    restatement must reproduce it bit for bit, the operator's stray stores for idx == -1 included."""
    n_cases = 0
    for c in _batch_hard_cases(golden):
        ap, an, ip, inn = O.batch_hard(c["xd"], c["labels"])
        np.testing.assert_array_equal(ap, c["ap"])
        np.testing.assert_array_equal(an, c["an"])
        np.testing.assert_array_equal(O.batch_hard_grad(ip, inn, c["dap"], c["dan"], stray_writes=True), c["dx"])
        # what the product computes (no stray stores) differs from the operator ONLY in the last column of the row before
        # an anchor whose search found nothing
        clean = O.batch_hard_grad(ip, inn, c["dap"], c["dan"])
        diff = np.argwhere(clean != c["dx"])
        n = len(ip)
        allowed = {(a - 1, n - 1) for a in range(1, n) if ip[a] < 0 or inn[a] < 0}
        assert {tuple(x) for x in diff} <= allowed
        n_cases += 1
    assert n_cases >= 6


def test_batch_hard_restatement_matches_live_reference_operator():
    from oracle import ref_ops
    if not ref_ops.available():
        pytest.skip("neither /root/reference nor a prebuilt oracle/_ref/libref_reid_ops.so")
    assert ref_ops.schema("BatchHard") == (2, 2) and ref_ops.schema("BatchHardGradient") == (4, 1)
    assert ref_ops.schema("PairWiseDistance") == (1, 1) and ref_ops.schema("PairWiseDistanceGradient") == (2, 1)
    # the gradient wiring the reference registers (batch_hard_op.cc:141-150, pairwise_distance_op.cc:14-24)
    assert ref_ops.gradient_def("BatchHard", 2, 2) == ("BatchHardGradient", ["I0", "I1", "GO0", "GO1"], ["I0_grad"])
    assert ref_ops.gradient_def("PairWiseDistance", 1, 1) == ("PairWiseDistanceGradient", ["I0", "GO0"], ["I0_grad"])
    rs = np.random.RandomState(3)
    for n, n_ids in [(50, 7), (17, 17), (130, 2), (3, 1)]:
        xd = np.abs(rs.randn(n, n)).astype(np.float32)
        xd[:, ::3] = np.round(xd[:, ::3] * 2) / 2
        labels = rs.randint(0, n_ids, size=n).astype(np.int32)
        dap, dan = rs.randn(n).astype(np.float32), rs.randn(n).astype(np.float32)
        ap, an = ref_ops.batch_hard(xd, labels)
        oap, oan, ip, inn = O.batch_hard(xd, labels)
        np.testing.assert_array_equal(ap, oap)
        np.testing.assert_array_equal(an, oan)
        dx, _ = ref_ops.batch_hard_grad(xd, labels, dap, dan)
        np.testing.assert_array_equal(dx, O.batch_hard_grad(ip, inn, dap, dan, stray_writes=True))
    with pytest.raises(RuntimeError, match=r"X.dim32\(0\) == X.dim32\(1\)"):           # CAFFE_ENFORCE_EQ of :16
        ref_ops.batch_hard(np.zeros((4, 5), np.float32), np.zeros(4, np.int32))


def test_pairwise_distance_restatement_matches_reference_operator_fixture(golden):
    """tests/golden/triplet_ref_pairwise.npz: outputs of the reference's CUDA operators PairWiseDistance /
    PairWiseDistanceGradient (pairwise_distance_op.cu, compiled unmodified for sm_100a and run on a B200 by
    `oracle/make_golden_triplet.py --cuda`).  float32 sums in another order (sequential over d there, atomics in the
    gradient), hence tolerances."""
    d = golden("triplet_ref_pairwise")
    i = 0
    while "x%d" % i in d:
        x, dz = d["x%d" % i], d["dz%d" % i]
        z = O.pairwise_distance(x)
        np.testing.assert_allclose(z, d["z%d" % i], rtol=1e-5, atol=1e-5)
        assert np.all(np.diag(d["z%d" % i]) == 0)
        scale = np.abs(d["dx%d" % i]).max() + 1e-6
        np.testing.assert_allclose(O.pairwise_distance_grad(x, dz), d["dx%d" % i], rtol=1e-4, atol=1e-5 * scale)
        i += 1
    assert i >= 5
