"""Parity of the device k-reciprocal re-ranking (csrc/rerank.cu through the C ABI) with the reference's own
`re_ranking` outputs (tests/golden/rerank_*.npz, made by oracle/make_golden_rerank.py from the UNMODIFIED
reid_dataset_evaluator.py:442-519).  The kernels keep the reference's float32 operation order; what may differ is
np.sum's pairwise order inside one normalisation and the last bit of exp, i.e. ~1e-6 relative on the result."""
import numpy as np
import pytest

from oracle import pps_oracle as O

pytestmark = pytest.mark.gpu

CASES = ["rerank_small", "rerank_tiny", "rerank_wide"]


@pytest.mark.parametrize("name", CASES)
def test_re_ranking_matches_reference_matrices(golden, name):
    import pps_b200
    d = golden(name)
    got = pps_b200.re_ranking(d["q_g"], d["q_q"], d["g_g"])
    assert got.shape == d["rerank"].shape and got.dtype == np.float32
    np.testing.assert_allclose(got, d["rerank"], rtol=2e-5, atol=2e-6)
    got = pps_b200.re_ranking(d["q_g"], d["q_q"], d["g_g"], k1=7, k2=1, lambda_value=0.5)
    np.testing.assert_allclose(got, d["rerank_k7_k2_1"], rtol=2e-5, atol=2e-6)
    # the scores evaluate() derives from the re-ranked matrix (:174-175)
    ids = (d["qid"], d["gid"], d["qcam"], d["gcam"])
    res = pps_b200.rank_distmat(pps_b200.re_ranking(d["q_g"], d["q_q"], d["g_g"]), *ids)
    assert abs(res.mean_ap() - float(d["mAP"])) < 1e-4
    assert np.max(np.abs(res.cmc(10, True) - d["cmc"])) <= 1.0 / len(d["qid"]) + 1e-12


@pytest.mark.parametrize("name", CASES)
def test_re_ranking_from_features(golden, name):
    """One [n, n] tensor-core product of the stacked features replaces the three compute_dist calls (:165-171)."""
    import torch
    import pps_b200
    d = golden(name)
    got = pps_b200.re_ranking_from_features(torch.from_numpy(d["q"]).cuda(), torch.from_numpy(d["g"]).cuda())
    assert got.is_cuda
    got = got.cpu().numpy()
    # distances differ from the reference's sgemm in the last bits, so a neighbour list can differ at a near tie:
    # the bulk must agree tightly, every element loosely
    err = np.abs(got - d["rerank"])
    assert np.mean(err < 1e-4) > 0.995, float(np.mean(err < 1e-4))
    assert err.max() < 0.2


def test_evaluate_with_re_ranking(golden):
    import pps_b200
    d = golden("rerank_small")
    feats = np.concatenate([d["q"], d["g"]], 0)
    ids = np.concatenate([d["qid"], d["gid"]])
    cams = np.concatenate([d["qcam"], d["gcam"]])
    marks = np.concatenate([np.zeros(len(d["qid"]), np.int64), np.ones(len(d["gid"]), np.int64)])
    plain = pps_b200.evaluate_arrays(feats, ids, cams, marks)
    rr = pps_b200.evaluate_arrays(feats, ids, cams, marks, to_re_rank=True)
    assert abs(rr[0] - float(d["mAP"])) < 2e-3            # the reference's re-ranked mAP
    assert rr[0] != plain[0] and rr[2] is None


def test_re_ranking_argument_errors():
    import pps_b200
    with pytest.raises(RuntimeError):
        pps_b200.re_ranking(np.zeros((3, 5), np.float32), np.zeros((2, 2), np.float32), np.zeros((5, 5), np.float32))
    with pytest.raises(RuntimeError):                      # k1 + 1 neighbours of 8 images
        pps_b200.re_ranking(np.ones((3, 5), np.float32), np.ones((3, 3), np.float32), np.ones((5, 5), np.float32))


def test_re_ranking_larger_set_runs_and_is_consistent():
    """A few thousand images: the whole re-ranking from features in one go; sanity on values + determinism."""
    import torch
    import pps_b200
    from pps_b200 import synthetic
    d = synthetic.make_reid_set(nq=300, ng=2500, dim=256, n_ids=100, n_cams=6, n_distractors=200, sigma=3.0, seed=5)
    q, g = torch.from_numpy(d["q"]).cuda(), torch.from_numpy(d["g"]).cuda()
    a = pps_b200.re_ranking_from_features(q, g)
    b = pps_b200.re_ranking_from_features(q, g)
    assert torch.equal(a, b)                               # no float atomics anywhere: bit-reproducible
    a = a.cpu().numpy()
    assert np.isfinite(a).all() and a.min() >= 0.0 and a.max() <= 1.0 + 1e-6
    sub = slice(0, 40)                                     # the oracle on a query subset would change the neighbourhoods:
    want = O.re_ranking(*(pps_b200.compute_dist(x, y) for x, y in ((d["q"], d["g"]), (d["q"], d["q"]), (d["g"], d["g"]))))
    err = np.abs(a - want)
    assert np.mean(err < 1e-4) > 0.995 and err[sub].max() < 0.2
    res = pps_b200.rank_distmat(a, d["qid"], d["gid"], d["qcam"], d["gcam"])
    plain = pps_b200.rank_eval(q, g, d["qid"], d["gid"], d["qcam"], d["gcam"])
    assert res.mean_ap() > plain.mean_ap()                 # re-ranking helps on this clustered synthetic set


def test_rerank_c_abi_degenerate_inputs():
    from pps_b200 import _lib
    lib = _lib.load()
    s = _lib.stream_ptr()
    assert lib.pps_rerank_vcap() == 256
    assert lib.pps_rerank_normalize(None, 0, 0, None, None, 0, s) == 0
    assert lib.pps_rerank_krecip(None, 21, 0, 20, None, 0, None, None, None, s) == 0
    assert lib.pps_rerank_krecip(None, 10, 0, 20, None, 0, None, None, None, s) == _lib.PPS_ERR_INVALID_ARG      # rank_cols < k1 + 1
    assert lib.pps_rerank_krecip(None, 40, 0, 39, None, 0, None, None, None, s) == _lib.PPS_ERR_UNSUPPORTED      # k1 + 1 > 32
    assert lib.pps_rerank_expand(None, 21, 0, 9, None, None, None, 2304, None, None, None, s) == _lib.PPS_ERR_INVALID_ARG
    assert lib.pps_rerank_jaccard(None, None, None, 256, None, None, None, 0, 5, None, 5, 0.3, None, 5, s) == 0


def test_re_ranking_no_query_expansion_and_small_k():
    import pps_b200
    rs = np.random.RandomState(8)
    f = rs.randn(60, 16).astype(np.float32)
    q, g = f[:10], f[10:]
    mats = (O.compute_dist(q, g), O.compute_dist(q, q), O.compute_dist(g, g))
    for k1, k2, lam in ((5, 1, 0.0), (3, 2, 1.0), (12, 3, 0.3)):
        got = pps_b200.re_ranking(*mats, k1=k1, k2=k2, lambda_value=lam)
        np.testing.assert_allclose(got, O.re_ranking(*mats, k1=k1, k2=k2, lambda_value=lam), rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("name", ["evaluate_small", "evaluate_mixed_order"])
def test_evaluate_matches_reference_evaluate_fixture(golden, name):
    """pps_b200.evaluate(json_dataset, all_feats, output_dir) against the UNMODIFIED reference evaluate()
    (reid_dataset_evaluator.py:29-209; oracle/make_golden_evaluate.py): image-name parsing, mark split, single-query
    scores with cfg.REID.RERANK = False and the re-ranked scores with True."""
    import pps_b200
    d = golden(name)

    class Dataset:
        def get_roidb(self, gt=True):
            return [{"image": str(im), "mark": int(m)} for im, m in zip(d["images"], d["marks"])]

    nq = int((d["marks"] == 0).sum())
    mAP, cmc, mq_mAP, mq_cmc = pps_b200.evaluate(Dataset(), d["feats"], None, verbose=False)
    assert mq_mAP is None and mq_cmc is None
    assert abs(mAP - float(d["mAP"])) < 1e-6
    assert np.max(np.abs(cmc - d["cmc"])) <= 1.0 / nq + 1e-12
    mAP, cmc, _, _ = pps_b200.evaluate(Dataset(), d["feats"], None, verbose=False, to_re_rank=True)
    assert abs(mAP - float(d["mAP_rerank"])) < 2e-3
    assert np.max(np.abs(cmc - d["cmc_rerank"])) <= 2.0 / nq + 1e-12
    res = pps_b200.reid_results((mAP, cmc, None, None), name="toy")
    assert res["toy"]["ReID"]["mAP"] == mAP and res["toy"]["ReID"]["CMC5"] == cmc[4] and res["toy"]["ReID"]["mq_mAP"] == -1
