"""Parity of the grouped tcgen05 embedding (pps_embed_tc + pps_l2_normalize_rows through the C ABI) with the
oracle's float64 restatement of reid_heads.py:34-127 (Conv1x1 + SpatialBN(test) + ReLU, Concat, Normalize).
Tolerance: 1e-5 relative to the row scale (the pooled-feature tolerance of BASELINE.json north_star)."""
import numpy as np
import pytest

from oracle import pps_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


def _params(K, E, C, seed):
    rs = np.random.RandomState(seed)
    w = (rs.randn(K, E, C) * np.sqrt(2.0 / C)).astype(np.float32)            # MSRAFill (reid_heads.py:48)
    p = dict(conv_bias=(0.1 * rs.randn(K, E)).astype(np.float32), bn_scale=(1 + 0.2 * rs.randn(K, E)).astype(np.float32),
             bn_bias=(0.3 * rs.randn(K, E)).astype(np.float32), bn_mean=(0.2 * rs.randn(K, E)).astype(np.float32),
             bn_var=(0.5 + rs.rand(K, E)).astype(np.float32))
    return w, p


def _head(w, p, precision="bf16x3"):
    import pps_b200
    alpha, beta = pps_b200.fold_bn(p["conv_bias"], p["bn_scale"], p["bn_bias"], p["bn_mean"], p["bn_var"])
    return pps_b200.ReidEmbedHead(w, alpha, beta, precision=precision)


@pytest.mark.parametrize("K,N,C", [(63, 37, 2048), (31, 300, 256), (3, 513, 100), (1, 1, 64)])
def test_embed_matches_oracle(torch, K, N, C):
    E = 128
    rs = np.random.RandomState(K + N)
    pooled = np.abs(rs.randn(K, N, C)).astype(np.float32)
    w, p = _params(K, E, C, seed=C)
    head = _head(w, p)
    raw = head(torch.from_numpy(pooled).cuda(), normalize=False).cpu().numpy()
    want_raw = O.reid_embed(pooled, w, normalize=False, **p)
    assert raw.shape == (N, K * E)
    scale = np.abs(want_raw).max()
    np.testing.assert_allclose(raw, want_raw, rtol=1e-5, atol=1e-5 * scale)
    assert (raw >= 0).all() and (raw == 0).any()                              # ReLU really clips something
    got = head(torch.from_numpy(pooled).cuda(), normalize=True).cpu().numpy()
    want = O.reid_embed(pooled, w, normalize=True, **p)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5 * np.abs(want).max())
    np.testing.assert_allclose(np.linalg.norm(got.astype(np.float64), axis=1), 1.0, atol=1e-6)


def test_embed_precisions_ordered(torch):
    K, N, C, E = 5, 200, 512, 128
    rs = np.random.RandomState(0)
    pooled = np.abs(rs.randn(K, N, C)).astype(np.float32)
    w, p = _params(K, E, C, seed=9)
    want = O.reid_embed(pooled, w, normalize=False, **p)
    errs = {}
    for prec in ("bf16x1", "bf16x3", "bf16x6"):
        got = _head(w, p, prec)(torch.from_numpy(pooled).cuda(), normalize=False).cpu().numpy()
        errs[prec] = np.abs(got - want).max() / np.abs(want).max()
    # the split passes recover the operand rounding; what is left is the fp32 accumulation inside the tensor core
    assert errs["bf16x1"] < 2e-2 and errs["bf16x3"] < 1e-5 and errs["bf16x6"] < 1e-5, errs
    assert errs["bf16x3"] < 0.1 * errs["bf16x1"]


def test_pool_then_embed_pipeline(torch):
    """conv5 maps -> pps_pool ([K, N, C] blobs) -> add_reid_outputs -> [N, 63*128] normalised feature."""
    import pps_b200
    rs = np.random.RandomState(4)
    x = np.maximum(rs.randn(6, 256, 24, 8), 0).astype(np.float32)
    cfg = pps_b200.ReIDPoolCfg(MAX_AVE_FEATURE=True)                          # every shipped PPS yaml
    blobs, dims = pps_b200.add_pps_part_head(torch.from_numpy(x).cuda(), 256, 1.0 / 16, cfg)
    assert len(blobs) == 63 and dims == [256] * 63
    w, p = _params(63, 128, 256, seed=11)
    feat = pps_b200.add_reid_outputs(blobs, _head(w, p)).cpu().numpy()
    pooled = np.transpose(O.pps_pool(x, 6, mode="max_ave"), (1, 0, 2))        # [K, N, C]
    want = O.reid_embed(pooled, w, normalize=True, **p)
    np.testing.assert_allclose(feat, want, rtol=1e-5, atol=1e-5 * np.abs(want).max())


@pytest.mark.parametrize("shape,n_parts,split,mode", [((6, 256, 24, 8), 6, None, "max_ave"), ((3, 128, 24, 8), 5, [5, 5, 4, 5, 5], "avg_max"),
                                                     ((2, 100, 12, 4), 3, None, "max_ave")])
def test_pool_into_planes_equals_pool_then_split(torch, shape, n_parts, split, mode):
    """embed_maps (pooling kernel writes the bf16 operand planes itself) == pps_pool -> split -> embed, bit for bit;
    C = 100 exercises the zero K padding of the planes."""
    import pps_b200
    rs = np.random.RandomState(shape[0])
    x = torch.from_numpy(np.maximum(rs.randn(*shape), 0).astype(np.float32)).cuda()
    K = (1 << n_parts) - 1
    w, p = _params(K, 128, shape[1], seed=3)
    head = _head(w, p)
    pooled = pps_b200.pps_pool(x, n_parts=n_parts, split=split, mode=mode, layout="knc")
    want = head(pooled, normalize=True)
    got = pps_b200.embed_maps(head, x, n_parts=n_parts, split=split, mode=mode, normalize=True)
    assert torch.equal(got, want)
    ref = O.reid_embed(np.transpose(O.pps_pool(x.cpu().numpy(), n_parts, split=split, mode=mode), (1, 0, 2)), w, normalize=True, **p)
    np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max())


def test_normalize_rows_edge_cases(torch):
    import pps_b200
    x = torch.zeros((3, 10), device="cuda")
    x[1] = torch.arange(10, device="cuda", dtype=torch.float32)
    y = pps_b200.l2_normalize_rows(x).cpu().numpy()
    assert (y[0] == 0).all() and (y[2] == 0).all()                            # |x| = 0 -> x / 1e-12 = 0, no NaN
    np.testing.assert_allclose(y[1], np.arange(10) / np.sqrt(285.0), rtol=1e-6)
    big = torch.randn((7, 8064), device="cuda")[:, :8061]                     # strided rows, odd width
    np.testing.assert_allclose(pps_b200.l2_normalize_rows(big).cpu().numpy(),
                               (big / big.norm(dim=1, keepdim=True)).cpu().numpy(), rtol=2e-6, atol=1e-8)


def test_embed_argument_errors(torch):
    w, p = _params(2, 128, 64, seed=1)
    head = _head(w, p)
    with pytest.raises(RuntimeError):
        head(torch.zeros((3, 4, 64), device="cuda"))                          # K mismatch
    with pytest.raises(RuntimeError):
        head(torch.zeros((2, 4, 64)))                                         # host tensor: no CPU path
    w2, p2 = _params(2, 64, 64, seed=1)
    with pytest.raises(RuntimeError):
        _head(w2, p2)(torch.zeros((2, 4, 64), device="cuda"))                 # E != 128 -> PPS_ERR_UNSUPPORTED


@pytest.mark.parametrize("name", ["embed_n3_c48", "embed_n6_c64", "embed_n4_raw"])
def test_embed_matches_reference_graph_fixture(torch, golden, name):
    """conv5 maps -> pool -> embed (-> normalise) on the device against what the reference's own add_reid_outputs
    builder produces for the same parameters (tests/golden/embed_*.npz, oracle/make_golden_embed.py)."""
    import pps_b200
    d = golden(name)
    n = int(d["n_parts"])
    p = {k: d[k] for k in ("conv_bias", "bn_scale", "bn_bias", "bn_mean", "bn_var")}
    head = _head(d["weight"], p)
    x = torch.from_numpy(d["x"]).cuda()
    normalize = bool(int(d["normalize"]))
    want = d["feature"]
    tol = 1e-5 * np.abs(want).max()
    got = pps_b200.embed_maps(head, x, n_parts=n, mode="max_ave", normalize=normalize).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=tol)
    blobs, _ = pps_b200.add_pps_part_head(x, int(x.shape[1]), 1.0 / 16, pps_b200.ReIDPoolCfg(BPM_STRIP_NUM=n, MAX_AVE_FEATURE=True))
    got2 = pps_b200.add_reid_outputs(blobs, head, normalize=normalize).cpu().numpy()
    np.testing.assert_allclose(got2, want, rtol=1e-5, atol=tol)
