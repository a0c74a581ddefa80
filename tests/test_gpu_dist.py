"""Parity of the CUDA distance (tcgen05 split-precision GEMM + fused norm/clamp/sqrt epilogue, and the
fp32 CUDA-core variant) with the reference's compute_dist (reid_dataset_evaluator.py:244-272).
Golden matrices were produced by the unmodified reference (oracle/make_golden.py).
Tolerance: 1e-4 relative on the distance (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES
from oracle import pps_oracle as O

pytestmark = pytest.mark.gpu


def _f64_dist(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    d2 = (a * a).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2.0 * a @ b.T
    return np.sqrt(np.maximum(d2, 0.0))


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("precision", ["f16x3", "bf16x3", "bf16x6", "fp32"])
def test_compute_dist_matches_reference_fixture(golden, name, precision):
    import pps_b200
    d = golden(name)
    dist = pps_b200.compute_dist(d["q"], d["g"], precision=precision)
    assert isinstance(dist, np.ndarray) and dist.dtype == np.float32 and dist.shape == d["dist"].shape
    ref = d["dist"]
    # duplicates (dup_ties) give d ~ 0 where sqrt amplifies the last-bit noise of the fp32 cancellation
    # |a|^2+|b|^2-2ab (the reference itself is off there): relative test away from 0, absolute near it.
    np.testing.assert_allclose(dist, ref, rtol=1e-4, atol=2e-3 if name == "dup_ties" else 1e-6)
    big = ref > 1e-2
    assert np.max(np.abs(dist[big] - ref[big]) / ref[big]) < 1e-4


def test_precision_ladder_against_float64():
    """bf16x1 is visibly coarse; bf16x3 is fp32-grade; bf16x6 and fp32 are tighter still."""
    import pps_b200
    rs = np.random.RandomState(0)
    a = rs.randn(300, 520).astype(np.float32)
    b = rs.randn(700, 520).astype(np.float32)
    exact = _f64_dist(a, b)
    err = {}
    for p in ("bf16x1", "bf16x3", "bf16x6", "f16x3", "fp32"):
        got = pps_b200.compute_dist(a, b, precision=p).astype(np.float64)
        err[p] = float(np.max(np.abs(got - exact) / exact))
    ref_err = float(np.max(np.abs(O.compute_dist(a, b).astype(np.float64) - exact) / exact))
    assert err["bf16x3"] < 1e-5 and err["bf16x6"] < 2e-6 and err["fp32"] < 2e-6 and err["f16x3"] < 2e-6
    assert err["bf16x1"] > err["bf16x3"] > err["f16x3"]
    assert err["bf16x3"] < 20 * max(ref_err, 1e-7)
    assert err["f16x3"] < 4 * max(ref_err, 1e-7)          # the default split is at the level of a float32 sgemm


@pytest.mark.parametrize("m1,m2,dim", [
    (1, 1, 1), (1, 300, 64), (129, 257, 65), (128, 256, 64), (127, 255, 63), (260, 1000, 2048),
    (5, 513, 8064 // 8), (300, 10, 200),
])
def test_tile_edges_and_ragged_shapes(m1, m2, dim):
    import pps_b200
    rs = np.random.RandomState(m1 * 7 + m2)
    a = rs.randn(m1, dim).astype(np.float32)
    b = rs.randn(m2, dim).astype(np.float32)
    got = pps_b200.compute_dist(a, b)
    np.testing.assert_allclose(got, _f64_dist(a, b), rtol=1e-4, atol=1e-5)


def test_empty_inputs():
    import pps_b200
    a = np.zeros((0, 16), np.float32)
    b = np.zeros((5, 16), np.float32)
    assert pps_b200.compute_dist(a, b).shape == (0, 5)
    assert pps_b200.compute_dist(b, a).shape == (5, 0)


def test_tensor_in_tensor_out_and_fp16_inputs():
    import torch
    import pps_b200
    rs = np.random.RandomState(3)
    a = rs.randn(200, 256).astype(np.float16)
    b = rs.randn(333, 256).astype(np.float16)
    out = pps_b200.compute_dist(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    assert out.is_cuda and out.dtype == torch.float32
    np.testing.assert_allclose(out.cpu().numpy(), _f64_dist(a.astype(np.float32), b.astype(np.float32)), rtol=1e-4, atol=1e-4)


def test_cosine_branch_returns_similarity():
    import pps_b200
    rs = np.random.RandomState(4)
    a = rs.randn(50, 96).astype(np.float32)
    b = rs.randn(70, 96).astype(np.float32)
    got = pps_b200.compute_dist(a, b, type="cosine")
    np.testing.assert_allclose(got, O.compute_dist(a, b, type="cosine"), rtol=1e-4, atol=1e-5)


def test_self_distance_clamps_at_zero():
    import pps_b200
    rs = np.random.RandomState(5)
    a = rs.randn(64, 128).astype(np.float32)
    # d(a, a) is sqrt of pure cancellation noise: ~|a| * sqrt(eps).  bf16x3 drops the p1.p1 term, which
    # for identical operands is a positive bias of ~1.3e-6 |a|^2; the tensor-core paths also carry the fp32
    # accumulator rounding of interleaved large and small plane terms (~2e-6 |a|^2); fp32 is at the reference level.
    na = float(np.sqrt((a.astype(np.float64) ** 2).sum(1)).max())
    for p, bound in (("f16x3", 3e-3), ("bf16x3", 3e-3), ("bf16x6", 3e-3), ("fp32", 1e-3)):
        d = pps_b200.compute_dist(a, a, precision=p)
        assert np.all(d >= 0) and np.all(np.isfinite(d))
        assert np.max(np.diag(d)) < bound * na, p
    assert np.max(np.diag(O.compute_dist(a, a))) < 1e-3 * na


def test_error_contract():
    import pps_b200
    with pytest.raises(RuntimeError, match="not aligned"):
        pps_b200.compute_dist(np.zeros((2, 8), np.float32), np.zeros((3, 9), np.float32))
    with pytest.raises(RuntimeError, match="2-D"):
        pps_b200.compute_dist(np.zeros(8, np.float32), np.zeros((3, 8), np.float32))
    with pytest.raises(AssertionError):
        pps_b200.compute_dist(np.zeros((2, 8), np.float32), np.zeros((3, 8), np.float32), type="manhattan")
    with pytest.raises(RuntimeError, match="precision"):
        pps_b200.compute_dist(np.zeros((2, 8), np.float32), np.zeros((3, 8), np.float32), precision="int8")


def test_market_shape_distance_rows_vs_oracle():
    """Full Market-1501 shape (3 368 x 19 732 x 2048) on the GPU; a 64-query slice against the oracle."""
    import torch
    import pps_b200
    from pps_b200 import synthetic
    d = synthetic.make_config("market1501")
    dist = pps_b200.compute_dist(torch.from_numpy(d["q"]).cuda(), torch.from_numpy(d["g"]).cuda())
    assert tuple(dist.shape) == (3368, 19732)
    sel = np.arange(0, 3368, 53)[:64]
    ref = O.compute_dist(d["q"][sel], d["g"])
    np.testing.assert_allclose(dist[torch.from_numpy(sel).cuda()].cpu().numpy(), ref, rtol=1e-4, atol=1e-6)
    # symmetry property at full size: d(q, g) == d(g, q)^T up to the fp32 summation-order noise of the
    # norm add (bit-identical dot products: the split planes and the K loop order are the same)
    dt = pps_b200.compute_dist(torch.from_numpy(d["g"][:2048]).cuda(), torch.from_numpy(d["q"]).cuda())
    assert torch.allclose(dist[:, :2048], dt.t(), rtol=1e-6, atol=1e-7)


def test_single_cta_kernel_still_agrees(monkeypatch):
    """The 1-CTA 128x256 kernel (PPS_DIST_KERNEL_1CTA) and the default 2-CTA 256x256 kernel compute the same
    split product; only the order in which the plane-pair terms enter the fp32 accumulator differs."""
    import pps_b200
    from pps_b200 import _lib, evaluator
    rs = np.random.RandomState(11)
    a = rs.randn(300, 520).astype(np.float32)
    b = rs.randn(1000, 520).astype(np.float32)
    d2 = pps_b200.compute_dist(a, b)
    monkeypatch.setattr(evaluator, "DIST_KERNEL_FLAGS", _lib.DIST_KERNEL_1CTA)
    d1 = pps_b200.compute_dist(a, b)
    np.testing.assert_allclose(d1, d2, rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(d1, _f64_dist(a, b), rtol=1e-5)


def test_full_concat_width_d8064():
    """D = 8 064 = 63 x 128, the real reid_feature_concat length (reid_heads.py:95-101): 126 k-blocks per tile."""
    import pps_b200
    rs = np.random.RandomState(63)
    a = rs.randn(70, 8064).astype(np.float32)
    b = rs.randn(333, 8064).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    got = pps_b200.compute_dist(a, b)
    np.testing.assert_allclose(got, O.compute_dist(a, b), rtol=1e-4, atol=1e-6)


def test_f16x3_row_scaling_covers_the_dynamic_range():
    """'f16x3' scales every row by its own power of two before the fp16 split: rows of wildly different magnitude (1e-6
    ... 1e6), rows with a huge spread inside, zero rows and the cosine branch all stay at the fp32 level."""
    import pps_b200
    rs = np.random.RandomState(17)
    a = rs.randn(130, 300).astype(np.float32) * np.exp(rs.uniform(-14, 14, size=(130, 1))).astype(np.float32)
    b = rs.randn(270, 300).astype(np.float32) * np.exp(rs.uniform(-14, 14, size=(270, 1))).astype(np.float32)
    b[5] = 0.0
    b[6] *= np.exp(rs.uniform(-20, 0, size=300)).astype(np.float32)       # elements far below the row maximum
    exact = _f64_dist(a, b)
    for kernel in ("2cta", "1cta"):
        from pps_b200 import _lib, evaluator
        old = evaluator.DIST_KERNEL_FLAGS
        evaluator.DIST_KERNEL_FLAGS = _lib.DIST_KERNEL_1CTA if kernel == "1cta" else 0
        try:
            got = pps_b200.compute_dist(a, b, precision="f16x3")
            cos = pps_b200.compute_dist(a, b, type="cosine", precision="f16x3")
        finally:
            evaluator.DIST_KERNEL_FLAGS = old
        ref = O.compute_dist(a, b)
        scale = np.maximum(np.linalg.norm(a.astype(np.float64), axis=1)[:, None], np.linalg.norm(b.astype(np.float64), axis=1)[None, :])
        # error relative to the larger of the two norms (what a float32 |a|^2 + |b|^2 - 2ab can resolve at all)
        assert np.max(np.abs(got - exact) / scale) < 4 * max(np.max(np.abs(ref - exact) / scale), 1e-7), kernel
        keep = np.ones(270, bool); keep[5] = False
        np.testing.assert_allclose(cos[:, keep], O.compute_dist(a, b[keep], type="cosine"), rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize("m1,m2,dim", [(257, 255, 64), (512, 1000, 192), (513, 300, 2048), (1300, 2049, 320), (3368, 700, 128)])
def test_cluster4_multicast_kernel_equals_two_cta_kernel(m1, m2, dim):
    """The opt-in cluster-of-4 variant for single-plane products (PPS_DIST_CLUSTER4: two CTA pairs share the B tile by TMA
    multicast, dist_tc4_kernel).  Same MMAs in the same order as the 2-CTA kernel: identical bits, with and without the
    top-k admission epilogue; against float64 at the fp16-product level."""
    import torch
    from pps_b200 import _lib, evaluator
    lib = _lib.load()
    rs = np.random.RandomState(m1 + m2)
    a = torch.from_numpy(rs.randn(m1, dim).astype(np.float16)).cuda()
    b = torch.from_numpy(rs.randn(m2, dim).astype(np.float16)).cuda()
    sa, sb = evaluator.SplitOperand(a, 1), evaluator.SplitOperand(b, 1)
    ld = (m2 + 3) // 4 * 4
    out = {}
    for name, flags in (("cl4", _lib.DIST_CLUSTER4), ("cl2", 0)):
        d = torch.zeros((m1, ld), dtype=torch.float32, device="cuda")
        _lib.check(lib.pps_dist_tc(_lib.ptr(sa.planes), _lib.ptr(sa.sqnorm), m1, 1, 0, _lib.ptr(sb.planes), _lib.ptr(sb.sqnorm), m2, 1, 0,
                                   dim, _lib.PREC_F16X1, flags, _lib.ptr(d), ld, _lib.stream_ptr()), "pps_dist_tc")
        out[name] = d[:, :m2].cpu().numpy()
    np.testing.assert_array_equal(out["cl4"], out["cl2"])
    np.testing.assert_allclose(out["cl4"], _f64_dist(a.float().cpu().numpy(), b.float().cpu().numpy()), rtol=1e-4, atol=1e-4)
    # admission epilogue: unbounded state -> every column of every row is a candidate until the buffer is full
    k, cap = 4, 64
    bound = torch.full((m1,), -1, dtype=torch.int32, device="cuda")           # 0xffffffff: unbounded
    res = {}
    for name, flags in (("cl4", _lib.DIST_CLUSTER4), ("cl2", 0)):
        cnt = torch.zeros(m1, dtype=torch.int32, device="cuda")
        cand = torch.zeros((m1, cap), dtype=torch.int64, device="cuda")
        d = torch.zeros((m1, ld), dtype=torch.float32, device="cuda")
        _lib.check(lib.pps_dist_topk_tc(_lib.ptr(sa.planes), _lib.ptr(sa.sqnorm), m1, 1, 0, _lib.ptr(sb.planes), _lib.ptr(sb.sqnorm), m2,
                                        1, 0, dim, _lib.PREC_F16X1, flags, _lib.ptr(d), ld, 1000, _lib.ptr(bound), _lib.ptr(cnt),
                                        _lib.ptr(cand), cap, _lib.stream_ptr()), "pps_dist_topk_tc")
        res[name] = (d[:, :m2].cpu().numpy(), cnt.cpu().numpy())
    np.testing.assert_array_equal(res["cl4"][0], out["cl2"])
    np.testing.assert_array_equal(res["cl4"][1], np.full(m1, m2))
    np.testing.assert_array_equal(res["cl2"][1], np.full(m1, m2))
