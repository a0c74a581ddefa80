"""Parity of the fused CUDA pooling (pps_pool_fwd through the C ABI) with the oracle's restatement of
bpm_heads.py:18-55 + pps_heads.py:38-142.  Tolerance: 1e-5 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

from oracle import pps_oracle as O

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


def _x(shape, seed=0, relu=True):
    rs = np.random.RandomState(seed)
    x = rs.randn(*shape)
    if relu:
        x = np.maximum(x, 0.0)
    return x.astype(np.float32)


def _run(torch, x, **kw):
    import pps_b200
    y = pps_b200.pps_pool(torch.from_numpy(x).cuda(), **kw)
    torch.cuda.synchronize()
    return y.cpu().numpy()


@pytest.mark.parametrize("mode", ["max_ave", "avg_max"])
def test_conv5_n6_all_63_combos(torch, mode):
    x = _x((5, 2048, 24, 8), seed=1)
    y = _run(torch, x, n_parts=6, mode=mode)
    ref = O.pps_pool(x, 6, mode=mode)
    assert y.shape == (5, 63, 2048)
    np.testing.assert_allclose(y, ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("mode", ["max_ave", "avg_max"])
def test_shipped_config_n5_uneven_split(torch, mode):
    # configs/market1501/pps_R-50_1x.yaml: BPM_STRIP_NUM 5 -> [5,5,4,5,5] (bpm_heads.py:29-32)
    x = _x((3, 256, 24, 8), seed=2)
    split = O.uniform_partition_split(5)
    y = _run(torch, x, n_parts=5, split=split, mode=mode)
    np.testing.assert_allclose(y, O.pps_pool(x, 5, split=split, mode=mode), rtol=RTOL, atol=ATOL)
    assert y.shape == (3, 31, 256)


def test_negative_inputs_and_max_semantics(torch):
    # not post-ReLU: max / mean must not assume non-negative values
    x = _x((2, 64, 24, 8), seed=3, relu=False) - 3.0
    for mode in ("max_ave", "avg_max"):
        np.testing.assert_allclose(_run(torch, x, n_parts=6, mode=mode), O.pps_pool(x, 6, mode=mode), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("shape,n_parts,split", [
    ((2, 256, 48, 16), 6, [8] * 6),          # pyramid level res3 (FPN_reid.py:411-416), scale 1/8
    ((2, 40, 24, 8), 6, None),               # C not a multiple of the 32-channel unit
    ((3, 33, 24, 8), 6, None),
    ((2, 16, 12, 4), 3, [4, 4, 4]),
    ((1, 8, 10, 8), 10, [1] * 10),           # PPS_POOL_MAX_PARTS strips -> 1023 combos
    ((2, 32, 24, 7), 6, None),               # W % 4 != 0 -> generic kernel
    ((2, 32, 7, 3), 2, [3, 4]),
    ((1, 4, 96, 96), 6, [16] * 6),           # plane larger than a ring slot -> generic kernel
    ((1, 32, 1, 4), 1, [1]),
])
def test_shapes_and_edge_cases(torch, shape, n_parts, split):
    x = _x(shape, seed=sum(shape))
    for mode in ("max_ave", "avg_max"):
        y = _run(torch, x, n_parts=n_parts, split=split, mode=mode)
        np.testing.assert_allclose(y, O.pps_pool(x, n_parts, split=split, mode=mode), rtol=RTOL, atol=ATOL)


def test_many_units_per_cta_persistent_loop(torch):
    # more units than SMs x ring depth: exercises ring wrap-around, phase flips and the double buffer
    x = _x((40, 512, 24, 8), seed=4)
    y = _run(torch, x, n_parts=6, mode="max_ave")
    np.testing.assert_allclose(y, O.pps_pool(x, 6, mode="max_ave"), rtol=RTOL, atol=ATOL)


def test_empty_batch(torch):
    import pps_b200
    y = pps_b200.pps_pool(torch.zeros((0, 64, 24, 8), device="cuda"), n_parts=6)
    assert tuple(y.shape) == (0, 63, 64)


def test_explicit_combos_pyramid_list_and_knc_layout(torch):
    import pps_b200
    x = _x((3, 64, 24, 8), seed=5)
    masks = [pps_b200.comb_to_mask(c) for c in pps_b200.pyramid_combs]
    y = _run(torch, x, n_parts=6, mode="max_ave", combos=masks, layout="knc")
    ref = O.pps_pool(x, 6, mode="max_ave", combos=masks)
    assert y.shape == (21, 3, 64)
    np.testing.assert_allclose(y.transpose(1, 0, 2), ref, rtol=RTOL, atol=ATOL)


def test_head_function_contract(torch):
    """add_pps_part_head_(blob_in, dim_in, spatial_scale) -> (blobs_out, dims_out): pps_heads.py:38-80."""
    import pps_b200
    x = _x((4, 128, 24, 8), seed=6)
    cfg = pps_b200.ReIDPoolCfg(BPM_STRIP_NUM=6, MAX_AVE_FEATURE=True)
    blobs, dims = pps_b200.add_pps_part_head(torch.from_numpy(x).cuda(), 128, 1.0 / 16, cfg)
    assert len(blobs) == 63 and dims == [128] * 63
    ref = O.pps_pool(x, 6, mode="max_ave")
    for k in (0, 1, 2, 30, 62):
        assert tuple(blobs[k].shape) == (4, 128, 1, 1)
        np.testing.assert_allclose(blobs[k].cpu().numpy()[:, :, 0, 0], ref[:, k, :], rtol=RTOL, atol=ATOL)


def test_multi_scale_heads(torch):
    """pps_heads.py:83-142: test -> level 0 only; train -> every level; FPN_SHARED -> batch concat."""
    import pps_b200
    levels = [_x((2, 256, 24, 8), 7), _x((2, 256, 24, 8), 8), _x((2, 256, 48, 16), 9)]
    scales = [1.0 / 16, 1.0 / 16, 1.0 / 8]
    blobs_in = [torch.from_numpy(v).cuda() for v in levels]
    refs = [O.pps_pool(v, 6, split=O.uniform_partition_split(6, 384, s), mode="max_ave") for v, s in zip(levels, scales)]

    cfg = pps_b200.ReIDPoolCfg(MAX_AVE_FEATURE=True, FPN_ON=True, train=False)
    blobs, dims = pps_b200.add_pps_part_head(blobs_in, [256] * 3, scales, cfg)
    assert len(blobs) == 63
    np.testing.assert_allclose(blobs[62].cpu().numpy()[:, :, 0, 0], refs[0][:, 62], rtol=RTOL, atol=ATOL)

    cfg = pps_b200.ReIDPoolCfg(MAX_AVE_FEATURE=True, FPN_ON=True, train=True)
    blobs, dims = pps_b200.add_pps_part_head(blobs_in, [256] * 3, scales, cfg)
    assert len(blobs) == 3 * 63 and len(dims) == 3 * 63
    for lvl in range(3):
        for k in (0, 17, 62):
            np.testing.assert_allclose(blobs[lvl * 63 + k].cpu().numpy()[:, :, 0, 0], refs[lvl][:, k], rtol=RTOL, atol=ATOL)

    cfg = pps_b200.ReIDPoolCfg(MAX_AVE_FEATURE=True, FPN_ON=True, train=True, FPN_SHARED=True)
    blobs, dims = pps_b200.add_pps_part_head(blobs_in, [256] * 3, scales, cfg)
    assert len(blobs) == 63 and tuple(blobs[0].shape) == (6, 256, 1, 1)
    for k in (0, 31, 62):
        want = np.concatenate([r[:, k] for r in refs], axis=0)
        np.testing.assert_allclose(blobs[k].cpu().numpy()[:, :, 0, 0], want, rtol=RTOL, atol=ATOL)


def test_error_contract(torch):
    """Bad shapes raise RuntimeError, as CAFFE_ENFORCE does (detectron/tests/test_zero_even_op.py:50-53)."""
    import pps_b200
    x = torch.zeros((1, 32, 24, 8), device="cuda")
    with pytest.raises(RuntimeError, match="shape"):
        pps_b200.pps_pool(x, n_parts=6, split=[4, 4, 4, 4, 4, 5])        # sum(split) != H
    with pytest.raises(RuntimeError):
        pps_b200.pps_pool(x, n_parts=6, split=[4] * 5)
    with pytest.raises(RuntimeError, match="ndim"):
        pps_b200.pps_pool(x[0], n_parts=6)
    with pytest.raises(RuntimeError, match="float32"):
        pps_b200.pps_pool(x.half(), n_parts=6)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        pps_b200.pps_pool(x.cpu(), n_parts=6)
    with pytest.raises(RuntimeError, match="shape"):
        pps_b200.pps_pool(x, n_parts=11, split=[2] * 10 + [4])
    with pytest.raises(RuntimeError, match="shape"):
        pps_b200.pps_pool(x, n_parts=6, combos=[0, 1])                   # mask 0 is not a combination


def test_full_size_property_linearity_in_batch(torch):
    """Market-shaped batch (1 024 maps = 1.6 GB): each image's result is independent of its batch
    neighbours (pool(batch)[i] == pool(batch[i:i+1])), and the first images match the oracle."""
    import pps_b200
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((1024, 2048, 24, 8), device="cuda", generator=g).clamp_min_(0)
    y = pps_b200.pps_pool(x, n_parts=6, mode="max_ave")
    for i in (0, 511, 1023):
        yi = pps_b200.pps_pool(x[i:i + 1].contiguous(), n_parts=6, mode="max_ave")
        assert torch.equal(y[i:i + 1], yi)
    ref = O.pps_pool(x[:2].cpu().numpy(), 6, mode="max_ave")
    np.testing.assert_allclose(y[:2].cpu().numpy(), ref, rtol=RTOL, atol=ATOL)


from conftest import POOL_GOLDEN_CASES, pool_fixture_expected  # noqa: E402


@pytest.mark.parametrize("name", POOL_GOLDEN_CASES)
def test_heads_match_reference_graph_fixture(torch, golden, name):
    """pps_b200.add_pps_part_head (fused CUDA kernel) against the blobs the reference's own graph builders return
    (tests/golden/pool_*.npz, oracle/make_golden_pool.py): count, order, shapes, values to 1e-5."""
    import pps_b200
    d = golden(name)
    cfg = pps_b200.ReIDPoolCfg(BPM_STRIP_NUM=int(d["strip_num"]), MAX_AVE_FEATURE=bool(int(d["max_ave"])),
                               FPN_ON=bool(int(d["fpn_on"])), FPN_SHARED=bool(int(d["fpn_shared"])), train=bool(int(d["train"])))
    levels = [torch.from_numpy(d["x%d" % j]).cuda() for j in range(int(d["n_levels"]))]
    scales = [float(v) for v in d["spatial_scale"]]
    if cfg.FPN_ON:
        blobs, dims = pps_b200.add_pps_part_head(levels, [int(x.shape[1]) for x in levels], scales, cfg)
    else:
        blobs, dims = pps_b200.add_pps_part_head(levels[0], int(levels[0].shape[1]), scales[0], cfg)
    want = [d["y%03d" % k] for k in range(int(d["n_out"]))]
    assert len(blobs) == len(want) and [int(v) for v in dims] == [int(v) for v in d["dims"]]
    for got, ref in zip(blobs, want):
        assert tuple(got.shape) == ref.shape
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=RTOL, atol=ATOL)


# ---- backward (pps_pool_bwd): the reference trains through this sub-graph; its multi-scale branch is train-only ----
@pytest.mark.parametrize("mode", ["max_ave", "avg_max"])
@pytest.mark.parametrize("shape,n_parts,split", [
    ((3, 70, 24, 8), 6, None), ((2, 33, 24, 8), 5, [5, 5, 4, 5, 5]), ((2, 40, 48, 16), 6, [8] * 6), ((1, 5, 7, 3), 3, [2, 4, 1]),
])
def test_pool_backward_matches_oracle_gradient(torch, mode, shape, n_parts, split):
    """dX of the fused kernel against the oracle's restated operator gradients (float64), 1e-5 relative to the gradient
    scale; the oracle gradient itself is pinned to finite differences in tests/test_oracle_golden.py."""
    import pps_b200
    rs = np.random.RandomState(sum(shape) + n_parts)
    x = rs.randn(*shape).astype(np.float32)
    K = (1 << n_parts) - 1
    dy = rs.randn(shape[0], K, shape[1]).astype(np.float32)
    dx = pps_b200.pps_pool_backward(torch.from_numpy(x).cuda(), torch.from_numpy(dy).cuda(), n_parts=n_parts, split=split,
                                    mode=mode).cpu().numpy()
    want = O.pps_pool_grad(x, dy, n_parts, split, mode)
    np.testing.assert_allclose(dx, want, rtol=1e-5, atol=1e-5 * np.abs(want).max())


def test_pool_backward_numeric_gradient_check(torch):
    """Gradient check in the style of detectron/tests/test_batch_permutation_op.py:43-50 (GradientChecker.CheckSimple):
    central differences of the float64 oracle forward against the CUDA backward, on the shipped configuration
    (n = 6, max_ave) and on the explicit pyramid list."""
    import pps_b200
    rs = np.random.RandomState(9)
    x = rs.randn(1, 4, 24, 8).astype(np.float32)
    for combos in (None, [pps_b200.comb_to_mask(c) for c in pps_b200.pyramid_combs]):
        K = 63 if combos is None else len(combos)
        dy = rs.randn(1, K, 4).astype(np.float32)
        dx = pps_b200.pps_pool_backward(torch.from_numpy(x).cuda(), torch.from_numpy(dy).cuda(), n_parts=6, mode="max_ave",
                                        combos=combos).cpu().numpy()
        f = lambda z: float((O.pps_pool(z, 6, mode="max_ave", combos=combos, dtype=np.float64) * dy).sum())
        x64 = x.astype(np.float64)
        num = np.zeros_like(x64)
        eps = 1e-4
        for idx in np.ndindex(*x.shape):
            xp, xm = x64.copy(), x64.copy()
            xp[idx] += eps
            xm[idx] -= eps
            num[idx] = (f(xp) - f(xm)) / (2 * eps)
        np.testing.assert_allclose(dx, num, rtol=1e-3, atol=1e-3 * np.abs(num).max())


def test_pool_autograd_through_the_multiscale_head(torch):
    """The train-time multi-scale branch (pps_heads.py:106-135) with gradients: the three pyramid levels, FPN_SHARED
    Concat(axis=0), loss = sum of all blobs weighted; dX per level equals the oracle gradient of that level."""
    import pps_b200
    rs = np.random.RandomState(13)
    shapes = [(2, 64, 24, 8), (2, 64, 24, 8), (2, 64, 48, 16)]
    scales = [1. / 16, 1. / 16, 1. / 8]
    xs = [rs.randn(*s).astype(np.float32) for s in shapes]
    for shared in (False, True):
        cfg = pps_b200.ReIDPoolCfg(BPM_STRIP_NUM=6, MAX_AVE_FEATURE=True, FPN_ON=True, FPN_SHARED=shared, train=True)
        ts = [torch.from_numpy(x).cuda().requires_grad_(True) for x in xs]
        blobs, _ = pps_b200.add_pps_part_head(ts, [64] * 3, scales, cfg)
        assert len(blobs) == (63 if shared else 189)
        w = [torch.from_numpy(rs.randn(*tuple(b.shape)).astype(np.float32)).cuda() for b in blobs]
        loss = sum((b * wi).sum() for b, wi in zip(blobs, w))
        loss.backward()
        for lvl, (x, t) in enumerate(zip(xs, ts)):
            split = O.uniform_partition_split(6, 384, scales[lvl])
            if shared:          # blob k = Concat over levels along the batch axis
                dy = np.stack([wi.cpu().numpy()[2 * lvl:2 * lvl + 2, :, 0, 0] for wi in w], axis=1)
            else:               # level-major list
                dy = np.stack([w[63 * lvl + k].cpu().numpy()[:, :, 0, 0] for k in range(63)], axis=1)
            want = O.pps_pool_grad(x, dy, 6, split, "max_ave")
            np.testing.assert_allclose(t.grad.cpu().numpy(), want, rtol=1e-5, atol=1e-5 * np.abs(want).max())


def test_pool_backward_error_contract(torch):
    import pps_b200
    x = torch.zeros((1, 4, 24, 8), device="cuda")
    with pytest.raises(RuntimeError, match="dy must be"):
        pps_b200.pps_pool_backward(x, torch.zeros((1, 62, 4), device="cuda"))
    with pytest.raises(RuntimeError, match="shape check"):
        pps_b200.pps_pool_backward(x, torch.zeros((1, 63, 4), device="cuda"), split=[4, 4, 4, 4, 4, 5])
