"""Triplet-mining ops against the oracle's restatement of
detectron/ops/pairwise_distance_op.cu and detectron/ops/batch_hard_op.cc, against fixtures written by the reference's own
operators (compiled unmodified: oracle/build_ref_ops.py, tests/golden/triplet_ref_*.npz) and, where the compiled library
is present, against those operators run live on the same device.  Test pattern follows the reference's op tests
(detectron/tests/test_batch_permutation_op.py: forward vs NumPy at rtol 1e-5, shape-contract errors)."""
import numpy as np
import pytest

from oracle import pps_oracle as O

pytestmark = pytest.mark.gpu


def _data(n, d, n_ids, seed):
    rs = np.random.RandomState(seed)
    x = rs.randn(n, d).astype(np.float32)
    labels = rs.randint(0, n_ids, size=n).astype(np.int32)
    return x, labels


@pytest.mark.parametrize("n,d", [(64, 128), (256, 128), (33, 7), (1, 5), (100, 300)])
def test_pairwise_distance_forward_and_gradient(n, d):
    import torch
    from pps_b200 import triplet
    x, _ = _data(n, d, 4, n + d)
    z = triplet.pairwise_distance(torch.from_numpy(x).cuda()).cpu().numpy()
    np.testing.assert_allclose(z, O.pairwise_distance(x), rtol=1e-5, atol=1e-5)
    assert np.all(np.diag(z) == 0)
    dz = np.random.RandomState(1).randn(n, n).astype(np.float32)
    dx = triplet.pairwise_distance_grad(torch.from_numpy(x).cuda(), torch.from_numpy(dz).cuda()).cpu().numpy()
    np.testing.assert_allclose(dx, O.pairwise_distance_grad(x, dz), rtol=1e-4, atol=1e-3)
    # against autograd of the same function
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    ((xt[:, None, :] - xt[None, :, :]) ** 2).sum(-1).backward(torch.from_numpy(dz).cuda())
    np.testing.assert_allclose(dx, xt.grad.cpu().numpy(), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("n,n_ids", [(64, 16), (256, 64), (37, 3), (8, 8), (5, 1)])
def test_batch_hard_forward_indices_and_gradient(n, n_ids):
    import torch
    from pps_b200 import triplet
    x, labels = _data(n, 32, n_ids, n)
    xd = O.pairwise_distance(x)
    xd[:, ::5] = np.round(xd[:, ::5])                 # exact ties: the first index must win
    ap, an, ip, inn = triplet.batch_hard(torch.from_numpy(xd).cuda(), torch.from_numpy(labels).cuda(), return_indices=True)
    oap, oan, oip, oin = O.batch_hard(xd, labels)
    np.testing.assert_array_equal(ap.cpu().numpy(), oap)
    np.testing.assert_array_equal(an.cpu().numpy(), oan)
    np.testing.assert_array_equal(ip.cpu().numpy(), oip)
    np.testing.assert_array_equal(inn.cpu().numpy(), oin)
    rs = np.random.RandomState(2)
    dap, dan = rs.randn(n).astype(np.float32), rs.randn(n).astype(np.float32)
    dx = triplet.batch_hard_grad(ip, inn, torch.from_numpy(dap).cuda(), torch.from_numpy(dan).cuda()).cpu().numpy()
    np.testing.assert_array_equal(dx, O.batch_hard_grad(oip, oin, dap, dan))


def test_fused_mining_equals_two_step():
    import torch
    from pps_b200 import triplet
    x, labels = _data(192, 128, 48, 9)
    xt, lt = torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda()
    ap, an, ip, inn = triplet.batch_hard_from_features(xt, lt)
    oap, oan, oip, oin = O.batch_hard(O.pairwise_distance(x), labels)
    np.testing.assert_allclose(ap.cpu().numpy(), oap, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(an.cpu().numpy(), oan, rtol=1e-5, atol=1e-5)
    assert np.mean(ip.cpu().numpy() == oip) > 0.98 and np.mean(inn.cpu().numpy() == oin) > 0.98


def test_error_contract():
    import torch
    from pps_b200 import triplet
    x = torch.zeros((4, 8), device="cuda")
    with pytest.raises(RuntimeError, match="dim"):
        triplet.pairwise_distance(x[0])                                   # CAFFE_ENFORCE_EQ(X.dim(), 2)
    with pytest.raises(RuntimeError, match="dim32"):
        triplet.batch_hard(torch.zeros((4, 5), device="cuda"), torch.zeros(4, dtype=torch.int32, device="cuda"))
    with pytest.raises(RuntimeError, match="dim32"):
        triplet.pairwise_distance_grad(x, torch.zeros((4, 3), device="cuda"))
    with pytest.raises(RuntimeError, match="int32"):
        triplet.batch_hard(torch.zeros((4, 4), device="cuda"), torch.zeros(4, dtype=torch.int64, device="cuda"))


def test_pairwise_distance_matches_reference_cuda_operator_live_and_fixture(golden):
    """The reference's PairWiseDistance / PairWiseDistanceGradient kernels themselves (pairwise_distance_op.cu:9-22,
    78-91, compiled unmodified for sm_100a) on this device, and the fixture they wrote on a B200."""
    import torch
    from pps_b200 import triplet
    from oracle import ref_ops
    d = golden("triplet_ref_pairwise")
    i = 0
    while "x%d" % i in d:
        x, dz = torch.from_numpy(d["x%d" % i]).cuda(), torch.from_numpy(d["dz%d" % i]).cuda()
        z = triplet.pairwise_distance(x)
        dx = triplet.pairwise_distance_grad(x, dz)
        scale = float(np.abs(d["dx%d" % i]).max()) + 1e-6
        np.testing.assert_allclose(z.cpu().numpy(), d["z%d" % i], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(dx.cpu().numpy(), d["dx%d" % i], rtol=1e-4, atol=1e-5 * scale)
        if ref_ops.available():
            np.testing.assert_allclose(z.cpu().numpy(), ref_ops.pairwise_distance(x).cpu().numpy(), rtol=1e-5, atol=1e-5)
            np.testing.assert_allclose(dx.cpu().numpy(), ref_ops.pairwise_distance_grad(x, dz).cpu().numpy(), rtol=1e-4,
                                       atol=1e-5 * scale)
        i += 1
    assert i >= 5
    if ref_ops.available():
        with pytest.raises(RuntimeError, match=r"X.dim\(\) == 2"):               # CAFFE_ENFORCE_EQ(X.dim(), 2)
            ref_ops._run("PairWiseDistance", 1, [(x.data_ptr(), (4,))], [(z.data_ptr(), 16)])


def test_batch_hard_matches_reference_operator_fixture(golden):
    """BatchHardOp / BatchHardGradientOp (batch_hard_op.cc, compiled unmodified) wrote the fixture: forward bit for bit;
    gradient bit for bit except the operator's stray stores for an anchor whose search found nothing (idx == -1 lands on the
    last column of the previous row), which the kernel leaves out."""
    import torch
    from pps_b200 import triplet
    d = golden("triplet_ref_batch_hard")
    i = 0
    while "xd%d" % i in d:
        xd, labels = d["xd%d" % i], d["labels%d" % i]
        n = xd.shape[0]
        ap, an, ip, inn = triplet.batch_hard(torch.from_numpy(xd).cuda(), torch.from_numpy(labels).cuda(), return_indices=True)
        np.testing.assert_array_equal(ap.cpu().numpy(), d["ap%d" % i])
        np.testing.assert_array_equal(an.cpu().numpy(), d["an%d" % i])
        dx = triplet.batch_hard_grad(ip, inn, torch.from_numpy(d["dap%d" % i]).cuda(),
                                     torch.from_numpy(d["dan%d" % i]).cuda()).cpu().numpy()
        ipn, inn_n = ip.cpu().numpy(), inn.cpu().numpy()
        allowed = {(a - 1, n - 1) for a in range(1, n) if ipn[a] < 0 or inn_n[a] < 0}
        assert {tuple(x) for x in np.argwhere(dx != d["dx%d" % i])} <= allowed
        np.testing.assert_array_equal(dx, O.batch_hard_grad(ipn, inn_n, d["dap%d" % i], d["dan%d" % i]))
        i += 1
    assert i >= 6
