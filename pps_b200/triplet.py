"""Triplet-mining ops of the reference's training graph (detectron/modeling/triplet_loss.py:145-158), host mirror.

    PairWiseDistance(X) -> Z            detectron/ops/pairwise_distance_op.{h,cc,cu}   (squared distances, CUDA only)
    BatchHard(Xdist, L) -> (AP, AN)     detectron/ops/batch_hard_op.{h,cc,cu}          (GPU = CPU fallback in the reference)

Same operator contracts (shape checks raise RuntimeError like CAFFE_ENFORCE); all arithmetic in csrc/triplet.cu.
"""
from __future__ import annotations

from . import _lib


def _check(x, ndim, dtype, name):
    torch = _lib.require_cuda()
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if x.dim() != ndim:
        raise RuntimeError("%s.dim() == %d required, got %d" % (name, ndim, x.dim()))
    if x.dtype != dtype:
        raise RuntimeError("%s must be %s" % (name, dtype))
    return x.contiguous()


def pairwise_distance(x):
    """Z[p, q] = sum_d (X[p, d] - X[q, d])^2   (pairwise_distance_op.cu:9-22; the sqrt there is commented out)."""
    torch = _lib.require_cuda()
    x = _check(x, 2, torch.float32, "X")
    n, d = int(x.shape[0]), int(x.shape[1])
    z = torch.empty((n, n), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().pps_pairwise_distance_fwd(_lib.ptr(x), n, d, _lib.ptr(z), _lib.stream_ptr()),
                   "pps_pairwise_distance_fwd")
    return z


def pairwise_distance_grad(x, dz):
    """dX of PairWiseDistance (pairwise_distance_op.cu:78-91, without its fp32 atomics: deterministic)."""
    torch = _lib.require_cuda()
    x = _check(x, 2, torch.float32, "X")
    dz = _check(dz, 2, torch.float32, "dZ")
    n, d = int(x.shape[0]), int(x.shape[1])
    if tuple(dz.shape) != (n, n):
        raise RuntimeError("dZ.dim32(0) == X.dim32(0) and dZ.dim32(1) == X.dim32(0) required")
    dx = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().pps_pairwise_distance_bwd(_lib.ptr(x), _lib.ptr(dz), n, d, _lib.ptr(dx), _lib.stream_ptr()),
                   "pps_pairwise_distance_bwd")
    return dx


def batch_hard(xdist, labels, return_indices=False):
    """(AP, AN): hardest positive / negative distance per anchor (batch_hard_op.cc:9-59)."""
    torch = _lib.require_cuda()
    xdist = _check(xdist, 2, torch.float32, "X")
    labels = _check(labels, 1, torch.int32, "L")
    n = int(xdist.shape[0])
    if int(xdist.shape[1]) != n or int(labels.shape[0]) != n:
        raise RuntimeError("X.dim32(0) == X.dim32(1) == L.dim32(0) required")
    ap = torch.empty(n, dtype=torch.float32, device=xdist.device)
    an = torch.empty_like(ap)
    ip = torch.empty(n, dtype=torch.int32, device=xdist.device)
    inn = torch.empty_like(ip)
    with torch.cuda.device(xdist.device):
        _lib.check(_lib.load().pps_batch_hard_fwd(_lib.ptr(xdist), _lib.ptr(labels), n, _lib.ptr(ap), _lib.ptr(an),
                                                  _lib.ptr(ip), _lib.ptr(inn), _lib.stream_ptr()), "pps_batch_hard_fwd")
    return (ap, an, ip, inn) if return_indices else (ap, an)


def batch_hard_from_features(x, labels):
    """PairWiseDistance + BatchHard in one kernel, without the [N, N] matrix -> (AP, AN, idx_p, idx_n)."""
    torch = _lib.require_cuda()
    x = _check(x, 2, torch.float32, "X")
    labels = _check(labels, 1, torch.int32, "L")
    n, d = int(x.shape[0]), int(x.shape[1])
    if int(labels.shape[0]) != n:
        raise RuntimeError("X.dim32(0) == L.dim32(0) required")
    ap = torch.empty(n, dtype=torch.float32, device=x.device)
    an = torch.empty_like(ap)
    ip = torch.empty(n, dtype=torch.int32, device=x.device)
    inn = torch.empty_like(ip)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().pps_batch_hard_fused_fwd(_lib.ptr(x), _lib.ptr(labels), n, d, _lib.ptr(ap), _lib.ptr(an),
                                                        _lib.ptr(ip), _lib.ptr(inn), _lib.stream_ptr()),
                   "pps_batch_hard_fused_fwd")
    return ap, an, ip, inn


def batch_hard_grad(idx_p, idx_n, dap, dan):
    """dX of BatchHard: dAP / dAN scattered to the mined indices of a zero [N, N] (batch_hard_op.cc:62-123)."""
    torch = _lib.require_cuda()
    idx_p, idx_n = _check(idx_p, 1, torch.int32, "idx_p"), _check(idx_n, 1, torch.int32, "idx_n")
    dap, dan = _check(dap, 1, torch.float32, "dAP"), _check(dan, 1, torch.float32, "dAN")
    n = int(idx_p.shape[0])
    if not (int(idx_n.shape[0]) == int(dap.shape[0]) == int(dan.shape[0]) == n):
        raise RuntimeError("X.dim32(0) == dAP.dim32(0) == dAN.dim32(0) required")
    dx = torch.empty((n, n), dtype=torch.float32, device=dap.device)
    with torch.cuda.device(dap.device):
        _lib.check(_lib.load().pps_batch_hard_bwd(_lib.ptr(idx_p), _lib.ptr(idx_n), _lib.ptr(dap), _lib.ptr(dan), n,
                                                  _lib.ptr(dx), _lib.stream_ptr()), "pps_batch_hard_bwd")
    return dx
