"""The step between pooling and distance: per-combination embedding, concat, L2 normalise.

Host mirror of ``add_reid_outputs`` at test time (detectron/modeling/reid_heads.py:34-127): every pooled
blob ``[N, C, 1, 1]`` goes through its own ``Conv1x1(C -> BPM_DIM) + SpatialBN(is_test) + ReLU`` (:41-76),
the K results are ``Concat(axis=1)`` into ``reid_feature_concat`` ``[N, K*BPM_DIM]`` (:95-101) and, with
``cfg.REID.NORMALIZE_FEATURE``, ``Normalize(axis=1)`` (:123-127, triplet_loss.py:18).  The classifier
``FC`` branch (:81-90) only feeds the training loss and is not part of the retrieval path.

All K branches are ONE grouped launch of the 2-CTA tcgen05 kernel (``pps_embed_tc``): the pooled
features stay in the ``[K, N, C]`` layout the pooling kernel writes, the affine + ReLU epilogue stores
straight into the concatenated ``[N, K*E]`` feature.  No CPU path.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _lib
from .evaluator import SplitOperand, _prec_code, _torch

BN_EPS = 1e-5          # Caffe2 SpatialBN default epsilon (model.SpatialBN is called without one, reid_heads.py:60)


def fold_bn(conv_bias, bn_scale, bn_bias, bn_mean, bn_var, eps: float = BN_EPS):
    """Conv bias + test-mode SpatialBN as one affine map per output channel: y = alpha * (w.x) + beta.

    SpatialBN(is_test): y = (z - mean) / sqrt(var + eps) * scale + bias with z = w.x + conv_bias.
    Arrays of any common shape (e.g. [K, E]); float64 arithmetic, float32 result.
    """
    f = lambda a: np.asarray(a, dtype=np.float64)
    alpha = f(bn_scale) / np.sqrt(f(bn_var) + eps)
    beta = (f(conv_bias) - f(bn_mean)) * alpha + f(bn_bias)
    return alpha.astype(np.float32), beta.astype(np.float32)


class ReidEmbedHead:
    """Weights of the K embedding branches, prepared once (bf16 residual planes in HBM).

    weight : [K, E, C] float32 (the K ``{prefix}_conv_w`` blobs ``[E, C, 1, 1]`` stacked in blob order)
    alpha, beta : [K, E] float32 (``fold_bn``)
    """

    def __init__(self, weight, alpha, beta, precision: str = "bf16x3", device=None):
        torch = _torch()
        _lib.require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        as_t = lambda a: (a if hasattr(a, "is_cuda") else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)))
        w = as_t(weight).to(dev, torch.float32)
        if w.dim() != 3:
            raise RuntimeError("weight: expected [K, E, C], got %s" % (tuple(w.shape),))
        self.K, self.E, self.C = (int(s) for s in w.shape)
        self.alpha = as_t(alpha).to(dev, torch.float32).reshape(-1).contiguous()
        self.beta = as_t(beta).to(dev, torch.float32).reshape(-1).contiguous()
        if self.alpha.numel() != self.K * self.E or self.beta.numel() != self.K * self.E:
            raise RuntimeError("alpha / beta: expected %d values" % (self.K * self.E))
        self.prec = _prec_code(precision)
        if self.prec not in (_lib.PREC_BF16X1, _lib.PREC_BF16X3, _lib.PREC_BF16X6):
            raise RuntimeError("embedding precision must be one of bf16x1 / bf16x3 / bf16x6")
        self.device = dev
        with torch.cuda.device(dev):
            self.w = SplitOperand(w.reshape(self.K * self.E, self.C).contiguous(), _lib.PLANES_FOR[self.prec])

    def __call__(self, pooled, normalize: bool = True, out=None):
        """pooled: [K, N, C] float32 CUDA tensor (``pps_pool(..., layout='knc')``) -> [N, K*E] float32."""
        torch = _torch()
        lib = _lib.load()
        if not hasattr(pooled, "is_cuda") or not pooled.is_cuda:
            raise RuntimeError("pooled: expected a CUDA tensor (there is no CPU path)")
        if pooled.dim() != 3 or int(pooled.shape[0]) != self.K or int(pooled.shape[2]) != self.C:
            raise RuntimeError("pooled: expected [K=%d, N, C=%d], got %s" % (self.K, self.C, tuple(pooled.shape)))
        if pooled.dtype != torch.float32:
            raise RuntimeError("pooled: expected float32")
        n = int(pooled.shape[1])
        feat_dim = self.K * self.E
        if out is None:
            out = torch.empty((n, feat_dim), dtype=torch.float32, device=pooled.device)
        elif tuple(out.shape) != (n, feat_dim) or out.dtype != torch.float32 or out.stride(1) != 1:
            raise RuntimeError("out: expected a float32 [N, K*E] tensor with unit column stride")
        if n == 0:
            return out
        with torch.cuda.device(pooled.device):
            x = SplitOperand(pooled.contiguous().reshape(self.K * n, self.C), self.w.planes_n)
            _lib.check(lib.pps_embed_tc(_lib.ptr(x.planes), x.planes_n, n, _lib.ptr(self.w.planes), self.w.planes_n,
                                        self.E, self.K, self.C, _lib.ptr(self.alpha), _lib.ptr(self.beta), self.prec,
                                        _lib.ptr(out), int(out.stride(0)), _lib.stream_ptr()), "pps_embed_tc")
            if normalize:
                l2_normalize_rows(out, out=out)
        return out


def embed_maps(head: "ReidEmbedHead", x, n_parts: int = 6, split=None, mode="max_ave", combos=None,
               normalize: bool = True, out=None):
    """conv5 maps [N, C, H, W] -> normalised ``reid_feature_concat`` [N, K*E] without the fp32 pooled intermediate:
    the pooling kernel writes the bf16 operand planes of the embedding product directly (``pps_pool_planes_fwd``),
    then one grouped tcgen05 launch (``pps_embed_tc``) and the row normalisation.  Same result as
    ``head(pps_pool(x, layout='knc'))`` bit for bit (the planes are the same numbers either way)."""
    import ctypes as C
    from . import pooling
    torch = _torch()
    lib = _lib.load()
    pooling._check_input(x)
    n, c, h, _ = (int(v) for v in x.shape)
    if c != head.C:
        raise RuntimeError("embed_maps: maps have %d channels, the head expects %d" % (c, head.C))
    if split is None:
        split = [h // n_parts] * n_parts
    k_out = (1 << n_parts) - 1 if combos is None else len(combos)
    if k_out != head.K:
        raise RuntimeError("embed_maps: %d combinations pooled, the head has %d branches" % (k_out, head.K))
    feat_dim = head.K * head.E
    if out is None:
        out = torch.empty((n, feat_dim), dtype=torch.float32, device=x.device)
    elif tuple(out.shape) != (n, feat_dim) or out.dtype != torch.float32 or out.stride(1) != 1:
        raise RuntimeError("out: expected a float32 [N, K*E] tensor with unit column stride")
    if n == 0:
        return out
    planes_n = head.w.planes_n
    split_arr = (C.c_int * n_parts)(*[int(v) for v in split])
    combos_arr, n_combos = (None, 0) if combos is None else ((C.c_int * len(combos))(*[int(m) for m in combos]), len(combos))
    with torch.cuda.device(x.device):
        nbytes = int(lib.pps_split_bytes(head.K * n, head.C, planes_n))
        planes = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        _lib.check(lib.pps_pool_planes_fwd(_lib.ptr(x), n, c, h, int(x.shape[3]), n_parts, split_arr, pooling._mode_code(mode),
                                           combos_arr, n_combos, _lib.ptr(planes), planes_n, _lib.stream_ptr()),
                   "pps_pool_planes_fwd")
        _lib.check(lib.pps_embed_tc(_lib.ptr(planes), planes_n, n, _lib.ptr(head.w.planes), head.w.planes_n, head.E, head.K,
                                    head.C, _lib.ptr(head.alpha), _lib.ptr(head.beta), head.prec, _lib.ptr(out),
                                    int(out.stride(0)), _lib.stream_ptr()), "pps_embed_tc")
        if normalize:
            l2_normalize_rows(out, out=out)
    return out


def l2_normalize_rows(x, out=None):
    """Caffe2 ``Normalize(axis=1)``: x / max(|x|_2, 1e-12) per row (triplet_loss.py:18)."""
    torch = _torch()
    lib = _lib.load()
    if not hasattr(x, "is_cuda") or not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1:
        raise RuntimeError("l2_normalize_rows: expected a 2-D float32 CUDA tensor with unit column stride")
    if out is None:
        out = torch.empty_like(x, memory_format=torch.contiguous_format)
    with torch.cuda.device(x.device):
        _lib.check(lib.pps_l2_normalize_rows(_lib.ptr(x), int(x.shape[0]), int(x.shape[1]), int(x.stride(0)),
                                             _lib.ptr(out), int(out.stride(0)), _lib.stream_ptr()),
                   "pps_l2_normalize_rows")
    return out


def add_reid_outputs(blob_in, head: ReidEmbedHead, normalize: Optional[bool] = True):
    """reid_heads.add_reid_outputs at test time: pooled blobs -> ``reid_feature_concat[_norm]`` [N, K*E].

    blob_in: the list of K ``[N, C, 1, 1]`` views ``add_pps_part_head`` returns (views of one [K, N, C]
    buffer are used in place), or that [K, N, C] tensor itself.
    """
    torch = _torch()
    if isinstance(blob_in, (list, tuple)):
        first = blob_in[0]
        n, c = int(first.shape[0]), int(first.shape[1])
        step = n * c * first.element_size()
        contiguous_views = all(b.is_contiguous() and b.data_ptr() == first.data_ptr() + i * step
                               for i, b in enumerate(blob_in))
        if contiguous_views:
            pooled = torch.as_strided(first, (len(blob_in), n, c), (n * c, c, 1))
        else:
            pooled = torch.stack([b.reshape(n, c) for b in blob_in], 0)
    else:
        pooled = blob_in
    return head(pooled, normalize=bool(normalize))
