"""Part-power-set pooling: host-side mirror of the reference's head functions.

Reference (paths relative to the reference repo):
  detectron/modeling/bpm_heads.py:18-55    add_uniform_partition  (strip split + global pools)
  detectron/modeling/pps_heads.py:38-80    add_pps_part_head_     (2^n - 1 subset combinations)
  detectron/modeling/pps_heads.py:83-142   add_pps_part_head      (FPN / multi-scale variant)
  detectron/modeling/FPN_reid.py:403-428   pyramid level list and scales

The reference builds a graph of ~200 stock Caffe2 ops per image; here one fused CUDA kernel
(csrc/pps_pool.cu, through ``pps_pool_fwd`` of the C ABI) produces the same blobs.  The head
function contract is kept: ``f(blob_in, dim_in, spatial_scale) -> (blobs_out, dims_out)`` with
``blobs_out`` the list of 2^n - 1 ``[N, C, 1, 1]`` tensors in ascending-mask order.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

from . import _lib

# the 21 contiguous subsets of 6 parts listed (and left disabled) at pps_heads.py:22-25,54-55
pyramid_combs = [[0], [1], [2], [3], [4], [5], [0, 1], [1, 2], [2, 3], [3, 4],
                 [4, 5], [0, 1, 2], [1, 2, 3], [2, 3, 4], [3, 4, 5],
                 [0, 1, 2, 3], [1, 2, 3, 4], [2, 3, 4, 5], [0, 1, 2, 3, 4],
                 [1, 2, 3, 4, 5], [0, 1, 2, 3, 4, 5]]


def comb_to_mask(comb: Sequence[int]) -> int:
    m = 0
    for j in comb:
        m |= 1 << int(j)
    return m


def mask_to_comb(mask: int, n_parts: int) -> List[int]:
    return [j for j in range(n_parts) if mask & (1 << j)]


@dataclass
class ReIDPoolCfg:
    """The cfg keys the pooling heads read (defaults: detectron/core/config.py:1016-1050, 713)."""
    BPM_STRIP_NUM: int = 6              # cfg.REID.BPM_STRIP_NUM
    MAX_AVE_FEATURE: bool = False       # cfg.REID.MAX_AVE_FEATURE (all shipped PPS yamls: True)
    SCALE: Tuple[int, int] = (128, 384)  # cfg.REID.SCALE = (W, H) of the input crop
    FPN_ON: bool = False                # cfg.FPN.FPN_ON
    FPN_SHARED: bool = False            # cfg.REID.FPN_SHARED
    train: bool = False                 # model.train
    PYRAMID_COMBS_ONLY: bool = False    # the filter commented out at pps_heads.py:54-55


def uniform_partition_split(strip_num: int, scale_h: int = 384, spatial_scale: float = 1.0 / 16) -> List[int]:
    """Rows per strip, exactly as bpm_heads.py:25-43 derives them."""
    tables = {7: [3, 3, 4, 4, 4, 3, 3], 5: [5, 5, 4, 5, 5], 9: [2, 3, 3, 3, 3, 3, 3, 2, 2],
              10: [2, 2, 2, 3, 3, 3, 3, 2, 2, 2]}
    if strip_num in tables and scale_h == 16 * 24:
        scale = 16 * spatial_scale
        return [int(s * scale) for s in tables[strip_num]]
    strip_h = int(scale_h * spatial_scale / strip_num)
    return [strip_h for _ in range(strip_num)]


def _mode_code(mode) -> int:
    if mode in ("max_ave", _lib.POOL_MAX_AVE, True):
        return _lib.POOL_MAX_AVE
    if mode in ("avg_max", _lib.POOL_AVG_MAX, False):
        return _lib.POOL_AVG_MAX
    raise RuntimeError("pps_pool: unknown mode %r (expected 'max_ave' or 'avg_max')" % (mode,))


def _launch(x, n_parts, split, mode, combos, out, out_elem_offset, sn, sk):
    """One pps_pool_fwd call writing at out.data_ptr() + 4*out_elem_offset with strides (sn, sk)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    N, Cc, H, W = (int(v) for v in x.shape)
    split_arr = (C.c_int * n_parts)(*[int(v) for v in split])
    if combos is None:
        combos_arr, n_combos = None, 0
    else:
        combos_arr, n_combos = (C.c_int * len(combos))(*[int(m) for m in combos]), len(combos)
    with torch.cuda.device(x.device):
        rc = lib.pps_pool_fwd(_lib.ptr(x), N, Cc, H, W, n_parts, split_arr, _mode_code(mode), combos_arr, n_combos,
                              C.c_void_p(out.data_ptr() + 4 * int(out_elem_offset)), int(sn), int(sk),
                              _lib.stream_ptr())
    _lib.check(rc, "pps_pool_fwd")


def _check_input(x):
    torch = _lib.require_cuda()
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError("pps_pool: x must be a CUDA tensor")
    if x.dtype != torch.float32:
        raise RuntimeError("pps_pool: x must be float32, got %s" % x.dtype)
    if x.dim() != 4:
        raise RuntimeError("pps_pool: x.ndim() == 4 required (NCHW), got %d" % x.dim())
    if not x.is_contiguous():
        raise RuntimeError("pps_pool: x must be contiguous NCHW")


def pps_pool(x, n_parts: int = 6, split: Optional[Sequence[int]] = None, mode="max_ave",
             combos: Optional[Sequence[int]] = None, layout: str = "nkc", out=None):
    """Fused strip pooling + part-power-set combination.

    x       : CUDA float32 tensor [N, C, H, W] (NCHW, contiguous).
    split   : rows per strip (default: H // n_parts each, bpm_heads.py:41-43).
    mode    : 'max_ave' -> Mean(strip averages) + Max(strip maxes)   (pps_heads.py:58-68)
              'avg_max' -> Max(strip averages)                       (pps_heads.py:69-76)
    combos  : optional list of subset masks; default all 1 .. 2^n - 1 ascending.
    layout  : 'nkc' -> [N, K, C]; 'knc' -> [K, N, C] (each out[k] is one reference blob).
    """
    torch = _lib.require_cuda()
    _check_input(x)
    N, Cc, H, W = (int(v) for v in x.shape)
    if split is None:
        split = [H // n_parts] * n_parts
    split = [int(s) for s in split]
    if len(split) != n_parts:
        raise RuntimeError("pps_pool: len(split) == n_parts required")
    K = (1 << n_parts) - 1 if combos is None else len(combos)
    if layout == "nkc":
        shape, sn, sk = (N, K, Cc), K * Cc, Cc
    elif layout == "knc":
        shape, sn, sk = (K, N, Cc), Cc, N * Cc
    else:
        raise RuntimeError("pps_pool: layout must be 'nkc' or 'knc'")
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=x.device)
    else:
        if tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != x.device:
            raise RuntimeError("pps_pool: out must be a contiguous float32 tensor of shape %s" % (shape,))
    _launch(x, n_parts, split, mode, combos, out, 0, sn, sk)
    return out


def pps_pool_backward(x, dy, n_parts: int = 6, split: Optional[Sequence[int]] = None, mode="max_ave",
                      combos: Optional[Sequence[int]] = None, layout: str = "nkc"):
    """Gradient of ``pps_pool`` w.r.t. ``x`` (pps_pool_bwd of the C ABI): ``dy`` has the shape / layout of the forward
    output.  Operator gradients as in Caffe2 (see include/pps_b200.h): ties between strips in ``Max`` all receive dY, the
    max pool routes to the first maximal element of the strip."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    _check_input(x)
    N, Cc, H, W = (int(v) for v in x.shape)
    if split is None:
        split = [H // n_parts] * n_parts
    split = [int(s) for s in split]
    if len(split) != n_parts:
        raise RuntimeError("pps_pool: len(split) == n_parts required")
    K = (1 << n_parts) - 1 if combos is None else len(combos)
    if layout == "nkc":
        shape, sn, sk = (N, K, Cc), K * Cc, Cc
    elif layout == "knc":
        shape, sn, sk = (K, N, Cc), Cc, N * Cc
    else:
        raise RuntimeError("pps_pool: layout must be 'nkc' or 'knc'")
    if (not isinstance(dy, torch.Tensor) or not dy.is_cuda or dy.dtype != torch.float32 or tuple(dy.shape) != shape
            or dy.device != x.device):
        raise RuntimeError("pps_pool_backward: dy must be a float32 CUDA tensor of shape %s" % (shape,))
    dy = dy.contiguous()
    dx = torch.empty_like(x)
    split_arr = (C.c_int * n_parts)(*split)
    if combos is None:
        combos_arr, n_combos = None, 0
    else:
        combos_arr, n_combos = (C.c_int * len(combos))(*[int(m) for m in combos]), len(combos)
    with torch.cuda.device(x.device):
        rc = lib.pps_pool_bwd(_lib.ptr(x), _lib.ptr(dy), N, Cc, H, W, n_parts, split_arr, _mode_code(mode), combos_arr,
                              n_combos, int(sn), int(sk), _lib.ptr(dx), _lib.stream_ptr())
    _lib.check(rc, "pps_pool_bwd")
    return dx


def pps_pool_autograd(x, n_parts: int = 6, split: Optional[Sequence[int]] = None, mode="max_ave",
                      combos: Optional[Sequence[int]] = None, layout: str = "nkc"):
    """``pps_pool`` as a differentiable torch op (forward pps_pool_fwd, backward pps_pool_bwd) - what the train-time
    multi-scale branch (pps_heads.py:106-135) needs."""
    torch = _lib.require_cuda()

    class _Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, inp):
            ctx.save_for_backward(inp)
            return pps_pool(inp, n_parts, split, mode, combos, layout)

        @staticmethod
        def backward(ctx, grad):
            (inp,) = ctx.saved_tensors
            return pps_pool_backward(inp, grad, n_parts, split, mode, combos, layout)

    return _Fn.apply(x)


def add_uniform_partition_split(cfg: ReIDPoolCfg, spatial_scale: float) -> List[int]:
    return uniform_partition_split(cfg.BPM_STRIP_NUM, cfg.SCALE[1], spatial_scale)


def add_pps_part_head_(blob_in, dim_in, spatial_scale, cfg: Optional[ReIDPoolCfg] = None, preprefix="pps"):
    """pps_heads.py:38-80 — returns (blobs_out, dims_out); blobs_out[k] is [N, C, 1, 1]."""
    cfg = cfg or ReIDPoolCfg()
    n = cfg.BPM_STRIP_NUM
    split = add_uniform_partition_split(cfg, spatial_scale)
    if sum(split) != int(blob_in.shape[2]):
        # Caffe2 Split enforces that the split sizes cover the axis
        raise RuntimeError("Split: sum(split) == input.dim(axis) failed: %d vs %d" % (sum(split), blob_in.shape[2]))
    combos = [comb_to_mask(c) for c in pyramid_combs] if cfg.PYRAMID_COMBS_ONLY else None
    pool = pps_pool_autograd if getattr(blob_in, "requires_grad", False) else pps_pool   # train: gradients flow to the map
    y = pool(blob_in, n, split, "max_ave" if cfg.MAX_AVE_FEATURE else "avg_max", combos=combos, layout="knc")
    blobs_out = [y[k].view(y.shape[1], y.shape[2], 1, 1) for k in range(y.shape[0])]
    dims_out = [dim_in] * len(blobs_out)
    return blobs_out, dims_out


def blob_names(n_parts: int, preprefix: str = "pps", combos: Optional[Sequence[int]] = None) -> List[str]:
    """Names the reference gives the outputs: preprefix + digits of the parts + '_pool2' (pps_heads.py:62-76)."""
    masks = combos if combos is not None else range(1, 1 << n_parts)
    return [preprefix + "".join(str(j) for j in mask_to_comb(m, n_parts)) + "_pool2" for m in masks]


def add_pps_part_head(blob_in, dim_in, spatial_scale, cfg: Optional[ReIDPoolCfg] = None, preprefix="pps"):
    """pps_heads.py:83-142 — the head entry point, including the multi-scale (FPN) variant.

    Without FPN: one map in, 2^n - 1 blobs out.  With FPN at test time only level 0 is pooled
    (:88-96).  With FPN at train time every level is pooled (:106-117); with FPN_SHARED the
    same-combination blobs of all levels are concatenated along the batch axis (:119-135).
    """
    cfg = cfg or ReIDPoolCfg()
    if not cfg.FPN_ON:
        return add_pps_part_head_(blob_in, dim_in, spatial_scale, cfg, preprefix)
    if not cfg.train:
        return add_pps_part_head_(blob_in[0], dim_in[0], spatial_scale[0], cfg, preprefix)
    if cfg.FPN_SHARED:
        # Concat(axis=0) of the same-combination blobs of all levels (:119-135): every level's
        # kernel writes straight into its batch slice of one [K, sum(N_i), C] buffer.
        torch = _lib.require_cuda()
        n = cfg.BPM_STRIP_NUM
        combos = [comb_to_mask(c) for c in pyramid_combs] if cfg.PYRAMID_COMBS_ONLY else None
        K = (1 << n) - 1 if combos is None else len(combos)
        Cc = int(blob_in[0].shape[1])
        for b in blob_in:
            _check_input(b)
            if int(b.shape[1]) != Cc:
                raise RuntimeError("Concat: all levels must have the same channel count when REID.FPN_SHARED")
        n_total = sum(int(b.shape[0]) for b in blob_in)
        if any(b.requires_grad for b in blob_in):
            # differentiable form: one autograd node per level, Concat(axis=0) of the batch slices
            mode = "max_ave" if cfg.MAX_AVE_FEATURE else "avg_max"
            ys = []
            for i, b in enumerate(blob_in):
                split = add_uniform_partition_split(cfg, spatial_scale[i])
                if sum(split) != int(b.shape[2]):
                    raise RuntimeError("Split: sum(split) == input.dim(axis) failed: %d vs %d" % (sum(split), b.shape[2]))
                ys.append(pps_pool_autograd(b, n, split, mode, combos=combos, layout="knc"))
            y = torch.cat(ys, dim=1)
            return [y[k].reshape(n_total, Cc, 1, 1) for k in range(K)], [dim_in[0]] * K
        y = torch.empty((K, n_total, Cc), dtype=torch.float32, device=blob_in[0].device)
        row = 0
        for i, b in enumerate(blob_in):
            split = add_uniform_partition_split(cfg, spatial_scale[i])
            if sum(split) != int(b.shape[2]):
                raise RuntimeError("Split: sum(split) == input.dim(axis) failed: %d vs %d" % (sum(split), b.shape[2]))
            _launch(b, n, split, "max_ave" if cfg.MAX_AVE_FEATURE else "avg_max", combos, y, row * Cc, Cc, n_total * Cc)
            row += int(b.shape[0])
        blobs_out = [y[k].view(n_total, Cc, 1, 1) for k in range(K)]
        return blobs_out, [dim_in[0]] * K
    blobs_outs, dims_outs = [], []
    for i in range(len(blob_in)):
        b, d = add_pps_part_head_(blob_in[i], dim_in[i], spatial_scale[i], cfg, preprefix + "_" + str(i) + "_")
        blobs_outs.extend(b)
        dims_outs.extend(d)
    return blobs_outs, dims_outs
