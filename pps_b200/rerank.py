"""k-reciprocal re-ranking on the device — host mirror of ``re_ranking`` and of the re-ranked branches of
``evaluate`` (detectron/datasets/reid_dataset_evaluator.py:442-519 and :161-207; ``cfg.REID.RERANK`` defaults to
True, detectron/core/config.py:1022).

The reference works on dense ``[N, N]`` float32 arrays (N = nq + ng) with Python loops over all N images; here the
neighbour lists, the k-reciprocal sets, the expanded V rows and the inverted index are sparse and every step is one
kernel of ``csrc/rerank.cu``.  The arithmetic is float32 in the reference's order, so the result differs from it
only through ``np.sum``'s pairwise order inside one normalisation, the last bit of ``exp`` and the order of exact
ties in ``np.argsort`` (an unstable sort there; (value, index) order here).  No CPU path.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .evaluator import _as_cuda_f32, _torch, compute_dist


def _rerank_device(m, nq, k1, k2, lambda_value):
    """m: [n, n] float32 CUDA tensor (assembled original distances).  Returns [nq, n - nq] CUDA tensor."""
    torch = _torch()
    lib = _lib.load()
    n = int(m.shape[0])
    ng = n - nq
    dev = m.device
    s = _lib.stream_ptr()
    vcap = int(lib.pps_rerank_vcap())
    if k1 + 1 > n:
        raise RuntimeError("re_ranking: k1 + 1 = %d neighbours asked of %d images" % (k1 + 1, n))
    od = torch.empty((n, n), dtype=torch.float32, device=dev)
    colmax = torch.empty(n, dtype=torch.float32, device=dev)
    _lib.check(lib.pps_rerank_normalize(_lib.ptr(m), int(m.stride(0)), n, _lib.ptr(colmax), _lib.ptr(od), n, s),
               "pps_rerank_normalize")
    # initial_rank[:, :k1 + 1]: the nearest columns of every row of od (np.argsort(original_dist), :456)
    rk = k1 + 1
    key = torch.empty((n, rk), dtype=torch.int64, device=dev)
    _lib.check(lib.pps_topk_init(_lib.ptr(key), n, rk, s), "pps_topk_init")
    # (the one-read sweep kernel with no pair lists: only its top-k half runs)
    no_pairs = torch.zeros(n + 1, dtype=torch.int32, device=dev)
    no_first = torch.zeros(n, dtype=torch.int32, device=dev)
    _lib.check(lib.pps_rank_sweep(_lib.ptr(od), n, n, n, 0, _lib.ptr(no_pairs), None, None, None, 0, None, _lib.ptr(no_first),
                                  _lib.ptr(key), rk, 0, s), "pps_rank_sweep")
    rank = torch.empty((n, rk), dtype=torch.int32, device=dev)
    _lib.check(lib.pps_topk_unpack(_lib.ptr(key), n, rk, None, _lib.ptr(rank), s), "pps_topk_unpack")
    v_idx = torch.empty((n, vcap), dtype=torch.int32, device=dev)
    v_val = torch.empty((n, vcap), dtype=torch.float32, device=dev)
    v_cnt = torch.empty(n, dtype=torch.int32, device=dev)
    _lib.check(lib.pps_rerank_krecip(_lib.ptr(rank), rk, n, k1, _lib.ptr(od), n, _lib.ptr(v_idx), _lib.ptr(v_val),
                                     _lib.ptr(v_cnt), s), "pps_rerank_krecip")
    if k2 != 1:
        cap = k2 * vcap
        q_idx = torch.empty((n, cap), dtype=torch.int32, device=dev)
        q_val = torch.empty((n, cap), dtype=torch.float32, device=dev)
        q_cnt = torch.empty(n, dtype=torch.int32, device=dev)
        _lib.check(lib.pps_rerank_expand(_lib.ptr(rank), rk, n, k2, _lib.ptr(v_idx), _lib.ptr(v_val), _lib.ptr(v_cnt), cap,
                                         _lib.ptr(q_idx), _lib.ptr(q_val), _lib.ptr(q_cnt), s), "pps_rerank_expand")
    else:
        cap, q_idx, q_val, q_cnt = vcap, v_idx, v_val, v_cnt
    nnz = int(q_cnt[nq:].sum().item()) if ng else 0          # sizes the inverted index (one small read-back)
    col_cnt = torch.empty(n, dtype=torch.int32, device=dev)
    col_off = torch.empty(n + 1, dtype=torch.int32, device=dev)
    cursor = torch.empty(n, dtype=torch.int32, device=dev)
    inv_row = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    inv_val = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
    _lib.check(lib.pps_rerank_invert(_lib.ptr(q_idx), _lib.ptr(q_val), _lib.ptr(q_cnt), cap, nq, n, _lib.ptr(col_cnt),
                                     _lib.ptr(col_off), _lib.ptr(cursor), _lib.ptr(inv_row), _lib.ptr(inv_val), s),
               "pps_rerank_invert")
    out = torch.empty((nq, ng), dtype=torch.float32, device=dev)
    _lib.check(lib.pps_rerank_jaccard(_lib.ptr(q_idx), _lib.ptr(q_val), _lib.ptr(q_cnt), cap, _lib.ptr(col_off),
                                      _lib.ptr(inv_row), _lib.ptr(inv_val), nq, ng, _lib.ptr(od), n, float(lambda_value),
                                      _lib.ptr(out), max(ng, 1), s), "pps_rerank_jaccard")
    return out


def re_ranking(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3):
    """reid_dataset_evaluator.py:442-519, same arguments.  numpy in -> numpy out; CUDA tensors in -> CUDA tensor out."""
    torch = _torch()
    qg, qg_np = _as_cuda_f32(q_g_dist, "q_g_dist")
    qq, _ = _as_cuda_f32(q_q_dist, "q_q_dist")
    gg, _ = _as_cuda_f32(g_g_dist, "g_g_dist")
    nq, ng = int(qg.shape[0]), int(qg.shape[1])
    if tuple(qq.shape) != (nq, nq) or tuple(gg.shape) != (ng, ng):
        raise RuntimeError("re_ranking: expected q_q_dist %s and g_g_dist %s, got %s and %s" % (
            (nq, nq), (ng, ng), tuple(qq.shape), tuple(gg.shape)))
    with torch.cuda.device(qg.device):
        n = nq + ng
        m = torch.empty((n, n), dtype=torch.float32, device=qg.device)      # np.concatenate of the four blocks (:447-452)
        m[:nq, :nq] = qq
        m[:nq, nq:] = qg
        m[nq:, :nq] = qg.t()
        m[nq:, nq:] = gg
        out = _rerank_device(m, nq, int(k1), int(k2), lambda_value)
    return out.cpu().numpy() if qg_np else out


def re_ranking_from_features(q_feats, g_feats, k1=20, k2=6, lambda_value=0.3, precision="f16x3"):
    """The re-ranked query x gallery distance straight from features: the three ``compute_dist`` calls of evaluate()
    (:165-171) are one [n, n] tensor-core product of the stacked features with themselves."""
    torch = _torch()
    q, q_np = _as_cuda_f32(q_feats, "q_feats")
    g, _ = _as_cuda_f32(g_feats, "g_feats")
    with torch.cuda.device(q.device):
        x = torch.cat([q.float(), g.float()], dim=0)
        m = compute_dist(x, x, precision=precision)
        out = _rerank_device(m, int(q.shape[0]), int(k1), int(k2), lambda_value)
    return out.cpu().numpy() if q_np else out
