"""CPU oracle for the PPS retrieval hot path — TEST INFRASTRUCTURE, not product code.

A NumPy restatement of the reference's algorithm, used only as the checker by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` (plus three study scripts
under ``tools/`` that measure error margins against it: map_margin_sweep, split_precision_sim, rerank_bench --cpu-images).
Nothing under ``pps_b200/`` imports it.

Pinning status
  * ranking (compute_dist / cmc / mean_ap): the reference ships NO golden vectors or tests for
    these functions.  The restatement is pinned instead against outputs of
    the unmodified reference functions run in the authoring container (oracle/ref_loader.py,
    fixtures in tests/golden/ made by oracle/make_golden.py): numpy 2.3.5, scikit-learn 1.9.0.
  * pooling: the reference's pooling exists only as a Caffe2 graph (pytorch v1.0.1 `Split/AveragePool/MaxPool/
    Mean/Max/Add`, not vendored, not importable here).  The GRAPH is pinned: oracle/ref_pool_loader.py executes the
    unmodified builders (bpm_heads.py:18-55, pps_heads.py:38-142) with eagerly evaluated operators and
    tests/golden/pool_*.npz (oracle/make_golden_pool.py) hold what they return; this file reproduces them bit for
    bit.  The arithmetic inside the six stock operators stays a restatement of their published semantics
    (**parity unpinned** in that sense).
  * triplet-mining ops: pinned to the reference's own operators - batch_hard_op.cc and
    pairwise_distance_op.{cc,cu} compiled unmodified against oracle/caffe2_shim (oracle/build_ref_ops.py ->
    oracle/_ref/libref_reid_ops.so), fixtures tests/golden/triplet_ref_*.npz (oracle/make_golden_triplet.py).

Third-party arithmetic the reference leans on: scikit-learn ``average_precision_score``
(reid_dataset_evaluator.py:434; the code asks for 0.18.1 at :398-407, the installed 1.9.0
computes the step-wise AP; both definitions are restated below), NumPy ``matmul``/``argsort``.
"""
from __future__ import annotations

import numpy as np

# ------------------------------------------------------------------------------------
# pooling  (bpm_heads.py:18-55, pps_heads.py:38-80)
# ------------------------------------------------------------------------------------


def uniform_partition_split(strip_num, scale_h=384, spatial_scale=1.0 / 16):
    """Rows per strip.  bpm_heads.py:25-43: fixed tables for 5/7/9/10 strips on a 384-row crop
    (scaled by 16*spatial_scale for pyramid levels), else int(H*scale/strip_num) each."""
    tables = {
        7: [3, 3, 4, 4, 4, 3, 3],
        5: [5, 5, 4, 5, 5],
        9: [2, 3, 3, 3, 3, 3, 3, 2, 2],
        10: [2, 2, 2, 3, 3, 3, 3, 2, 2, 2],
    }
    if strip_num in tables and scale_h == 16 * 24:
        return [int(s * (16 * spatial_scale)) for s in tables[strip_num]]
    return [int(scale_h * spatial_scale / strip_num)] * strip_num


def strip_pools(x, split, dtype=np.float32):
    """Split(axis=2) then global AveragePool / MaxPool per strip (bpm_heads.py:45-55).
    x: [N, C, H, W] -> avg [n, N, C], max [n, N, C]."""
    x = np.asarray(x)
    assert x.ndim == 4 and sum(split) == x.shape[2], "Split: sum(split) must equal H"
    avgs, maxs, r = [], [], 0
    for h in split:
        s = x[:, :, r:r + h, :].astype(dtype)
        avgs.append(s.mean(axis=(2, 3), dtype=dtype))
        maxs.append(s.max(axis=(2, 3)))
        r += h
    return np.stack(avgs), np.stack(maxs)


def pps_pool(x, n_parts=6, split=None, mode="max_ave", combos=None, dtype=np.float32):
    """All part-combination features, ascending mask order (pps_heads.py:47-76) -> [N, K, C].

    max_ave (:58-68): Caffe2 ``Mean`` = inputs summed in order then scaled by 1/count;
                      ``Max`` = elementwise max; ``Add``.
    avg_max (:69-76): ``Max`` over the strip averages.
    """
    x = np.asarray(x)
    if split is None:
        split = [x.shape[2] // n_parts] * n_parts
    avg, mx = strip_pools(x, split, dtype)
    masks = list(combos) if combos is not None else list(range(1, 1 << n_parts))
    out = np.empty((x.shape[0], len(masks), x.shape[1]), dtype=dtype)
    for k, m in enumerate(masks):
        parts = [j for j in range(n_parts) if m & (1 << j)]
        if mode == "max_ave":
            s = avg[parts[0]].copy()
            for j in parts[1:]:
                s = s + avg[j]
            if len(parts) > 1:
                s = s * dtype(1.0 / len(parts))
            top = mx[parts[0]]
            for j in parts[1:]:
                top = np.maximum(top, mx[j])
            out[:, k, :] = s + top
        elif mode == "avg_max":
            top = avg[parts[0]]
            for j in parts[1:]:
                top = np.maximum(top, avg[j])
            out[:, k, :] = top
        else:
            raise ValueError(mode)
    return out


# ------------------------------------------------------------------------------------
# distance  (reid_dataset_evaluator.py:244-272)
# ------------------------------------------------------------------------------------


def compute_dist(array1, array2, type="euclidean"):
    """:266-271: |a|^2 + |b|^2 - 2 a.b^T, negatives clamped to 0, sqrt.  dtype follows the input.
    'cosine' (:259-263) in the reference dies on an undefined ``normalize``; the helper it means is
    test_engine.py:52-55 (row L2 normalisation), and the result is the similarity a.b^T."""
    assert type in ["cosine", "euclidean"]
    if type == "cosine":
        a = array1 / (np.linalg.norm(array1, ord=2, axis=1, keepdims=True) + 1e-12)
        b = array2 / (np.linalg.norm(array2, ord=2, axis=1, keepdims=True) + 1e-12)
        return np.matmul(a, b.T)
    sq1 = np.sum(np.square(array1), axis=1)[:, np.newaxis]
    sq2 = np.sum(np.square(array2), axis=1)[np.newaxis, :]
    d2 = -2 * np.matmul(array1, array2.T) + sq1 + sq2
    d2[d2 < 0] = 0
    return np.sqrt(d2)


# ------------------------------------------------------------------------------------
# masks / ranking  (reid_dataset_evaluator.py:283-439)
# ------------------------------------------------------------------------------------


def valid_mask(query_id, query_cam, gallery_ids, gallery_cams):
    """:327-328 / :427-428: drop gallery items with the query's id AND camera."""
    return (gallery_ids != query_id) | (gallery_cams != query_cam)


def average_precision_step(y_true, y_score):
    """scikit-learn >= 0.19 ``average_precision_score`` for one binary ranking: thresholds are the
    distinct scores; AP = sum_k (R_k - R_{k-1}) P_k  (tie-grouped, no interpolation)."""
    y_true = np.asarray(y_true).astype(bool)
    y_score = np.asarray(y_score)
    order = np.argsort(-y_score, kind="stable")
    ys, yt = y_score[order], y_true[order]
    last_of_group = np.r_[np.nonzero(np.diff(ys))[0], len(ys) - 1]
    tps = np.cumsum(yt)[last_of_group].astype(np.float64)
    fps = (1 + last_of_group) - tps
    precision = tps / (tps + fps)
    recall = tps / tps[-1]
    return float(np.sum(np.diff(np.r_[0.0, recall]) * precision))


def average_precision_trapezoid(y_true, y_score):
    """scikit-learn 0.18.1 ``average_precision_score`` (what :398-407 asks for): area under the
    precision-recall curve by the trapezoidal rule, curve starting at (recall 0, precision 1)."""
    y_true = np.asarray(y_true).astype(bool)
    y_score = np.asarray(y_score)
    order = np.argsort(-y_score, kind="stable")
    ys, yt = y_score[order], y_true[order]
    last_of_group = np.r_[np.nonzero(np.diff(ys))[0], len(ys) - 1]
    tps = np.cumsum(yt)[last_of_group].astype(np.float64)
    fps = (1 + last_of_group) - tps
    precision = tps / (tps + fps)
    recall = tps / tps[-1]
    # precision_recall_curve stops at full recall and prepends the (0, 1) end point
    stop = int(np.searchsorted(tps, tps[-1]))
    precision = np.r_[1.0, precision[:stop + 1]]
    recall = np.r_[0.0, recall[:stop + 1]]
    trapezoid = getattr(np, "trapezoid", None) or np.trapz
    return float(trapezoid(precision, recall))


def mean_ap(distmat, query_ids, gallery_ids, query_cams, gallery_cams, average=True, ap_fn=None):
    """:366-439.  ``ap_fn`` defaults to scikit-learn's installed ``average_precision_score`` exactly
    as the reference calls it; pass ``average_precision_step`` / ``average_precision_trapezoid`` for
    the restated definitions."""
    if ap_fn is None:
        from sklearn.metrics import average_precision_score as ap_fn
    m, n = distmat.shape
    order = np.argsort(distmat, axis=1)
    aps = np.zeros(m)
    is_valid_query = np.zeros(m)
    for i in range(m):
        idx = order[i]
        keep = valid_mask(query_ids[i], query_cams[i], gallery_ids[idx], gallery_cams[idx])
        y_true = (gallery_ids[idx] == query_ids[i])[keep]
        if not np.any(y_true):
            continue
        y_score = -distmat[i][idx][keep]
        is_valid_query[i] = 1
        aps[i] = ap_fn(y_true, y_score)
    if len(aps) == 0:
        raise RuntimeError("No valid query")
    if average:
        return float(np.sum(aps)) / np.sum(is_valid_query)
    return aps, is_valid_query


def cmc(distmat, query_ids, gallery_ids, query_cams, gallery_cams, topk=100, first_match_break=False,
        average=True, stable=False, separate_camera_set=False):
    """:283-363 with single_gallery_shot=False (the evaluator also fixes separate_camera_set=False, :35-37; True removes
    EVERY gallery item of the query's camera, :329-331).  ``stable=True`` sorts ties by gallery index (the reference's
    np.argsort is unstable, so tie order there is implementation-defined)."""
    m, n = distmat.shape
    order = np.argsort(distmat, axis=1, kind="stable" if stable else None)
    ret = np.zeros([m, topk])
    is_valid_query = np.zeros(m)
    num_valid = 0
    for i in range(m):
        idx = order[i]
        keep = valid_mask(query_ids[i], query_cams[i], gallery_ids[idx], gallery_cams[idx])
        if separate_camera_set:
            keep = keep & (gallery_cams[idx] != query_cams[i])
        hits = (gallery_ids[idx] == query_ids[i])[keep]
        if not np.any(hits):
            continue
        is_valid_query[i] = 1
        where = np.nonzero(hits)[0]
        delta = 1.0 / len(where)
        for j, k in enumerate(where):
            if k - j >= topk:
                break
            if first_match_break:
                ret[i, k - j] += 1
                break
            ret[i, k - j] += delta
        num_valid += 1
    if num_valid == 0:
        raise RuntimeError("No valid query")
    ret = ret.cumsum(axis=1)
    if average:
        return np.sum(ret, axis=0) / num_valid
    return ret, is_valid_query


# ------------------------------------------------------------------------------------
# count-based restatement (what the GPU kernels compute) — an independent cross-check
# ------------------------------------------------------------------------------------


def rank_counts(distmat, query_ids, gallery_ids, query_cams, gallery_cams):
    """Per query: AP from <=-counts, 0-based rank of the first match under (distance, index)
    order, and for every positive the number of valid non-matches at distance <= its own.
    Returns (ap[m], is_valid[m], first_rank[m], neg_before: list of arrays in gallery-index order)."""
    m, n = distmat.shape
    ap = np.zeros(m)
    valid_q = np.zeros(m, dtype=np.uint8)
    first = np.full(m, -1, dtype=np.int32)
    neg_before = []
    gidx = np.arange(n)
    for i in range(m):
        same = gallery_ids == query_ids[i]
        pos = same & (gallery_cams != query_cams[i])
        keep = ~(same & ~pos)
        if not pos.any():
            neg_before.append(np.zeros(0, dtype=np.int64))
            continue
        d = distmat[i]
        dv = np.sort(d[keep])
        dp = d[pos]
        n_le = np.searchsorted(dv, dp, side="right")
        p_le = np.searchsorted(np.sort(dp), dp, side="right")
        ap[i] = float(np.mean(p_le / n_le))
        valid_q[i] = 1
        neg_before.append((n_le - p_le).astype(np.int64))
        pidx = gidx[pos]
        best = np.lexsort((pidx, dp))[0]
        dstar, gstar = dp[best], pidx[best]
        before = keep & ((d < dstar) | ((d == dstar) & (gidx < gstar)))
        first[i] = int(before.sum())
    return ap, valid_q, first, neg_before


def topk_filtered(distmat, query_ids, gallery_ids, query_cams, gallery_cams, k):
    """k nearest valid gallery items per query, ties by gallery index -> (index [m,k], dist [m,k])."""
    m, n = distmat.shape
    idx_out = np.full((m, k), -1, dtype=np.int32)
    d_out = np.full((m, k), np.inf, dtype=np.float32)
    for i in range(m):
        keep = valid_mask(query_ids[i], query_cams[i], gallery_ids, gallery_cams)
        cand = np.nonzero(keep)[0]
        order = cand[np.argsort(distmat[i][cand], kind="stable")][:k]
        idx_out[i, :len(order)] = order
        d_out[i, :len(order)] = distmat[i][order]
    return idx_out, d_out


# ------------------------------------------------------------------------------------
# training-side triplet mining ops.  The reference has no test for them and PairWiseDistance has no
# CPU implementation (pairwise_distance_op.cc:5-24 holds schema + gradient only); these restatements are PINNED to the
# reference's own operators, compiled unmodified against a small interface shim (oracle/build_ref_ops.py, oracle/ref_ops.py):
# fixtures tests/golden/triplet_ref_*.npz written by oracle/make_golden_triplet.py (BatchHard here on the CPU,
# PairWiseDistance on a B200), plus live comparisons wherever the compiled library is present.
# ------------------------------------------------------------------------------------


def pairwise_distance(x):
    """pairwise_distance_op.cu:9-22: Z[p,q] = sum_d (X[p,d]-X[q,d])^2, float32, sequential over d."""
    x = np.asarray(x, dtype=np.float32)
    diff = x[:, None, :] - x[None, :, :]
    return np.sum(diff * diff, axis=2, dtype=np.float32)


def pairwise_distance_grad(x, dz):
    """pairwise_distance_op.cu:78-91: for every (p,q): dX[p] += 2 (x_p-x_q) dZ[p,q]; dX[q] -= 2 (x_p-x_q) dZ[p,q]."""
    x = np.asarray(x, dtype=np.float64)
    dz = np.asarray(dz, dtype=np.float64)
    diff = x[:, None, :] - x[None, :, :]                       # [p, q, d]
    g = 2.0 * diff * dz[:, :, None]
    return (g.sum(axis=1) - g.sum(axis=0)).astype(np.float32)


def batch_hard(xdist, labels):
    """batch_hard_op.cc:9-59 (and the index search of :62-123): strict compares in index order.
    Returns ap, an, idx_p, idx_n (-1 where no candidate improved on the initial 0 / FLT_MAX)."""
    xdist = np.asarray(xdist, dtype=np.float32)
    n = xdist.shape[0]
    ap = np.zeros(n, np.float32); an = np.full(n, np.finfo(np.float32).max, np.float32)
    ip = np.full(n, -1, np.int32); inn = np.full(n, -1, np.int32)
    for a in range(n):
        for j in range(n):
            if labels[j] == labels[a]:
                if ap[a] < xdist[a, j]:
                    ap[a], ip[a] = xdist[a, j], j
            else:
                if an[a] > xdist[a, j]:
                    an[a], inn[a] = xdist[a, j], j
    return ap, an, ip, inn


def batch_hard_grad(idx_p, idx_n, dap, dan, stray_writes=False):
    """batch_hard_op.cc:62-123.  The operator stores `dX[a * N + idx]` also when idx stayed -1 (an anchor alone in its class,
    or a batch of one class): that lands on element (a - 1, N - 1) - the last column of the PREVIOUS row - and, for a = 0,
    one float before the buffer.  stray_writes=False (what pps_b200 computes): those stores are left out.
    stray_writes=True: the operator's stores exactly, in its order (the a = 0 store outside the buffer excepted), which is
    what the compiled reference is compared with."""
    n = len(idx_p)
    dx = np.zeros(n * n, np.float32)
    for a in range(n):
        for idx, val in ((idx_p[a], dap[a]), (idx_n[a], dan[a])):
            flat = a * n + int(idx)
            if idx >= 0 or (stray_writes and flat >= 0):
                dx[flat] = val
    return dx.reshape(n, n)


# ------------------------------------------------------------------------------------
# per-combination embedding + concat + normalise: a Caffe2 graph (Conv, SpatialBN, Relu, Concat,
# Normalize of pytorch v1.0.1) with no reference test.  The graph is pinned by executing the reference's own builder
# eagerly (oracle/ref_pool_loader.run_reid_outputs -> tests/golden/embed_*.npz); the operator arithmetic is restated.
# ------------------------------------------------------------------------------------


def reid_embed(pooled, weight, conv_bias, bn_scale, bn_bias, bn_mean, bn_var, eps=1e-5, normalize=True,
               dtype=np.float64):
    """reid_heads.py:34-127 at test time.

    pooled [K, N, C] (blob k = pooled[k] as [N, C, 1, 1]); weight [K, E, C]; the five per-channel vectors [K, E].
    Branch k: z = pooled[k] @ weight[k].T + conv_bias[k]            (Conv 1x1, :41-57)
              y = (z - mean) / sqrt(var + eps) * scale + bias       (SpatialBN is_test, :58-60)
              y = max(y, 0)                                         (Relu, :76)
    Concat(axis=1) of the K branches (:95-101), then x / max(|x|_2, 1e-12) per row (Normalize, :123-127).
    """
    pooled, weight = np.asarray(pooled, dtype=dtype), np.asarray(weight, dtype=dtype)
    K, N, _ = pooled.shape
    outs = []
    for k in range(K):
        z = pooled[k] @ weight[k].T + np.asarray(conv_bias[k], dtype=dtype)
        y = (z - np.asarray(bn_mean[k], dtype)) / np.sqrt(np.asarray(bn_var[k], dtype) + eps)
        y = y * np.asarray(bn_scale[k], dtype) + np.asarray(bn_bias[k], dtype)
        outs.append(np.maximum(y, 0))
    feat = np.concatenate(outs, axis=1)
    if normalize:
        feat = feat / np.maximum(np.sqrt((feat * feat).sum(axis=1, keepdims=True)), 1e-12)
    return feat


# ------------------------------------------------------------------------------------
# k-reciprocal re-ranking — pinned against the reference's own re_ranking
# (tests/golden/rerank_*.npz made by oracle/make_golden_rerank.py)
# ------------------------------------------------------------------------------------


def re_ranking(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3):
    """reid_dataset_evaluator.py:442-519, restated step by step (float32 throughout, like the reference).

    Differences from the reference text: only the first k1 + 1 columns of the argsort are kept (all it reads),
    the inverted index covers gallery rows only (all that survives the final slice), `np.argsort(kind='stable')`
    pins the order of exact ties (the reference's default sort leaves it undefined).
    """
    q_g = np.asarray(q_g_dist, dtype=np.float32)
    nq, ng = q_g.shape
    m = np.concatenate([np.concatenate([np.asarray(q_q_dist, np.float32), q_g], axis=1),
                        np.concatenate([q_g.T, np.asarray(g_g_dist, np.float32)], axis=1)], axis=0)      # :447-452
    m = np.power(m, 2).astype(np.float32)                                                                # :453
    od = np.transpose(1. * m / np.max(m, axis=0)).astype(np.float32)                                     # :454
    n = nq + ng
    rank = np.argsort(od, axis=1, kind="stable")[:, :k1 + 1].astype(np.int32)                            # :456
    kh = int(np.around(k1 / 2.)) + 1
    rows = []                                            # sparse V: (indices ascending, float32 values) per image
    for i in range(n):
        fwd = rank[i, :k1 + 1]
        back = rank[fwd, :k1 + 1]
        kr = fwd[np.where(back == i)[0]]                                                                 # :463-466
        exp = kr
        for cand in kr:
            cf = rank[cand, :kh]
            cb = rank[cf, :kh]
            ck = cf[np.where(cb == cand)[0]]
            if len(np.intersect1d(ck, kr)) > 2. / 3 * len(ck):                                           # :477-481
                exp = np.append(exp, ck)
        exp = np.unique(exp)                                                                             # :485
        w = np.exp(-od[i, exp])                                                                          # :486
        rows.append((exp.astype(np.int64), (1. * w / np.sum(w)).astype(np.float32)))                     # :487
    if k2 != 1:                                                                                          # :489-494
        dense = np.zeros((n, n), dtype=np.float32)
        for i, (idx, val) in enumerate(rows):
            dense[i, idx] = val
        qe = np.zeros_like(dense)
        for i in range(n):
            qe[i, :] = np.mean(dense[rank[i, :k2], :], axis=0)
        dense = qe
    else:
        dense = np.zeros((n, n), dtype=np.float32)
        for i, (idx, val) in enumerate(rows):
            dense[i, idx] = val
    inv = [np.where(dense[nq:, c] != 0)[0] for c in range(n)]                                            # :496-498 (gallery rows)
    out = np.zeros((nq, ng), dtype=np.float32)
    lam = np.float32(lambda_value)
    for i in range(nq):                                                                                  # :502-509
        temp = np.zeros(ng, dtype=np.float32)
        nz = np.where(dense[i, :] != 0)[0]
        for c in nz:
            r = inv[c]
            temp[r] = temp[r] + np.minimum(dense[i, c], dense[nq + r, c])
        jac = 1 - temp / (2. - temp)
        out[i] = jac * (1 - lambda_value) + od[i, nq:] * lambda_value                                    # :511-512
    return out.astype(np.float32)


# ------------------------------------------------------------------------------------
# pooling backward — gradients of the stock Caffe2 operators of the pooling sub-graph, restated (float64):
# AveragePoolGradient (global) spreads dY / (h W); MaxPoolGradient routes dY to the maximal element (first one in
# row-major order on ties); MeanGradient = dY / N to every input; MaxGradient passes dY to every input equal to the output;
# Add passes dY to both.  **parity unpinned** like the forward operator arithmetic (Caffe2 is not importable); the test
# pins it against float64 finite differences of pps_pool instead (the reference's own gradient-check style,
# detectron/tests/test_batch_permutation_op.py:43-50).
# ------------------------------------------------------------------------------------


def pps_pool_grad(x, dy, n_parts=6, split=None, mode="max_ave", combos=None):
    """dX [N, C, H, W] (float64) for dY [N, K, C]."""
    x = np.asarray(x, dtype=np.float64)
    dy = np.asarray(dy, dtype=np.float64)
    N, C, H, W = x.shape
    if split is None:
        split = [H // n_parts] * n_parts
    avg, mx = strip_pools(x, split, np.float64)                      # [n, N, C]
    masks = list(combos) if combos is not None else list(range(1, 1 << n_parts))
    d_avg = np.zeros_like(avg)
    d_max = np.zeros_like(mx)
    for k, m in enumerate(masks):
        parts = [j for j in range(n_parts) if m & (1 << j)]
        g = dy[:, k, :]
        if mode == "max_ave":
            top = np.max(mx[parts], axis=0)
            for j in parts:
                d_avg[j] += g / len(parts)
                d_max[j] += g * (mx[j] == top)
        else:
            top = np.max(avg[parts], axis=0)
            for j in parts:
                d_avg[j] += g * (avg[j] == top)
    dx = np.zeros_like(x)
    r = 0
    for j, h in enumerate(split):
        dx[:, :, r:r + h, :] += (d_avg[j] / (h * W))[:, :, None, None]
        if mode == "max_ave":
            flat = x[:, :, r:r + h, :].reshape(N, C, h * W)
            arg = np.argmax(flat, axis=2)                            # first maximum
            sub = np.zeros((N, C, h * W))
            np.put_along_axis(sub, arg[:, :, None], d_max[j][:, :, None], axis=2)
            dx[:, :, r:r + h, :] += sub.reshape(N, C, h, W)
        r += h
    return dx
