"""Tie proofs for the ranking parity checks - TEST INFRASTRUCTURE (used by tests/ and by bench.py's oracle-subset
checks, never by the product).

BASELINE.json: "junk masks, ranks and CMC must be bit-exact apart from ties within the stated tolerance"
(|delta d| <= 1e-4 d).  The GPU path and the reference's float32 sgemm path round the same distances
differently in the last bits, so an integer output may differ from the oracle's ONLY where two oracle distances
are closer than the tolerance.  These helpers do not accept "most queries agree": every difference has to be
explained by such a tie, otherwise the test fails.
"""
import numpy as np

TIE_RTOL = 1e-4          # north_star: distances must match to 1e-4 relative


def _pair_masks(qid, qcam, gid, gcam):
    same = gid == qid
    junk = same & (gcam == qcam)
    return same & ~junk, ~junk          # positives, valid (kept) items


def assert_first_rank_parity(first_rank, dist_oracle, qid, qcam, gid, gcam, rtol=TIE_RTOL, what=""):
    """first_rank[i] (0-based position of the first correct match in the valid-filtered ranking, -1 = invalid query)
    against the oracle distance rows dist_oracle[i]: the rank may differ from the oracle's only by items whose oracle
    distance lies within rtol of the nearest positive's.  Returns the number of queries whose rank differs."""
    moved = 0
    for i in range(len(first_rank)):
        pos, keep = _pair_masks(qid[i], qcam[i], gid, gcam)
        d = dist_oracle[i].astype(np.float64)
        if not pos.any():
            assert first_rank[i] == -1, "%s query %d: oracle says invalid, got rank %d" % (what, i, first_rank[i])
            continue
        dstar = d[pos].min()
        lo = int(np.sum(keep & (d < dstar * (1 - rtol))))                  # certainly before the first match
        hi = int(np.sum(keep & (d <= dstar * (1 + rtol)))) - 1              # everything that could precede it
        pidx = np.nonzero(pos)[0]
        gstar = pidx[np.lexsort((pidx, d[pos]))[0]]
        exact = int(np.sum(keep & ((d < dstar) | ((d == dstar) & (np.arange(len(d)) < gstar)))))
        r = int(first_rank[i])
        assert lo <= r <= hi, ("%s query %d: first-match rank %d outside the tie window [%d, %d] of the oracle "
                               "(oracle rank %d)" % (what, i, r, lo, hi, exact))
        moved += int(r != exact)
    return moved


def assert_topk_parity(topk_index, topk_dist, dist_oracle, qid, qcam, gid, gcam, rtol=TIE_RTOL, what=""):
    """topk_index / topk_dist [m, k] against the oracle rows: distances within rtol of the oracle's k smallest valid
    distances position by position; an index may differ from the oracle's at a position only if the two items' oracle
    distances are within rtol (a tie); no junk item, no duplicates.  Returns the number of differing positions."""
    differing = 0
    m, k = topk_index.shape
    for i in range(m):
        _, keep = _pair_masks(qid[i], qcam[i], gid, gcam)
        d = dist_oracle[i]
        cand = np.nonzero(keep)[0]
        dc = d[cand]
        if len(dc) > 8 * k:                                   # long rows: select, then order only what was selected
            kth = np.partition(dc, k - 1)[k - 1]
            pool = np.nonzero(dc <= kth)[0]
        else:
            pool = np.arange(len(dc))
        order = cand[pool[np.lexsort((pool, dc[pool]))]][:k]  # (distance, gallery index) order, like a stable argsort
        n = len(order)
        got = topk_index[i, :n]
        assert np.all(topk_index[i, n:] == -1), "%s query %d: more entries than valid gallery items" % (what, i)
        assert np.all(keep[got]), "%s query %d: a junk item in the top-k" % (what, i)
        assert len(np.unique(got)) == n, "%s query %d: duplicate index in the top-k" % (what, i)
        want_d = d[order].astype(np.float64)
        np.testing.assert_allclose(topk_dist[i, :n], want_d, rtol=rtol, atol=1e-7,
                                   err_msg="%s query %d: top-k distances" % (what, i))
        diff = np.nonzero(got != order)[0]
        if len(diff):
            a, b = d[got[diff]].astype(np.float64), want_d[diff]
            assert np.all(np.abs(a - b) <= rtol * np.maximum(a, b) + 1e-12), \
                "%s query %d: top-k index differs from the oracle outside a distance tie" % (what, i)
            differing += len(diff)
    return differing


def assert_counts_exact_on_own_distances(res, dist_gpu, qid, qcam, gid, gcam, O, sel=None, topk=0):
    """The integer machinery, separated from the last-bit distance noise: the oracle's count-based restatement run on
    the GPU's OWN distance rows must reproduce the GPU's valid flags and first-match ranks bit for bit, its AP to
    1e-12 and its top-k exactly."""
    sel = np.arange(len(res.ap)) if sel is None else np.asarray(sel)
    ap, valid, first, _ = O.rank_counts(dist_gpu, qid[sel], gid, qcam[sel], gcam)
    np.testing.assert_array_equal(res.is_valid[sel], valid)
    np.testing.assert_array_equal(res.first_rank[sel], first)
    np.testing.assert_allclose(res.ap[sel], ap, rtol=0, atol=1e-12)
    if topk:
        ti, td = O.topk_filtered(dist_gpu, qid[sel], gid, qcam[sel], gcam, topk)
        np.testing.assert_array_equal(res.topk_index[sel], ti)
        np.testing.assert_array_equal(res.topk_dist[sel], td)


def subset_check(res, sel, dist_oracle, qid, qcam, gid, gcam, topk=0, map_tol=1e-6, rtol=TIE_RTOL):
    """bench.py's oracle-subset check: the GPU result ``res`` (RankResult over ALL queries) on the query slice ``sel``
    against oracle distance rows ``dist_oracle`` [len(sel), ng] (reference arithmetic, full gallery).  Count-based (no
    argsort of multi-million-column rows).  Returns a dict; ``ok`` is False if anything is outside north_star's
    tolerances: valid flags identical, mean AP of the slice within ``map_tol``, every first-match rank / top-k index
    that differs from the oracle's proven to be a distance tie within ``rtol``."""
    sel = np.asarray(sel)
    out = {"queries": int(len(sel)), "ok": True, "errors": []}
    ap = np.zeros(len(sel))
    valid = np.zeros(len(sel), dtype=np.uint8)
    for r, i in enumerate(sel):
        pos, keep = _pair_masks(qid[i], qcam[i], gid, gcam)
        if not pos.any():
            continue
        d = dist_oracle[r]
        dp = d[pos]
        dk = d[keep]
        n_le = np.array([np.count_nonzero(dk <= t) for t in dp])
        p_le = np.searchsorted(np.sort(dp), dp, side="right")
        ap[r] = float(np.mean(p_le / n_le))
        valid[r] = 1
    out["valid"] = int(valid.sum())
    if not np.array_equal(res.is_valid[sel], valid):
        out["ok"] = False
        out["errors"].append("valid flags differ")
    nv = max(int(valid.sum()), 1)
    out["mAP_cpu_subset"] = float(ap.sum() / nv)
    out["mAP_gpu_subset"] = float(res.ap[sel].sum() / nv)
    out["mAP_cpu_subset_diff"] = abs(out["mAP_cpu_subset"] - out["mAP_gpu_subset"])
    out["max_abs_ap_diff"] = float(np.abs(res.ap[sel] - ap).max()) if len(sel) else 0.0
    if out["mAP_cpu_subset_diff"] > map_tol:
        out["ok"] = False
        out["errors"].append("mean AP of the slice differs by %.3g" % out["mAP_cpu_subset_diff"])
    try:
        moved = assert_first_rank_parity(res.first_rank[sel], dist_oracle, qid[sel], qcam[sel], gid, gcam, rtol=rtol)
        out["first_rank_equal"] = int(len(sel) - moved)
        out["first_rank_moved_inside_tie"] = int(moved)
    except AssertionError as e:
        out["ok"] = False
        out["errors"].append(str(e)[:200])
    if topk and res.topk_index is not None:
        try:
            differing = assert_topk_parity(res.topk_index[sel], res.topk_dist[sel], dist_oracle, qid[sel], qcam[sel], gid, gcam,
                                           rtol=rtol)
            out["topk_index_equal_frac"] = 1.0 - differing / float(max(len(sel) * topk, 1))
        except AssertionError as e:
            out["ok"] = False
            out["errors"].append(str(e)[:200])
    return out
