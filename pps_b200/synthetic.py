"""Deterministic synthetic inputs shaped like the BASELINE configs.

Identity model: ``centers = randn(n_ids, D)``; item = ``centers[id] + sigma * randn(D)``,
L2-normalised (mirrors REID.NORMALIZE_FEATURE), float32.  Gallery ids are uniform over the
identities plus a block of id-0 distractors that matches no query (Market-1501 keeps its
distractors in the gallery: tools/dataset/transform_market1501.py:214).  Person ids start at 1.
"""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # name: nq, ng, dim, n_ids, n_cams, n_distractors
    "cuhk03": dict(nq=1400, ng=5332, dim=2048, n_ids=700, n_cams=2, n_distractors=0),
    "market1501": dict(nq=3368, ng=19732, dim=2048, n_ids=750, n_cams=6, n_distractors=2800),
    "duke": dict(nq=2228, ng=17661, dim=2048, n_ids=702, n_cams=8, n_distractors=2000),
}


def make_reid_set(nq, ng, dim, n_ids, n_cams, n_distractors=0, sigma=4.0, seed=0, dtype=np.float32,
                  chunk_rows=4096):
    """Returns dict(q, g, qid, gid, qcam, gcam).  Features are generated in row chunks so that a
    multi-million-row gallery never needs a float64 copy of itself."""
    rs = np.random.RandomState(seed)
    centers = rs.randn(n_ids, dim)
    qid = rs.randint(1, n_ids + 1, size=nq).astype(np.int64)
    n_real = ng - n_distractors
    gid = np.concatenate([rs.randint(1, n_ids + 1, size=n_real), np.zeros(n_distractors, dtype=np.int64)]).astype(np.int64)
    perm = rs.permutation(ng)
    gid = gid[perm]
    qcam = rs.randint(0, n_cams, size=nq).astype(np.int64)
    gcam = rs.randint(0, n_cams, size=ng).astype(np.int64)

    def feats(ids):
        out = np.empty((len(ids), dim), dtype=dtype)
        for r0 in range(0, len(ids), chunk_rows):
            sl = ids[r0:r0 + chunk_rows]
            base = np.where(sl[:, None] > 0, centers[np.maximum(sl, 1) - 1], 0.0)
            x = base + sigma * rs.randn(len(sl), dim)
            x /= np.linalg.norm(x, axis=1, keepdims=True)
            out[r0:r0 + chunk_rows] = x.astype(dtype)
        return out

    return dict(q=feats(qid), g=feats(gid), qid=qid, gid=gid, qcam=qcam, gcam=gcam)


def make_config(name, seed=0, **overrides):
    cfg = dict(CONFIGS[name])
    cfg.update(overrides)
    return make_reid_set(seed=seed, **cfg)


def make_conv5(n, c=2048, h=24, w=8, seed=0):
    """Post-ReLU conv5 maps (ResNet.py:195 ends res5 with a ReLU, so values are non-negative)."""
    rs = np.random.RandomState(seed)
    return np.maximum(rs.randn(n, c, h, w), 0.0).astype(np.float32)
