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


def make_distractor_ids(nq, ng, n_real=16932, n_ids=750, n_cams=6, seed=0):
    """Global id / camera arrays of BASELINE configs[3] / configs[4]: ``n_real`` gallery rows scattered over the
    gallery carry Market-like identities, every other row is an id-0 distractor that matches no query."""
    rs = np.random.RandomState(seed)
    qid = rs.randint(1, n_ids + 1, size=nq).astype(np.int64)
    qcam = rs.randint(0, n_cams, size=nq).astype(np.int64)
    gid = np.zeros(ng, dtype=np.int64)
    real_rows = np.sort(rs.permutation(ng)[:min(n_real, ng)])
    gid[real_rows] = rs.randint(1, n_ids + 1, size=len(real_rows))
    gcam = rs.randint(0, n_cams, size=ng).astype(np.int64)
    return qid, qcam, gid, gcam


def iter_features_device(ids, dim, n_ids, sigma, seed, device, dtype, block_rows=65536):
    """Identity-model features generated ON THE DEVICE in row blocks (a 41 GB gallery never exists on the host):
    centers (seed 1234, shared by queries and gallery) + sigma * noise (``seed``), L2-normalised, cast to ``dtype``.
    Yields (row0, block).  The values depend only on (ids, seed, block_rows), so a gallery shard can be generated - or
    regenerated block by block for a check - on its own."""
    import torch
    gen = torch.Generator(device=device).manual_seed(1234)
    centers = torch.randn((n_ids + 1, dim), device=device, generator=gen)
    centers[0] = 0
    g = torch.Generator(device=device).manual_seed(int(seed))
    ids_t = torch.from_numpy(np.ascontiguousarray(ids)).to(device)
    for r0 in range(0, len(ids), block_rows):
        sl = ids_t[r0:r0 + block_rows]
        x = centers[sl] + sigma * torch.randn((len(sl), dim), device=device, generator=g)
        x = x / x.norm(dim=1, keepdim=True)
        yield r0, x.to(dtype)


def make_features_device(ids, dim, n_ids, sigma, seed, device, dtype, block_rows=65536):
    """All blocks of iter_features_device in one [len(ids), dim] tensor."""
    import torch
    out = torch.empty((len(ids), dim), dtype=dtype, device=device)
    for r0, x in iter_features_device(ids, dim, n_ids, sigma, seed, device, dtype, block_rows):
        out[r0:r0 + x.shape[0]] = x
    return out
