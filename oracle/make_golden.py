"""Generate tests/golden/*.npz by running the UNMODIFIED reference functions
(oracle/ref_loader.py) on small seeded inputs.  Run in the authoring container:

    python -m oracle.make_golden

Each fixture stores the inputs (so nothing depends on RNG reproducibility) and the reference's
outputs: distance matrix, mAP, per-query AP, CMC (both first_match_break settings).
Environment at generation time is recorded in the fixture (numpy / scikit-learn versions):
the mAP definition follows the installed scikit-learn (>= 0.19 step-wise AP).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from pps_b200 import synthetic  # noqa: E402

CASES = {
    # name: generator kwargs
    "small_mid": dict(nq=96, ng=700, dim=128, n_ids=24, n_cams=3, n_distractors=60, sigma=3.0, seed=11),
    "ragged_dim": dict(nq=37, ng=301, dim=100, n_ids=9, n_cams=2, n_distractors=0, sigma=2.5, seed=12),
    "many_pos": dict(nq=20, ng=900, dim=64, n_ids=3, n_cams=4, n_distractors=10, sigma=2.5, seed=13),
    "dup_ties": dict(nq=40, ng=400, dim=96, n_ids=12, n_cams=3, n_distractors=40, sigma=2.5, seed=15, duplicate=120),
    "some_invalid": dict(nq=50, ng=200, dim=72, n_ids=40, n_cams=2, n_distractors=100, sigma=2.0, seed=14),
}


def main():
    import sklearn
    ref = ref_loader.load()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, kw in CASES.items():
        kw = dict(kw)
        dup = kw.pop("duplicate", 0)
        d = synthetic.make_reid_set(**kw)
        if dup:
            # exact distance ties: the last `dup` gallery rows repeat the first `dup` features
            # (ids / cameras stay their own), so AP must be tie-grouped to match
            d["g"][-dup:] = d["g"][:dup]
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink):
            dist = ref.compute_dist(d["q"], d["g"], type="euclidean")
            args = dict(query_ids=d["qid"], gallery_ids=d["gid"], query_cams=d["qcam"], gallery_cams=d["gcam"])
            m_ap = ref.mean_ap(distmat=dist, **args)
            aps, valid = ref.mean_ap(distmat=dist, average=False, **args)
            cmc_fmb = ref.cmc(distmat=dist, topk=10, separate_camera_set=False, single_gallery_shot=False,
                              first_match_break=True, **args)
            cmc_all = ref.cmc(distmat=dist, topk=20, first_match_break=False, **args)
            cmc_rows, cmc_valid = ref.cmc(distmat=dist, topk=10, first_match_break=True, average=False, **args)
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"),
            q=d["q"], g=d["g"], qid=d["qid"], gid=d["gid"], qcam=d["qcam"], gcam=d["gcam"],
            dist=dist.astype(np.float32), mAP=np.float64(m_ap), aps=aps, valid=valid,
            cmc_fmb=cmc_fmb, cmc_all=cmc_all, cmc_rows=cmc_rows, cmc_valid=cmc_valid,
            numpy_version=np.__version__, sklearn_version=sklearn.__version__)
        print("%-14s nq=%d ng=%d dim=%d  mAP=%.6f  cmc1=%.4f valid=%d/%d" % (
            name, kw["nq"], kw["ng"], kw["dim"], m_ap, cmc_fmb[0], int(valid.sum()), kw["nq"]))


if __name__ == "__main__":
    main()
