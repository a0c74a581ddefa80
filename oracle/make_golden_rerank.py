"""Generate tests/golden/rerank_*.npz by running the UNMODIFIED reference `re_ranking`
(reid_dataset_evaluator.py:442-519, through oracle/ref_loader.py) on small seeded inputs:

    python -m oracle.make_golden_rerank

Stored: the features, the three distance matrices the reference's evaluate() feeds it (:165-171, from the
reference's own compute_dist), its re-ranked query x gallery matrix and the scores evaluate() derives from it
(:174-175).  Cases avoid exact distance ties: the reference's argsort leaves their order undefined.
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
    "rerank_small": dict(nq=60, ng=420, dim=96, n_ids=20, n_cams=3, n_distractors=40, sigma=3.0, seed=21),
    "rerank_tiny": dict(nq=9, ng=70, dim=40, n_ids=5, n_cams=2, n_distractors=8, sigma=2.0, seed=22),
    "rerank_wide": dict(nq=130, ng=300, dim=64, n_ids=30, n_cams=4, n_distractors=0, sigma=3.5, seed=23),
}


def main():
    ref = ref_loader.load()
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, kw in CASES.items():
        d = synthetic.make_reid_set(**kw)
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink):
            q_g = ref.compute_dist(d["q"], d["g"], type="euclidean")
            q_q = ref.compute_dist(d["q"], d["q"], type="euclidean")
            g_g = ref.compute_dist(d["g"], d["g"], type="euclidean")
            rr = ref.re_ranking(q_g, q_q, g_g)
            rr_k = ref.re_ranking(q_g, q_q, g_g, k1=7, k2=1, lambda_value=0.5)
            args = dict(query_ids=d["qid"], gallery_ids=d["gid"], query_cams=d["qcam"], gallery_cams=d["gcam"])
            m_ap = ref.mean_ap(distmat=rr, **args)
            cmc = ref.cmc(distmat=rr, topk=10, separate_camera_set=False, single_gallery_shot=False,
                          first_match_break=True, **args)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), q=d["q"], g=d["g"], qid=d["qid"], gid=d["gid"],
                            qcam=d["qcam"], gcam=d["gcam"], q_g=q_g.astype(np.float32), q_q=q_q.astype(np.float32),
                            g_g=g_g.astype(np.float32), rerank=rr.astype(np.float32), rerank_k7_k2_1=rr_k.astype(np.float32),
                            mAP=np.float64(m_ap), cmc=cmc, numpy_version=np.__version__)
        print("%-14s nq=%d ng=%d  re-ranked mAP=%.6f cmc1=%.4f" % (name, kw["nq"], kw["ng"], m_ap, cmc[0]))


if __name__ == "__main__":
    main()
