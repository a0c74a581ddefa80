"""Generate tests/golden/evaluate_*.npz by running the UNMODIFIED ``evaluate`` of the reference
(detectron/datasets/reid_dataset_evaluator.py:29-209, through oracle/ref_loader.py) on a small synthetic roidb:

    python -m oracle.make_golden_evaluate

Two runs per case: cfg.REID.RERANK = False (single-query scores, :104-125) and True (the returned scores are those of
the k-reciprocal re-ranked distances, :161-175).  The multi-query branch (:131-159) cannot be run unmodified under
Python 3 (``zip(*keys)[0]`` at :152 is a Python 2 idiom), so the fixtures carry no mark-2 images.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from pps_b200 import synthetic  # noqa: E402

CASES = {
    "evaluate_small": dict(nq=60, ng=420, dim=96, n_ids=20, n_cams=3, n_distractors=40, sigma=3.0, seed=21),
    "evaluate_mixed_order": dict(nq=35, ng=260, dim=64, n_ids=12, n_cams=4, n_distractors=20, sigma=2.5, seed=27),
}


class _Dataset:
    def __init__(self, roidb):
        self._roidb = roidb

    def get_roidb(self, gt=True):
        return self._roidb


def main():
    ref = ref_loader.load()
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, kw in CASES.items():
        d = synthetic.make_reid_set(**kw)
        feats = np.concatenate([d["q"], d["g"]], 0)
        ids = np.concatenate([d["qid"], d["gid"]])
        cams = np.concatenate([d["qcam"], d["gcam"]])
        marks = np.concatenate([np.zeros(len(d["qid"]), np.int64), np.ones(len(d["gid"]), np.int64)])
        if name == "evaluate_mixed_order":                       # query / gallery images interleaved in the roidb
            perm = np.random.RandomState(5).permutation(len(ids))
            feats, ids, cams, marks = feats[perm], ids[perm], cams[perm], marks[perm]
        images = ["/data/x/%08d_%04d_%08d.jpg" % (int(p), int(c), k) for k, (p, c) in enumerate(zip(ids, cams))]
        roidb = [{"image": im, "mark": int(m)} for im, m in zip(images, marks)]
        out = {}
        for rr in (False, True):
            ref.cfg.REID = types.SimpleNamespace(RERANK=rr, VIS=False)   # VIS (:108): the matplotlib visualisation, config.py default False
            with contextlib.redirect_stdout(io.StringIO()):
                mAP, cmc, mq_mAP, mq_cmc = ref.evaluate(_Dataset(roidb), feats, None)
            assert mq_mAP is None and mq_cmc is None
            out["mAP_rerank" if rr else "mAP"] = np.float64(mAP)
            out["cmc_rerank" if rr else "cmc"] = np.asarray(cmc, np.float64)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), feats=feats.astype(np.float32), images=np.array(images),
                            marks=marks, **out)
        print("%-22s images=%d  mAP=%.6f cmc1=%.4f | re-ranked mAP=%.6f cmc1=%.4f" % (
            name, len(images), out["mAP"], out["cmc"][0], out["mAP_rerank"], out["cmc_rerank"][0]))


if __name__ == "__main__":
    main()
