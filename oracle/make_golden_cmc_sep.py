"""tests/golden/cmc_sep_<case>.npz: the UNMODIFIED reference cmc (reid_dataset_evaluator.py:283-363) with
separate_camera_set=True on the inputs of the existing fixtures (distance matrix, ids, cameras of tests/golden/<case>.npz).

    python -m oracle.make_golden_cmc_sep

The branch is never reached by the reference's own evaluate() (:35-37 fixes the flag to False) but it is part of the cmc
signature the drop-in keeps.  single_gallery_shot=True is NOT pinned: that branch draws np.random.choice 100 times per query
and calls the removed np.bool (:274-279), i.e. it does not run under the installed numpy."""
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

CASES = ["small_mid", "ragged_dim", "many_pos", "some_invalid"]     # (dup_ties: the reference's unstable argsort decides ties)


def main():
    ref = ref_loader.load()
    gdir = os.path.join(ROOT, "tests", "golden")
    for name in CASES:
        d = dict(np.load(os.path.join(gdir, name + ".npz")))
        args = dict(query_ids=d["qid"], gallery_ids=d["gid"], query_cams=d["qcam"], gallery_cams=d["gcam"])
        with contextlib.redirect_stdout(io.StringIO()):
            fmb = ref.cmc(distmat=d["dist"], topk=10, separate_camera_set=True, first_match_break=True, **args)
            allm = ref.cmc(distmat=d["dist"], topk=20, separate_camera_set=True, first_match_break=False, **args)
            rows, valid = ref.cmc(distmat=d["dist"], topk=10, separate_camera_set=True, first_match_break=True, average=False, **args)
        np.savez_compressed(os.path.join(gdir, "cmc_sep_" + name + ".npz"), cmc_fmb=fmb, cmc_all=allm, cmc_rows=rows,
                            cmc_valid=valid)
        print("%-14s cmc1 %.4f (plain %.4f)  valid %d" % (name, fmb[0], d["cmc_fmb"][0], int(valid.sum())))


if __name__ == "__main__":
    main()
