#!/usr/bin/env python
"""mAP margin of the split-precision distance ON THE DEVICE against the reference's CPU path, over seeds and noise
levels: Market-1501-shaped sets, |mAP(GPU) - mAP(CPU float32 sgemm + count-based AP)| for every
(seed, sigma, precision), plus the number of first-match ranks that differ (each proven to be a distance tie).

    python tools/map_margin_sweep.py --seeds 0 1 2 3 4 --sigmas 3 4 5 --precisions f16x3 bf16x3 > profiles/r02_map_margin.jsonl
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="market1501")
    ap.add_argument("--seeds", type=int, nargs="+", default=[0, 1, 2, 3, 4])
    ap.add_argument("--sigmas", type=float, nargs="+", default=[3.0, 4.0, 5.0])
    ap.add_argument("--precisions", nargs="+", default=["f16x3", "bf16x3"])
    a = ap.parse_args()
    import torch
    import pps_b200
    from pps_b200 import synthetic
    from oracle import pps_oracle as O
    from oracle import parity as P
    worst = {p: 0.0 for p in a.precisions}
    for sigma in a.sigmas:
        for seed in a.seeds:
            d = synthetic.make_config(a.config, seed=seed, sigma=sigma)
            ids = (d["qid"], d["gid"], d["qcam"], d["gcam"])
            dist = O.compute_dist(d["q"], d["g"])
            ap_ref, valid, first, _ = O.rank_counts(dist, *ids)
            m_ref = float(ap_ref.sum() / valid.sum())
            q, g = torch.from_numpy(d["q"]).cuda(), torch.from_numpy(d["g"]).cuda()
            rec = {"config": a.config, "seed": seed, "sigma": sigma, "mAP_cpu": m_ref}
            for prec in a.precisions:
                res = pps_b200.rank_eval(q, g, *ids, precision=prec)
                moved = P.assert_first_rank_parity(res.first_rank, dist, d["qid"], d["qcam"], d["gid"], d["gcam"])
                diff = abs(res.mean_ap() - m_ref)
                worst[prec] = max(worst[prec], diff)
                rec[prec] = {"mAP_abs_diff": diff, "max_ap_diff": float(np.abs(res.ap - ap_ref).max()),
                             "first_rank_moved_inside_tie": int(moved), "valid_equal": bool(np.array_equal(res.is_valid, valid))}
            print(json.dumps(rec), flush=True)
    print(json.dumps({"summary": "max |mAP_gpu - mAP_cpu| over the sweep", **worst}), flush=True)


if __name__ == "__main__":
    main()
