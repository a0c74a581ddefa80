#!/usr/bin/env python
"""k-reciprocal re-ranking at a BASELINE shape (default Market-1501: 3 368 + 19 732 images, 2048-d): device time of
re_ranking_from_features and of its steps, and the re-ranked mAP / CMC next to the plain ones."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pps_b200
from pps_b200 import _lib, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="market1501")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--cpu-images", type=int, default=0, help="also time the oracle port of the reference on this many images")
a = ap.parse_args()
d = synthetic.make_config(a.workload)
q, g = torch.from_numpy(d["q"]).cuda(), torch.from_numpy(d["g"]).cuda()
ids = (d["qid"], d["gid"], d["qcam"], d["gcam"])
plain = pps_b200.rank_eval(q, g, *ids)


def run():
    return pps_b200.re_ranking_from_features(q, g)


out = run()
torch.cuda.synchronize()
n0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    out = run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
launches = (_lib.launch_count() - n0) / a.iters
res = pps_b200.rank_distmat(out, *ids)
line = {"tool": "rerank_bench", "workload": a.workload, "nq": int(q.shape[0]), "ng": int(g.shape[0]), "dim": int(q.shape[1]),
        "ms_per_rerank": ms, "gpu_launches": launches, "k1": 20, "k2": 6, "lambda": 0.3,
        "mAP_plain": plain.mean_ap(), "cmc1_plain": float(plain.cmc(10, True)[0]),
        "mAP_reranked": res.mean_ap(), "cmc1_reranked": float(res.cmc(10, True)[0])}
if a.cpu_images:
    from oracle import pps_oracle as O
    nq_s = max(1, a.cpu_images * int(q.shape[0]) // (int(q.shape[0]) + int(g.shape[0])))
    ng_s = a.cpu_images - nq_s
    qs, gs = d["q"][:nq_s], d["g"][:ng_s]
    t0 = time.perf_counter()
    O.re_ranking(O.compute_dist(qs, gs), O.compute_dist(qs, qs), O.compute_dist(gs, gs))
    line["cpu_port_s"] = time.perf_counter() - t0
    line["cpu_port_images"] = a.cpu_images
print(json.dumps(line))
