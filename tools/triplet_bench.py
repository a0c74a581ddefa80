#!/usr/bin/env python
"""Triplet-mining ops at the shapes the reference's training graph feeds them (triplet_loss.py:145-158:
P x K = 16 x 4 ... 64 x 4 images, 128-d embeddings, one op pair per part-combination head): device time per call of
PairWiseDistance, BatchHard on the matrix, the fused mining kernel, and the gradients, with CUDA events."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pps_b200 import triplet


def timed(fn, iters=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3      # us


out = []
for n, d in ((64, 128), (128, 128), (256, 128), (256, 2048)):
    g = torch.Generator(device="cuda").manual_seed(n)
    x = torch.randn((n, d), device="cuda", generator=g)
    labels = (torch.arange(n, device="cuda") // 4).to(torch.int32)
    z = triplet.pairwise_distance(x)
    ap, an, ip, inn = triplet.batch_hard(z, labels, return_indices=True)
    dz = torch.randn((n, n), device="cuda", generator=g)
    row = {"N": n, "D": d,
           "pairwise_distance_us": timed(lambda: triplet.pairwise_distance(x)),
           "batch_hard_us": timed(lambda: triplet.batch_hard(z, labels)),
           "fused_mining_us": timed(lambda: triplet.batch_hard_from_features(x, labels)),
           "pairwise_distance_grad_us": timed(lambda: triplet.pairwise_distance_grad(x, dz)),
           "batch_hard_grad_us": timed(lambda: triplet.batch_hard_grad(ip, inn, ap, an)),
           "torch_cdist_sq_us": timed(lambda: torch.cdist(x, x).pow(2))}
    out.append(row)
print(json.dumps({"tool": "triplet_bench", "note": "per-call device time incl. the Python wrapper's allocations; all shapes are launch-latency bound",
                  "rows": out}))
