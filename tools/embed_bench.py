#!/usr/bin/env python
"""pool -> embed -> normalise at the Market-1501 shape: conv5 maps [n, 2048, 24, 8] -> 63 pooled blobs -> 63 x (2048 -> 128)
embedding + BN + ReLU -> concat [n, 8064] -> L2 normalise.  Device time per stage and roofline fractions."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pps_b200
from pps_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=1024)
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
n, C, K, E = a.images, 2048, 63, 128
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((n, C, 24, 8), device="cuda", generator=g).clamp_min_(0)
w = torch.randn((K, E, C), device="cuda", generator=g) * (2.0 / C) ** 0.5
alpha = (1 + 0.1 * torch.randn((K, E), device="cuda", generator=g))
beta = 0.1 * torch.randn((K, E), device="cuda", generator=g)
head = pps_b200.ReidEmbedHead(w, alpha, beta, precision=a.precision)
pooled = torch.empty((K, n, C), device="cuda")
feat = torch.empty((n, K * E), device="cuda")


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters


ms_pool = timed(lambda: pps_b200.pps_pool(x, n_parts=6, mode="max_ave", layout="knc", out=pooled))
ms_embed = timed(lambda: head(pooled, normalize=False, out=feat))
ms_all = timed(lambda: head(pps_b200.pps_pool(x, n_parts=6, mode="max_ave", layout="knc", out=pooled), normalize=True, out=feat))
ms_fused = timed(lambda: pps_b200.embed_maps(head, x, n_parts=6, mode="max_ave", normalize=True, out=feat))
ms_norm = timed(lambda: pps_b200.l2_normalize_rows(feat, out=feat))
terms = {"bf16x1": 1, "bf16x3": 3, "bf16x6": 6}[a.precision]
fl = 2.0 * n * K * C * E
planes = {"bf16x1": 1, "bf16x3": 2, "bf16x6": 3}[a.precision]
split_bytes = K * n * C * (4 + 2 * planes)
print(json.dumps({
    "tool": "embed_bench", "images": n, "precision": a.precision,
    "pool_ms": ms_pool, "pool_gbs": n * (4.0 * C * 24 * 8 + 4.0 * K * C) / ms_pool / 1e6,
    "embed_ms_incl_split": ms_embed, "embed_algorithmic_tflops": fl / ms_embed / 1e9, "embed_issued_tflops": fl * terms / ms_embed / 1e9,
    "split_bytes": split_bytes, "normalize_ms": ms_norm, "normalize_gbs": 2.0 * n * K * E * 4 / ms_norm / 1e6,
    "pipeline_ms": ms_all, "images_per_s": n / (ms_all * 1e-3),
    "pipeline_planes_ms": ms_fused, "images_per_s_planes": n / (ms_fused * 1e-3),
    "pipeline_planes_gbs": n * (4.0 * C * 24 * 8 + 2.0 * planes * K * C + 4.0 * K * E) / ms_fused / 1e6,
    "peaks": {"hbm_gbs": peaks.get("hbm_gbs"), "bf16_tflops_sustained": peaks.get("bf16_tflops_sustained")}}))
