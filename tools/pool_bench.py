#!/usr/bin/env python
"""Micro-benchmark of pps_pool_fwd: GB/s of algorithmic bytes (4*C*H*W in + 4*K*C out per image)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pps_b200

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=1024)
ap.add_argument("--c", type=int, default=2048)
ap.add_argument("--h", type=int, default=24)
ap.add_argument("--w", type=int, default=8)
ap.add_argument("--parts", type=int, default=6)
ap.add_argument("--mode", default="max_ave")
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
x = torch.randn((a.images, a.c, a.h, a.w), device="cuda").clamp_min_(0)
K = (1 << a.parts) - 1
y = torch.empty((a.images, K, a.c), device="cuda")
split = [a.h // a.parts] * a.parts
for _ in range(3):
    pps_b200.pps_pool(x, n_parts=a.parts, split=split, mode=a.mode, out=y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    pps_b200.pps_pool(x, n_parts=a.parts, split=split, mode=a.mode, out=y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
b = a.images * (4.0 * a.c * a.h * a.w + 4.0 * K * a.c)
print("pool [%d,%d,%d,%d] n=%d %s: %.3f ms  %.1f GB/s  %.0f images/s" % (a.images, a.c, a.h, a.w, a.parts, a.mode, ms,
                                                                         b / ms / 1e6, a.images / ms * 1e3))
