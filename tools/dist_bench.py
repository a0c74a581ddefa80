#!/usr/bin/env python
"""Micro-benchmark of pps_dist_tc alone (operands already split): ms, algorithmic and issued TFLOP/s."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pps_b200 import _lib, evaluator

ap = argparse.ArgumentParser()
ap.add_argument("--m1", type=int, default=3368)
ap.add_argument("--m2", type=int, default=19732)
ap.add_argument("--dim", type=int, default=2048)
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--kernel", default="2cta", choices=["2cta", "1cta"])
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--debug-flags", type=lambda x: int(x, 0), default=0)
ap.add_argument("--dtype", default="fp32", choices=["fp32", "fp16"], help="fp16: inputs are fp16 rows (single exact-product pass)")
a = ap.parse_args()
lib = _lib.load()
prec = _lib.PREC_F16X1 if a.dtype == "fp16" else _lib.PRECISIONS[a.precision]
planes = _lib.PLANES_FOR[prec]
terms = {1: 1, 3: 3, 6: 6, 16: 1, 19: 3}[prec]
q = torch.randn((a.m1, a.dim), device="cuda")
g = torch.randn((a.m2, a.dim), device="cuda")
if a.dtype == "fp16":
    q, g = q.half(), g.half()
scaled = prec == _lib.PREC_F16X3
sq, sg = evaluator.SplitOperand(q, planes, scaled), evaluator.SplitOperand(g, planes, scaled)
ldd = (a.m2 + 3) // 4 * 4
out = torch.empty((a.m1, ldd), device="cuda")
flags = (_lib.DIST_KERNEL_1CTA if a.kernel == "1cta" else 0) | a.debug_flags


def run():
    _lib.check(lib.pps_dist_tc(_lib.ptr(sq.planes), _lib.ptr(sq.sqnorm), a.m1, planes, 0, _lib.ptr(sg.planes),
                               _lib.ptr(sg.sqnorm), a.m2, planes, 0, a.dim, prec, flags, _lib.ptr(out), ldd,
                               _lib.stream_ptr()), "pps_dist_tc")


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
alg = 2.0 * a.m1 * a.m2 * a.dim / ms / 1e9
print("dist %s %s %dx%dx%d: %.3f ms  %.1f TFLOP/s algorithmic, %.1f issued" % (
    a.kernel, "fp16-inputs" if a.dtype == "fp16" else a.precision, a.m1, a.m2, a.dim, ms, alg, alg * terms))
