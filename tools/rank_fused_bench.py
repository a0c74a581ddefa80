#!/usr/bin/env python
"""Micro-benchmark of pps_dist_rank_tc (distance + counting epilogue, no matrix) against pps_dist_tc (matrix out)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pps_b200 import _lib, evaluator

ap = argparse.ArgumentParser()
ap.add_argument("--m1", type=int, default=3368)
ap.add_argument("--m2", type=int, default=262144)
ap.add_argument("--dim", type=int, default=2048)
ap.add_argument("--dtype", default="fp16")
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--p-caps", default="8,24,48,64")
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
lib = _lib.load()
dt = torch.float16 if a.dtype == "fp16" else torch.float32
prec = _lib.PREC_F16X1 if a.dtype == "fp16" else _lib.PRECISIONS[a.precision]
planes = _lib.PLANES_FOR[prec]
q = torch.randn((a.m1, a.dim), device="cuda")
g = torch.randn((a.m2, a.dim), device="cuda")
q, g = (q / q.norm(dim=1, keepdim=True)).to(dt), (g / g.norm(dim=1, keepdim=True)).to(dt)
sq, sg = evaluator.SplitOperand(q, planes), evaluator.SplitOperand(g, planes)
ldd = (a.m2 + 3) // 4 * 4
out = torch.empty((a.m1, ldd), device="cuda")


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


def plain():
    _lib.check(lib.pps_dist_tc(_lib.ptr(sq.planes), _lib.ptr(sq.sqnorm), a.m1, planes, 0, _lib.ptr(sg.planes),
                               _lib.ptr(sg.sqnorm), a.m2, planes, 0, a.dim, prec, 0, _lib.ptr(out), ldd,
                               _lib.stream_ptr()), "pps_dist_tc")


ms = timed(plain)
fl = 2.0 * a.m1 * a.m2 * a.dim
print("dist_tc   (matrix out)        : %.3f ms  %.0f TFLOP/s alg" % (ms, fl / ms / 1e9))
for p_cap in [int(x) for x in a.p_caps.split(",")]:
    elems = int(lib.pps_rank_tab_elems(a.m1, p_cap))
    groups = elems // (p_cap * 128)
    # thresholds: p_cap sorted values per row around the distance distribution (sqrt(2) for normalised random rows)
    t = (1.414 + 0.03 * torch.randn((groups, 128, p_cap), device="cuda")).sort(dim=2).values
    thr = t.permute(0, 2, 1).contiguous().view(-1)
    cnt = torch.zeros(elems, dtype=torch.int32, device="cuda")
    dstar = thr.view(groups, p_cap, 128)[:, 0, :].contiguous().view(-1)[:a.m1].contiguous()
    gstar = torch.zeros(a.m1, dtype=torch.int32, device="cuda")
    cfirst = torch.zeros(a.m1, dtype=torch.int32, device="cuda")

    def fused():
        _lib.check(lib.pps_dist_rank_tc(_lib.ptr(sq.planes), _lib.ptr(sq.sqnorm), a.m1, planes, 0, _lib.ptr(sg.planes),
                                        _lib.ptr(sg.sqnorm), a.m2, planes, 0, a.dim, prec, 0, 0, p_cap, _lib.ptr(thr),
                                        _lib.ptr(cnt), _lib.ptr(dstar), _lib.ptr(gstar), _lib.ptr(cfirst),
                                        _lib.stream_ptr()), "pps_dist_rank_tc")
    ms = timed(fused)
    tot = int(cnt.view(groups, p_cap, 128).sum().item())
    print("dist_rank_tc p_cap=%2d          : %.3f ms  %.0f TFLOP/s alg   (counted %.3f of the elements)" % (
        p_cap, ms, fl / ms / 1e9, tot / float((a.iters + 2) * a.m1 * a.m2)))
