#!/usr/bin/env python
"""Operand-split kernels alone: GB/s (read 4*D + written 2*kpad*planes bytes per row) for the bf16 and the scaled-fp16 split.
    PPS_SPLIT_F16_VARIANT=0|1 python tools/split_bench.py [--rows 19732 --dim 2048]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=19732)
ap.add_argument("--dim", type=int, default=2048)
a = ap.parse_args()
import torch
from pps_b200 import _lib
lib = _lib.load()
x = torch.randn((a.rows, a.dim), device="cuda")
kpad = (a.dim + 63) // 64 * 64
planes = torch.empty(2 * a.rows * kpad, dtype=torch.float16, device="cuda")
sq = torch.empty(2 * a.rows, dtype=torch.float32, device="cuda")
junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {"rows": a.rows, "dim": a.dim, "variant": os.environ.get("PPS_SPLIT_F16_VARIANT", "0")}
for name, arg in (("bf16_2planes", 2), ("f16_scaled", 2 | _lib.SPLIT_F16_SCALED)):
    ts = []
    for it in range(12):
        junk.zero_()                       # flush L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.pps_split_rows(_lib.ptr(x), _lib.DTYPE_F32, a.rows, a.dim, a.dim, arg, _lib.ptr(planes), _lib.ptr(sq),
                                      _lib.stream_ptr()), "split")
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts[2:])[len(ts[2:]) // 2]
    out[name] = {"ms": ms, "GBps": a.rows * (4.0 * a.dim + 4.0 * kpad) / ms / 1e6}
print(json.dumps(out))
