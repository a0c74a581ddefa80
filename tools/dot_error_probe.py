#!/usr/bin/env python
"""How accurate is the tensor-core dot product, and where does its error come from?

Market-shaped features (a subset), every precision: rms / max error of q.g against the float64 product, for
  * the full split product as the distance kernel accumulates it (all terms into ONE fp32 TMEM accumulator),
  * the leading plane alone against the float64 product OF THAT PLANE (pure accumulation error of K / 16 MMA steps),
  * numpy's float32 sgemm (what the reference runs).
    python tools/dot_error_probe.py [--nq 512 --ng 8192]
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
    ap.add_argument("--nq", type=int, default=512)
    ap.add_argument("--ng", type=int, default=8192)
    ap.add_argument("--dim", type=int, default=2048)
    ap.add_argument("--relu", action="store_true", help="non-negative features (post-ReLU embeddings): large common mean")
    a = ap.parse_args()
    import torch
    from pps_b200 import _lib, evaluator, synthetic
    d = synthetic.make_reid_set(nq=a.nq, ng=a.ng, dim=a.dim, n_ids=200, n_cams=6, n_distractors=a.ng // 8, sigma=4.0, seed=0)
    q, g = d["q"], d["g"]
    if a.relu:
        q, g = np.abs(q), np.abs(g)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        g /= np.linalg.norm(g, axis=1, keepdims=True)
    exact = q.astype(np.float64) @ g.astype(np.float64).T
    out = {"shape": [a.nq, a.ng, a.dim], "relu": a.relu, "dot_rms": float(np.sqrt(np.mean(exact ** 2)))}
    err = lambda x: {"rms": float(np.sqrt(np.mean((x - exact) ** 2))), "max": float(np.abs(x - exact).max()),
                     "mean_signed": float(np.mean(x - exact))}
    out["numpy_sgemm"] = err((q @ g.T).astype(np.float64))
    lib = _lib.load()
    tq, tg = torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda()
    res = torch.empty((a.nq, a.ng), dtype=torch.float32, device="cuda")
    for prec in ("bf16x1", "bf16x3", "bf16x6", "f16x3"):
        code = _lib.PRECISIONS[prec]
        sq = evaluator.SplitOperand(tq, _lib.PLANES_FOR[code], prec == "f16x3")
        sg = evaluator.SplitOperand(tg, _lib.PLANES_FOR[code], prec == "f16x3")
        evaluator.dist_block(sq, sg, code, res, flags=_lib.DIST_DOT)
        out[prec] = err(res.cpu().numpy().astype(np.float64))
        if prec == "f16x3":
            # leading plane alone through the single-pass fp16 kernel: accumulation error of K/16 steps, nothing else
            _lib.check(lib.pps_dist_tc(_lib.ptr(sq.planes), _lib.ptr(sq.sqnorm), a.nq, 2, 0, _lib.ptr(sg.planes), _lib.ptr(sg.sqnorm),
                                       a.ng, 2, 0, a.dim, _lib.PREC_F16X1, _lib.DIST_DOT, _lib.ptr(res), a.ng, _lib.stream_ptr()),
                       "pps_dist_tc")
            kpad = (a.dim + 63) // 64 * 64
            hq = sq.planes.view(torch.float16)[:a.nq * kpad].view(a.nq, kpad).double().cpu().numpy()
            hg = sg.planes.view(torch.float16)[:a.ng * kpad].view(a.ng, kpad).double().cpu().numpy()
            plane_exact = hq @ hg.T
            got = res.cpu().numpy().astype(np.float64)
            iq = sq.sqnorm[a.nq:2 * a.nq].double().cpu().numpy()
            ig = sg.sqnorm[a.ng:2 * a.ng].double().cpu().numpy()
            sc = iq[:, None] * ig[None, :]
            e = (got - plane_exact) * sc
            out["f16_leading_plane_accumulation_only"] = {"rms": float(np.sqrt(np.mean(e ** 2))), "max": float(np.abs(e).max()),
                                                          "mean_signed": float(np.mean(e)),
                                                          "mean_signed_rel_to_dot": float(np.mean(e / np.maximum(np.abs(plane_exact * sc), 1e-30)))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
