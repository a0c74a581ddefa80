#!/usr/bin/env python
"""CPU emulation of the split-precision distance (no GPU): how far does the mAP of a split product sit from the
reference's float32 sgemm path, per seed and noise level?

    python tools/split_precision_sim.py --seeds 0 1 2 --sigmas 4 [--config market1501]

Schemes (plane products are formed exactly in float64 and rounded once to float32, i.e. the emulation isolates the
error of the SPLIT; the fp32 accumulation error of the tensor core comes on top and is the same for every scheme):
  bf16x3  x = p0 + p1, planes = bf16 roundings of the residual; p0.p0 + p0.p1 + p1.p0
  f16x3   rows scaled by a power of two so that max|x| lies in [2^13, 2^14), planes = fp16 roundings of the residual
          (22 mantissa bits instead of 16), same three terms, dot rescaled exactly
  exact   float64 product rounded to float32
Everything is compared with oracle.compute_dist (float32 numpy sgemm, the reference's arithmetic) through the
count-based AP (oracle.rank_counts), so the numbers are |mAP(scheme) - mAP(reference path)|.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def bf16_round(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = (u + 0x7FFF + ((u >> 16) & 1)) & np.uint32(0xFFFF0000)
    return r.view(np.float32)


def split_bf16(x):
    p0 = bf16_round(x)
    p1 = bf16_round(x - p0)
    return p0.astype(np.float64), p1.astype(np.float64), np.ones(len(x))


def split_f16(x):
    mx = np.abs(x).max(axis=1)
    e = np.where(mx > 0, 13 - np.floor(np.log2(np.maximum(mx, 1e-300))), 0.0)
    s = np.exp2(e).astype(np.float32)
    xs = x * s[:, None]
    p0 = xs.astype(np.float16).astype(np.float32)
    p1 = (xs - p0).astype(np.float16).astype(np.float32)
    return p0.astype(np.float64), p1.astype(np.float64), 1.0 / s.astype(np.float64)


def dist_from_dot(dot32, an, bn):
    d2 = (np.float32(-2.0) * dot32 + an[:, None]) + bn[None, :]
    np.maximum(d2, 0, out=d2)
    return np.sqrt(d2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="market1501")
    ap.add_argument("--seeds", type=int, nargs="+", default=[0])
    ap.add_argument("--sigmas", type=float, nargs="+", default=[4.0])
    ap.add_argument("--nq", type=int, default=0, help="use only the first nq queries (0 = all)")
    a = ap.parse_args()
    from oracle import pps_oracle as O
    from pps_b200 import synthetic
    out = []
    for sigma in a.sigmas:
        for seed in a.seeds:
            d = synthetic.make_config(a.config, seed=seed, sigma=sigma)
            if a.nq:
                for k in ("q", "qid", "qcam"):
                    d[k] = d[k][:a.nq]
            ids = (d["qid"], d["gid"], d["qcam"], d["gcam"])
            q, g = d["q"], d["g"]
            ref = O.compute_dist(q, g)
            ap_ref, valid, first_ref, _ = O.rank_counts(ref, *ids)
            nv = valid.sum()
            an = np.sum(q.astype(np.float64) ** 2, axis=1).astype(np.float32)
            bn = np.sum(g.astype(np.float64) ** 2, axis=1).astype(np.float32)
            rec = {"config": a.config, "seed": seed, "sigma": sigma, "mAP_ref": float(ap_ref.sum() / nv)}
            dots = {"exact": (q.astype(np.float64) @ g.astype(np.float64).T)}
            for name, fn in (("bf16x3", split_bf16), ("f16x3", split_f16)):
                a0, a1, ia = fn(q)
                b0, b1, ib = fn(g)
                dot = a0 @ b0.T
                dot += a0 @ b1.T
                dot += a1 @ b0.T
                dot *= ia[:, None]
                dot *= ib[None, :]
                dots[name] = dot
            for name, dot in dots.items():
                dm = dist_from_dot(dot.astype(np.float32), an, bn)
                ap_s, _, first_s, _ = O.rank_counts(dm, *ids)
                rec[name] = {"mAP_abs_diff": float(abs(ap_s.sum() / nv - ap_ref.sum() / nv)),
                             "max_ap_diff": float(np.abs(ap_s - ap_ref).max()),
                             "first_rank_differs": int((first_s != first_ref).sum()),
                             "max_rel_dist_err": float(np.max(np.abs(dm - ref) / np.maximum(ref, 1e-6))),
                             "dot_rms_err_vs_exact": float(np.sqrt(np.mean((dot - dots["exact"]) ** 2)))}
            rec["sgemm_dot_rms_err_vs_exact"] = float(np.sqrt(np.mean(((q @ g.T).astype(np.float64) - dots["exact"]) ** 2)))
            print(json.dumps(rec), flush=True)
            out.append(rec)
    return 0


if __name__ == "__main__":
    sys.exit(main())
