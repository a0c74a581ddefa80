"""Generate tests/golden/embed_*.npz by executing the UNMODIFIED ``add_reid_outputs`` of the reference
(detectron/modeling/reid_heads.py:34-127, test mode) through oracle/ref_pool_loader.run_reid_outputs, one image at a
time like the reference's feature extraction (batch size 1; its Reshape(shape=[1, -1]) is per forward pass):

    python -m oracle.make_golden_embed

Stored: the conv5 maps, the head configuration, the parameters of every branch (conv weight / bias, BN scale / bias /
running mean / running variance, named as Caffe2's helpers name them) and ``reid_feature_concat[_norm]`` per image.
The pooled blobs fed to the head come from the pooling restatement (pinned by tests/golden/pool_*.npz).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import pps_oracle as O  # noqa: E402
from oracle import ref_pool_loader as R  # noqa: E402
from pps_b200 import pooling  # noqa: E402

CASES = {
    "embed_n3_c48": dict(n_parts=3, C=48, N=3, normalize=True, seed=41),
    "embed_n6_c64": dict(n_parts=6, C=64, N=2, normalize=True, seed=42),
    "embed_n4_raw": dict(n_parts=4, C=100, N=2, normalize=False, seed=43),
}


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    E = 128
    for name, kw in CASES.items():
        rs = np.random.RandomState(kw["seed"])
        n_parts, C, N = kw["n_parts"], kw["C"], kw["N"]
        K = (1 << n_parts) - 1
        x = np.maximum(rs.randn(N, C, 24, 8), 0).astype(np.float32)
        names = pooling.blob_names(n_parts, "pps")
        w = (rs.randn(K, E, C) * np.sqrt(2.0 / C)).astype(np.float32)
        p = dict(conv_bias=(0.1 * rs.randn(K, E)).astype(np.float32), bn_scale=(1 + 0.2 * rs.randn(K, E)).astype(np.float32),
                 bn_bias=(0.3 * rs.randn(K, E)).astype(np.float32), bn_mean=(0.2 * rs.randn(K, E)).astype(np.float32),
                 bn_var=(0.5 + rs.rand(K, E)).astype(np.float32))
        params = {}
        for k, nm in enumerate(names):
            pre = nm.split("_")[0]                                   # reid_heads.get_prefix
            params[pre + "_conv_w"] = w[k].reshape(E, C, 1, 1)
            params[pre + "_conv_b"] = p["conv_bias"][k]
            params[pre + "_bn_s"], params[pre + "_bn_b"] = p["bn_scale"][k], p["bn_bias"][k]
            params[pre + "_bn_rm"], params[pre + "_bn_riv"] = p["bn_mean"][k], p["bn_var"][k]
        split = [24 // n_parts] * n_parts
        pooled = O.pps_pool(x, n_parts, split=split, mode="max_ave")          # [N, K, C]
        rows, n_ops = [], 0
        for i in range(N):
            blobs = [pooled[i:i + 1, k, :].reshape(1, C, 1, 1) for k in range(K)]
            f, ops = R.run_reid_outputs(blobs, names, params, bpm_dim=E, normalize=kw["normalize"])
            rows.append(f.reshape(-1))
            n_ops = len(ops)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), x=x, weight=w, n_parts=np.int64(n_parts),
                            normalize=np.int64(kw["normalize"]), feature=np.stack(rows), n_ops=np.int64(n_ops), **p)
        print("%-14s K=%d C=%d N=%d -> feature %s, %d graph ops per image" % (name, K, C, N, np.stack(rows).shape, n_ops))


if __name__ == "__main__":
    main()
