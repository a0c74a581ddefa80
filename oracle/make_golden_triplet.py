#!/usr/bin/env python
"""TEST INFRASTRUCTURE: fixtures for the triplet-mining ops written by the REFERENCE'S OWN operators (compiled unmodified,
oracle/build_ref_ops.py):

    python oracle/make_golden_triplet.py                 # here (no GPU): tests/golden/triplet_ref_batch_hard.npz
    python oracle/make_golden_triplet.py --cuda OUT.npz  # on a B200: PairWiseDistance / ...Gradient outputs -> OUT.npz
                                                         # (committed as tests/golden/triplet_ref_pairwise.npz)

Inputs are seeded; the files hold inputs and the reference's outputs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_ops  # noqa: E402

BATCH_HARD_CASES = [(64, 16), (256, 64), (37, 3), (8, 8), (5, 1), (1, 1), (96, 96)]
PAIRWISE_CASES = [(64, 128), (256, 128), (33, 7), (1, 5), (100, 300)]


def batch_hard_inputs(n, n_ids, seed):
    rs = np.random.RandomState(seed)
    x = rs.randn(n, 32).astype(np.float32)
    labels = rs.randint(0, n_ids, size=n).astype(np.int32)
    diff = x[:, None, :] - x[None, :, :]
    xd = np.sum(diff * diff, axis=2, dtype=np.float32)
    xd[:, ::5] = np.round(xd[:, ::5])            # exact ties: the first index must win
    return xd, labels, rs.randn(n).astype(np.float32), rs.randn(n).astype(np.float32)


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--cuda":
        import torch
        out = {}
        for i, (n, d) in enumerate(PAIRWISE_CASES):
            rs = np.random.RandomState(100 + i)
            x = rs.randn(n, d).astype(np.float32)
            dz = rs.randn(n, n).astype(np.float32)
            xt, dzt = torch.from_numpy(x).cuda(), torch.from_numpy(dz).cuda()
            out["x%d" % i], out["dz%d" % i] = x, dz
            out["z%d" % i] = ref_ops.pairwise_distance(xt).cpu().numpy()
            out["dx%d" % i] = ref_ops.pairwise_distance_grad(xt, dzt).cpu().numpy()
            torch.cuda.synchronize()
        np.savez_compressed(sys.argv[2], **out)
        print("wrote", sys.argv[2])
        return
    out = {}
    for i, (n, n_ids) in enumerate(BATCH_HARD_CASES):
        xd, labels, dap, dan = batch_hard_inputs(n, n_ids, 10 + i)
        ap, an = ref_ops.batch_hard(xd, labels)
        dx, _ = ref_ops.batch_hard_grad(xd, labels, dap, dan)
        out.update({"xd%d" % i: xd, "labels%d" % i: labels, "dap%d" % i: dap, "dan%d" % i: dan,
                    "ap%d" % i: ap, "an%d" % i: an, "dx%d" % i: dx})
    path = os.path.join(ROOT, "tests", "golden", "triplet_ref_batch_hard.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
