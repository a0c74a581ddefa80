"""Generate tests/golden/pool_*.npz by executing the UNMODIFIED graph builders of the reference's pooling heads
(detectron/modeling/bpm_heads.py:18-55, pps_heads.py:38-142) through oracle/ref_pool_loader.py:

    python -m oracle.make_golden_pool

Each fixture stores the input maps, the head configuration and what ``add_pps_part_head`` returns: blob names (in the
reference's order), the evaluated blobs and dims_out.  See ref_pool_loader for what this pins (the graph the heads
emit) and what it restates (the arithmetic inside the six stock Caffe2 operators).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_pool_loader as R  # noqa: E402


def maps(shape, seed, relu=True):
    x = np.random.RandomState(seed).randn(*shape)
    return (np.maximum(x, 0) if relu else x - 1.0).astype(np.float32)


CASES = {
    # name: (list of map shapes, kwargs of run_pps_head)
    "pool_n6_max_ave": ([(3, 96, 24, 8)], dict(strip_num=6, max_ave=True)),
    "pool_n6_avg_max": ([(3, 96, 24, 8)], dict(strip_num=6, max_ave=False)),
    "pool_n5_shipped": ([(2, 64, 24, 8)], dict(strip_num=5, max_ave=True)),               # configs/*/pps*.yaml: [5,5,4,5,5]
    "pool_n7_table": ([(2, 40, 24, 8)], dict(strip_num=7, max_ave=False)),
    "pool_n4_negative": ([(2, 33, 24, 6)], dict(strip_num=4, max_ave=True)),               # not post-ReLU, odd C, W % 4 != 0
    "pool_fpn_test": ([(2, 32, 24, 8), (2, 32, 24, 8), (2, 32, 48, 16)],
                      dict(strip_num=6, max_ave=True, fpn_on=True, train=False, spatial_scale=[1 / 16., 1 / 16., 1 / 8.])),
    "pool_fpn_train": ([(2, 32, 24, 8), (2, 32, 24, 8), (2, 32, 48, 16)],
                       dict(strip_num=6, max_ave=True, fpn_on=True, train=True, spatial_scale=[1 / 16., 1 / 16., 1 / 8.])),
    "pool_fpn_shared": ([(2, 32, 24, 8), (3, 32, 24, 8), (1, 32, 48, 16)],
                        dict(strip_num=6, max_ave=True, fpn_on=True, fpn_shared=True, train=True,
                             spatial_scale=[1 / 16., 1 / 16., 1 / 8.])),
}


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    for i, (name, (shapes, kw)) in enumerate(CASES.items()):
        xs = [maps(s, 100 + 10 * i + j, relu=(name != "pool_n4_negative")) for j, s in enumerate(shapes)]
        arg = xs if kw.get("fpn_on") else xs[0]
        names, arrs, dims, ops = R.run_pps_head(arg, **kw)
        store = {"x%d" % j: x for j, x in enumerate(xs)}
        store.update({"y%03d" % k: a for k, a in enumerate(arrs)})
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), names=np.array(names), dims=np.array(dims, np.int64),
                            n_levels=np.int64(len(xs)), n_out=np.int64(len(arrs)), n_ops=np.int64(len(ops)),
                            strip_num=np.int64(kw["strip_num"]), max_ave=np.int64(kw["max_ave"]),
                            fpn_on=np.int64(kw.get("fpn_on", False)), fpn_shared=np.int64(kw.get("fpn_shared", False)),
                            train=np.int64(kw.get("train", False)),
                            spatial_scale=np.array(kw.get("spatial_scale", [1 / 16.]) if kw.get("fpn_on") else [1 / 16.]), **store)
        print("%-18s %d level(s) -> %3d blobs %s ... %s, %d graph ops" % (name, len(xs), len(arrs), names[0], names[-1], len(ops)))


if __name__ == "__main__":
    main()
