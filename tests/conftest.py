import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    return load


GOLDEN_CASES = ["small_mid", "ragged_dim", "many_pos", "dup_ties", "some_invalid"]

POOL_GOLDEN_CASES = ["pool_n6_max_ave", "pool_n6_avg_max", "pool_n5_shipped", "pool_n7_table", "pool_n4_negative",
                     "pool_fpn_test", "pool_fpn_train", "pool_fpn_shared"]


def pool_fixture_expected(d, pool_fn, split_fn):
    """What the pooling of fixture ``d`` must be according to ``pool_fn(x, n, split, mode) -> [N, K, C]``: the list of
    blobs ``add_pps_part_head`` returns (pps_heads.py:83-142), each [N, C, 1, 1]."""
    import numpy as np
    n, mode = int(d["strip_num"]), ("max_ave" if int(d["max_ave"]) else "avg_max")
    levels = [d["x%d" % j] for j in range(int(d["n_levels"]))]
    scales = [float(v) for v in d["spatial_scale"]]

    def level(x, ss):
        y = pool_fn(x, n, split_fn(n, 384, ss), mode)                 # [N, K, C]
        return [y[:, k, :].reshape(y.shape[0], y.shape[2], 1, 1) for k in range(y.shape[1])]

    if not int(d["fpn_on"]):
        return level(levels[0], scales[0])
    if not int(d["train"]):
        return level(levels[0], scales[0])                            # test time: level 0 only (:88-96)
    per_level = [level(x, ss) for x, ss in zip(levels, scales)]
    if not int(d["fpn_shared"]):
        return [b for lv in per_level for b in lv]                    # level-major (:106-117)
    return [np.concatenate([lv[k] for lv in per_level], axis=0) for k in range(len(per_level[0]))]   # Concat(axis=0) (:119-135)
