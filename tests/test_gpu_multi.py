"""Sharded-gallery ranking over NCCL on real GPUs (needs >= 2 devices; skipped otherwise).
The merged result must equal the single-GPU result bit for bit (integer counters, same kernels)."""
import os
import socket

import numpy as np
import pytest

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, golden_path, out_dir):
    import torch
    import torch.distributed as dist
    import pps_b200
    from pps_b200 import evaluator
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    d = dict(np.load(golden_path))
    ng = d["g"].shape[0]
    row0, rows = evaluator.gallery_shard(ng, rank, world)
    q = torch.from_numpy(d["q"]).cuda()
    g = torch.from_numpy(d["g"][row0:row0 + rows]).cuda()
    res = pps_b200.rank_eval(q, g, d["qid"], d["gid"], d["qcam"], d["gcam"], topk=12, want_neg_before=True,
                             gallery_offset=row0, group=dist.group.WORLD)
    # the C step path (no top-k, no neg_before): begin -> all-gather -> distance -> all-reduce -> count -> all-reduce -> end
    fast = pps_b200.rank_eval(q, g, d["qid"], d["gid"], d["qcam"], d["gcam"], gallery_offset=row0, group=dist.group.WORLD)
    # sharded AND chunked: thresholds from the compacted same-id rows of each shard
    small = pps_b200.rank_eval(q, g, d["qid"], d["gid"], d["qcam"], d["gcam"], topk=12, gallery_offset=row0,
                               group=dist.group.WORLD, max_block_bytes=int(q.shape[0]) * 256 * 4)
    assert np.array_equal(small.ap, res.ap) and np.array_equal(small.first_rank, res.first_rank)
    assert np.array_equal(small.topk_index, res.topk_index)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), ap=res.ap, valid=res.is_valid, first=res.first_rank,
             neg_before=res.neg_before, ti=res.topk_index, td=res.topk_dist, fast_ap=fast.ap, fast_valid=fast.is_valid,
             fast_first=fast.first_rank)
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["small_mid", "dup_ties", "many_pos"])
def test_sharded_gallery_equals_single_gpu(tmp_path, name):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    import pps_b200
    world = 2
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    mp.spawn(_worker, args=(world, _free_port(), path, str(tmp_path)), nprocs=world, join=True)
    d = dict(np.load(path))
    one = pps_b200.rank_eval(torch.from_numpy(d["q"]).cuda(), torch.from_numpy(d["g"]).cuda(), d["qid"], d["gid"],
                             d["qcam"], d["gcam"], topk=12, want_neg_before=True)
    for r in range(world):
        o = dict(np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)))
        np.testing.assert_array_equal(o["ap"], one.ap)
        np.testing.assert_array_equal(o["valid"], one.is_valid)
        np.testing.assert_array_equal(o["first"], one.first_rank)
        np.testing.assert_array_equal(o["neg_before"][:one.pairs.n_pairs], one.neg_before[:one.pairs.n_pairs])
        np.testing.assert_array_equal(o["ti"], one.topk_index)
        np.testing.assert_array_equal(o["td"], one.topk_dist)
        np.testing.assert_array_equal(o["fast_ap"], one.ap)
        np.testing.assert_array_equal(o["fast_valid"], one.is_valid)
        np.testing.assert_array_equal(o["fast_first"], one.first_rank)
    assert abs(float(np.sum(one.ap)) / np.sum(one.is_valid) - float(d["mAP"])) < (5e-3 if name == "dup_ties" else 1e-6)
