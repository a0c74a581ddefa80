"""world_size-2 `gloo` tests of the host-side logic of the sharded ranking path (run on CPU).

Two exchange protocols exist: the launch-by-launch Python path (three sum-all-reduces + evaluator.merge_topk_keys) and the
C pass (csrc/pass.cu: all-gather of the thresholds, all-gather of the packed [top-k keys | cnt_first | cnt_le | flags]
buffers, own reduce / merge).  Both are exercised below with the kernels emulated by the oracle.

The CUDA kernels cannot run here, so each rank emulates what its kernels produce for its gallery shard
with the oracle (pair distances of the positives that live on the shard, integer <=-counts, first-match
counters, packed top-k keys) and then goes through the product's own exchange + aggregation code:
all-reduce SUM of thresholds and counters, evaluator.merge_topk_keys, RankResult averaging.
The merged result must equal the unsharded oracle bit for bit (counters are integers)."""
import os
import socket

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import pps_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, golden_path, out_dir):
    import torch
    import torch.distributed as dist
    from pps_b200 import evaluator
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = dict(np.load(golden_path))
    D = d["dist"]
    nq, ng = D.shape
    row0, rows = evaluator.gallery_shard(ng, rank, world)
    p = evaluator.PairLists(d["qid"], d["qcam"], d["gid"], d["gcam"])       # global pair lists on every rank
    n = p.n_pairs
    # sweep 1 (emulated kernel: pps_rank_gather on the local block) + exchange
    pair_d = torch.zeros(n, dtype=torch.float32)
    local = (p.g[:n] >= row0) & (p.g[:n] < row0 + rows)
    pair_d[torch.from_numpy(np.nonzero(local)[0])] = torch.from_numpy(D[p.q[:n][local], p.g[:n][local]])
    dist.all_reduce(pair_d, op=dist.ReduceOp.SUM)
    thr = pair_d.numpy()
    # sweep 2 (emulated kernel: pps_rank_count on the local block) + exchange
    blk = D[:, row0:row0 + rows]
    cnt_le = np.zeros(n, dtype=np.int32)
    cnt_first = np.zeros(nq, dtype=np.int32)
    for i in range(nq):
        e = np.arange(p.off[i], p.off[i + 1])
        pos = e[p.pos[e] == 1]
        if len(pos) == 0:
            continue
        cnt_le[pos] = (blk[i][None, :] <= thr[pos][:, None]).sum(axis=1)
        best = pos[np.lexsort((p.g[pos], thr[pos]))[0]]
        cols = np.arange(row0, row0 + rows)
        cnt_first[i] = int(((blk[i] < thr[best]) | ((blk[i] == thr[best]) & (cols < p.g[best]))).sum())
    cnt_le_t, cnt_first_t = torch.from_numpy(cnt_le), torch.from_numpy(cnt_first)
    dist.all_reduce(cnt_le_t, op=dist.ReduceOp.SUM)
    dist.all_reduce(cnt_first_t, op=dist.ReduceOp.SUM)
    # top-k keys of the local shard (valid-filtered), packed like pps_topk_update does, then the product's merge
    k = 12
    keys = np.full((nq, k), -1, dtype=np.int64)
    for i in range(nq):
        keep = O.valid_mask(d["qid"][i], d["qcam"][i], d["gid"][row0:row0 + rows], d["gcam"][row0:row0 + rows])
        cols = np.nonzero(keep)[0]
        packed = (blk[i][cols].view(np.uint32).astype(np.int64) << 32) | (cols + row0).astype(np.int64)
        packed.sort()
        keys[i, :min(k, len(packed))] = packed[:k]
    merged = evaluator.merge_topk_keys(torch.from_numpy(keys), k, dist.group.WORLD).numpy()
    # finalize (what pps_rank_finalize computes) from the reduced integers
    cnt_le, cnt_first = cnt_le_t.numpy(), cnt_first_t.numpy()
    ap = np.zeros(nq); valid = np.zeros(nq, np.uint8); first = np.full(nq, -1, np.int32)
    for i in range(nq):
        e = np.arange(p.off[i], p.off[i + 1])
        pos, junk = e[p.pos[e] == 1], e[p.pos[e] == 0]
        if len(pos) == 0:
            continue
        acc = 0.0
        for x in pos:
            c_pos = int((thr[pos] <= thr[x]).sum()); c_junk = int((thr[junk] <= thr[x]).sum())
            acc += c_pos / float(cnt_le[x] - c_junk)
        ap[i], valid[i] = acc / len(pos), 1
        best = pos[np.lexsort((p.g[pos], thr[pos]))[0]]
        jb = int(((thr[junk] < thr[best]) | ((thr[junk] == thr[best]) & (p.g[junk] < p.g[best]))).sum())
        first[i] = cnt_first[i] - jb
    res = evaluator.RankResult(ap, valid, first, None, p)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), ap=ap, valid=valid, first=first, mAP=res.mean_ap(),
             cmc=res.cmc(10, True), merged=merged, row0=row0, rows=rows)
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["small_mid", "dup_ties"])
def test_sharded_exchange_equals_unsharded(tmp_path, name):
    import torch.multiprocessing as mp
    world = 2
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    mp.spawn(_worker, args=(world, _free_port(), path, str(tmp_path)), nprocs=world, join=True)
    d = dict(np.load(path))
    ap, valid, first, _ = O.rank_counts(d["dist"], d["qid"], d["gid"], d["qcam"], d["gcam"])
    ti, td = O.topk_filtered(d["dist"], d["qid"], d["gid"], d["qcam"], d["gcam"], 12)
    outs = [dict(np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))) for r in range(world)]
    assert int(outs[0]["row0"]) == 0 and int(outs[0]["rows"]) + int(outs[1]["rows"]) == d["dist"].shape[1]
    for o in outs:
        np.testing.assert_allclose(o["ap"], ap, rtol=0, atol=1e-15)
        np.testing.assert_array_equal(o["valid"], valid)
        np.testing.assert_array_equal(o["first"], first)
        assert abs(float(o["mAP"]) - float(d["mAP"])) < 1e-12
        merged = o["merged"]
        np.testing.assert_array_equal((merged & 0xffffffff).astype(np.int32), ti)
        np.testing.assert_array_equal((merged >> 32).astype(np.uint32).view(np.float32), td)
    np.testing.assert_array_equal(outs[0]["merged"], outs[1]["merged"])


def test_gallery_shard_covers_rows_like_array_split():
    from pps_b200 import evaluator
    for ng, world in [(19732, 8), (10, 3), (5, 8), (0, 2), (1000000, 4)]:
        blocks = [evaluator.gallery_shard(ng, r, world) for r in range(world)]
        want = np.array_split(np.arange(ng), world)
        for (row0, rows), w in zip(blocks, want):
            assert rows == len(w) and (rows == 0 or row0 == w[0])


def _pass_worker(rank, world, port, golden_path, out_dir):
    """The pass protocol (pps_pass_begin / _count / _end) over gloo: every rank packs what its kernels would produce in the
    layout of csrc/ctx.cuh (PassState::keys / cnt_first / cnt_le / flags_dev), the two all-gathers move the bytes, and
    the reductions the library does with pass_reduce_counters_kernel / pass_merge_keys_kernel are restated in numpy."""
    import torch
    import torch.distributed as dist
    from pps_b200 import evaluator
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = dict(np.load(golden_path))
    D = d["dist"]
    nq, ng = D.shape
    k = 12
    row0, rows = evaluator.gallery_shard(ng, rank, world)
    p = evaluator.PairLists(d["qid"], d["qcam"], d["gid"], d["gcam"])       # global lists: identical on every rank
    n = p.n_pairs
    # exchange 1: thresholds of the pairs that live in this shard (others 0) -> all-gather -> sum (exact: one contributor)
    x1 = np.zeros(n, dtype=np.float32)
    local = (p.g[:n] >= row0) & (p.g[:n] < row0 + rows)
    x1[local] = D[p.q[:n][local], p.g[:n][local]]
    g1 = torch.zeros(world * n, dtype=torch.int32)
    dist.all_gather_into_tensor(g1, torch.from_numpy(x1.view(np.int32).copy()))
    thr = g1.numpy().reshape(world, n).sum(axis=0, dtype=np.int32).view(np.float32)
    np.testing.assert_array_equal(thr, D[p.q[:n], p.g[:n]])
    # local kernels (emulated): counters and top-k keys of the shard
    blk = D[:, row0:row0 + rows]
    cols = np.arange(row0, row0 + rows)
    cnt_le = np.zeros(n, dtype=np.uint32)
    cnt_first = np.zeros(nq, dtype=np.uint32)
    keys = np.full((nq, k), -1, dtype=np.int64)
    for i in range(nq):
        e = np.arange(p.off[i], p.off[i + 1])
        pos = e[p.pos[e] == 1]
        if len(pos):
            cnt_le[pos] = (blk[i][None, :] <= thr[pos][:, None]).sum(axis=1)
            best = pos[np.lexsort((p.g[pos], thr[pos]))[0]]
            eq = blk[i] == thr[best]
            cnt_first[i] = np.uint32((int((eq & (cols < p.g[best])).sum()) - int(eq.sum())) & 0xffffffff)   # mod 2^32, as the kernel
        keep = O.valid_mask(d["qid"][i], d["qcam"][i], d["gid"][row0:row0 + rows], d["gcam"][row0:row0 + rows])
        c = np.nonzero(keep)[0]
        packed = (blk[i][c].view(np.uint32).astype(np.int64) << 32) | (c + row0).astype(np.int64)
        packed.sort()
        keys[i, :min(k, len(packed))] = packed[:k]
    # exchange 2: [keys nq*k u64 | cnt_first nq u32 | cnt_le n u32 | flags 2 u32], padded to 16 bytes, one all-gather
    flags = np.array([1 if rank == 1 else 0, 0], dtype=np.uint32)          # a candidate overflow on ONE rank only
    body = keys.tobytes() + cnt_first.tobytes() + cnt_le.tobytes() + flags.tobytes()
    nbytes = (len(body) + 15) & ~15
    assert nbytes == ((nq * k * 8 + (nq + n + 2) * 4 + 15) & ~15)
    mine = torch.frombuffer(bytearray(body.ljust(nbytes, b"\0")), dtype=torch.uint8)
    gathered = torch.zeros(world * nbytes, dtype=torch.uint8)
    dist.all_gather_into_tensor(gathered, mine)
    G = gathered.numpy().reshape(world, nbytes)
    off = nq * k * 8
    words = np.stack([np.frombuffer(G[r, off:off + (nq + n + 2) * 4].tobytes(), dtype=np.uint32) for r in range(world)])
    red = words.sum(axis=0, dtype=np.uint32)                               # pass_reduce_counters_kernel
    cnt_first_all, cnt_le_all, flag_all = red[:nq], red[nq:nq + n], red[nq + n:]
    assert flag_all[0] == 1                                                # every rank sees the other's overflow flag
    allk = np.concatenate([np.frombuffer(G[r, :off].tobytes(), dtype=np.uint64).reshape(nq, k) for r in range(world)], axis=1)
    merged = np.sort(allk, axis=1)[:, :k]                                  # pass_merge_keys_kernel (unsigned order, ~0 = empty)
    # finalize from the reduced integers (pps_rank_finalize)
    ap = np.zeros(nq); valid = np.zeros(nq, np.uint8); first = np.full(nq, -1, np.int32)
    for i in range(nq):
        e = np.arange(p.off[i], p.off[i + 1])
        pos, junk = e[p.pos[e] == 1], e[p.pos[e] == 0]
        if len(pos) == 0:
            continue
        acc = 0.0
        for x in pos:
            acc += int((thr[pos] <= thr[x]).sum()) / float(int(cnt_le_all[x]) - int((thr[junk] <= thr[x]).sum()))
        ap[i], valid[i] = acc / len(pos), 1
        best = pos[np.lexsort((p.g[pos], thr[pos]))[0]]
        jb = int(((thr[junk] < thr[best]) | ((thr[junk] == thr[best]) & (p.g[junk] < p.g[best]))).sum())
        first[i] = np.int32((int(cnt_le_all[best]) + int(cnt_first_all[i])) & 0xffffffff) - jb
    np.savez(os.path.join(out_dir, "pass%d.npz" % rank), ap=ap, valid=valid, first=first, merged=merged.view(np.int64))
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["small_mid", "dup_ties"])
def test_pass_protocol_two_all_gathers(tmp_path, name):
    import torch.multiprocessing as mp
    world = 2
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    mp.spawn(_pass_worker, args=(world, _free_port(), path, str(tmp_path)), nprocs=world, join=True)
    d = dict(np.load(path))
    ap, valid, first, _ = O.rank_counts(d["dist"], d["qid"], d["gid"], d["qcam"], d["gcam"])
    ti, td = O.topk_filtered(d["dist"], d["qid"], d["gid"], d["qcam"], d["gcam"], 12)
    for r in range(world):
        o = dict(np.load(os.path.join(str(tmp_path), "pass%d.npz" % r)))
        np.testing.assert_allclose(o["ap"], ap, rtol=0, atol=1e-15)
        np.testing.assert_array_equal(o["valid"], valid)
        np.testing.assert_array_equal(o["first"], first)
        merged = o["merged"]
        np.testing.assert_array_equal((merged & 0xffffffff).astype(np.int32), ti)
        np.testing.assert_array_equal((merged >> 32).astype(np.uint32).view(np.float32), td)
