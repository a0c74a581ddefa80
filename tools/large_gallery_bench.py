#!/usr/bin/env python
"""BASELINE configs[3] / configs[4]: Market-shaped queries against a gallery of real identities plus a very large
block of id-0 distractors (500 k ... 10 M rows, fp32 or fp16), top-k retrieval + exact positive ranks.

    python tools/large_gallery_bench.py --ng 10000000 --dtype fp16 --topk 100
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/large_gallery_bench.py --ng 10000000 --dtype fp16 --topk 100          # gallery sharded over N GPUs

The gallery is generated ON THE DEVICE in row blocks (a 41 GB fp16 gallery never exists on the host): the first
`--n-real` rows of the global gallery carry Market-like identities (centers + noise, L2-normalised), the rest are
distractors; ids / cameras are global numpy arrays every rank builds identically.  The pass is timed with CUDA
events, max over ranks; a few queries are then re-ranked by an independent torch fp32 computation over the whole
gallery (size-independent check: AP, first-match rank and top-k against a plain sort)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ng", type=int, default=2_000_000, help="GLOBAL gallery rows")
    ap.add_argument("--nq", type=int, default=3368)
    ap.add_argument("--dim", type=int, default=2048)
    ap.add_argument("--n-real", type=int, default=16932)
    ap.add_argument("--n-ids", type=int, default=750)
    ap.add_argument("--n-cams", type=int, default=6)
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "fp32"])
    ap.add_argument("--precision", default="f16x3")
    ap.add_argument("--topk", type=int, default=100)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--check-queries", type=int, default=8)
    ap.add_argument("--max-block-gib", type=float, default=8.0)
    ap.add_argument("--sigma", type=float, default=4.0, help="noise of the identity model (4: hard, mAP ~0.01-0.05 at this scale)")
    ap.add_argument("--two-sweep", action="store_true", help="thresholds from a full first sweep (the old form)")
    a = ap.parse_args()

    import torch
    import torch.distributed as dist
    from pps_b200 import _lib, evaluator

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    tdt = torch.float16 if a.dtype == "fp16" else torch.float32

    rs = np.random.RandomState(0)
    qid = rs.randint(1, a.n_ids + 1, size=a.nq).astype(np.int64)
    qcam = rs.randint(0, a.n_cams, size=a.nq).astype(np.int64)
    gid = np.zeros(a.ng, dtype=np.int64)
    real_rows = np.sort(rs.permutation(a.ng)[:a.n_real])            # real identities scattered over the whole gallery
    gid[real_rows] = rs.randint(1, a.n_ids + 1, size=a.n_real)
    gcam = rs.randint(0, a.n_cams, size=a.ng).astype(np.int64)
    row0, ngl = evaluator.gallery_shard(a.ng, rank, world)

    gen = torch.Generator(device=dev).manual_seed(1234)
    centers = torch.randn((a.n_ids + 1, a.dim), device=dev, generator=gen)
    centers[0] = 0

    def feats(ids_np, seed):
        g = torch.Generator(device=dev).manual_seed(seed)
        out = torch.empty((len(ids_np), a.dim), dtype=tdt, device=dev)
        ids_t = torch.from_numpy(ids_np).to(dev)
        for r0 in range(0, len(ids_np), 65536):
            sl = ids_t[r0:r0 + 65536]
            x = centers[sl] + a.sigma * torch.randn((len(sl), a.dim), device=dev, generator=g)
            x = x / x.norm(dim=1, keepdim=True)
            out[r0:r0 + 65536] = x.to(tdt)
        return out

    t0 = time.time()
    q = feats(qid, 7)
    g = feats(gid[row0:row0 + ngl], 1000 + rank)
    torch.cuda.synchronize()
    gen_s = time.time() - t0

    eng = evaluator.RankEngine(qid, gid, qcam, gcam, nq=a.nq, ng_local=ngl, dim=a.dim, gallery_offset=row0,
                               precision=a.precision, topk=a.topk, group=group, device=dev,
                               max_block_bytes=int(a.max_block_gib * (1 << 30)), in_dtype=tdt)
    eng.compact_thresholds = not a.two_sweep
    eng.use_c_path = False

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    res = None
    for _ in range(a.warmup):
        res = eng.run(q, g)
    barrier()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        res = eng.run(q, g)
    e1.record()
    barrier()
    launches = _lib.launch_count() - n0
    t = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])

    # ---- independent check on a few queries: torch fp32 over the whole (sharded) gallery + a plain sort ----
    check = None
    nchk = min(a.check_queries, a.nq)
    if nchk:
        qs = q[:nchk].float()
        d_local = torch.empty((nchk, ngl), dtype=torch.float32, device=dev)
        for r0 in range(0, ngl, 1 << 20):
            gb = g[r0:r0 + (1 << 20)].float()
            d2 = (qs * qs).sum(1, keepdim=True) + (gb * gb).sum(1)[None, :] - 2.0 * (qs @ gb.T)
            d_local[:, r0:r0 + gb.shape[0]] = d2.clamp_min(0).sqrt()
        if world > 1:
            parts = [torch.empty((nchk, evaluator.gallery_shard(a.ng, r, world)[1]), dtype=torch.float32, device=dev)
                     for r in range(world)]
            dist.all_gather(parts, d_local)
            d_all = torch.cat(parts, dim=1)
        else:
            d_all = d_local
        if rank == 0:
            dh = d_all.cpu().numpy()
            ap_err, first_ok, topk_ok, n_valid = 0.0, 0, 0.0, 0
            for i in range(nchk):
                same = gid == qid[i]
                junk = same & (gcam == qcam[i])
                pos = same & ~junk
                if not pos.any():
                    continue
                n_valid += 1
                order = np.argsort(dh[i], kind="stable")
                keep = ~junk[order]
                ranked_pos = pos[order][keep]
                hits = np.nonzero(ranked_pos)[0]
                ap_ref = np.mean((np.arange(len(hits)) + 1.0) / (hits + 1.0))
                ap_err = max(ap_err, abs(ap_ref - float(res.ap[i])))
                first_ok += int(hits[0] == int(res.first_rank[i]))
                if a.topk:
                    ref_top = order[keep][:a.topk]
                    topk_ok += len(np.intersect1d(ref_top, res.topk_index[i])) / float(a.topk)
            check = {"queries": nchk, "valid": n_valid, "max_abs_ap_diff": ap_err,
                     "first_rank_equal": first_ok, "topk_overlap": topk_ok / max(n_valid, 1) if a.topk else None}

    if rank == 0:
        pairs = float(a.nq) * float(a.ng)
        flops = 2.0 * a.dim * pairs
        line = {"tool": "large_gallery_bench", "n_gpus": world, "nq": a.nq, "ng": a.ng, "dim": a.dim, "dtype": a.dtype, "sigma": a.sigma,
                "precision": "f16x1" if a.dtype == "fp16" else a.precision, "topk": a.topk,
                "rows_per_gpu": ngl, "chunks_per_gpu": len(eng._chunk_list()), "chunk_rows": eng.chunk,
                "threshold_pass": "two-sweep" if a.two_sweep else "compacted same-id rows (%d on rank 0)" % eng.threshold_rows,
                "ms_per_pass": ms, "pairs_per_s": pairs / (ms * 1e-3), "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
                "gpu_launches_per_pass": launches / max(a.steps, 1), "gallery_gen_s": gen_s,
                "mAP": res.mean_ap(), "cmc1": float(res.cmc(10, True)[0]), "check": check}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
