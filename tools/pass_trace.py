#!/usr/bin/env python
"""Where does a sharded pass spend its time?  CUDA events (device timeline) and host clocks around the five segments of
RankEngine._run_pass - pps_pass_begin | all-reduce | pps_pass_count | all-gather | pps_pass_end - averaged over steps,
max / mean over ranks.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pass_trace.py [--c3]
"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--c3", action="store_true", help="configs[3] strong-scaling shape (519 732 rows fp32, top-100) instead of the weak Market shards")
ap.add_argument("--steps", type=int, default=30)
a = ap.parse_args()
import torch
import torch.distributed as dist
from pps_b200 import evaluator, synthetic
world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
group = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
nq, dim = 3368, 2048
if a.c3:
    ng, topk = 519732, 100
else:
    ng, topk = 19732 * world, 0
qid, qcam, gid, gcam = synthetic.make_distractor_ids(nq, ng)
row0, ngl = evaluator.gallery_shard(ng, rank, world)
q = synthetic.make_features_device(qid, dim, 750, 4.0, 7, dev, torch.float32)
g = synthetic.make_features_device(gid[row0:row0 + ngl], dim, 750, 4.0, 1000 + rank, dev, torch.float32)
eng = evaluator.RankEngine(qid, gid, qcam, gcam, nq=nq, ng_local=ngl, dim=dim, gallery_offset=row0, topk=topk, group=group, device=dev)
eng.use_c_path = False
for _ in range(5):
    eng.run(q, g)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
seg_dev, seg_host, tot = {}, {}, []
for _ in range(a.steps):
    eng.trace = []
    e0 = torch.cuda.Event(enable_timing=True); e0.record()
    eng.run(q, g)
    e1 = torch.cuda.Event(enable_timing=True); e1.record()
    torch.cuda.synchronize()
    tr = eng.trace
    tot.append(e0.elapsed_time(e1))
    for (n0, ev0, t0), (n1, ev1, t1) in zip(tr[:-1], tr[1:]):
        seg_dev.setdefault(n1, []).append(ev0.elapsed_time(ev1))
        seg_host.setdefault(n1, []).append(1e3 * (t1 - t0))
eng.trace = None
names = list(seg_dev.keys())
vec = torch.tensor([np.mean(tot)] + [np.mean(seg_dev[n]) for n in names] + [np.mean(seg_host[n]) for n in names], dtype=torch.float64, device=dev)
mx, mean = vec.clone(), vec.clone()
if world > 1:
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(mean, op=dist.ReduceOp.SUM)
    mean /= world
if rank == 0:
    k = len(names)
    print(json.dumps({"world": world, "shape": "c3" if a.c3 else "weak-market", "ms_per_pass_mean": float(mean[0]), "ms_per_pass_max": float(mx[0]),
                      "device_ms_between_marks_mean": {n: float(mean[1 + i]) for i, n in enumerate(names)},
                      "device_ms_between_marks_max": {n: float(mx[1 + i]) for i, n in enumerate(names)},
                      "host_ms_between_marks_mean": {n: float(mean[1 + k + i]) for i, n in enumerate(names)}}))
if world > 1:
    dist.destroy_process_group()
