#!/usr/bin/env python
"""A short configs[4]-shaped pass (fp16 rows, top-100) for profiling the counting epilogue of the distance kernel:

    python tools/fused_pass_profile.py --ng 1300000                     # prints the pass time and what the pass decided
    ncu --set full --clock-control none --import-source on -k regex:dist_tc2_kernel -s 2 -c 1 -o gpurun_out/prof \
        python tools/fused_pass_profile.py --ng 1300000 --passes 1      # launch 3 = the first block after the short one

Launch order of dist_tc2_kernel in a pass: first (short) block, threshold product, then one launch per remaining block -
with fp16 rows those are the counting-epilogue instantiation (EPI_RANK: nothing written)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--ng", type=int, default=1_300_000)
ap.add_argument("--nq", type=int, default=3368)
ap.add_argument("--dim", type=int, default=2048)
ap.add_argument("--topk", type=int, default=100)
ap.add_argument("--passes", type=int, default=3)
ap.add_argument("--no-fused-count", action="store_true")
a = ap.parse_args()

import torch
from pps_b200 import evaluator, synthetic

dev = torch.device("cuda")
qid, qcam, gid, gcam = synthetic.make_distractor_ids(a.nq, a.ng)
q = synthetic.make_features_device(qid, a.dim, 750, 4.0, 7, dev, torch.float16)
g = synthetic.make_features_device(gid, a.dim, 750, 4.0, 1000, dev, torch.float16)
eng = evaluator.RankEngine(qid, gid, qcam, gcam, nq=a.nq, ng_local=a.ng, dim=a.dim, topk=a.topk, device=dev,
                           in_dtype=torch.float16)
eng.fused_count = not a.no_fused_count
for i in range(a.passes):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = eng.run(q, g)
    e1.record()
    torch.cuda.synchronize()
    print("pass %d: %.2f ms, blocks %d, counting epilogue %s, mAP %.6f" % (
        i, e0.elapsed_time(e1), eng.pass_blocks, eng.used_fused_count, float(res.ap[res.is_valid].mean())))
