#!/usr/bin/env python
"""TEST INFRASTRUCTURE: compile the reference's re-ID custom ops UNMODIFIED, from where they lie under /root/reference, into
oracle/_ref/libref_reid_ops.so (git-ignored; travels to the GPU box with the snapshot):

    /root/reference/detectron/ops/batch_hard_op.cc             BatchHard / BatchHardGradient          (CPU operators)
    /root/reference/detectron/ops/pairwise_distance_op.cc      schema + gradient definition
    /root/reference/detectron/ops/pairwise_distance_op.cu      PairWiseDistance / ...Gradient        (CUDA operators)

They need the Caffe2 operator interface of pytorch v1.0.1, which is neither under /root/reference nor in this image (the
reference's own build is cmake inside a pytorch checkout: unbuildable here).  oracle/caffe2_shim/ supplies the handful of
containers / macros those three files touch (Tensor = dims + caller buffer, Operator<Context>::Input / Output, enforce
macros, registries, math::Set); every arithmetic statement that runs is the reference's.  No reference source is copied
into the repo: nvcc reads the files in place.  batch_hard_op.cu (a GPUFallbackOp registration) is not compiled: it only
routes the CUDA device type to the CPU operator above."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_OPS = "/root/reference/detectron/ops"
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libref_reid_ops.so")
SOURCES = ["batch_hard_op.cc", "pairwise_distance_op.cc", "pairwise_distance_op.cu"]


def available():
    return all(os.path.exists(os.path.join(REF_OPS, s)) for s in SOURCES) and shutil.which("nvcc") is not None


def build(force=False):
    """Returns the path of the library, or None where /root/reference does not exist (the GPU box uses the prebuilt file)."""
    if not available():
        return OUT if os.path.exists(OUT) else None
    srcs = [os.path.join(REF_OPS, s) for s in SOURCES] + [os.path.join(HERE, "ref_ops_harness.cu")]
    deps = srcs + [os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(HERE, "caffe2_shim")) for f in fs] + [__file__]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []
    for s in srcs:
        o = os.path.join(OUT_DIR, os.path.basename(s) + ".o")
        cmd = ["nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
               "-I", os.path.join(HERE, "caffe2_shim"), "-I", REF_OPS, "-x", "cu", "-c", s, "-o", o]
        subprocess.run(cmd, check=True)
        objs.append(o)
    subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs, check=True)
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p if p else "reference sources not present and no prebuilt library")
