"""TEST INFRASTRUCTURE: the reference's own re-ID custom ops (BatchHard, BatchHardGradient on the CPU; PairWiseDistance,
PairWiseDistanceGradient on CUDA), compiled UNMODIFIED by oracle/build_ref_ops.py, behind NumPy / torch wrappers.

Only tests/ and oracle/ scripts load this (the product path never does).  `available()` is False where neither
/root/reference nor a prebuilt oracle/_ref/libref_reid_ops.so exists."""
import ctypes as C
import os

import numpy as np

from . import build_ref_ops

_lib = None


def available():
    return build_ref_ops.available() or os.path.exists(build_ref_ops.OUT)


def load():
    global _lib
    if _lib is None:
        path = build_ref_ops.build()
        if path is None:
            raise RuntimeError("reference op library not built and /root/reference not present")
        _lib = C.CDLL(path)
        _lib.ref_op_run.restype = C.c_int
        _lib.ref_op_run.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int),
                                    C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int), C.POINTER(C.c_int64), C.c_void_p, C.c_char_p, C.c_int]
        _lib.ref_op_gradient_def.restype = C.c_int
        _lib.ref_op_gradient_def.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
        _lib.ref_op_schema.restype = C.c_int
        _lib.ref_op_schema.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return _lib


def _run(name, device, ins, outs, stream=None):
    """ins: [(pointer, shape)], outs: [(pointer, capacity bytes)] -> list of output shapes.  Raises RuntimeError with the
    reference's enforce text on failure (the behaviour a Caffe2 user sees)."""
    lib = load()
    n_in, n_out = len(ins), len(outs)
    in_ptr = (C.c_void_p * n_in)(*[p for p, _ in ins])
    in_nd = (C.c_int * n_in)(*[len(s) for _, s in ins])
    in_dims = (C.c_int64 * (4 * n_in))()
    for i, (_, s) in enumerate(ins):
        for k, d in enumerate(s):
            in_dims[4 * i + k] = int(d)
    out_ptr = (C.c_void_p * n_out)(*[p for p, _ in outs])
    out_cap = (C.c_int64 * n_out)(*[int(c) for _, c in outs])
    out_nd = (C.c_int * n_out)()
    out_dims = (C.c_int64 * (4 * n_out))()
    err = C.create_string_buffer(512)
    rc = lib.ref_op_run(name.encode(), device, n_in, in_ptr, in_nd, in_dims, n_out, out_ptr, out_cap, out_nd, out_dims,
                        stream, err, 512)
    if rc != 0:
        raise RuntimeError("%s: %s" % (name, err.value.decode()))
    return [tuple(out_dims[4 * i + k] for k in range(out_nd[i])) for i in range(n_out)]


def schema(name):
    n_in, n_out = C.c_int(0), C.c_int(0)
    if load().ref_op_schema(name.encode(), C.byref(n_in), C.byref(n_out)) != 0:
        raise KeyError(name)
    return n_in.value, n_out.value


def gradient_def(name, n_in, n_out):
    buf = C.create_string_buffer(512)
    if load().ref_op_gradient_def(name.encode(), n_in, n_out, buf, 512) != 0:
        raise KeyError(name)
    typ, ins, outs = buf.value.decode().split("|")
    return typ, ins.split(","), outs.split(",")


# ---- CPU operators (batch_hard_op.cc) ----
def batch_hard(xdist, labels):
    """BatchHardOp<float, CPUContext>::RunOnDevice (batch_hard_op.cc:9-59): (AP, AN)."""
    x = np.ascontiguousarray(xdist, dtype=np.float32)
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    n = x.shape[0]
    ap, an = np.empty(max(n, 1), np.float32), np.empty(max(n, 1), np.float32)
    shapes = _run("BatchHard", 0, [(x.ctypes.data, x.shape), (lab.ctypes.data, lab.shape)],
                  [(ap.ctypes.data, ap.nbytes), (an.ctypes.data, an.nbytes)])
    assert shapes == [(n,), (n,)], shapes
    return ap[:n], an[:n]


def batch_hard_grad(xdist, labels, dap, dan):
    """BatchHardGradientOp<float, CPUContext>::RunOnDevice (batch_hard_op.cc:62-123): dX [N, N].  The operator writes
    dX[a * N + idx] with idx = -1 when no candidate improved on the initial value (an anchor alone in its class, or without
    any other class): one element BEFORE row a.  The buffer gets guard words on both sides; `stray` returns what landed in
    the front guard (row 0's stray write)."""
    x = np.ascontiguousarray(xdist, dtype=np.float32)
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    dap = np.ascontiguousarray(dap, dtype=np.float32)
    dan = np.ascontiguousarray(dan, dtype=np.float32)
    n = x.shape[0]
    buf = np.full(n * n + 16, np.nan, np.float32)
    body = buf[8:8 + n * n]
    shapes = _run("BatchHardGradient", 0,
                  [(x.ctypes.data, x.shape), (lab.ctypes.data, lab.shape), (dap.ctypes.data, dap.shape), (dan.ctypes.data, dan.shape)],
                  [(body.ctypes.data, body.nbytes)])
    assert shapes == [(n, n)], shapes
    return body.reshape(n, n).copy(), buf[:8].copy()


# ---- CUDA operators (pairwise_distance_op.cu); torch tensors on the current device ----
def pairwise_distance(x):
    import torch
    x = x.contiguous().float()
    n = int(x.shape[0])
    z = torch.empty((n, n), dtype=torch.float32, device=x.device)
    s = torch.cuda.current_stream().cuda_stream
    shapes = _run("PairWiseDistance", 1, [(x.data_ptr(), tuple(x.shape))], [(z.data_ptr(), z.numel() * 4)], stream=s)
    assert shapes == [(n, n)], shapes
    return z


def pairwise_distance_grad(x, dz):
    import torch
    x = x.contiguous().float()
    dz = dz.contiguous().float()
    dx = torch.empty_like(x)
    s = torch.cuda.current_stream().cuda_stream
    shapes = _run("PairWiseDistanceGradient", 1, [(x.data_ptr(), tuple(x.shape)), (dz.data_ptr(), tuple(dz.shape))],
                  [(dx.data_ptr(), dx.numel() * 4)], stream=s)
    assert shapes == [tuple(x.shape)], shapes
    return dx
