"""Run the UNMODIFIED graph-building code of the reference's pooling heads (authoring container only).

The reference has no pooling op: ``detectron/modeling/bpm_heads.py`` and ``pps_heads.py`` emit a Caffe2 sub-graph
(``Split``, global ``AveragePool`` / ``MaxPool``, ``Mean``, ``Max``, ``Add``, ``Concat``).  Caffe2 itself is not
importable here, but the graph BUILDERS are plain Python: this module executes them, from where they lie under
/root/reference, against a stand-in ``model`` whose operators are evaluated eagerly in NumPy float32.  What that pins
to the reference's own code is everything the heads decide — the split tables, the enumeration order of the 2^n - 1
combinations, which strip blobs feed which ``Mean`` / ``Max`` / ``Add``, the blob names, the FPN / FPN_SHARED
branches; what stays a restatement is the arithmetic inside the six stock operators (their published semantics:
global average = sum / count, ``Mean`` = sum of the inputs in order times 1/N, elementwise ``Max`` / ``Add``).
Nothing is copied; the GPU box only sees the fixtures written by oracle/make_golden_pool.py.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("PPS_REFERENCE_ROOT", "/root/reference")
MODELING = os.path.join(REFERENCE_ROOT, "detectron", "modeling")


def available() -> bool:
    return os.path.exists(os.path.join(MODELING, "pps_heads.py")) and os.path.exists(os.path.join(MODELING, "bpm_heads.py"))


class _Net:
    """The slice of caffe2's net / CNNModelHelper API the two head files call, evaluated eagerly."""

    def __init__(self, blobs, train):
        self.ws = dict(blobs)          # blob name -> float32 array
        self.train = train
        self.net = self
        self.ops = []                  # (type, inputs, outputs) in emission order

    def _get(self, name):
        return self.ws[str(name)]

    def Split(self, blob_in, blobs_out, split=None, axis=1):
        x = self._get(blob_in)
        assert sum(split) == x.shape[axis], "Split: sum(split) == input.dim(axis)"      # caffe2 enforces this
        off = 0
        for name, s in zip(blobs_out, split):
            idx = [slice(None)] * x.ndim
            idx[axis] = slice(off, off + s)
            self.ws[name] = np.ascontiguousarray(x[tuple(idx)])
            off += s
        self.ops.append(("Split", [str(blob_in)], list(blobs_out)))
        return blobs_out

    def AveragePool(self, blob_in, blob_out, global_pooling=False, **kw):
        assert global_pooling
        x = self._get(blob_in)
        cnt = np.float32(x.shape[2] * x.shape[3])
        self.ws[blob_out] = (x.sum(axis=(2, 3), dtype=np.float32, keepdims=True) / cnt).astype(np.float32)
        self.ops.append(("AveragePool", [str(blob_in)], [blob_out]))
        return blob_out

    def MaxPool(self, blob_in, blob_out, global_pooling=False, **kw):
        assert global_pooling
        self.ws[blob_out] = self._get(blob_in).max(axis=(2, 3), keepdims=True)
        self.ops.append(("MaxPool", [str(blob_in)], [blob_out]))
        return blob_out

    def Mean(self, blobs_in, blob_out):
        acc = self._get(blobs_in[0]).astype(np.float32).copy()
        for b in blobs_in[1:]:
            acc = acc + self._get(b)
        if len(blobs_in) > 1:
            acc = acc * np.float32(1.0 / len(blobs_in))                                  # caffe2 Mean: Scale(1.0f / InputSize())
        self.ws[blob_out] = acc.astype(np.float32)
        self.ops.append(("Mean", [str(b) for b in blobs_in], [blob_out]))
        return blob_out

    def Max(self, blobs_in, blob_out):
        acc = self._get(blobs_in[0]).copy()
        for b in blobs_in[1:]:
            acc = np.maximum(acc, self._get(b))
        self.ws[blob_out] = acc
        self.ops.append(("Max", [str(b) for b in blobs_in], [blob_out]))
        return blob_out

    def Add(self, blobs_in, blob_out):
        self.ws[blob_out] = (self._get(blobs_in[0]) + self._get(blobs_in[1])).astype(np.float32)
        self.ops.append(("Add", [str(b) for b in blobs_in], [blob_out]))
        return blob_out

    def Concat(self, blobs_in, blobs_out, axis=1):
        self.ws[blobs_out[0]] = np.concatenate([self._get(b) for b in blobs_in], axis=axis)
        self.ops.append(("Concat", [str(b) for b in blobs_in], list(blobs_out)))
        return blobs_out[0], blobs_out[1]


def _load(cfg):
    stubs = {
        "caffe2": {}, "caffe2.python": {"workspace": types.SimpleNamespace()},
        "detectron": {}, "detectron.core": {}, "detectron.core.config": {"cfg": cfg},
        "detectron.utils": {}, "detectron.utils.c2": {"const_fill": lambda *a, **k: None, "gauss_fill": lambda *a, **k: None},
        "detectron.utils.net": {"get_group_gn": lambda *a, **k: None}, "detectron.utils.blob": {},
        "detectron.modeling": {}, "detectron.modeling.ResNet": {"add_stage": lambda *a, **k: None},
        "detectron.modeling.init": {},
    }
    saved = {}
    for name, attrs in stubs.items():
        saved[name] = sys.modules.get(name)
        mod = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod

    def load_file(modname, fname):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(MODELING, fname))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    try:
        bpm = load_file("detectron.modeling.bpm_heads", "bpm_heads.py")
        sys.modules["detectron.modeling"].bpm_heads = bpm
        pps = load_file("detectron.modeling.pps_heads", "pps_heads.py")
    finally:
        for name, old in saved.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old
        for name in ("detectron.modeling.bpm_heads", "detectron.modeling.pps_heads"):
            sys.modules.pop(name, None)
    return pps


def run_pps_head(maps, strip_num=6, max_ave=True, scale=(128, 384), spatial_scale=1.0 / 16, fpn_on=False, fpn_shared=False,
                 train=False, preprefix="pps"):
    """maps: one [N, C, H, W] float32 array, or (FPN) a list of them with a list of spatial scales.
    Returns (names, arrays, dims_out, ops): the blobs ``add_pps_part_head`` returns, evaluated."""
    if not available():
        raise RuntimeError("reference not found under %s" % MODELING)
    cfg = types.SimpleNamespace(
        REID=types.SimpleNamespace(BPM_STRIP_NUM=strip_num, MAX_AVE_FEATURE=max_ave, SCALE=tuple(scale), FPN_SHARED=fpn_shared),
        FPN=types.SimpleNamespace(FPN_ON=fpn_on))
    pps = _load(cfg)
    if fpn_on:
        names = ["level%d" % i for i in range(len(maps))]
        model = _Net({n: np.asarray(m, np.float32) for n, m in zip(names, maps)}, train)
        blob_in, dim_in, ss = names, [int(m.shape[1]) for m in maps], list(spatial_scale)
    else:
        model = _Net({"conv5": np.asarray(maps, np.float32)}, train)
        blob_in, dim_in, ss = "conv5", int(maps.shape[1]), spatial_scale
    with contextlib.redirect_stdout(io.StringIO()):                 # the builder print()s every combination
        blobs_out, dims_out = pps.add_pps_part_head(model, blob_in, dim_in, ss, preprefix=preprefix)
    return [str(b) for b in blobs_out], [model.ws[str(b)] for b in blobs_out], list(dims_out), model.ops


# ------------------------------------------------------------------------------------
# the embedding head between pooling and distance: reid_heads.add_reid_outputs at test time
# ------------------------------------------------------------------------------------
class _EmbedNet(_Net):
    """+ the operators detectron/modeling/reid_heads.py:34-127 emits (Conv 1x1, SpatialBN in test mode, Relu,
    DropoutIfTraining, FC, Reshape, Normalize), with parameters looked up by the blob names Caffe2's helpers give them:
    '<out>_w' / '<out>_b' for Conv and FC, '<out>_s' / '_b' / '_rm' / '_riv' for SpatialBN."""

    def __init__(self, blobs, params, num_classes, train=False):
        super().__init__(blobs, train)
        self.params = params
        self.num_classes = num_classes

    def Conv(self, blob_in, blob_out, dim_in, dim_out, kernel, stride=1, pad=0, no_bias=0, **kw):
        assert kernel == 1 and stride == 1 and pad == 0
        x = self._get(blob_in)                                        # [N, C, 1, 1]
        w, b = self.params[blob_out + "_w"], self.params[blob_out + "_b"]
        assert w.shape == (dim_out, dim_in, 1, 1)
        y = np.einsum("nchw,ec->nehw", x.astype(np.float64), w[:, :, 0, 0].astype(np.float64)) + b.astype(np.float64)[None, :, None, None]
        self.ws[blob_out] = y                                         # float64 through the branch, rounded at the end
        self.ops.append(("Conv", [str(blob_in)], [blob_out]))
        return blob_out

    def SpatialBN(self, blob_in, blob_out, dim, is_test=True, epsilon=1e-5, **kw):
        assert is_test
        x = self._get(blob_in)
        p = lambda s: self.params[blob_out + s].astype(np.float64)[None, :, None, None]
        self.ws[blob_out] = (x - p("_rm")) / np.sqrt(p("_riv") + epsilon) * p("_s") + p("_b")
        self.ops.append(("SpatialBN", [str(blob_in)], [blob_out]))
        return blob_out

    def Relu(self, blob_in, blob_out):
        self.ws[str(blob_out)] = np.maximum(self._get(blob_in), 0)
        self.ops.append(("Relu", [str(blob_in)], [str(blob_out)]))
        return blob_out

    def DropoutIfTraining(self, blob_in, ratio):
        return blob_in

    def FC(self, blob_in, blob_out, dim_in, dim_out, **kw):
        x = self._get(blob_in)
        self.ws[blob_out] = np.zeros((x.shape[0], dim_out))           # classifier logits: not on the retrieval path
        self.ops.append(("FC", [str(blob_in)], [blob_out]))
        return blob_out

    def Reshape(self, blob_in, blobs_out, shape=None):
        self.ws[blobs_out[0]] = self._get(blob_in).reshape(shape)
        self.ops.append(("Reshape", [str(blob_in)], list(blobs_out)))
        return blobs_out[0], blobs_out[1]

    def Normalize(self, blob_in, blob_out, axis=-1):
        x = self._get(blob_in)
        nrm = np.maximum(np.sqrt((x * x).sum(axis=axis, keepdims=True)), 1e-12)      # caffe2 normalize_op.h: kEps = 1e-12
        self.ws[blob_out] = x / nrm
        self.ops.append(("Normalize", [str(blob_in)], [blob_out]))
        return blob_out


def run_reid_outputs(pooled_blobs, names, params, bpm_dim=128, normalize=True, num_classes=751):
    """Execute the unmodified ``add_reid_outputs`` (reid_heads.py:34-127, model.train = False) on ONE image's pooled
    blobs (the reference extracts features with batch size 1, and its Reshape(shape=[1, -1]) flattens whatever batch it
    gets).  pooled_blobs: list of [1, C, 1, 1] arrays named ``names``.  Returns (feature [1, K*bpm_dim] float64, ops)."""
    if not available():
        raise RuntimeError("reference not found under %s" % MODELING)
    cfg = types.SimpleNamespace(
        REID=types.SimpleNamespace(BPM_DIM=bpm_dim, DROPOUT_FEATURE=False, NORMALIZE_FEATURE=normalize, CRM=False),
        MODEL=types.SimpleNamespace(USE_GN=False))
    stubs = {
        "detectron": {}, "detectron.core": {}, "detectron.core.config": {"cfg": cfg},
        "detectron.utils": {}, "detectron.utils.c2": {"const_fill": lambda *a, **k: None, "gauss_fill": lambda *a, **k: None,
                                                     "UnscopeName": lambda s: s[s.rfind("/") + 1:]},
        "detectron.utils.blob": {}, "detectron.utils.reid": {}, "detectron.modeling": {}, "detectron.modeling.init": {},
        "detectron.modeling.crm_heads": {},
    }
    saved = {}
    for name, attrs in stubs.items():
        saved[name] = sys.modules.get(name)
        mod = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod
    loaded = []

    def load_file(modname, fname):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(MODELING, fname))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        loaded.append(modname)
        spec.loader.exec_module(mod)
        return mod

    try:
        load_file("detectron.modeling.triplet_loss", "triplet_loss.py")
        heads = load_file("detectron.modeling.reid_heads", "reid_heads.py")      # fresh module: its feature_list is a global
    finally:
        for name, old in saved.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old
        for name in loaded:
            sys.modules.pop(name, None)
    model = _EmbedNet({n: np.asarray(b, np.float32) for n, b in zip(names, pooled_blobs)}, params, num_classes, train=False)
    dims = [int(b.shape[1]) for b in pooled_blobs]
    with contextlib.redirect_stdout(io.StringIO()):
        heads.add_reid_outputs(model, list(names), dims, preprefix="reid")
    out = model.ws["reid_feature_concat_norm" if normalize else "reid_feature_concat"]
    return np.asarray(out, np.float64), model.ops
