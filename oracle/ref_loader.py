"""Load the UNMODIFIED reference evaluator: from /root/reference where it exists (authoring container), else from
the staged copy under baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box, where
/root/reference does not exist).  ``stage()`` - called by ``__graft_entry__.build()`` in the authoring container -
copies the one file byte for byte; nothing of the reference enters the repository's history.

``reid_dataset_evaluator.py`` imports pycocotools and detectron.* at module top (:19-24), none
of which ``compute_dist`` / ``cmc`` / ``mean_ap`` touch; empty stub modules satisfy those
imports, then the file is executed from where it lies.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PPS_REFERENCE_ROOT", "/root/reference")
_REL = os.path.join("detectron", "datasets", "reid_dataset_evaluator.py")
STAGED_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
STAGED = os.path.join(STAGED_ROOT, _REL)
EVALUATOR = os.path.join(REFERENCE_ROOT, _REL)
if not os.path.exists(EVALUATOR) and os.path.exists(STAGED):
    EVALUATOR = STAGED


def available() -> bool:
    return os.path.exists(EVALUATOR)


def stage() -> bool:
    """Copy the unmodified evaluator file from /root/reference to baseline/_ref/ (no-op without /root/reference).
    Returns True if a staged copy exists afterwards."""
    import shutil
    src = os.path.join(REFERENCE_ROOT, _REL)
    if os.path.exists(src):
        os.makedirs(os.path.dirname(STAGED), exist_ok=True)
        if not os.path.exists(STAGED) or open(src, "rb").read() != open(STAGED, "rb").read():
            shutil.copyfile(src, STAGED)
    return os.path.exists(STAGED)


def load():
    """Returns the reference module object (functions compute_dist, cmc, mean_ap, ...)."""
    if not available():
        raise RuntimeError("reference not found at %s" % EVALUATOR)
    stubs = {
        "pycocotools": {}, "pycocotools.cocoeval": {"COCOeval": object},
        "detectron": {}, "detectron.core": {},
        "detectron.core.config": {"cfg": types.SimpleNamespace(), "get_output_dir": lambda *a, **k: "/tmp"},
        "detectron.utils": {}, "detectron.utils.io": {"save_object": lambda *a, **k: None},
        "detectron.utils.boxes": {},
    }
    saved = {}
    for name, attrs in stubs.items():
        saved[name] = sys.modules.get(name)
        mod = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod
    try:
        spec = importlib.util.spec_from_file_location("_pps_reference_evaluator", EVALUATOR)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for name, old in saved.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old
    return mod
