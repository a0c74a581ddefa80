"""Load the UNMODIFIED reference evaluator from /root/reference (authoring container only).

``reid_dataset_evaluator.py`` imports pycocotools and detectron.* at module top (:19-24), none
of which ``compute_dist`` / ``cmc`` / ``mean_ap`` touch; empty stub modules satisfy those
imports, then the file is executed from where it lies.  Nothing is copied.  The GPU box has
no /root/reference: there only the committed fixtures (tests/golden/) are used.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PPS_REFERENCE_ROOT", "/root/reference")
EVALUATOR = os.path.join(REFERENCE_ROOT, "detectron", "datasets", "reid_dataset_evaluator.py")


def available() -> bool:
    return os.path.exists(EVALUATOR)


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
