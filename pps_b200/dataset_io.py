"""On-disk inputs of a standalone ranking run: the reference's cached ``features.npy`` plus its COCO-style json.

Reference: detectron/core/test_engine.py:225-226 writes ``features.npy`` ([num_images, D] float32, rows in roidb
order); the roidb comes from a COCO-format json (detectron/datasets/json_dataset.py:89-190): images sorted by id,
``image`` = file name ``{pid:08d}_{cam:04d}_{k:08d}.jpg`` (tools/dataset/transform_market1501.py:60), ``mark`` of the
image's single annotation (0 query / 1 gallery / 2 multi-query, json_dataset.py:188-189).

    python -m pps_b200.dataset_io --features features.npy --annotations market1501_test.json [--precision f16x3]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np


class JsonReidDataset:
    """The part of JsonDataset the evaluator touches: ``get_roidb(gt=True)`` -> entries with 'image' and 'mark'."""

    def __init__(self, annotation_file: str, image_directory: str = "", name: str = None):
        self.name = name or os.path.splitext(os.path.basename(annotation_file))[0]
        with open(annotation_file) as f:
            data = json.load(f)
        marks = {}
        for ann in data.get("annotations", []):
            if "mark" in ann:
                if ann["image_id"] in marks:
                    raise RuntimeError("image %r has more than one annotation (json_dataset.py:186 asserts one)" % ann["image_id"])
                marks[ann["image_id"]] = int(ann["mark"])
        images = sorted(data["images"], key=lambda im: im["id"])          # json_dataset.py:104-106: sorted image ids
        self._roidb = []
        for im in images:
            if im["id"] not in marks:
                raise RuntimeError("image %r has no annotation with a 'mark'" % im["id"])
            self._roidb.append(dict(image=os.path.join(image_directory, im["file_name"]), mark=marks[im["id"]], id=im["id"]))

    def get_roidb(self, gt=True):
        return self._roidb


def load_features(path: str) -> np.ndarray:
    feats = np.load(path, mmap_mode="r")
    if feats.ndim != 2:
        raise RuntimeError("features must be [num_images, D], got shape %s" % (feats.shape,))
    return np.array(feats, dtype=np.float32, order="C")      # a writable in-memory copy of the mapped file


def main(argv=None):
    ap = argparse.ArgumentParser(description="CMC / mAP of cached re-ID features on B200 (PPS evaluate())")
    ap.add_argument("--features", required=True, help="features.npy written by the reference's test engine")
    ap.add_argument("--annotations", required=True, help="COCO-style json of the test split (with per-annotation 'mark')")
    ap.add_argument("--precision", default="f16x3", choices=["bf16x1", "bf16x3", "bf16x6", "f16x3"])
    ap.add_argument("--rerank", action="store_true", help="k-reciprocal re-ranking (the reference's cfg.REID.RERANK)")
    ap.add_argument("--output", default=None, help="write the result dict as json here")
    args = ap.parse_args(argv)
    from . import evaluator
    ds = JsonReidDataset(args.annotations)
    feats = load_features(args.features)
    if feats.shape[0] != len(ds.get_roidb()):
        raise RuntimeError("%d feature rows but %d images in %s" % (feats.shape[0], len(ds.get_roidb()), args.annotations))
    result = evaluator.evaluate(ds, feats, None, precision=args.precision, verbose=True, to_re_rank=args.rerank)
    res = evaluator.reid_results(result, ds.name)
    if args.output:
        with open(args.output, "w") as f:
            json.dump(res, f, indent=1, default=float)
    return res


if __name__ == "__main__":
    main()
    sys.exit(0)
