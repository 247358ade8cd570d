"""Drop-in for the reference's ``step/eval_ins_seg.py`` (AP at IoU 0.5 of ``result/ins_seg/*.npy`` against the VOC
instance ground truth) without chainercv: same inputs, same printed quantity (step/eval_ins_seg.py:7-23).  chainercv is
absent from this image, so the restated metric (irn_b200/step/_voc_eval.py::instance_ap) is checked on hand-computable
cases only: "parity unpinned" for this evaluator."""
import os

import numpy as np

from . import _voc_eval


def run(args):
    ids = _voc_eval.voc_seg_ids(args.voc12_root, args.chainer_eval_set)
    gt = [_voc_eval.voc_instances(args.voc12_root, i) for i in ids]
    pred = [np.load(os.path.join(args.ins_seg_out_dir, i + ".npy"), allow_pickle=True).item() for i in ids]
    out = _voc_eval.instance_ap([p["mask"] for p in pred], [p["class"] for p in pred], [p["score"] for p in pred],
                                [g[0] for g in gt], [g[1] for g in gt], iou_thresh=0.5)
    print("0.5iou:", out)
    return out
