"""Ground-truth access and the confusion-matrix IoU the reference's evaluation steps use, without chainercv
(N2 in SURVEY.md section 8(f)).

chainercv's VOCSemanticSegmentationDataset reads `ImageSets/Segmentation/<split>.txt` and the palette PNGs under
`SegmentationClass/` (index 255 = ignore -> -1); calc_semantic_segmentation_confusion counts (gt, pred) pairs over the
pixels with gt >= 0.  Both are restated here in numpy.
"""
import os

import numpy as np
from PIL import Image


def voc_seg_ids(voc12_root, split):
    with open(os.path.join(voc12_root, "ImageSets", "Segmentation", split + ".txt")) as f:
        return [l.strip() for l in f if l.strip()]


def voc_seg_label(voc12_root, img_id):
    lab = np.asarray(Image.open(os.path.join(voc12_root, "SegmentationClass", img_id + ".png"))).astype(np.int32)
    lab[lab == 255] = -1
    return lab


def confusion(preds, labels, n_class=21):
    """rows = ground truth, cols = prediction; pixels with label < 0 are ignored (step/eval_sem_seg.py:18)."""
    conf = np.zeros((n_class, n_class), np.int64)
    for p, l in zip(preds, labels):
        p = np.asarray(p).reshape(-1).astype(np.int64)
        l = np.asarray(l).reshape(-1).astype(np.int64)
        ok = l >= 0
        if (p[ok] >= n_class).any() or (l[ok] >= n_class).any():
            raise ValueError("label outside [0, %d)" % n_class)
        conf += np.bincount(n_class * l[ok] + p[ok], minlength=n_class * n_class).reshape(n_class, n_class)
    return conf


def iou_from_confusion(conf):
    """step/eval_sem_seg.py:20-26: per-class IoU plus the false-positive / false-negative rates it prints."""
    gtj, resj, diag = conf.sum(axis=1), conf.sum(axis=0), np.diag(conf)
    with np.errstate(divide="ignore", invalid="ignore"):
        denom = gtj + resj - diag
        return {"iou": diag / denom, "fp": 1.0 - gtj / denom, "fn": 1.0 - resj / denom}
