"""N2: the evaluator restatements (no chainercv) on a tiny synthetic VOC tree, against the oracle's confusion / IoU."""
import os
import types

import numpy as np
from PIL import Image

from irn_b200.step import _voc_eval, eval_cam, eval_sem_seg
from oracle import steps


def test_eval_steps_on_synthetic_voc(tmp_path):
    rng = np.random.default_rng(0)
    root = tmp_path / "voc"
    os.makedirs(root / "ImageSets" / "Segmentation")
    os.makedirs(root / "SegmentationClass")
    os.makedirs(tmp_path / "sem")
    os.makedirs(tmp_path / "cam")
    ids = ["2007_000032", "2007_000039"]
    (root / "ImageSets" / "Segmentation" / "train.txt").write_text("\n".join(ids) + "\n")
    gts, preds, cam_preds = [], [], []
    for i in ids:
        gt = rng.choice([0, 3, 15, 255], size=(40, 50), p=[0.5, 0.2, 0.2, 0.1]).astype(np.uint8)
        im = Image.new("P", (gt.shape[1], gt.shape[0]))          # VOC ground truth is a palette PNG whose indices are the labels
        im.putdata(gt.reshape(-1).tolist())
        im.putpalette([v for k in range(256) for v in (k, k, k)])
        im.save(root / "SegmentationClass" / (i + ".png"))
        pr = rng.choice([0, 3, 15], size=(40, 50)).astype(np.uint8)
        Image.fromarray(pr).save(tmp_path / "sem" / (i + ".png"))
        high = rng.random((2, 40, 50)).astype(np.float32)
        np.save(tmp_path / "cam" / (i + ".npy"), {"keys": np.array([2, 14]), "cam": None, "high_res": high})
        g = gt.astype(np.int32)
        g[g == 255] = -1
        gts.append(g)
        preds.append(pr)
        full = np.concatenate([np.full((1, 40, 50), 0.15, np.float32), high], 0)
        cam_preds.append(np.array([0, 3, 15])[np.argmax(full, 0)])
    args = types.SimpleNamespace(voc12_root=str(root), chainer_eval_set="train", sem_seg_out_dir=str(tmp_path / "sem"),
                                 cam_out_dir=str(tmp_path / "cam"), cam_eval_thres=0.15)
    out = eval_sem_seg.run(args)
    iou, miou = steps.confusion_miou(preds, gts)
    assert np.allclose(out["iou"], iou, equal_nan=True) and abs(out["miou"] - miou) < 1e-12
    out = eval_cam.run(args)
    iou, miou = steps.confusion_miou(cam_preds, gts)
    assert np.allclose(out["iou"], iou, equal_nan=True) and abs(out["miou"] - miou) < 1e-12
    conf = _voc_eval.confusion(preds, gts)
    assert conf.sum() == sum((g >= 0).sum() for g in gts)
