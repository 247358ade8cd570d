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


def _palette_png(path, arr):
    im = Image.new("P", (arr.shape[1], arr.shape[0]))
    im.putdata(arr.reshape(-1).tolist())
    im.putpalette([v for k in range(256) for v in (k, k, k)])
    im.save(path)


def test_eval_ins_seg_hand_computed_ap(tmp_path):
    """step/eval_ins_seg.py restated without chainercv: AP at IoU 0.5 on cases small enough to compute by hand."""
    from irn_b200.step import eval_ins_seg
    root = tmp_path / "voc"
    for d in ("ImageSets/Segmentation", "SegmentationClass", "SegmentationObject"):
        os.makedirs(root / d)
    os.makedirs(tmp_path / "ins")
    ids = ["2007_000001", "2007_000002"]
    (root / "ImageSets" / "Segmentation" / "train.txt").write_text("\n".join(ids) + "\n")
    H, W = 20, 30
    # image 1: two instances of class 3 (VOC id 4), one of class 7 (VOC id 8); a 255 boundary strip
    obj = np.zeros((H, W), np.uint8); cls = np.zeros((H, W), np.uint8)
    obj[2:8, 2:10] = 1; cls[2:8, 2:10] = 4
    obj[10:18, 2:10] = 2; cls[10:18, 2:10] = 4
    obj[2:18, 15:28] = 3; cls[2:18, 15:28] = 8
    obj[9, :] = 255; cls[9, :] = 255
    _palette_png(root / "SegmentationObject" / (ids[0] + ".png"), obj)
    _palette_png(root / "SegmentationClass" / (ids[0] + ".png"), cls)
    m = lambda y0, y1, x0, x1: np.pad(np.ones((y1 - y0, x1 - x0), bool), ((y0, H - y1), (x0, W - x1)))
    # predictions, class 3: exact hit (0.9), duplicate of the same object (0.8: false positive), miss (0.7), hit on the second (0.6)
    # class 7: one detection with IoU < 0.5 (false positive)
    np.save(tmp_path / "ins" / (ids[0] + ".npy"), {
        "mask": np.stack([m(2, 8, 2, 10), m(2, 8, 2, 9), m(0, 3, 20, 30), m(10, 18, 2, 10), m(2, 6, 15, 20)]),
        "class": np.array([3, 3, 3, 3, 7]), "score": np.array([0.9, 0.8, 0.7, 0.6, 0.5], np.float32)})
    # image 2: one instance of class 7, detected
    obj2 = np.zeros((H, W), np.uint8); cls2 = np.zeros((H, W), np.uint8)
    obj2[5:15, 5:25] = 1; cls2[5:15, 5:25] = 8
    _palette_png(root / "SegmentationObject" / (ids[1] + ".png"), obj2)
    _palette_png(root / "SegmentationClass" / (ids[1] + ".png"), cls2)
    np.save(tmp_path / "ins" / (ids[1] + ".npy"), {"mask": m(5, 15, 5, 24)[None], "class": np.array([7]), "score": np.array([0.95], np.float32)})
    masks, labels = _voc_eval.voc_instances(str(root), ids[0])
    assert masks.shape == (3, H, W) and labels.tolist() == [3, 3, 7]
    assert not masks[:, 9].any()                       # the boundary row belongs to no instance
    args = types.SimpleNamespace(voc12_root=str(root), chainer_eval_set="train", ins_seg_out_dir=str(tmp_path / "ins"))
    out = eval_ins_seg.run(args)
    # class 3: matches by descending score = [1, 0, 0, 1], 2 positives -> precision envelope: recall 0.5 at precision 1, recall 1 at 0.5
    assert abs(out["ap"][3] - (0.5 * 1.0 + 0.5 * 0.5)) < 1e-12
    # class 7: scores [0.95 (hit), 0.5 (miss)], 2 positives -> recall 0.5 at precision 1, nothing beyond
    assert abs(out["ap"][7] - 0.5) < 1e-12
    assert np.isnan(out["ap"][0]) and abs(out["map"] - 0.625) < 1e-12
