"""The three label-generation steps end to end through their reference-compatible entry points
(`step.<name>.run(args)` with `--cam_network irn_b200.cam --irn_network irn_b200.irn`) on a tiny synthetic VOC tree,
against what the unmodified reference produced for the same files (tests/golden/steps.npz)."""
import os
import types

import numpy as np
import pytest
import torch
from PIL import Image

from conftest import golden_path
from irn_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def voc_tree(tmp_path_factory, cuda_dev):
    g = np.load(golden_path("steps.npz"))
    root = tmp_path_factory.mktemp("voc")
    os.makedirs(root / "JPEGImages")
    ids = [str(s) for s in g["ids"]]
    labels = {}
    for i, name in enumerate(ids):
        # the decoded pixels the reference saw, stored losslessly (PIL sniffs the format, the .jpg suffix is only a name)
        Image.fromarray(g["img%d" % i]).save(root / "JPEGImages" / (name + ".jpg"), format="PNG")
        labels[int(name.replace("_", ""))] = g["label%d" % i]
    np.save(root / "cls_labels.npy", labels, allow_pickle=True)
    (root / "list.txt").write_text("\n".join(ids) + "\n")
    for d in ("sess", "cam", "sem", "ins"):
        os.makedirs(root / d)
    torch.save(synth.cam_state_dict(), root / "sess" / "res50_cam.pth.pth")
    torch.save(synth.irn_state_dict(), root / "sess" / "res50_irn.pth")
    args = types.SimpleNamespace(
        num_workers=0, voc12_root=str(root), train_list=str(root / "list.txt"), infer_list=str(root / "list.txt"),
        cam_network="irn_b200.cam", irn_network="irn_b200.irn", cam_scales=(1.0, 0.5, 1.5, 2.0),
        cam_weights_name=str(root / "sess" / "res50_cam.pth"), irn_weights_name=str(root / "sess" / "res50_irn.pth"),
        cam_out_dir=str(root / "cam"), sem_seg_out_dir=str(root / "sem"), ins_seg_out_dir=str(root / "ins"),
        beta=10, exp_times=8, sem_seg_bg_thres=0.25, ins_seg_bg_thres=0.25, synthetic=0)
    from irn_b200.voc12 import dataloader
    dataloader._cls_labels["voc12/cls_labels.npy"] = labels     # the reference loads this file from the cwd at import time
    return g, ids, args


def test_make_cam_outputs(voc_tree):
    g, ids, args = voc_tree
    from irn_b200.step import make_cam
    make_cam.run(args)
    for i, name in enumerate(ids):
        d = np.load(os.path.join(args.cam_out_dir, name + ".npy"), allow_pickle=True).item()
        assert set(d) == {"keys", "cam", "high_res"}
        assert isinstance(d["cam"], torch.Tensor) and isinstance(d["high_res"], np.ndarray)
        assert np.array_equal(d["keys"].numpy(), g["cam_keys%d" % i])
        assert np.abs(d["cam"].numpy() - g["cam_cam%d" % i]).max() < 1e-4      # max-normalised: absolute = relative to 1
        assert np.abs(d["high_res"] - g["cam_high%d" % i]).max() < 1e-4


def test_make_cam_host_pyramid_identical(voc_tree):
    """--device_pyramid False (PIL pyramids built by the loader, the reference's data path) writes the same bytes."""
    g, ids, args = voc_tree
    from irn_b200.step import make_cam
    host_args = types.SimpleNamespace(**vars(args))
    host_args.device_pyramid = False
    host_args.cam_out_dir = os.path.join(os.path.dirname(args.cam_out_dir), "cam_host")
    os.makedirs(host_args.cam_out_dir, exist_ok=True)
    make_cam.run(host_args)
    make_cam.run(args)
    for name in ids:
        a = np.load(os.path.join(args.cam_out_dir, name + ".npy"), allow_pickle=True).item()
        b = np.load(os.path.join(host_args.cam_out_dir, name + ".npy"), allow_pickle=True).item()
        assert np.array_equal(a["keys"].numpy(), b["keys"].numpy())
        assert np.array_equal(a["cam"].numpy(), b["cam"].numpy()) and np.array_equal(a["high_res"], b["high_res"])


def test_make_sem_seg_labels_outputs(voc_tree):
    g, ids, args = voc_tree
    from irn_b200.step import make_sem_seg_labels
    make_sem_seg_labels.run(args)
    for i, name in enumerate(ids):
        lab = np.asarray(Image.open(os.path.join(args.sem_seg_out_dir, name + ".png")))
        ref = g["sem%d" % i]
        assert lab.dtype == np.uint8 and lab.shape == ref.shape
        assert (lab != ref).mean() < 5e-3, "label disagreement %g" % (lab != ref).mean()


def test_make_ins_seg_labels_outputs(voc_tree):
    g, ids, args = voc_tree
    from irn_b200.step import make_ins_seg_labels
    make_ins_seg_labels.run(args)
    for i, name in enumerate(ids):
        d = np.load(os.path.join(args.ins_seg_out_dir, name + ".npy"), allow_pickle=True).item()
        shape = tuple(g["ins_mask_shape%d" % i])
        ref_mask = np.unpackbits(g["ins_mask%d" % i], axis=-1, count=shape[-1]).astype(bool).reshape(shape)
        assert set(d) == {"score", "mask", "class"}
        assert d["mask"].dtype == bool and d["mask"].shape[1:] == shape[1:]
        # detections as a labelled image: identical up to boundary pixels
        def paint(masks, classes):
            out = np.zeros(shape[1:], np.int32)
            for m, c in zip(masks, classes):
                out[m] = int(c) + 1
            return out
        a, b = paint(d["mask"], d["class"]), paint(ref_mask, g["ins_class%d" % i])
        assert (a != b).mean() < 1e-2
        assert sorted(set(np.asarray(d["class"]).tolist())) == sorted(set(g["ins_class%d" % i].tolist()))


def test_run_sample_cli_synthetic(tmp_path, cuda_dev):
    """`python run_sample.py --synthetic N` (reference flag names, default output dirs) writes the three result trees."""
    import subprocess
    import sys
    from conftest import ROOT
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "run_sample.py"), "--synthetic", "2", "--num_workers", "0", "--exp_times", "6"],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    for i in range(2):
        name = "2007_%06d" % i
        d = np.load(tmp_path / "result" / "cam" / (name + ".npy"), allow_pickle=True).item()
        assert d["cam"].shape[1:] == (128, 128) and d["high_res"].shape[1:] == (512, 512)
        lab = np.asarray(Image.open(tmp_path / "result" / "sem_seg" / (name + ".png")))
        assert lab.shape == (512, 512) and lab.dtype == np.uint8
        ins = np.load(tmp_path / "result" / "ins_seg" / (name + ".npy"), allow_pickle=True).item()
        assert set(ins) == {"score", "mask", "class"} and ins["mask"].shape[1:] == (512, 512)
