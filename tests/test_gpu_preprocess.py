"""C1 on the device against Pillow / the reference's loader body: integer + table work, exact equality."""
import numpy as np
import pytest
import torch
from PIL import Image

from irn_b200 import preprocess, synth
from irn_b200.voc12 import dataloader as dl
from oracle import resize as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W,oh,ow", [(512, 512, 256, 256), (512, 512, 768, 768), (375, 500, 188, 250), (375, 500, 750, 1000),
                                       (333, 500, 500, 750), (7, 5, 14, 10), (3, 3, 1, 1), (64, 64, 64, 96), (64, 64, 32, 64),
                                       (40, 40, 40, 40)])
def test_resize_equals_pillow(cuda_dev, H, W, oh, ow):
    rng = np.random.default_rng(H + W)
    imgs = np.stack([rng.integers(0, 256, (H, W, 3), dtype=np.uint8), synth.image(2, H, W), synth.image(9, H, W)])
    f32, u8 = preprocess.resize_normalize(torch.from_numpy(imgs).to(cuda_dev), (oh, ow), want_u8=True)
    lut = R.normalize_lut()
    for b in range(3):
        ref = np.asarray(Image.fromarray(imgs[b]).resize((ow, oh), Image.BICUBIC)) if (oh, ow) != (H, W) else imgs[b]
        assert np.array_equal(u8[b].cpu().numpy(), ref)
        chw = np.stack([lut[c][ref[..., c]] for c in range(3)])
        assert np.array_equal(f32[2 * b].cpu().numpy(), chw)
        assert np.array_equal(f32[2 * b + 1].cpu().numpy(), chw[..., ::-1])


@pytest.mark.parametrize("H,W", [(512, 512), (375, 500), (281, 500)])
def test_msf_batch_equals_reference_loader_body(cuda_dev, H, W):
    """voc12/dataloader.py:191-201: the device pyramid equals the host loader's, bit for bit, at the reference's four scales."""
    scales = (1.0, 0.5, 1.5, 2.0)
    imgs = np.stack([synth.image(i, H, W) for i in range(2)])
    got = preprocess.msf_batch(torch.from_numpy(imgs).to(cuda_dev), scales)
    for b in range(2):
        ref = dl.multi_scale_flip(imgs[b], scales)
        for k in range(len(scales)):
            g = got[k][2 * b:2 * b + 2].cpu().numpy()
            assert g.shape == ref[k].shape and np.array_equal(g, ref[k])


def test_pipeline_run_u8_equals_run(cuda_dev):
    """PseudoLabelPipeline.run_u8 (device pre-processing) gives the labels of run() on host-prepared pyramids."""
    from irn_b200.cam import CAM
    from irn_b200.irn import EdgeDisplacement
    from irn_b200.pipeline import PseudoLabelPipeline, preprocess_batch
    cam, irn = CAM(), EdgeDisplacement()
    cam.load_state_dict(synth.cam_state_dict(), strict=True)
    irn.load_state_dict(synth.irn_state_dict(), strict=False)
    cam.cuda(), irn.cuda()
    H = W = 128
    imgs = [synth.image(i, H, W) for i in range(3)]
    labels = torch.from_numpy(np.stack([synth.label(i) for i in range(3)]))
    pipe = PseudoLabelPipeline(cam, irn, cuda_dev)
    a = pipe.run(preprocess_batch(imgs, pin=False), labels, (H, W), want_highres=False)["labels"].cpu().numpy()
    b = pipe.run_u8(torch.from_numpy(np.stack(imgs)), labels, want_highres=False)["labels"].cpu().numpy()
    assert np.array_equal(a, b)
