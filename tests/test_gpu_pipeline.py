"""The batched device-resident pipeline equals the per-image reference-style chain (same kernels, batched)."""
import numpy as np
import pytest
import torch

from irn_b200 import cam_ops, indexing, synth
from irn_b200.cam import CAM
from irn_b200.irn import EdgeDisplacement
from irn_b200.pipeline import PseudoLabelPipeline, preprocess_batch

pytestmark = pytest.mark.gpu


def test_pipeline_equals_per_image_chain(cuda_dev):
    cam, irn = CAM(), EdgeDisplacement()
    cam.load_state_dict(synth.cam_state_dict(), strict=True)
    irn.load_state_dict(synth.irn_state_dict(), strict=False)
    H, W, N = 96, 128, 5
    scales = (1.0, 0.5, 1.5, 2.0)
    imgs = [synth.image(40 + i, H, W) for i in range(N)]
    labels = np.stack([synth.label(40 + i) for i in range(N)])
    labels[3] = 0                                         # an image with no class present
    host = preprocess_batch(imgs, scales, pin=True)
    pipe = PseudoLabelPipeline(cam, irn, cuda_dev, scales, cam_sub_batch=2, rw_sub_batch=2)
    out_h = pipe.run(host, labels, (H, W))               # host tensors: side-stream copies
    out_d = pipe.run([x.to(cuda_dev) for x in host], labels, (H, W))
    assert torch.equal(out_h["labels"], out_d["labels"])
    for i in range(N):
        per_scale = [cam(x[2 * i:2 * i + 2].to(cuda_dev)) for x in host]
        keys, strided, _ = cam_ops.merge_cams(per_scale, (H, W), labels[i])
        assert np.array_equal(keys.numpy(), out_d["keys"][i])
        if len(keys) == 0:
            assert int(out_d["labels"][i].max()) == 0
            continue
        assert torch.equal(strided, out_d["cams"][i])
        edge, _ = irn(host[0][2 * i:2 * i + 2].to(cuda_dev))
        rw = indexing.propagate_to_edge(strided, edge, beta=10, exp_times=8, radius=5)
        lab, _, _ = indexing.rw_labels(rw, keys.numpy(), (H, W), 0.25)
        assert torch.equal(lab, out_d["labels"][i])
