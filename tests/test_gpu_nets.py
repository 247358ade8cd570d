"""CAM / EdgeDisplacement forward on the GPU against the reference's outputs (golden fixtures made
by the unmodified reference on CPU, IEEE fp32).  Tolerance: 1e-4 max-abs on the max-normalised CAM
(BASELINE.json north_star), stated per assertion."""
import numpy as np
import pytest
import torch

from conftest import golden_path
from irn_b200 import synth
from irn_b200.cam import CAM
from irn_b200.irn import EdgeDisplacement

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cam_model(cuda_dev):
    m = CAM()
    m.load_state_dict(synth.cam_state_dict(), strict=True)
    m.eval()
    return m


@pytest.fixture(scope="module")
def irn_model(cuda_dev):
    m = EdgeDisplacement()
    m.load_state_dict(synth.irn_state_dict(), strict=False)
    m.eval()
    return m


def test_cam_forward_vs_reference(cuda_dev, cam_model):
    g = np.load(golden_path("cam_forward.npz"))
    for i in range(3):
        y = cam_model(torch.from_numpy(g["x%d" % i]).to(cuda_dev)).cpu().numpy()
        ref = g["y%d" % i]
        assert y.shape == ref.shape
        scale = ref.max()
        assert np.abs(y - ref).max() / scale < 1e-4, "normalised CAM err %g" % (np.abs(y - ref).max() / scale)


def test_cam_batch_equals_single(cuda_dev, cam_model):
    g = np.load(golden_path("cam_forward.npz"))
    x = torch.from_numpy(g["x0"]).to(cuda_dev)
    xb = torch.cat([x, x.flip(0), x], 0)          # three pairs
    yb = cam_model.forward_batch(xb)
    y = cam_model(x)
    assert torch.equal(yb[0], y) and torch.equal(yb[2], y)
    # flipping the pair flips the CAM (SURVEY.md section 4 property)
    assert torch.allclose(yb[1], y.flip(-1), atol=1e-5)


def test_edge_displacement_vs_reference(cuda_dev, irn_model):
    g = np.load(golden_path("irn_forward.npz"))
    for i in range(3):
        if ("x%d" % i) in g:
            x = g["x%d" % i]
        else:
            x = synth.normalize_image(synth.image(200 + i, 512, 512))
            x = np.stack([x, x[..., ::-1].copy()])
        e, d = irn_model(torch.from_numpy(x).to(cuda_dev))
        assert e.shape == g["edge%d" % i].shape and d.shape == g["dp%d" % i].shape
        assert np.abs(e.cpu().numpy() - g["edge%d" % i]).max() < 1e-4
        assert np.abs(d.cpu().numpy() - g["dp%d" % i]).max() < 1e-3    # dp is in pixels (|dp| up to ~3): 1e-3 abs ~ 3e-4 relative


def test_state_dict_keys_match_reference_format(cam_model, irn_model):
    assert len(cam_model.state_dict()) == 956 and len(irn_model.state_dict()) == 1035
