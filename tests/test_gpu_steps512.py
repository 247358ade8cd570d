"""Parity at the BENCHMARK's size: one 512x512 synthetic image (BASELINE.json configs: inputs of 256 / 512 / 768 / 1024
pixels for the four CAM scales, 128x128 walk grid) against what the unmodified reference produced for it on the CPU
(tests/golden/steps512.npz, written by tests/golden/make_golden.py --only steps512: the reference's own three `_work`
loops incl. the dense 16384^2 walk).  Tolerances are north_star's: 1e-4 on float outputs (max-normalised CAMs, edge,
walk), stated per assertion."""
import numpy as np
import pytest
import torch

from conftest import golden_path, check_detections, unpack_masks, record
from irn_b200 import cam_ops, indexing, instance, preprocess, synth
from irn_b200.cam import CAM
from irn_b200.irn import EdgeDisplacement
from irn_b200.pipeline import PseudoLabelPipeline

pytestmark = pytest.mark.gpu
SCALES = (1.0, 0.5, 1.5, 2.0)


@pytest.fixture(scope="module")
def world(cuda_dev):
    g = np.load(golden_path("steps512.npz"))
    cam = CAM()
    cam.load_state_dict(synth.cam_state_dict(), strict=True)
    irn = EdgeDisplacement()
    irn.load_state_dict(synth.irn_state_dict(), strict=False)
    cam.cuda(cuda_dev), irn.cuda(cuda_dev)
    img = synth.image(int(g["seed"]), 512, 512)
    x = torch.from_numpy(img[None]).to(cuda_dev)
    pyr = preprocess.msf_batch(x, SCALES)
    return g, cam, irn, x, pyr


@pytest.mark.parametrize("mode", [1, 2], ids=["tf32x3", "f16x3"])
def test_cam_every_scale_vs_reference(world, mode):
    """CAM.forward at the benchmark's four input sizes (net/resnet50_cam.py:55-70 via step/make_cam.py:35), in both
    tensor-core arithmetic modes; the f16x3 mode (production) must keep a 3x margin to the bar."""
    g, cam, _, _, pyr = world
    cam.set_conv_mode(mode)
    bound = 1e-4 if mode == 1 else 1e-4 / 3
    for p in pyr:
        ref = g["camscale_%d" % p.shape[-1]]
        y = cam(p).cpu().numpy()
        assert y.shape == ref.shape
        err = float(np.abs(y - ref).max() / ref.max())
        record("cam_512_per_scale", input=int(p.shape[-1]), normalised_err=err, conv_mode=mode)
        assert err < bound, "scale input %d: normalised CAM err %g" % (p.shape[-1], err)
    cam.set_conv_mode(None)


def test_make_cam_merge_vs_reference(world):
    g, cam, _, _, pyr = world
    outs = [cam(p) for p in pyr]
    keys, lo, hi = cam_ops.merge_cams(outs, (512, 512), torch.from_numpy(g["label"]))
    assert np.array_equal(keys.numpy(), g["cam_keys"])
    e_lo = float(np.abs(lo.cpu().numpy() - g["cam_cam"]).max())
    e_hi = float(np.abs(hi.cpu().numpy()[:, 1::4, 2::4] - g["cam_high_s4"]).max())
    record("make_cam_512", strided_err=e_lo, highres_err=e_hi)
    assert e_lo < 1e-4 and e_hi < 1e-4                       # max-normalised maps: absolute = relative to 1
    assert np.abs(hi.cpu().numpy().reshape(len(keys), -1).max(1) - g["cam_high_max"]).max() < 1e-4


def test_edge_displacement_512_vs_reference(world):
    g, _, irn, _, pyr = world
    e, d = irn(pyr[0])
    e_err = float(np.abs(e.cpu().numpy() - g["edge"]).max())
    d_err = float(np.abs(d.cpu().numpy() - g["dp"]).max())
    record("edge_displacement_512", edge_err=e_err, dp_err=d_err, dp_absmax=float(np.abs(g["dp"]).max()))
    assert e_err < 1e-4
    assert d_err < 1e-4


def test_walk_128x128_on_reference_cams(world):
    """propagate_to_edge (misc/indexing.py:141-167) on the reference's own CAMs and edge map, 256 steps at 128x128: against
    the reference's dense fp32 walk (16384^2 matrix squared 8 times)."""
    g, _, _, x, _ = world
    dev = x.device
    rw = indexing.propagate_to_edge(torch.from_numpy(g["cam_cam"]).to(dev), torch.from_numpy(g["edge"]).to(dev), radius=5, beta=10, exp_times=8)
    err = float(np.abs(rw.cpu().numpy() - g["walk_sem"]).max())
    record("walk_512_sem", err=err, ref_max=float(g["walk_sem"].max()))
    assert rw.shape == g["walk_sem"].shape and err < 1e-4


def test_whole_chain_512_vs_reference(world):
    """uint8 image -> labels through the batched pipeline (the bench's code path) against the reference's PNG and its
    instance dict; the mIoU of our label map scored against the reference's is within 0.1 pt of 100."""
    from oracle import steps as osteps
    g, cam, irn, x, _ = world
    pipe = PseudoLabelPipeline(cam, irn, x.device, SCALES)
    out = pipe.run_u8(x, g["label"][None])
    lab = out["labels"][0].cpu().numpy()
    dis = float((lab != g["sem"]).mean())
    _, miou = osteps.confusion_miou([lab], [g["sem"]])
    record("chain_512_sem", disagreement=dis, miou_vs_reference_labels=miou)
    assert dis < 2.5e-3 and (1.0 - miou) * 100 < 0.1
    dets = pipe.instance_stage(out["cams"], out["keys"], out["edge"], out["dp"], (512, 512))
    worst = check_detections(dets[0], g["ins_score"], unpack_masks(g), g["ins_class"], score_tol=1e-4, pixel_tol=2.5e-3)
    record("chain_512_ins", detections=len(dets[0]["score"]), worst_mask_disagreement=worst)
