"""The CPU oracle against the fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import glob
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import golden_path, check_detections, unpack_masks
from oracle import indexing as oi
from oracle import nets, steps
from irn_b200 import synth


def _sha(pi):
    h = hashlib.sha256()
    for p in pi.path_indices:
        h.update(np.ascontiguousarray(p, dtype=np.int64).tobytes())
    h.update(np.ascontiguousarray(pi.src_indices, dtype=np.int64).tobytes())
    h.update(np.ascontiguousarray(pi.dst_indices, dtype=np.int64).tobytes())
    h.update(np.ascontiguousarray(pi.search_dst, dtype=np.int64).tobytes())
    return h.hexdigest()


PI = json.load(open(golden_path("path_index.json")))


@pytest.mark.parametrize("key", sorted(PI))
def test_path_index_bit_exact(key):
    g = PI[key]
    pi = oi.PathIndex(g["radius"], tuple(g["size"]))
    assert [list(p.shape) for p in pi.path_indices] == g["group_shapes"]
    assert np.asarray(pi.search_dst).tolist() == g["search_dst"]
    assert _sha(pi) == g["sha256"]


def test_path_index_survey_known_answer():
    # sha256 of PathIndex(5, (133,138)) measured on the reference
    assert PI["r5_133x138"]["sha256"] == "9d6c7172b84f0f3fa9a97ecb3523c2f385ff606614983e632b666704cb2f6042"
    assert PI["r5_133x138"]["src_head"] == [4, 5, 6, 7, 8]


def test_edge_to_affinity_matches_reference():
    g = np.load(golden_path("affinity_12x17.npz"))
    edge, aff = g["edge"], g["aff"][0]
    h, w, r = 12, 17, 5
    pi = oi.PathIndex(r, (h + r, w + 2 * r))
    ep = np.ones((h + r, w + 2 * r), np.float32)
    ep[:h, r:r + w] = edge[0]
    mine = oi.edge_to_affinity(ep, pi.path_indices)
    assert np.array_equal(mine, aff)
    # and the stencil form on the un-padded grid: window = rows 0..h, cols -1..w -> interior [0:h, 1:w+1]
    W, offs = oi.stencil_weights(edge, 5, beta=1)
    win = aff.reshape(34, h + 1, w + 2)[:, :h, 1:w + 1]
    assert np.array_equal(W, win)


@pytest.mark.parametrize("tag", ["r10", "r5", "r10b"])
def test_to_affinity_oracle_matches_reference(tag):
    """N4: the restated AffinityDisplacementLoss.to_affinity (net/resnet50_irn.py:162-175) against the unmodified method's output
    and against the gradient autograd computed through its index_select + max_pool2d (tests/golden/make_golden.py:gen_to_affinity)."""
    g = np.load(golden_path("to_affinity.npz"))
    edge, r = g[tag + "_edge"][:, 0], int(g[tag + "_radius"])
    B, h, w = edge.shape
    pi = oi.PathIndex(r, (h, w))
    aff, arg = oi.to_affinity(edge, pi.path_indices)
    assert np.array_equal(aff, g[tag + "_aff"])                       # gather / max / 1-x: exact
    grad_aff = np.random.RandomState(int(g[tag + "_seed"])).standard_normal(aff.shape).astype(np.float32)
    ge = oi.to_affinity_backward(grad_aff, arg, h * w).reshape(B, 1, h, w)
    ref = g[tag + "_grad_edge"]
    assert np.abs(ge - ref).max() < 1e-5 * max(1.0, np.abs(ref).max())    # fp32 scatter-add in the reference vs the exact sum


@pytest.mark.parametrize("path", sorted(glob.glob(golden_path("rw_*.npz"))), ids=os.path.basename)
def test_random_walk_oracle(path):
    g = np.load(path)
    h, w = g["x"].shape[-2:]
    n_iter = 2 ** int(g["exp_times"])
    if h * w <= 48 * 48:   # the faithful dense restatement (O((hw)^2) memory)
        dense = oi.propagate_to_edge(g["x"], g["edge"], 5, 10, int(g["exp_times"]))
        assert np.abs(dense - g["rw"]).max() < 2e-6
    st = oi.propagate_stencil(g["x"], g["edge"], 5, 10, n_iter)
    # the reference's own fp32 squaring error against the exact operator: SURVEY.md App. B (< 1e-4)
    assert np.abs(st - g["rw"]).max() < 1e-4


def test_cam_forward_oracle():
    g = np.load(golden_path("cam_forward.npz"))
    sd = synth.cam_state_dict()
    with torch.no_grad():
        for i in range(3):
            y = nets.cam_forward(torch.from_numpy(g["x%d" % i]), sd).numpy()
            assert y.shape == g["y%d" % i].shape
            assert np.abs(y - g["y%d" % i]).max() < 1e-5


def test_edge_displacement_oracle():
    g = np.load(golden_path("irn_forward.npz"))
    sd = synth.irn_state_dict()
    with torch.no_grad():
        for i in range(2):
            e, d = nets.edge_displacement(torch.from_numpy(g["x%d" % i]), sd)
            assert np.abs(e.numpy() - g["edge%d" % i]).max() < 1e-5
            assert np.abs(d.numpy() - g["dp%d" % i]).max() < 1e-4


def test_instance_functions_oracle():
    g = np.load(golden_path("instance_fns.npz"))
    for i in range(3):
        dp = g["dp%d" % i]
        cen = steps.find_centroids(dp)
        assert np.array_equal(cen, g["centroids%d" % i])
        shape = tuple(g["instances_shape%d" % i])
        inst = np.unpackbits(g["instances%d" % i], axis=-1, count=shape[-1]).astype(bool).reshape(shape)
        assert np.array_equal(steps.cluster_centroids(cen, dp), inst)


def test_step_bodies_oracle():
    """cam_merge / sem_seg_labels restatements against the reference's own _work loops."""
    from PIL import Image
    g = np.load(golden_path("steps.npz"))
    cam_sd, irn_sd = synth.cam_state_dict(), synth.irn_state_dict()
    for i in range(len(g["ids"])):
        img = g["img%d" % i]
        H, W = img.shape[:2]
        outs = []
        with torch.no_grad():
            for s in (1.0, 0.5, 1.5, 2.0):   # step/make_cam.py + voc12/dataloader.py:185-205
                if s == 1.0:
                    im = img
                else:
                    im = np.asarray(Image.fromarray(img).resize((int(np.round(W * s)), int(np.round(H * s))), Image.BICUBIC))
                x = synth.normalize_image(im)
                outs.append(nets.cam_forward(torch.from_numpy(np.stack([x, x[..., ::-1].copy()])), cam_sd))
            keys, low, high = steps.cam_merge(outs, (H, W), torch.from_numpy(g["label%d" % i]))
            assert np.array_equal(keys.numpy(), g["cam_keys%d" % i])
            assert np.abs(low.numpy() - g["cam_cam%d" % i]).max() < 1e-5
            assert np.abs(high.numpy() - g["cam_high%d" % i]).max() < 1e-5
            x = synth.normalize_image(img)
            edge, dp = nets.edge_displacement(torch.from_numpy(np.stack([x, x[..., ::-1].copy()])), irn_sd)
            rw = oi.propagate_stencil(g["cam_cam%d" % i], edge.numpy(), 5, 10, 256)
            lab = steps.sem_seg_labels(torch.from_numpy(rw.astype(np.float32)), g["cam_keys%d" % i], (H, W))
        ref = g["sem%d" % i]
        assert lab.shape == ref.shape
        assert (lab != ref).mean() < 2e-3   # pixels on a decision boundary may flip (float walk differs at 1e-5)


def test_detect_instance_oracle_vs_reference_outputs():
    """oracle.steps.detect_instance / ins_seg_labels (step/make_ins_seg_labels.py:82-105,131-150) against the dicts the
    reference's own `_work` loop saved: detection count, class order, scores (1e-4) and masks."""
    g = np.load(golden_path("steps.npz"))
    irn_sd = synth.irn_state_dict()
    for i in range(len(g["ids"])):
        img = g["img%d" % i]
        H, W = img.shape[:2]
        x = synth.normalize_image(img)
        with torch.no_grad():
            edge, dp = nets.edge_displacement(torch.from_numpy(np.stack([x, x[..., ::-1].copy()])), irn_sd)
        det = steps.ins_seg_labels(g["cam_cam%d" % i], g["cam_keys%d" % i], edge.numpy(), dp.numpy(), (H, W))
        check_detections(det, g["ins_score%d" % i], unpack_masks(g, str(i)), g["ins_class%d" % i])


def test_detect_instance_oracle_exact_on_reference_masks():
    """detect_instance alone, fed the reference's own masks back as the argmax one-hot: segments, order and areas must
    reproduce exactly (integer work)."""
    g = np.load(golden_path("steps.npz"))
    for i in range(len(g["ids"])):
        ref_mask, ref_class, ref_score = unpack_masks(g, str(i)), g["ins_class%d" % i], g["ins_score%d" % i]
        # rebuild the per-channel masks: union of the reference's segments per (class, instance) channel is not stored, but
        # every segment is one connected component, so feeding each segment as its own channel must return it unchanged
        scores = ref_mask.astype(np.float32) * ref_score[:, None, None]
        det = steps.detect_instance(scores, ref_mask, ref_class, max_fragment_size=0)
        assert np.array_equal(det["mask"], ref_mask) and np.array_equal(det["class"], ref_class)
        assert np.array_equal(det["score"][ref_score > 0], ref_score[ref_score > 0])


def test_oracle_at_benchmark_size_512():
    """The oracle's label tail and ins-seg tail at the benchmark's size, fed the reference's own 128x128 walk outputs /
    CAMs / edge / dp for one 512x512 image (tests/golden/steps512.npz): labels identical, detections identical."""
    g = np.load(golden_path("steps512.npz"))
    lab = steps.sem_seg_labels(torch.from_numpy(g["walk_sem"]), g["cam_keys"], (512, 512))
    assert (lab != g["sem"]).mean() == 0.0
    # exact-operator walk vs the reference's dense fp32 walk at 128x128 / 256 steps (< 1e-4)
    st = oi.propagate_stencil(g["cam_cam"], g["edge"], 5, 10, 256)
    assert np.abs(st.reshape(g["walk_sem"].shape) - g["walk_sem"]).max() < 1e-4
    det = steps.ins_seg_labels(g["cam_cam"], g["cam_keys"], g["edge"], g["dp"], (512, 512),
                               walk=lambda seeds, edge: g["walk_ins"][:, 0])
    check_detections(det, g["ins_score"], unpack_masks(g), g["ins_class"], score_tol=1e-6, pixel_tol=1e-9)


def test_oracle_nets_at_benchmark_size_512():
    """oracle.nets CAM forward at the four benchmark input sizes (256/512/768/1024) and EdgeDisplacement at 512, against the
    reference's own forwards inside its make_cam / make_sem_seg loops."""
    from oracle import pipeline as opipe
    g = np.load(golden_path("steps512.npz"))
    img = synth.image(int(g["seed"]), 512, 512)
    cam_sd, irn_sd = synth.cam_state_dict(), synth.irn_state_dict()
    xs = opipe.msf_inputs(img, (1.0, 0.5, 1.5, 2.0))
    with torch.no_grad():
        for x in xs:
            y = nets.cam_forward(x, cam_sd).numpy()
            ref = g["camscale_%d" % x.shape[-1]]
            assert np.abs(y - ref).max() / ref.max() < 1e-5
        e, d = nets.edge_displacement(xs[0], irn_sd)
    assert np.abs(e.numpy() - g["edge"]).max() < 1e-5 and np.abs(d.numpy() - g["dp"]).max() < 1e-4


def test_split_indices():
    parts = steps.split_indices(10, 3)
    assert [p.tolist() for p in parts] == [[0, 3, 6, 9], [1, 4, 7], [2, 5, 8]]
