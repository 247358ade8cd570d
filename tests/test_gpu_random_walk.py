"""Parity of the CUDA random walk / label kernels with the reference (golden fixtures) and the
oracle.  Float tolerance: 1e-4 max-abs (BASELINE.json north_star); PathIndex/affinity exact."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import golden_path
from irn_b200 import indexing, synth
from oracle import indexing as oi
from oracle import steps

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_edge_to_affinity_exact(cuda_dev):
    g = np.load(golden_path("affinity_12x17.npz"))
    h, w = 12, 17
    aff = indexing.edge_to_affinity(_t(g["edge"], cuda_dev), 5).cpu().numpy()[0]
    ref = g["aff"][0].reshape(34, h + 1, w + 2)[:, :h, 1:w + 1]     # reference window -> image interior
    assert np.array_equal(aff, ref)
    for r in (3, 10):   # other radii against the oracle (training config uses radius 10)
        e = synth.edge_map(37, 45, "uniform", r)
        W, _ = oi.stencil_weights(e, r, beta=1)
        assert np.array_equal(indexing.edge_to_affinity(_t(e, cuda_dev), r).cpu().numpy()[0], W)


@pytest.mark.parametrize("tag", ["r10", "r5", "r10b"])
def test_to_affinity_forward_backward_vs_reference(cuda_dev, tag):
    """N4: indexing.to_affinity (the body for AffinityDisplacementLoss.to_affinity, net/resnet50_irn.py:162-175) against the
    unmodified reference method: forward bit-exact, gradient (autograd through index_select + max_pool2d in the reference)
    within fp32 summation order; "r10b" has saturated edges (ties: the first maximum along the path takes the gradient)."""
    g = np.load(golden_path("to_affinity.npz"))
    edge, r = g[tag + "_edge"], int(g[tag + "_radius"])
    e = _t(edge, cuda_dev).requires_grad_(True)
    aff = indexing.to_affinity(e, radius=r)
    assert np.array_equal(aff.detach().cpu().numpy(), g[tag + "_aff"])
    grad_aff = np.random.RandomState(int(g[tag + "_seed"])).standard_normal(tuple(aff.shape)).astype(np.float32)
    aff.backward(_t(grad_aff, cuda_dev))
    ref = g[tag + "_grad_edge"]
    err = float(np.abs(e.grad.cpu().numpy() - ref).max())
    assert err < 1e-5 * max(1.0, float(np.abs(ref).max())), "grad_edge max-abs err %g" % err
    # the oracle's exact (fp64) scatter and its argmax agree with the kernel's routing
    B, _, h, w = edge.shape
    _, arg = oi.to_affinity(edge[:, 0], oi.PathIndex(r, (h, w)).path_indices)
    truth = oi.to_affinity_backward(grad_aff, arg, h * w).reshape(B, 1, h, w)
    assert np.abs(e.grad.cpu().numpy() - truth).max() < 1e-5 * max(1.0, float(np.abs(truth).max()))
    # through a PathIndex object, no gradient requested (inference): same values, [B,H,W] input accepted
    pi = indexing.PathIndex(r, (h, w))
    again = indexing.to_affinity(_t(edge[:, 0], cuda_dev), pi)
    assert np.array_equal(again.cpu().numpy(), g[tag + "_aff"])


@pytest.mark.parametrize("path", sorted(glob.glob(golden_path("rw_*.npz"))), ids=os.path.basename)
@pytest.mark.parametrize("variant", [0, 1, 2, 4])
def test_random_walk_vs_reference(cuda_dev, path, variant):
    g = np.load(path)
    x, edge = g["x"], g["edge"]
    C, h, w = x.shape
    out = indexing.random_walk_batch(_t(x, cuda_dev), _t(edge, cuda_dev), [0, C], 5, 10, 2 ** int(g["exp_times"]), variant=variant)
    err = np.abs(out.cpu().numpy().reshape(C, 1, h, w) - g["rw"]).max()
    assert err < TOL, "max-abs err %g vs reference" % err
    truth = oi.propagate_stencil(x, edge, 5, 10, 2 ** int(g["exp_times"]))
    assert np.abs(out.cpu().numpy().reshape(C, 1, h, w) - truth).max() < 2e-6   # fp64 state: only fp32 output rounding


def test_propagate_to_edge_signature(cuda_dev):
    g = np.load(golden_path("rw_20x28_c3_e4_sigmoid4.npz"))
    rw = indexing.propagate_to_edge(_t(g["x"], cuda_dev), _t(g["edge"], cuda_dev), radius=5, beta=10, exp_times=4)
    assert rw.shape == (3, 1, 20, 28) and rw.is_cuda and rw.dtype == torch.float32
    assert np.abs(rw.cpu().numpy() - g["rw"]).max() < TOL
    # 4-D seeds [K, I, h, w] flatten to K*I channels (step/make_ins_seg_labels.py:135)
    x4 = torch.stack([_t(g["x"], cuda_dev), 0.5 * _t(g["x"], cuda_dev)], 1)
    rw4 = indexing.propagate_to_edge(x4, _t(g["edge"], cuda_dev), exp_times=4)
    assert rw4.shape == (6, 1, 20, 28)
    assert torch.allclose(rw4[0::2], rw, atol=1e-6) and torch.allclose(rw4[1::2], 0.5 * rw, atol=1e-6)


@pytest.mark.parametrize("h,w", [(1, 1), (3, 50), (8, 128), (9, 5), (17, 31), (33, 127), (64, 64), (65, 40), (128, 128), (125, 94),
                                 (129, 60), (40, 130)])
def test_walk_shapes_and_properties(cuda_dev, h, w):
    """Ragged / tiny / full-size grids: fused cluster kernel (every cluster size 1..16) == step kernels == generic kernel ==
    fp64 oracle; convexity.  Grids beyond 128 fall back to the per-step kernel."""
    C = 3
    e = synth.edge_map(h, w, "uniform", h + w)
    x = synth.seeds(C, h, w, h)
    run = lambda v, n=16: indexing.random_walk_batch(_t(x, cuda_dev), _t(e, cuda_dev), [0, C], n_iter=n, variant=v).cpu().numpy()
    a = run(0)
    fits = h <= 128 and w <= 128
    assert indexing.last_walk_was_fused() == fits
    b, s2 = run(1), run(2)
    assert np.abs(a - b).max() < 1e-6
    assert np.array_equal(a, s2)   # fused and per-step kernels do identical arithmetic
    if fits:
        assert np.array_equal(a, run(4))
        for n in (0, 1, 3):   # odd / zero step counts exercise both state buffers
            assert np.array_equal(run(4, n), run(2, n))
    else:
        with pytest.raises(Exception):
            run(4)
    truth = oi.propagate_stencil(x, e, 5, 10, 16).reshape(C, h, w)
    assert np.abs(a - truth).max() < 1e-6
    x0 = x * (1 - e)
    assert a.min() >= x0.min() - 1e-6 and a.max() <= x0.max() + 1e-6   # every step is a convex combination


def test_walk_batch_mixed_channels(cuda_dev):
    """A batch with 0..6 channels per image equals per-image calls (channel chunking, empty images)."""
    h, w = 40, 36
    counts = [2, 0, 1, 6, 3, 5]
    offs = np.concatenate([[0], np.cumsum(counts)])
    edges = np.concatenate([synth.edge_map(h, w, "bimodal", i) for i in range(len(counts))], 0)
    x = synth.seeds(int(offs[-1]), h, w, 99)
    out = indexing.random_walk_batch(_t(x, cuda_dev), _t(edges, cuda_dev), offs, n_iter=32).cpu().numpy()
    for i, c in enumerate(counts):
        if c == 0:
            continue
        xi = x[offs[i]:offs[i + 1]]
        one = indexing.random_walk_batch(_t(xi, cuda_dev), _t(edges[i:i + 1], cuda_dev), [0, c], n_iter=32).cpu().numpy()
        assert np.array_equal(one, out[offs[i]:offs[i + 1]])
        truth = oi.propagate_stencil(xi, edges[i], 5, 10, 32).reshape(c, h, w)
        assert np.abs(one - truth).max() < 1e-6


def test_walk_dispatch_many_channels(cuda_dev):
    """Instance-path shape (classes x instances = many channels per image): variant 0 picks the per-step kernel, which shares
    the weight reads between 4 channels; the fused kernel gives the same bits."""
    h, w, C = 64, 64, 40
    e = synth.edge_map(h, w, "bimodal", 7)
    x = synth.seeds(C, h, w, 7)
    a = indexing.random_walk_batch(_t(x, cuda_dev), _t(e, cuda_dev), [0, C], n_iter=8, variant=0).cpu().numpy()
    assert not indexing.last_walk_was_fused()
    b = indexing.random_walk_batch(_t(x, cuda_dev), _t(e, cuda_dev), [0, C], n_iter=8, variant=4).cpu().numpy()
    assert indexing.last_walk_was_fused()
    assert np.array_equal(a, b)


def test_walk_linearity_and_fixed_point(cuda_dev):
    h, w = 64, 64
    e = synth.edge_map(h, w, "bimodal", 3)
    x = synth.seeds(2, h, w, 3)
    f = lambda v: indexing.random_walk_batch(_t(v, cuda_dev), _t(e, cuda_dev), [0, v.shape[0]], n_iter=64).cpu().numpy()
    a = f(x)
    assert np.abs(f((x[0:1] + x[1:2]) * 0.5) - 0.5 * (a[0:1] + a[1:2])).max() < 1e-6
    ones = np.ones((1, h, w), np.float32)
    zero_edge = np.zeros((1, h, w), np.float32)
    c = indexing.random_walk_batch(_t(ones, cuda_dev), _t(zero_edge, cuda_dev), [0, 1], n_iter=64).cpu().numpy()
    assert np.abs(c - 1).max() < 1e-6        # rows of the transition matrix sum to 1


def test_other_radius_generic_kernel(cuda_dev):
    h, w, C = 30, 33, 2
    e = synth.edge_map(h, w, "uniform", 5)
    x = synth.seeds(C, h, w, 5)
    for r in (3, 7):
        a = indexing.random_walk_batch(_t(x, cuda_dev), _t(e, cuda_dev), [0, C], radius=r, n_iter=8).cpu().numpy()
        truth = oi.propagate_stencil(x, e, r, 10, 8).reshape(C, h, w)
        assert np.abs(a - truth).max() < 1e-6


def test_labels_kernel(cuda_dev):
    rng = np.random.default_rng(0)
    for (C, h, w, H, W) in [(2, 24, 32, 96, 128), (1, 31, 23, 121, 90), (5, 19, 25, 75, 100), (3, 128, 128, 512, 512)]:
        rw = rng.random((C, 1, h, w)).astype(np.float32)
        rw[:, :, : h // 2] *= 0.2   # make the background plane win somewhere
        keys = sorted(rng.choice(20, C, replace=False).tolist())
        lab, idx, sc = indexing.rw_labels(_t(rw, cuda_dev), keys, (H, W), 0.25, want_index=True, want_scores=True)
        up = steps.upsample4_norm(torch.from_numpy(rw), (H, W)).numpy()
        assert np.abs(sc.cpu().numpy() - up).max() < 1e-6
        ref = steps.sem_seg_labels(torch.from_numpy(rw), keys, (H, W), 0.25)
        got = lab.cpu().numpy()
        bad = got != ref
        if bad.any():   # only pixels whose top-2 margin is at rounding level may differ
            full = np.concatenate([np.full((1, H, W), 0.25, np.float32), up], 0)
            srt = np.sort(full, 0)
            assert (srt[-1] - srt[-2])[bad].max() < 1e-6
        assert bad.mean() < 1e-4
        k = np.pad(np.asarray(keys) + 1, (1, 0))
        assert np.array_equal(k[idx.cpu().numpy()], got)


def test_sem_seg_labels_end_to_end_vs_reference_png(cuda_dev):
    """CAM npy + edge (oracle nets on the golden image) -> CUDA walk + labels vs the reference's PNG."""
    from oracle import nets
    g = np.load(golden_path("steps.npz"))
    irn_sd = synth.irn_state_dict()
    for i in range(len(g["ids"])):
        img = g["img%d" % i]
        H, W = img.shape[:2]
        x = synth.normalize_image(img)
        with torch.no_grad():
            edge, _ = nets.edge_displacement(torch.from_numpy(np.stack([x, x[..., ::-1].copy()])), irn_sd)
        rw = indexing.propagate_to_edge(_t(g["cam_cam%d" % i], cuda_dev), edge.to(cuda_dev), beta=10, exp_times=8, radius=5)
        lab, _, _ = indexing.rw_labels(rw, g["cam_keys%d" % i], (H, W), 0.25)
        assert (lab.cpu().numpy() != g["sem%d" % i]).mean() < 2e-3
