"""Instance path kernels: integer outputs bit-exact against the reference's own functions (golden fixtures) and
the oracle."""
import numpy as np
import pytest
import torch

from conftest import golden_path
from irn_b200 import instance, synth
from oracle import steps

pytestmark = pytest.mark.gpu


def test_find_centroids_bit_exact(cuda_dev):
    g = np.load(golden_path("instance_fns.npz"))
    for i in range(3):
        cen = instance.find_centroids_with_refinement(torch.from_numpy(g["dp%d" % i]).to(cuda_dev))
        assert np.array_equal(cen.cpu().numpy(), g["centroids%d" % i])
    # fewer iterations / noisy field against the oracle
    dp = synth.displacement(37, 61, 4, seed=9) + np.random.default_rng(1).normal(0, 0.5, (2, 37, 61)).astype(np.float32)
    cen = instance.find_centroids_with_refinement(torch.from_numpy(dp).to(cuda_dev), iterations=50)
    assert np.array_equal(cen.cpu().numpy(), steps.find_centroids(dp, 50))


def test_cluster_centroids_exact(cuda_dev):
    g = np.load(golden_path("instance_fns.npz"))
    for i in range(3):
        shape = tuple(g["instances_shape%d" % i])
        ref = np.unpackbits(g["instances%d" % i], axis=-1, count=shape[-1]).astype(bool).reshape(shape)
        inst, n = instance.cluster_centroids(torch.from_numpy(g["centroids%d" % i]).to(cuda_dev), torch.from_numpy(g["dp%d" % i]).to(cuda_dev))
        assert n == shape[0]
        got = steps.one_hot(inst.cpu().numpy(), n)
        assert np.array_equal(got, ref)


def test_connected_components_vs_scipy(cuda_dev):
    rng = np.random.default_rng(3)
    for (h, w, nv) in [(64, 64, 1), (97, 131, 3), (512, 512, 4), (1, 17, 2), (33, 1, 1)]:
        vals = (rng.random((h, w)) < 0.6).astype(np.int32) * rng.integers(1, nv + 1, (h, w)).astype(np.int32)
        lab = instance.connected_components(torch.from_numpy(vals).to(cuda_dev)).cpu().numpy()
        assert ((lab == 0) == (vals == 0)).all()
        for v in range(1, nv + 1):   # per value, same partition as scipy and raster-order numbering
            ref = steps.cc_label(vals == v)
            m = vals == v
            ids = np.unique(lab[m])
            assert len(ids) == ref.max()
            assert np.array_equal(np.searchsorted(ids, lab[m]) + 1, ref[m])


def test_instance_seeds_and_detect(cuda_dev):
    rng = np.random.default_rng(5)
    K, I, h, w = 2, 3, 20, 24
    cams = rng.random((K, h, w)).astype(np.float32)
    inst = rng.integers(0, I, (h, w)).astype(np.int32)
    out = instance.separate_score_by_mask(torch.from_numpy(cams).to(cuda_dev), torch.from_numpy(inst).to(cuda_dev), I).cpu().numpy()
    ref = cams[:, None] * steps.one_hot(inst, I)[None].astype(np.float32)
    assert np.array_equal(out, ref)
    # detect_instance against the oracle restatement
    H, W, C = 60, 72, 4
    idx = (rng.integers(0, C + 1, (H // 6, W // 6)).repeat(6, 0).repeat(6, 1)).astype(np.int32)
    sc = rng.random((C, H, W)).astype(np.float32)
    got = instance.detect_instance(torch.from_numpy(sc).to(cuda_dev), torch.from_numpy(idx).to(cuda_dev), [3, 3, 7, 7], max_fragment_size=40)
    ref = steps.detect_instance(sc, steps.one_hot(idx, C + 1)[1:], [3, 3, 7, 7], max_fragment_size=40)
    assert np.array_equal(got["mask"], ref["mask"]) and np.array_equal(got["class"], ref["class"])
    assert np.array_equal(got["score"].astype(np.float32), np.asarray(ref["score"], np.float32))
