"""C1 pre-processing, CPU side: the oracle's restatement of Pillow's 8-bit bicubic resampler against the installed
Pillow (the third-party code misc/imutils.py:8-22 calls), the normalisation table against TorchvisionNormalize, and
the C ABI's host-side coefficient / table functions against the oracle.  All integer / table work: exact equality."""
import numpy as np
import pytest
from PIL import Image

from irn_b200 import preprocess, synth
from irn_b200.voc12 import dataloader as dl
from oracle import resize as R

SIZES = [(512, 512, 256, 256), (512, 512, 768, 768), (512, 512, 1024, 1024), (375, 500, 188, 250), (375, 500, 562, 750),
         (375, 500, 750, 1000), (333, 500, 166, 250), (500, 281, 250, 140), (7, 5, 14, 10), (16, 16, 8, 8), (100, 37, 51, 74),
         (3, 3, 1, 1), (2, 9, 5, 4), (64, 64, 64, 96), (64, 64, 32, 64)]


@pytest.mark.parametrize("H,W,oh,ow", SIZES)
def test_oracle_resize_equals_pillow(H, W, oh, ow):
    rng = np.random.default_rng(H * 1000 + W)
    for img in (rng.integers(0, 256, (H, W, 3), dtype=np.uint8), synth.image(3, H, W)):
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
        assert np.array_equal(R.pil_bicubic_resize_u8(img, oh, ow), ref)


def test_oracle_resize_random_size_pairs():
    rng = np.random.default_rng(7)
    for _ in range(120):
        H, W = (int(v) for v in rng.integers(1, 160, 2))
        oh, ow = (int(v) for v in rng.integers(1, 240, 2))
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
        assert np.array_equal(R.pil_bicubic_resize_u8(img, oh, ow), ref), (H, W, oh, ow)


def test_oracle_msf_equals_reference_loader_body():
    img = synth.image(5, 375, 500)
    scales = (1.0, 0.5, 1.5, 2.0)
    for a, b in zip(dl.multi_scale_flip(img, scales), R.msf_preprocess(img, scales)):
        assert a.dtype == b.dtype == np.float32 and np.array_equal(a, b)
    norm = dl.TorchvisionNormalize()
    ramp = np.repeat(np.arange(256, dtype=np.uint8)[:, None, None], 3, 2)
    assert np.array_equal(norm(ramp)[:, 0].T, R.normalize_lut())


@pytest.mark.parametrize("n_in,n_out", [(512, 256), (512, 768), (512, 1024), (375, 188), (500, 250), (375, 562), (500, 1000),
                                        (7, 14), (5, 10), (3, 1), (9, 4), (1, 7), (1000, 3), (333, 166), (281, 140)])
def test_abi_coefficients_equal_oracle(n_in, n_out):
    b, k = preprocess.resize_coeffs(n_in, n_out)
    rb, rk = R.pil_bicubic_coeffs(n_in, n_out)
    assert b.shape == rb.shape and k.shape == rk.shape
    assert np.array_equal(b, rb) and np.array_equal(k, rk)


def test_abi_lut_equals_oracle():
    assert np.array_equal(preprocess.normalize_lut(), R.normalize_lut())
    assert preprocess.rescaled_size(375, 500, 0.5) == R.rescaled_size(375, 500, 0.5) == (188, 250)


def test_oracle_resize_equals_recorded_pillow():
    """The committed outputs of Pillow 12.2.0 (tests/golden/make_resize_golden.py): the pin that travels with the repo."""
    from conftest import golden_path
    g = np.load(golden_path("resize_pillow.npz"))
    for i, (H, W, oh, ow) in enumerate(g["cases"]):
        assert np.array_equal(R.pil_bicubic_resize_u8(g["img%d" % i], int(oh), int(ow)), g["out%d" % i])


def test_rescaled_size_matches_imutils_rounding():
    """misc/imutils.py:19-22 rounds half to even (np.round); the device path must request the same sizes."""
    from irn_b200.misc import imutils
    for H, W in [(375, 500), (333, 500), (281, 500), (500, 375), (3, 5), (512, 512)]:
        img = np.zeros((H, W, 3), np.uint8)
        for s in (0.5, 1.5, 2.0, 0.75, 1.25):
            assert imutils.pil_rescale(img, s, 3).shape[:2] == preprocess.rescaled_size(H, W, s)


def test_abi_resize_argument_errors():
    from irn_b200 import _lib
    L = _lib.lib()
    assert L.irn_resize_ksize(0, 4) == 0 and L.irn_resize_ksize(4, 0) == 0
    assert L.irn_resize_ksize(512, 256) == 9 and L.irn_resize_ksize(512, 1024) == 5     # support 2*scale (down) / 2 (up), both sides + centre
    b = np.zeros((4, 2), np.int32)
    k = np.zeros((4, 9), np.int32)
    assert L.irn_resize_coeffs(0, 4, b.ctypes.data, k.ctypes.data) != 0
    assert L.irn_resize_coeffs(8, 4, None, k.ctypes.data) != 0
    assert b"irn_resize_coeffs" in L.irn_last_error()
    assert L.irn_normalize_lut(None, None, None) != 0
    assert L.irn_resize_workspace_bytes(None, 1) == 0
