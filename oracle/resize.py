"""Oracle (test infrastructure only): restatement of the image pre-processing of row C1.

 * pil_bicubic_coeffs / pil_bicubic_resize_u8
       misc/imutils.py:8-22 calls ``Image.fromarray(img).resize(size[::-1], Image.BICUBIC)``.  The arithmetic
       lives in Pillow (third-party, not under /root/reference; the reference pins nothing:
       ``imageio>=2.5.0`` pulls whatever Pillow is current; this image has Pillow 12.2.0).  Restated from
       Pillow's published algorithm (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
       ImagingResampleHorizontal_8bpc / Vertical_8bpc): separable two-pass convolution, horizontal first,
       coefficients evaluated in double, normalised, rounded half away from zero to 22 fractional bits, 8-bit
       intermediate, accumulators start at 2^21, arithmetic shift, clamp to [0, 255].
       Pinned: bit-exact against the installed Pillow in tests/test_preprocess_cpu.py (many size pairs).
 * normalize_lut / msf_preprocess
       voc12/dataloader.py:65-78 (TorchvisionNormalize: float64 arithmetic, one rounding to float32) and
       :191-201 (per scale: rescale, normalise, HWC->CHW, stack with the W-flip).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_bicubic_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box.
    Returns (bounds int32 [out,2] = (first source index, tap count), kk int32 [out,ksize])."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size     # box corners are C floats
    filterscale = scale if scale >= 1.0 else 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, bounds, kk, axis):
    """One 8-bit resampling pass along `axis` (0 = vertical, 1 = horizontal) of a uint8 [H,W,C] image."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    n_out = bounds.shape[0]
    out = np.empty((n_out,) + src.shape[1:], np.uint8)
    for i in range(n_out):
        x0, n = int(bounds[i, 0]), int(bounds[i, 1])
        acc = np.tensordot(kk[i, :n].astype(np.int64), src[x0:x0 + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[i] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_bicubic_resize_u8(img, out_h, out_w):
    """uint8 [H,W,C] -> uint8 [out_h,out_w,C], == np.asarray(Image.fromarray(img).resize((out_w,out_h), BICUBIC))."""
    H, W = img.shape[:2]
    if (out_h, out_w) == (H, W):
        return img.copy()
    out = img
    if out_w != W:
        out = _pass(out, *pil_bicubic_coeffs(W, out_w), axis=1)
    if out_h != H:
        out = _pass(out, *pil_bicubic_coeffs(H, out_h), axis=0)
    return out


def normalize_lut(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """fp32 [3,256]: TorchvisionNormalize on every possible byte (float64 arithmetic, rounded once)."""
    u = np.arange(256, dtype=np.float64)
    return np.stack([((u / 255. - mean[c]) / std[c]).astype(np.float32) for c in range(3)])


def rescaled_size(H, W, scale):
    # misc/imutils.py:19-22
    return int(np.round(H * scale)), int(np.round(W * scale))


def msf_preprocess(img_u8, scales, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """voc12/dataloader.py:191-201 on a decoded uint8 [H,W,3] image -> list of fp32 [2,3,h_s,w_s]."""
    lut = normalize_lut(mean, std)
    out = []
    H, W = img_u8.shape[:2]
    for s in scales:
        im = img_u8 if s == 1 else pil_bicubic_resize_u8(img_u8, *rescaled_size(H, W, s))
        chw = np.stack([lut[c][im[..., c]] for c in range(3)])
        out.append(np.stack([chw, chw[..., ::-1]], 0))
    return out
