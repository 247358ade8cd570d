"""Numerical probe (CPU, not a test): CAM-level error of operand-splitting schemes for the
tensor-core convolutions, emulated with fp64 accumulation against the fp32 oracle network.
  3xtf32 : hi/lo tf32, hi*hi + hi*lo + lo*hi                       (current kernels)
  tf32+bf: tf32 hi*hi + bf16(hi)*bf16(lo) + bf16(lo)*bf16(hi)      (cross terms on the 2x-rate bf16 path)
  bf16x3 : b0*b0 + b0*b1 + b1*b0 with x = b0 + b1 + ..., bf16 parts
  bf16x6 : all products down to 2^-16
Run: python tests/probe_split_precision.py
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
from irn_b200 import synth
from oracle import nets

torch.set_num_threads(32)


def rt(x, keep):   # round-to-nearest(-ish, ties away) to `keep` explicit mantissa bits of fp32
    drop = 23 - keep
    u = x.contiguous().view(torch.int32)
    u = (u + (1 << (drop - 1))) & ~((1 << drop) - 1)
    return u.view(torch.float32)


def tf32(x): return rt(x, 10)
def bf16(x): return rt(x, 7)


real_conv = F.conv2d


def make(scheme):
    def conv(x, w, b=None, **kw):
        xd = lambda t: t.double()
        if scheme == "3xtf32":
            xh, wh = tf32(x), tf32(w); xl, wl = tf32(x - xh), tf32(w - wh)
            terms = [(xh, wh), (xh, wl), (xl, wh)]
        elif scheme == "tf32+bf":
            xh, wh = tf32(x), tf32(w); xl, wl = bf16(x - xh), bf16(w - wh)
            terms = [(xh, wh), (bf16(xh), wl), (xl, bf16(wh))]
        elif scheme == "tf32+bf1":   # cross terms as ONE bf16 product pair with the residual of the bf16 hi: x = xb + (x-xb)
            xh, wh = tf32(x), tf32(w)
            terms = [(xh, wh), (bf16(x), bf16(w - wh)), (bf16(x - xh), bf16(w))]
        elif scheme in ("bf16x3", "bf16x6"):
            x0, w0 = bf16(x), bf16(w); x1, w1 = bf16(x - x0), bf16(w - w0)
            terms = [(x0, w0), (x0, w1), (x1, w0)]
            if scheme == "bf16x6":
                x2, w2 = bf16(x - x0 - x1), bf16(w - w0 - w1)
                terms += [(x1, w1), (x0, w2), (x2, w0)]
        elif scheme == "tf32x1":
            terms = [(tf32(x), tf32(w))]
        y = sum(real_conv(xd(a), xd(c), None, **kw) for a, c in terms).float()
        if b is not None:
            y = y + b.view(1, -1, 1, 1)
        return y
    return conv


class Shim:
    def __init__(self, conv): self.conv2d = conv
    def __getattr__(self, k): return getattr(F, k)


def main():
    sd = synth.cam_state_dict()
    errs = {}
    for seed, (H, W) in enumerate([(64, 96), (128, 160)]):
        xi = synth.normalize_image(synth.image(5 + seed, H, W))
        xi = torch.from_numpy(np.stack([xi, xi[..., ::-1].copy()]))
        with torch.no_grad():
            nets.F = F
            ref = nets.cam_forward(xi, sd)
            for s in ["3xtf32", "bf16x3"]:
                nets.F = Shim(make(s))
                out = nets.cam_forward(xi, sd)
                errs.setdefault(s, []).append(float((out - ref).abs().max() / ref.max()))
            nets.F = F
    for s, e in errs.items():
        print("%-9s CAM normalised max-abs err per image: %s" % (s, ["%.3g" % v for v in e]))
    sd = synth.irn_state_dict()
    xi = synth.normalize_image(synth.image(7, 96, 128))
    xi = torch.from_numpy(np.stack([xi, xi[..., ::-1].copy()]))
    with torch.no_grad():
        nets.F = F
        e0, d0 = nets.irn_forward(xi, sd)
        for s in ["3xtf32", "tf32+bf", "bf16x3"]:
            nets.F = Shim(make(s))
            e1, d1 = nets.irn_forward(xi, sd)
            print("%-9s IRNet edge max-abs err %.3g   dp max-abs err %.3g (|dp| max %.3g)" % (s, float((e1 - e0).abs().max()), float((d1 - d0).abs().max()), float(d0.abs().max())))
        nets.F = F


if __name__ == "__main__":
    main()
