"""Writes tests/golden/resize_pillow.npz: outputs of the Pillow that is installed in the build container (the third-party
code behind misc/imutils.py:8-22; the reference pins no version) for a few small images, so that the oracle's restatement
(oracle/resize.py) stays pinned to a recorded Pillow even on a box whose Pillow differs.  Run: python tests/golden/make_resize_golden.py"""
import os
import sys

import numpy as np
import PIL
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from irn_b200 import synth  # noqa: E402

CASES = [(48, 64, 24, 32), (48, 64, 72, 96), (48, 64, 96, 128), (37, 50, 18, 25), (37, 50, 56, 75), (9, 7, 18, 14), (5, 5, 1, 1)]


def main():
    out = {"pillow_version": np.array(PIL.__version__), "cases": np.array(CASES, np.int32)}
    rng = np.random.default_rng(2024)
    for i, (H, W, oh, ow) in enumerate(CASES):
        img = synth.image(i, H, W) if i % 2 == 0 else rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        out["img%d" % i] = img
        out["out%d" % i] = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "resize_pillow.npz"), **out)
    print("wrote resize_pillow.npz for Pillow", PIL.__version__)


if __name__ == "__main__":
    main()
