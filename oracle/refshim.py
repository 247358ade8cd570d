"""Import the UNMODIFIED reference from /root/reference on a CPU-only box.

Used only by tests/golden/make_golden.py (build container; /root/reference does not exist on
the GPU box and nothing run there imports this module).  No reference source is copied:
the shims below only make its hard-wired `.cuda()` calls, weight download and absent
third-party imports harmless.
"""
import os
import sys
import types

REF = os.environ.get("IRN_REFERENCE", "/root/reference")


def install(seed_state_dict_fn=None):
    """Patch the environment and put the reference on sys.path.  Returns nothing; import
    `misc.indexing`, `net.resnet50_cam`, `step.make_cam`, ... afterwards."""
    import numpy as np
    import torch

    if not os.path.isdir(REF):
        raise RuntimeError("reference tree not found at %s" % REF)
    if REF not in sys.path:
        sys.path.insert(0, REF)

    # D4: `.cuda()` everywhere -> identity on a CPU box
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.device_count = lambda: 1

    class _NullDev:
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False
    torch.cuda.device = _NullDev
    torch.cuda.empty_cache = lambda: None

    # D11: absent third-party modules
    for name in ("pydensecrf", "pydensecrf.densecrf", "pydensecrf.utils"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.unary_from_labels = None
            sys.modules[name] = m
    if "imageio" not in sys.modules:
        from PIL import Image
        m = types.ModuleType("imageio")
        m.saved = {}

        def imread(path):
            return np.asarray(Image.open(path))

        def imsave(path, arr):
            m.saved[path] = np.asarray(arr).copy()
            Image.fromarray(np.asarray(arr)).save(path)
        m.imread = imread
        m.imsave = imsave
        m.imwrite = imsave
        sys.modules["imageio"] = m
    if "skimage" not in sys.modules:
        import scipy.ndimage as ndi
        sk = types.ModuleType("skimage")
        meas = types.ModuleType("skimage.measure")

        def label(a, connectivity=1, background=0):
            assert connectivity == 1
            return ndi.label(np.asarray(a) != background)[0]
        meas.label = label
        sk.measure = meas
        sys.modules["skimage"] = sk
        sys.modules["skimage.measure"] = meas

    # np.bool (misc/pyutils.py:86) disappeared in numpy >= 1.24
    if not hasattr(np, "bool"):
        np.bool = bool

    # D8: no network -> the ImageNet download returns a seeded dict
    import net.resnet50 as r50

    def fake_load_url(*a, **k):
        if seed_state_dict_fn is not None:
            return seed_state_dict_fn()
        m = r50.ResNet(r50.Bottleneck, [3, 4, 6, 3], strides=(2, 2, 2, 1))
        sd = m.state_dict()
        sd["fc.weight"] = torch.zeros(1)
        sd["fc.bias"] = torch.zeros(1)
        return sd
    r50.model_zoo.load_url = fake_load_url
