"""Oracle (test infrastructure only): the reference's make_cam -> make_sem_seg_labels chain for ONE image on the
CPU, restated with oracle.nets / oracle.steps / oracle.indexing.  Used by bench.py's cpu_baseline and
`--impl reference` legs and by the parity tests -- never by the product path."""
import time

import numpy as np
import torch
from PIL import Image

from . import indexing as oi
from . import nets, steps

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def _normalize(img):
    a = np.asarray(img)
    out = np.empty(a.shape, np.float32)
    for c in range(3):
        out[..., c] = (a[..., c] / 255. - MEAN[c]) / STD[c]     # voc12/dataloader.py:65-78
    return np.ascontiguousarray(out.transpose(2, 0, 1))


def msf_inputs(img_u8, scales):
    """voc12/dataloader.py:185-205: per scale PIL bicubic, normalise, CHW, stack with the W-flip."""
    H, W = img_u8.shape[:2]
    out = []
    for s in scales:
        im = img_u8 if s == 1 else np.asarray(Image.fromarray(img_u8).resize((int(np.round(W * s)), int(np.round(H * s))), Image.BICUBIC))
        x = _normalize(im)
        out.append(torch.from_numpy(np.stack([x, x[..., ::-1].copy()])))
    return out


def pseudo_label(img_u8, label, cam_sd, irn_sd, scales=(1.0, 0.5, 1.5, 2.0), beta=10, exp_times=8, bg_thres=0.25, walk="dense"):
    """Returns (label map uint8 [H,W], dict of stage seconds).  walk='dense' is the reference's algorithm
    (misc/indexing.py:112-139: (hw)^2 matrix squared exp_times times); walk='stencil' is the same operator iterated
    in float64 (minutes -> milliseconds; used only when the dense matrix would not fit the time budget)."""
    t = {}
    H, W = img_u8.shape[:2]
    with torch.no_grad():
        t0 = time.perf_counter()
        xs = msf_inputs(img_u8, scales)
        t["preprocess"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        outs = [nets.cam_forward(x, cam_sd) for x in xs]
        keys, cam, high = steps.cam_merge(outs, (H, W), torch.as_tensor(label))
        t["cam"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        edge, dp = nets.edge_displacement(xs[scales.index(1.0)], irn_sd)
        t["irn"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        if walk == "dense":
            rw = oi.propagate_to_edge(cam.numpy(), edge.numpy(), 5, beta, exp_times)
        else:
            rw = oi.propagate_stencil(cam.numpy(), edge.numpy(), 5, beta, 2 ** exp_times).astype(np.float32)
        t["walk"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        lab = steps.sem_seg_labels(torch.from_numpy(np.ascontiguousarray(rw, dtype=np.float32)), keys.numpy(), (H, W), bg_thres)
        t["labels"] = time.perf_counter() - t0
    return lab, t, {"keys": keys.numpy(), "cam": cam.numpy(), "edge": edge.numpy()}
