"""Instance path (P1-P4) host wrappers over the C ABI: step/make_ins_seg_labels.py:18-105."""

import numpy as np
import torch

from . import _lib


def find_centroids_with_refinement(displacement, iterations=300):
    """step/make_ins_seg_labels.py:18-56.  displacement cuda fp32 [2,h,w] -> cuda int32 [2,h,w] (y, x)."""
    _lib.require_cuda(displacement)
    dp = displacement.contiguous().float()
    _, h, w = dp.shape
    out = torch.empty((2, h, w), dtype=torch.int32, device=dp.device)
    with torch.cuda.device(dp.device):
        _lib.check(_lib.lib().irn_find_centroids(_lib.ptr(dp), _lib.ptr(out), h, w, int(iterations), _lib.stream_ptr()), "irn_find_centroids")
    return out


def connected_components(values):
    """4-connected components of equal non-zero values.  values cuda int32 [h,w] -> labels int32 [h,w]
    (0 = background, else 1 + first raster index of the component)."""
    _lib.require_cuda(values)
    v = values.contiguous().to(torch.int32)
    h, w = v.shape
    labels = torch.empty_like(v)
    scratch = torch.empty(h * w, dtype=torch.int32, device=v.device)
    with torch.cuda.device(v.device):
        _lib.check(_lib.lib().irn_connected_components(_lib.ptr(v), _lib.ptr(labels), h, w, _lib.ptr(scratch), _lib.stream_ptr()),
                   "irn_connected_components")
    return labels


def cluster_centroids(centroids, displacement, thres=2.5):
    """step/make_ins_seg_labels.py:58-75.  Returns (instance_map cuda int32 [h,w] in 0..I-1, I); the reference's
    bool [I,h,w] is one_hot(instance_map)."""
    _lib.require_cuda(centroids, displacement)
    L = _lib.lib()
    dp = displacement.contiguous().float()
    cen = centroids.contiguous().to(torch.int32)
    _, h, w = dp.shape
    inst = torch.empty((h, w), dtype=torch.int32, device=dp.device)
    count = torch.zeros(1, dtype=torch.int32, device=dp.device)
    scratch = torch.empty(L.irn_cluster_scratch_bytes(h, w), dtype=torch.uint8, device=dp.device)
    with torch.cuda.device(dp.device):
        _lib.check(L.irn_cluster_centroids(_lib.ptr(dp), _lib.ptr(cen), float(thres), _lib.ptr(inst), _lib.ptr(count), h, w, _lib.ptr(scratch),
                                           _lib.stream_ptr()), "irn_cluster_centroids")
    return inst, int(count.item())


def separate_score_by_mask(cams, instance_map, n_instances):
    """step/make_ins_seg_labels.py:77-80 ("separte_score_by_mask"): cams cuda [K,h,w] -> [K,I,h,w]."""
    _lib.require_cuda(cams, instance_map)
    c = cams.contiguous().float()
    K, h, w = c.shape
    out = torch.empty((K, n_instances, h, w), dtype=torch.float32, device=c.device)
    with torch.cuda.device(c.device):
        _lib.check(_lib.lib().irn_instance_seeds(_lib.ptr(c), _lib.ptr(instance_map.contiguous()), K, int(n_instances), h, w, _lib.ptr(out),
                                                 _lib.stream_ptr()), "irn_instance_seeds")
    return out


def detect_instance(scores, index, class_ids, max_fragment_size=0):
    """step/make_ins_seg_labels.py:82-105.  scores cuda fp32 [C,H,W] (normalised upsampled walk), index cuda int32 [H,W]
    (argmax with background 0), class_ids: int sequence of length C.  Returns the reference's dict of numpy arrays
    {'score' f32[M], 'mask' bool[M,H,W], 'class' int64[M]} ordered by (channel, raster order of the segment).
    Components, per-segment area / max score and the mask planes are built on the device; the host only orders the
    (few) segments."""
    _lib.require_cuda(scores, index)
    L = _lib.lib()
    C, H, W = scores.shape
    dev = scores.device
    index = index.contiguous()
    labels = connected_components(index)
    area = torch.empty(H * W + 1, dtype=torch.int32, device=dev)
    mx = torch.empty(H * W + 1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.irn_segment_stats(_lib.ptr(labels), _lib.ptr(index), _lib.ptr(scores.contiguous()), H, W, _lib.ptr(area),
                                       _lib.ptr(mx), _lib.stream_ptr()), "irn_segment_stats")
    area_h = area.cpu().numpy()
    seg_ids = np.nonzero(area_h)[0].astype(np.int32)                # ascending = raster order of first pixels
    if seg_ids.size == 0:
        return {"score": np.stack([], 0), "mask": None, "class": None}   # np.stack([]) raises like the reference
    seg_area = area_h[seg_ids]
    ids_dev = torch.from_numpy(seg_ids).to(dev)
    seg_max = mx[ids_dev.long()].cpu().numpy().view(np.float32)
    seg_chan = index.reshape(-1)[(ids_dev - 1).long()].cpu().numpy() - 1   # channel of the segment's first pixel
    order = np.lexsort((seg_ids, seg_chan))
    M = int(order.size)
    masks = torch.empty((M, H, W), dtype=torch.uint8, device=dev)
    ordered = torch.from_numpy(np.ascontiguousarray(seg_ids[order])).to(dev)
    with torch.cuda.device(dev):
        _lib.check(L.irn_segment_masks(_lib.ptr(labels), _lib.ptr(ordered), M, H, W, _lib.ptr(masks), _lib.stream_ptr()), "irn_segment_masks")
    score = np.where(seg_area[order] < max_fragment_size, np.float32(0), seg_max[order]).astype(np.float32)
    cls = np.asarray(class_ids)[seg_chan[order]]
    return {"score": score, "mask": masks.cpu().numpy().view(np.bool_), "class": cls}
