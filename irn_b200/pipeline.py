"""Batched pseudo-label pipeline: the reference's make_cam -> make_sem_seg_labels chain
(step/make_cam.py:28-56, step/make_sem_seg_labels.py:28-51) for a batch of equally-sized images, kept on the
device end to end (the reference round-trips every image through a .npy file and one-image kernels).

    multi-scale CAM (C2-C4) -> EdgeDisplacement (I1/I2) -> random walk (R1-R6) -> label map (S1)
"""
import numpy as np
import torch

from . import cam_ops, indexing, preprocess
from .voc12 import dataloader as voc_data


def preprocess_batch(images_u8, scales=(1.0, 0.5, 1.5, 2.0), pin=True):
    """C1 on the host (voc12/dataloader.py:185-205): list of uint8 [H,W,3] of one size -> list over scales of
    float32 tensors [2N,3,h_s,w_s] (image, flipped image interleaved), pinned for async H2D."""
    per_scale = [[] for _ in scales]
    for img in images_u8:
        ms = voc_data.multi_scale_flip(img, scales)
        ms = ms if isinstance(ms, list) else [ms]
        for s, a in enumerate(ms):
            per_scale[s].append(a)
    out = []
    for lst in per_scale:
        t = torch.from_numpy(np.ascontiguousarray(np.concatenate(lst, 0)))
        out.append(t.pin_memory() if pin and torch.cuda.is_available() else t)
    return out


class PseudoLabelPipeline:
    def __init__(self, cam_model, irn_model, device, scales=(1.0, 0.5, 1.5, 2.0), beta=10, exp_times=8, bg_thres=0.25,
                 cam_sub_batch=8, rw_sub_batch=64):
        self.cam, self.irn, self.device = cam_model, irn_model, device
        self.scales, self.beta, self.exp_times, self.bg = scales, beta, exp_times, bg_thres
        self.cam_sub, self.rw_sub = cam_sub_batch, rw_sub_batch
        self._copy_stream = None
        self._staging = {}
        if 1.0 not in scales:
            raise ValueError("the IRNet pass uses the scale-1.0 input (step/make_sem_seg_labels.py:64-66)")

    @torch.no_grad()
    def run_u8(self, images_u8, labels, want_highres=True):
        """Same as run(), starting from decoded images: uint8 [N,H,W,3], cuda or (pinned) host.  The per-scale bicubic
        rescale, normalisation and flip stack of the reference's loader (voc12/dataloader.py:191-201) run on the device,
        bit-exact (irn_b200.preprocess), so a step uploads 0.79 MB per 512x512 image instead of 47 MB of fp32 pyramids."""
        dev = self.device
        x = images_u8
        if not x.is_cuda:
            buf = self._staging.get("u8")
            if buf is None or buf.shape != x.shape:
                buf = torch.empty(x.shape, dtype=torch.uint8, device=dev)
                self._staging["u8"] = buf
            buf.copy_(x, non_blocking=True)
            x = buf
        N, H, W = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        inputs = []
        for k, s in enumerate(self.scales):
            oh, ow = (H, W) if s == 1 else preprocess.rescaled_size(H, W, s)
            buf = self._staging.get(k)
            if buf is None or tuple(buf.shape) != (2 * N, 3, oh, ow):
                buf = torch.empty((2 * N, 3, oh, ow), dtype=torch.float32, device=dev)
                self._staging[k] = buf
            inputs.append(preprocess.resize_normalize(x, (oh, ow), out=buf))
        return self.run(inputs, labels, (H, W), want_highres)

    @torch.no_grad()
    def run(self, inputs, labels, size, want_highres=True):
        """inputs: list over scales of [2N,3,h_s,w_s] fp32 (cuda, or pinned host -> copied here);
        labels: [N,20] multi-hot; size=(H,W).  Returns dict with 'labels' uint8 cuda [N,H,W], 'keys' list,
        'cams' list of cuda [K_i,h4,w4] and 'high_res' list (or None)."""
        dev = self.device
        # host inputs: copy on a side stream, one event per scale, so the forward of scale k overlaps the H2D of scale k+1..
        main = torch.cuda.current_stream(dev)
        if any(not x.is_cuda for x in inputs):
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=dev)
            self._copy_stream.wait_stream(main)          # buffers freed on the main stream may be recycled for the copies
        order = sorted(range(len(inputs)), key=lambda k: inputs[k].numel())     # smallest scale first: least exposed copy
        xs, ready = [None] * len(inputs), [None] * len(inputs)
        for k in order:
            x = inputs[k]
            if x.is_cuda:
                xs[k] = x
            else:
                # persistent device staging buffers (no allocator traffic across streams); the wait_stream above guarantees the
                # previous step has finished reading them
                buf = self._staging.get(k)
                if buf is None or buf.shape != x.shape:
                    buf = torch.empty(x.shape, dtype=torch.float32, device=dev)
                    self._staging[k] = buf
                with torch.cuda.stream(self._copy_stream):
                    buf.copy_(x, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self._copy_stream)
                xs[k] = buf
                ready[k] = ev
        N = xs[0].shape[0] // 2
        # ---- C2/C3: CAM forward per scale, in sub-batches of image pairs
        cams = [None] * len(xs)
        for k in order:                                   # compute in copy order; `cams` stays in self.scales order for the merge
            x, s = xs[k], self.scales[k]
            if ready[k] is not None:
                main.wait_event(ready[k])
            # sub-batch so that every forward sees about the same number of pixels (cam_sub images at scale 2.0):
            # small scales batch more images to keep all SMs busy, large scales bound the activation arena
            sub = max(1, int(self.cam_sub * (2.0 / s) ** 2))
            outs = [self.cam.forward_batch(x[2 * i:2 * min(i + sub, N)]) for i in range(0, N, sub)]
            cams[k] = torch.cat(outs, 0)
        # ---- C4: merge + normalise per image (classes present differ per image)
        keys, strided, highres = [], [], []
        for i in range(N):
            k, lo, hi = cam_ops.merge_cams([c[i] for c in cams], size, labels[i])
            keys.append(k.numpy())
            strided.append(lo)
            highres.append(hi if want_highres else None)
        # ---- I1/I2: edge + displacement
        x1 = xs[self.scales.index(1.0)]
        edges = []
        irn_sub = self.cam_sub * 4
        for i in range(0, N, irn_sub):
            e, _ = self.irn.forward_batch(x1[2 * i:2 * min(i + irn_sub, N)])
            edges.append(e[:, 0])
        edges = torch.cat(edges, 0)
        # ---- R1-R6: batched walk (the fused cluster kernel takes every (image, class) of the sub-batch in one launch)
        counts = [int(s.shape[0]) for s in strided]
        rws = []
        for i in range(0, N, self.rw_sub):
            j = min(i + self.rw_sub, N)
            offs = np.concatenate([[0], np.cumsum(counts[i:j])])
            seeds = torch.cat(strided[i:j], 0)
            rws.append(indexing.random_walk_batch(seeds, edges[i:j], offs, 5, self.beta, 2 ** self.exp_times))
        rw = torch.cat(rws, 0)
        # ---- S1: label maps
        out = torch.empty((N, size[0], size[1]), dtype=torch.uint8, device=dev)
        o = 0
        for i in range(N):
            if counts[i]:
                lab, _, _ = indexing.rw_labels(rw[o:o + counts[i]], keys[i], size, self.bg)
                out[i] = lab
            else:
                out[i].zero_()
            o += counts[i]
        return {"labels": out, "keys": keys, "cams": strided, "high_res": highres if want_highres else None, "edge": edges}
