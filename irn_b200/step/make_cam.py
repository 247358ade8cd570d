"""Drop-in for the reference's ``step/make_cam.py``: same ``run(args)`` namespace and the same ``.npy`` output --
a pickled ``{"keys": LongTensor[K], "cam": FloatTensor[K,h/4,w/4], "high_res": ndarray[K,H,W]}`` per image
(step/make_cam.py:55-56) -- with the arithmetic in libirn_b200."""
import os

import numpy as np
import torch

from .. import cam_ops
from . import _common


def cam_one_image(model, pack, args):
    """step/make_cam.py:28-56: CAM forward per scale, merge + normalise, save."""
    per_scale = [model(x[0].cuda(non_blocking=True)) for x in pack["img"]]
    keys, strided, highres = cam_ops.merge_cams(per_scale, pack["size"], pack["label"][0])
    np.save(os.path.join(args.cam_out_dir, pack["name"][0] + ".npy"),
            {"keys": keys, "cam": strided.cpu(), "high_res": highres.cpu().numpy()})


def _save(ctx, names, keys, strided, highres, out_dir):
    """Pool thread: one device->host copy per tensor kind for the whole batch, then one np.save per image."""
    counts = [int(k.size) for k in keys]
    lo, hi = ctx.writer.to_host([strided, highres])
    o = 0
    for name, k, c in zip(names, keys, counts):
        # views of the batch's host buffers: pickling copies them, so the buffers live exactly as long as the last write
        ctx.writer.submit_file(np.save, os.path.join(out_dir, name + ".npy"),
                               {"keys": torch.from_numpy(k.astype(np.int64)), "cam": lo[o:o + c].clone(), "high_res": hi[o:o + c].numpy()})
        o += c


def cam_batch(ctx, packs):
    """The same body for a bucket of equally-sized decoded images (irn_b200.pipeline stages C1-C4)."""
    args = ctx.args
    with ctx.phase("stack + upload images"):
        x = ctx.stack_images(packs)
    labels = torch.cat([p["label"] for p in packs], 0)
    with ctx.phase("pyramids + cam forward + merge (issue)"):
        xs = ctx.pipe.pyramids(x, ctx.scales)
        keys, strided, highres = ctx.pipe.cam_stage(xs, labels, packs[0]["size"], want_highres=True, scales=ctx.scales)
    with ctx.phase("hand to writer"):
        ctx.writer.submit(_save, ctx, [p["name"][0] for p in packs], keys, strided, highres, args.cam_out_dir)


def _work(process_id, model, dataset, args):
    _common.work_loop(process_id, model, dataset, args, cam_one_image, cam_batch)


def run(args):
    # the reference appends '.pth' to --cam_weights_name here (step/make_cam.py:64) but not for the IRN weights
    _common.run_step(args, _work, args.cam_network, "CAM", args.cam_weights_name + ".pth", True, args.train_list, args.cam_scales)
