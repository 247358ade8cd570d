"""Drop-in for the reference's ``step/make_sem_seg_labels.py``: IRNet edge map, random walk of the stored CAMs,
x4 upsample / background threshold / argmax, uint8 PNG per image (step/make_sem_seg_labels.py:28-51)."""
import os

import numpy as np
from PIL import Image

from .. import indexing
from ..voc12 import dataloader as voc_data
from . import _common


def sem_seg_one_image(model, pack, args):
    name = voc_data.decode_int_filename(pack["name"][0])
    edge, _ = model(pack["img"][0].cuda(non_blocking=True))
    stored = np.load(os.path.join(args.cam_out_dir, name + ".npy"), allow_pickle=True).item()
    walk = indexing.propagate_to_edge(stored["cam"].cuda(), edge, beta=float(args.beta), exp_times=int(args.exp_times), radius=5)
    labels, _, _ = indexing.rw_labels(walk, np.asarray(stored["keys"]), pack["size"], float(args.sem_seg_bg_thres))
    Image.fromarray(labels.cpu().numpy()).save(os.path.join(args.sem_seg_out_dir, name + ".png"))


def _save(ctx, names, labels, out_dir):
    (lab,) = ctx.writer.to_host([labels])
    lab = lab.numpy()
    for i, name in enumerate(names):
        # compress_level 1: same pixels in the file, a quarter of zlib's time on busy label maps (the default 6 made PNG encoding
        # the pacer of the whole step: 21 ms per 512x512 map against ~1 ms of GPU time)
        ctx.writer.submit_file(lambda a, p: Image.fromarray(a).save(p, compress_level=1), lab[i], os.path.join(out_dir, name + ".png"))


def sem_seg_batch(ctx, packs):
    """step/make_sem_seg_labels.py:28-51 for a bucket of equally-sized images: one IRNet forward, one batched walk."""
    args = ctx.args
    names = [voc_data.decode_int_filename(p["name"][0]) for p in packs]
    with ctx.phase("cam dicts"):
        keys, cams = _common.load_cam_dicts(ctx, packs, names, args.cam_out_dir)
    with ctx.phase("stack + upload images"):
        x = ctx.stack_images(packs)
    with ctx.phase("irn forward (issue)"):
        x1 = ctx.pipe.pyramids(x, (1.0,))[0]
        edges, _ = ctx.pipe.irn_stage(x1)
    with ctx.phase("upload cams"):
        seeds = _common.to_device_list(ctx, cams)
    with ctx.phase("walk (issue)"):
        rw, counts = ctx.pipe.walk_stage(seeds, edges)
    with ctx.phase("labels (issue)"):
        labels = ctx.pipe.label_stage(rw, counts, keys, packs[0]["size"], float(args.sem_seg_bg_thres))
    with ctx.phase("hand to writer"):
        ctx.writer.submit(_save, ctx, names, labels, args.sem_seg_out_dir)


def _work(process_id, model, dataset, args):
    _common.work_loop(process_id, model, dataset, args, sem_seg_one_image, sem_seg_batch)


def run(args):
    _common.run_step(args, _work, args.irn_network, "EdgeDisplacement", args.irn_weights_name, False, args.infer_list, (1.0,), opening="[",
                     cam_dir=args.cam_out_dir)
