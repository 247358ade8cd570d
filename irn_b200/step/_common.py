"""Shared plumbing of the three label-generation steps.

The reference repeats the same skeleton in step/make_cam.py, step/make_sem_seg_labels.py and
step/make_ins_seg_labels.py: build the model class named on the command line, load its checkpoint, split the
image list over the visible GPUs with the stride partition, spawn one worker per GPU, loop over a batch-size-1
DataLoader.  Here that skeleton exists once; each step module only supplies its per-image function.
"""
import importlib
import os

import torch
from torch.utils.data import DataLoader
from torch.utils.data._utils.collate import default_collate

from .. import preprocess
from ..misc import torchutils
from ..voc12 import dataloader as voc_data


def collate_one(batch):
    """batch_size=1 collation like the reference's DataLoader, except that `size` stays a pair of python ints
    (torch's default collate makes it [tensor([H]), tensor([W])], which modern numpy can no longer use as a slice
    bound, step/make_sem_seg_labels.py:29,43)."""
    out = default_collate(batch)
    out["size"] = (int(batch[0]["size"][0]), int(batch[0]["size"][1]))
    return out


def progress(process_id, n_gpus, it, n_items):
    """The reference prints `iter % (len(databin)//20)` which divides by zero for shards < 20 images
    (SURVEY.md D9); same output, guarded."""
    step = max(n_items // 20, 1)
    if process_id == n_gpus - 1 and it % step == 0:
        print("%d " % ((5 * it + 1) // step), end="", flush=True)


def device_pyramid(args):
    """--device_pyramid (default on): loader workers only decode; the per-scale rescale / normalise / flip stack of
    voc12/dataloader.py:191-201 is built on the GPU, bit-identical (irn_b200.preprocess)."""
    return bool(getattr(args, "device_pyramid", True))


def make_dataset(args, list_path, scales):
    """VOC images from --voc12_root, or seeded synthetic ones with --synthetic N."""
    if getattr(args, "synthetic", 0):
        names = list_path if os.path.exists(list_path) else None
        return voc_data.SyntheticMSF(int(args.synthetic), scales=scales, name_list=names, decode_only=device_pyramid(args))
    return voc_data.VOC12ClassificationDatasetMSF(list_path, voc12_root=args.voc12_root, scales=scales,
                                                  decode_only=device_pyramid(args))


def attach_pyramid(pack, scales):
    """Turn a decode-only item into what the reference's loader yields: pack['img'] = [1,2,3,h,w] per scale (a single
    tensor when there is one scale, voc12/dataloader.py:200-201), already on the current device."""
    if "img_u8" not in pack:
        return pack
    pyr = preprocess.msf_batch(pack["img_u8"].cuda(non_blocking=True), scales)
    pack["img"] = pyr[0][None] if len(scales) == 1 else [p[None] for p in pyr]
    return pack


def work_loop(process_id, model, dataset, args, per_image):
    """One GPU's share: the reference's `_work(process_id, model, dataset, args)` signature and loop order
    (step/make_cam.py:16-59), with the per-image body supplied by the step."""
    shard = dataset[process_id]
    n_gpus = torch.cuda.device_count()
    loader = DataLoader(shard, shuffle=False, num_workers=args.num_workers // max(n_gpus, 1), pin_memory=False, collate_fn=collate_one)
    with torch.no_grad(), torch.cuda.device(process_id):
        model.cuda()
        scales = getattr(getattr(shard, "dataset", shard), "scales", (1.0,))
        for it, pack in enumerate(loader):
            per_image(model, attach_pyramid(pack, scales), args)
            progress(process_id, n_gpus, it, len(shard))


def run_step(args, work, module_name, class_name, weights_path, strict, list_path, scales, opening="[ "):
    """The reference's `run(args)` (e.g. step/make_cam.py:62-77): model class resolved by name, checkpoint loaded,
    stride partition (misc/torchutils.py:66-68), one process per GPU (a single GPU runs in-process)."""
    model = getattr(importlib.import_module(module_name), class_name)()
    model.load_state_dict(torch.load(weights_path), strict=strict)
    model.eval()
    n_gpus = torch.cuda.device_count()
    if n_gpus <= 0:
        raise RuntimeError("irn_b200 steps need at least one CUDA device (there is no CPU fallback)")
    shards = torchutils.split_dataset(make_dataset(args, list_path, scales), n_gpus)
    print(opening, end="")
    if n_gpus == 1:
        work(0, model, shards, args)
    else:
        torch.multiprocessing.spawn(work, nprocs=n_gpus, args=(model, shards, args), join=True)
    print("]")
    torch.cuda.empty_cache()
