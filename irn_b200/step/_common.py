"""Shared plumbing of the three label-generation steps: collation, progress printing, process spawn."""
import torch
from torch.utils.data._utils.collate import default_collate


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


def spawn(work, n_gpus, args_tuple):
    """One process per GPU like the reference (multiprocessing.spawn, step/make_cam.py:74); a single visible GPU
    runs in-process."""
    if n_gpus <= 0:
        raise RuntimeError("irn_b200 steps need at least one CUDA device (there is no CPU fallback)")
    if n_gpus == 1:
        work(0, *args_tuple)
    else:
        torch.multiprocessing.spawn(work, nprocs=n_gpus, args=args_tuple, join=True)
