"""Drop-in for the reference's ``step/make_sem_seg_labels.py``: EdgeDisplacement forward, random walk of the
stored CAMs along the edge map, x4 upsample / background threshold / argmax -> uint8 PNG
(step/make_sem_seg_labels.py:28-51)."""
import importlib
import os

import numpy as np
import torch
from PIL import Image
from torch import cuda
from torch.utils.data import DataLoader

from .. import indexing
from ..misc import torchutils
from ..voc12 import dataloader as voc_data
from . import _common
from .make_cam import make_dataset


def _work(process_id, model, dataset, args):
    n_gpus = torch.cuda.device_count()
    databin = dataset[process_id]
    loader = DataLoader(databin, shuffle=False, num_workers=args.num_workers // n_gpus, pin_memory=False, collate_fn=_common.collate_one)
    with torch.no_grad(), cuda.device(process_id):
        model.cuda()
        for it, pack in enumerate(loader):
            img_name = voc_data.decode_int_filename(pack["name"][0])
            size = pack["size"]
            edge, dp = model(pack["img"][0].cuda(non_blocking=True))
            cam_dict = np.load(args.cam_out_dir + "/" + img_name + ".npy", allow_pickle=True).item()
            cams = cam_dict["cam"].cuda()
            keys = np.asarray(cam_dict["keys"])
            rw = indexing.propagate_to_edge(cams, edge, beta=float(args.beta), exp_times=int(args.exp_times), radius=5)
            labels, _, _ = indexing.rw_labels(rw, keys, size, float(args.sem_seg_bg_thres))
            Image.fromarray(labels.cpu().numpy()).save(os.path.join(args.sem_seg_out_dir, img_name + ".png"))
            _common.progress(process_id, n_gpus, it, len(databin))


def run(args):
    model = getattr(importlib.import_module(args.irn_network), "EdgeDisplacement")()
    model.load_state_dict(torch.load(args.irn_weights_name), strict=False)
    model.eval()
    n_gpus = torch.cuda.device_count()
    dataset = make_dataset(args, args.infer_list, (1.0,))
    dataset = torchutils.split_dataset(dataset, n_gpus)
    print("[", end="")
    _common.spawn(_work, n_gpus, (model, dataset, args))
    print("]")
    torch.cuda.empty_cache()
