"""Drop-in for the reference's ``step/make_cam.py``: same ``run(args)`` namespace, same ``.npy`` output
(``{"keys": LongTensor[K], "cam": FloatTensor[K,h/4,w/4], "high_res": ndarray[K,H,W]}``, step/make_cam.py:55-56),
CUDA arithmetic from libirn_b200."""
import importlib
import os

import numpy as np
import torch
from torch import cuda
from torch.utils.data import DataLoader

from .. import cam_ops
from ..misc import torchutils
from ..voc12 import dataloader as voc_data
from . import _common


def _work(process_id, model, dataset, args):
    databin = dataset[process_id]
    n_gpus = torch.cuda.device_count()
    loader = DataLoader(databin, shuffle=False, num_workers=args.num_workers // n_gpus, pin_memory=False, collate_fn=_common.collate_one)
    with torch.no_grad(), cuda.device(process_id):
        model.cuda()
        for it, pack in enumerate(loader):
            img_name = pack["name"][0]
            label = pack["label"][0]
            size = pack["size"]
            outputs = [model(img[0].cuda(non_blocking=True)) for img in pack["img"]]   # step/make_cam.py:35-36
            keys, strided_cam, highres_cam = cam_ops.merge_cams(outputs, size, label)   # step/make_cam.py:38-52
            np.save(os.path.join(args.cam_out_dir, img_name + ".npy"),
                    {"keys": keys, "cam": strided_cam.cpu(), "high_res": highres_cam.cpu().numpy()})
            _common.progress(process_id, n_gpus, it, len(databin))


def make_dataset(args, list_path, scales):
    if getattr(args, "synthetic", 0):
        return voc_data.SyntheticMSF(int(args.synthetic), scales=scales, name_list=list_path if os.path.exists(list_path) else None)
    return voc_data.VOC12ClassificationDatasetMSF(list_path, voc12_root=args.voc12_root, scales=scales)


def run(args):
    model = getattr(importlib.import_module(args.cam_network), "CAM")()
    model.load_state_dict(torch.load(args.cam_weights_name + ".pth"), strict=True)   # step/make_cam.py:64 (sic: '.pth' appended)
    model.eval()
    n_gpus = torch.cuda.device_count()
    dataset = make_dataset(args, args.train_list, args.cam_scales)
    dataset = torchutils.split_dataset(dataset, n_gpus)
    print("[ ", end="")
    _common.spawn(_work, n_gpus, (model, dataset, args))
    print("]")
    torch.cuda.empty_cache()
