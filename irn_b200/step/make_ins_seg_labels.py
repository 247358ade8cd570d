"""Drop-in for the reference's ``step/make_ins_seg_labels.py`` (step/make_ins_seg_labels.py:108-171): displacement
field -> centroids -> instance clusters, per-instance CAM seeds through the random walk, argmax, per-segment
detection dict saved as ``.npy`` ({'score','mask','class'})."""
import importlib
import os

import numpy as np
import torch
from torch import cuda
from torch.utils.data import DataLoader

from .. import indexing, instance
from ..misc import torchutils
from . import _common
from .make_cam import make_dataset


def _work(process_id, model, dataset, args):
    n_gpus = torch.cuda.device_count()
    databin = dataset[process_id]
    loader = DataLoader(databin, shuffle=False, num_workers=args.num_workers // n_gpus, pin_memory=False, collate_fn=_common.collate_one)
    with torch.no_grad(), cuda.device(process_id):
        model.cuda()
        for it, pack in enumerate(loader):
            img_name = pack["name"][0]
            size = pack["size"]
            edge, dp = model(pack["img"][0].cuda(non_blocking=True))
            cam_dict = np.load(args.cam_out_dir + "/" + img_name + ".npy", allow_pickle=True).item()
            cams = cam_dict["cam"].cuda()
            keys = np.asarray(cam_dict["keys"])
            centroids = instance.find_centroids_with_refinement(dp)
            instance_map, n_inst = instance.cluster_centroids(centroids, dp)
            instance_cam = instance.separate_score_by_mask(cams, instance_map, n_inst)
            rw = indexing.propagate_to_edge(instance_cam, edge, beta=float(args.beta), exp_times=int(args.exp_times), radius=5)
            _, index, scores = indexing.rw_labels(rw, None, size, float(args.ins_seg_bg_thres), want_index=True, want_scores=True)
            class_ids = np.repeat(keys, n_inst)                                   # step/make_ins_seg_labels.py:147
            detected = instance.detect_instance(scores, index, class_ids, max_fragment_size=size[0] * size[1] * 0.01)
            np.save(os.path.join(args.ins_seg_out_dir, img_name + ".npy"), detected)
            _common.progress(process_id, n_gpus, it, len(databin))


def run(args):
    model = getattr(importlib.import_module(args.irn_network), "EdgeDisplacement")()
    model.load_state_dict(torch.load(args.irn_weights_name), strict=False)
    model.eval()
    n_gpus = torch.cuda.device_count()
    dataset = make_dataset(args, args.infer_list, (1.0,))
    dataset = torchutils.split_dataset(dataset, n_gpus)
    print("[ ", end="")
    _common.spawn(_work, n_gpus, (model, dataset, args))
    print("]")
