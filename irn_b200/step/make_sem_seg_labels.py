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


def _work(process_id, model, dataset, args):
    _common.work_loop(process_id, model, dataset, args, sem_seg_one_image)


def run(args):
    _common.run_step(args, _work, args.irn_network, "EdgeDisplacement", args.irn_weights_name, False, args.infer_list, (1.0,), opening="[")
