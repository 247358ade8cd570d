"""Drop-in for the reference's ``step/make_cam.py``: same ``run(args)`` namespace and the same ``.npy`` output --
a pickled ``{"keys": LongTensor[K], "cam": FloatTensor[K,h/4,w/4], "high_res": ndarray[K,H,W]}`` per image
(step/make_cam.py:55-56) -- with the arithmetic in libirn_b200."""
import os

import numpy as np

from .. import cam_ops
from . import _common


def cam_one_image(model, pack, args):
    """step/make_cam.py:28-56: CAM forward per scale, merge + normalise, save."""
    per_scale = [model(x[0].cuda(non_blocking=True)) for x in pack["img"]]
    keys, strided, highres = cam_ops.merge_cams(per_scale, pack["size"], pack["label"][0])
    np.save(os.path.join(args.cam_out_dir, pack["name"][0] + ".npy"),
            {"keys": keys, "cam": strided.cpu(), "high_res": highres.cpu().numpy()})


def _work(process_id, model, dataset, args):
    _common.work_loop(process_id, model, dataset, args, cam_one_image)


def run(args):
    # the reference appends '.pth' to --cam_weights_name here (step/make_cam.py:64) but not for the IRN weights
    _common.run_step(args, _work, args.cam_network, "CAM", args.cam_weights_name + ".pth", True, args.train_list, args.cam_scales)
