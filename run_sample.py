#!/usr/bin/env python
"""Drop-in driver for the pseudo-label generation steps of the reference's ``run_sample.py``.

Same flag names and defaults (run_sample.py:11-72) and the same output directories, so the reference's evaluation steps
(step/eval_cam.py, step/eval_sem_seg.py, step/eval_ins_seg.py) consume ``result/cam/*.npy``, ``result/sem_seg/*.png`` and
``result/ins_seg/*.npy`` unchanged.  Only the three hot-path steps are implemented here (make_cam, make_ins_seg_labels,
make_sem_seg_labels); the training / CRF / evaluation passes are the reference's own and are skipped with a note.

Differences: --cam_network / --irn_network default to the B200 modules; flags the reference declares without a type
(--beta, --exp_times, --*_bg_thres) are parsed as numbers; --synthetic N runs on N seeded synthetic images instead of VOC.
"""
import argparse
import os

from irn_b200.misc import pyutils


def str2bool(v):
    return str(v).lower() in ("1", "true", "yes", "y")


if __name__ == "__main__":
    p = argparse.ArgumentParser()
    # Environment
    p.add_argument("--num_workers", default=os.cpu_count() // 2, type=int)
    p.add_argument("--voc12_root", default="", type=str, help="VOC 2012 devkit (must contain ./JPEGImages); not needed with --synthetic")
    p.add_argument("--synthetic", default=0, type=int, help="run on this many seeded synthetic 512x512 images")
    # Dataset
    p.add_argument("--train_list", default="voc12/train_aug.txt", type=str)
    p.add_argument("--val_list", default="voc12/val.txt", type=str)
    p.add_argument("--infer_list", default="voc12/train.txt", type=str)
    p.add_argument("--chainer_eval_set", default="train", type=str)
    # Class Activation Map
    p.add_argument("--cam_network", default="irn_b200.cam", type=str)
    p.add_argument("--cam_crop_size", default=512, type=int)
    p.add_argument("--cam_batch_size", default=16, type=int)
    p.add_argument("--cam_num_epoches", default=5, type=int)
    p.add_argument("--cam_learning_rate", default=0.1, type=float)
    p.add_argument("--cam_weight_decay", default=1e-4, type=float)
    p.add_argument("--cam_eval_thres", default=0.15, type=float)
    p.add_argument("--cam_scales", default=(1.0, 0.5, 1.5, 2.0), type=lambda s: tuple(float(v) for v in s.split(",")),
                   help="Multi-scale inferences (comma separated)")
    # Mining Inter-pixel Relations
    p.add_argument("--conf_fg_thres", default=0.30, type=float)
    p.add_argument("--conf_bg_thres", default=0.05, type=float)
    # Inter-pixel Relation Network (IRNet)
    p.add_argument("--irn_network", default="irn_b200.irn", type=str)
    p.add_argument("--irn_crop_size", default=512, type=int)
    p.add_argument("--irn_batch_size", default=32, type=int)
    p.add_argument("--irn_num_epoches", default=3, type=int)
    p.add_argument("--irn_learning_rate", default=0.1, type=float)
    p.add_argument("--irn_weight_decay", default=1e-4, type=float)
    # Random Walk Params
    p.add_argument("--beta", default=10, type=float)
    p.add_argument("--exp_times", default=8, type=int,
                   help="The random walk is performed 2^{exp_times} times.")
    p.add_argument("--ins_seg_bg_thres", default=0.25, type=float)
    p.add_argument("--sem_seg_bg_thres", default=0.25, type=float)
    # Output Path
    p.add_argument("--log_name", default="sample_train_eval", type=str)
    p.add_argument("--cam_weights_name", default="sess/res50_cam.pth", type=str)
    p.add_argument("--irn_weights_name", default="sess/res50_irn.pth", type=str)
    p.add_argument("--cam_out_dir", default="result/cam", type=str)
    p.add_argument("--ir_label_out_dir", default="result/ir_label", type=str)
    p.add_argument("--sem_seg_out_dir", default="result/sem_seg", type=str)
    p.add_argument("--ins_seg_out_dir", default="result/ins_seg", type=str)
    # Step
    for name in ("train_cam", "make_cam", "eval_cam", "cam_to_ir_label", "train_irn", "make_ins_seg", "eval_ins_seg", "make_sem_seg",
                 "eval_sem_seg"):
        p.add_argument("--%s_pass" % name, default=True, type=str2bool)
    args = p.parse_args()

    os.makedirs("sess", exist_ok=True)
    for d in (args.cam_out_dir, args.ir_label_out_dir, args.sem_seg_out_dir, args.ins_seg_out_dir):
        os.makedirs(d, exist_ok=True)
    pyutils.Logger(args.log_name + ".log")
    print(vars(args))

    if args.synthetic:
        # no VOC data / trained checkpoints on this box: seeded synthetic checkpoints in the reference's format
        import torch
        from irn_b200 import synth
        if not os.path.exists(args.cam_weights_name + ".pth"):
            torch.save(synth.cam_state_dict(), args.cam_weights_name + ".pth")
        if not os.path.exists(args.irn_weights_name):
            torch.save(synth.irn_state_dict(), args.irn_weights_name)

    for name in ("train_cam", "eval_cam", "cam_to_ir_label", "train_irn", "eval_ins_seg", "eval_sem_seg"):
        if getattr(args, name + "_pass"):
            print("[irn_b200] step.%s is outside the B200 hot path: run the reference's own step for it" % name)

    if args.make_cam_pass:
        from irn_b200.step import make_cam
        timer = pyutils.Timer("step.make_cam:")
        make_cam.run(args)
    if args.make_ins_seg_pass:
        from irn_b200.step import make_ins_seg_labels
        timer = pyutils.Timer("step.make_ins_seg_labels:")
        make_ins_seg_labels.run(args)
    if args.make_sem_seg_pass:
        from irn_b200.step import make_sem_seg_labels
        timer = pyutils.Timer("step.make_sem_seg_labels:")
        make_sem_seg_labels.run(args)
