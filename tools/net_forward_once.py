"""One 4-scale CAM forward + one IRNet forward for a few synthetic 512x512 images, exactly as the pipeline sub-batches them
(development aid: run under `ncu --metrics ...` to get per-kernel time / tensor-pipe / DRAM bytes of every layer).
    python tools/net_forward_once.py [n_images=8] [conv_mode=-1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from irn_b200 import _lib, synth
from irn_b200.cam import CAM
from irn_b200.irn import EdgeDisplacement
from irn_b200.pipeline import PseudoLabelPipeline

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mode = int(sys.argv[2]) if len(sys.argv) > 2 else -1
dev = torch.device("cuda:0")
cam, irn = CAM(), EdgeDisplacement()
cam.load_state_dict(synth.cam_state_dict(), strict=True)
irn.load_state_dict(synth.irn_state_dict(), strict=False)
cam.cuda(dev), irn.cuda(dev)
if mode >= 0:
    for m in (cam, irn):
        _lib.check(_lib.lib().irn_net_set_conv_mode(m._get_plan(dev).handle, mode))
pipe = PseudoLabelPipeline(cam, irn, dev)
x = torch.from_numpy(np.stack([synth.image(i) for i in range(n)])).to(dev)
labels = np.stack([synth.label(i) for i in range(n)])
with torch.no_grad():
    for rep in range(int(os.environ.get("REPS", "2"))):     # first pass warms plans / tensor maps; profile the second (ncu -s)
        xs = pipe.pyramids(x)
        pipe.cam_stage(xs, labels, (512, 512), want_highres=False)
        pipe.irn_stage(xs[0])
        torch.cuda.synchronize()
        print("pass", rep, "launches so far", _lib.lib().irn_total_launch_count(), flush=True)
