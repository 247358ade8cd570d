"""Flatten a reference-format state dict into the parameter blob libirn_b200's plan
constructors read (order documented in include/irn_b200.h, irn_cam_net_create)."""
import numpy as np

from ._params import TRUNK, EDGE_HEADS, DP_HEADS

_BN = ("weight", "bias", "running_mean", "running_var")


def _t(sd, key):
    return sd[key].detach().cpu().float().contiguous().numpy().reshape(-1)


def _trunk(sd, prefix="resnet50."):
    out = [_t(sd, prefix + "conv1.weight")] + [_t(sd, prefix + "bn1." + k) for k in _BN]
    for li, (planes, blocks, stride) in enumerate(TRUNK, start=1):
        for b in range(blocks):
            p = "%slayer%d.%d." % (prefix, li, b)
            for i in (1, 2, 3):
                out.append(_t(sd, p + "conv%d.weight" % i))
                out += [_t(sd, p + "bn%d.%s" % (i, k)) for k in _BN]
            if b == 0:
                out.append(_t(sd, p + "downsample.0.weight"))
                out += [_t(sd, p + "downsample.1." + k) for k in _BN]
    return out


def pack_cam(sd):
    return np.ascontiguousarray(np.concatenate(_trunk(sd) + [_t(sd, "classifier.weight")]), dtype=np.float32)


def pack_irn(sd):
    out = _trunk(sd)
    for name, *_ in EDGE_HEADS:
        out += [_t(sd, name + ".0.weight"), _t(sd, name + ".1.weight"), _t(sd, name + ".1.bias")]
    out += [_t(sd, "fc_edge6.weight"), _t(sd, "fc_edge6.bias")]
    for name, *_ in DP_HEADS:
        out += [_t(sd, name + ".0.weight"), _t(sd, name + ".1.weight"), _t(sd, name + ".1.bias")]
    out += [_t(sd, "fc_dp7.3.weight"), _t(sd, "mean_shift.running_mean")]
    return np.ascontiguousarray(np.concatenate(out), dtype=np.float32)
