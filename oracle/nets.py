"""Oracle (test infrastructure only): torch-CPU fp32 functional restatement of the reference
networks on the hot path, driven directly by a reference-format ``state_dict``.

 * ResNet-50 trunk, strides (2,2,2,1)          net/resnet50.py:17-91
 * CAM head                                     net/resnet50_cam.py:55-70
 * IRNet edge / displacement heads + MeanShift  net/resnet50_irn.py:23-133
 * EdgeDisplacement wrapper                     net/resnet50_irn.py:216-234

Pinned by tests/golden/cam_*.npz and irn_*.npz (unmodified reference outputs).
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
LAYERS = (3, 4, 6, 3)
PLANES = (64, 128, 256, 512)
STRIDES = (1, 2, 2, 1)      # layer1..4 with the reference's strides=(2,2,2,1): conv1 carries strides[0]


def _bn(x, sd, p):
    # FixedBatchNorm: always inference statistics (net/resnet50.py:11-14)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], training=False, eps=BN_EPS)


def _bottleneck(x, sd, p, stride):
    # net/resnet50.py:34-54; the stride sits on the 3x3 conv2 and on the 1x1 downsample
    y = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"]), sd, p + ".bn1"))
    y = F.relu(_bn(F.conv2d(y, sd[p + ".conv2.weight"], stride=stride, padding=1), sd, p + ".bn2"))
    y = _bn(F.conv2d(y, sd[p + ".conv3.weight"]), sd, p + ".bn3")
    if (p + ".downsample.0.weight") in sd:
        x = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride), sd, p + ".downsample.1")
    return F.relu(y + x)


def trunk(x, sd, prefix="resnet50."):
    """Returns (x1 after maxpool, x2 layer1, x3 layer2, x4 layer3, x5 layer4)."""
    y = F.conv2d(x, sd[prefix + "conv1.weight"], stride=2, padding=3)
    y = F.relu(_bn(y, sd, prefix + "bn1"))
    feats = [F.max_pool2d(y, 3, 2, 1)]
    y = feats[0]
    for li, (n, s) in enumerate(zip(LAYERS, STRIDES), start=1):
        for b in range(n):
            y = _bottleneck(y, sd, "%slayer%d.%d" % (prefix, li, b), s if b == 0 else 1)
        feats.append(y)
    return feats


def cam_forward(x, sd):
    """x [2,3,h,w] (image, flipped image) -> [20, ceil(h/16), ceil(w/16)]
    (net/resnet50_cam.py:55-70)."""
    f = trunk(x, sd)[-1]
    y = F.relu(F.conv2d(f, sd["classifier.weight"]))
    return y[0] + y[1].flip(-1)


def _head(x, sd, p, groups, up=None):
    # conv1x1 -> GroupNorm -> [bilinear upsample] -> ReLU   (net/resnet50_irn.py:23-93)
    y = F.conv2d(x, sd[p + ".0.weight"])
    y = F.group_norm(y, groups, sd[p + ".1.weight"], sd[p + ".1.bias"], eps=1e-5)
    if up:
        y = F.interpolate(y, scale_factor=up, mode="bilinear", align_corners=False)
    return F.relu(y)


def irn_forward(x, sd, eval_mode=True):
    """Net.forward (net/resnet50_irn.py:110-133): x [B,3,H,W] -> edge_out [B,1,H/4,W/4],
    dp_out [B,2,H/4,W/4]."""
    x1, x2, x3, x4, x5 = trunk(x, sd)
    e1 = _head(x1, sd, "fc_edge1", 4)
    e2 = _head(x2, sd, "fc_edge2", 4)
    hh, ww = e2.shape[2], e2.shape[3]
    e3 = _head(x3, sd, "fc_edge3", 4, 2)[..., :hh, :ww]
    e4 = _head(x4, sd, "fc_edge4", 4, 4)[..., :hh, :ww]
    e5 = _head(x5, sd, "fc_edge5", 4, 4)[..., :hh, :ww]
    edge = F.conv2d(torch.cat([e1, e2, e3, e4, e5], 1), sd["fc_edge6.weight"], sd["fc_edge6.bias"])

    d1 = _head(x1, sd, "fc_dp1", 8)
    d2 = _head(x2, sd, "fc_dp2", 16)
    d3 = _head(x3, sd, "fc_dp3", 16)
    h3, w3 = d3.shape[2], d3.shape[3]
    d4 = _head(x4, sd, "fc_dp4", 16, 2)[..., :h3, :w3]
    d5 = _head(x5, sd, "fc_dp5", 16, 2)[..., :h3, :w3]
    up3 = _head(torch.cat([d3, d4, d5], 1), sd, "fc_dp6", 16, 2)[..., :d2.shape[2], :d2.shape[3]]
    y = _head(torch.cat([d1, d2, up3], 1), sd, "fc_dp7", 16)
    dp = F.conv2d(y, sd["fc_dp7.3.weight"])
    if eval_mode:   # MeanShift (net/resnet50_irn.py:105-108)
        dp = dp - sd["mean_shift.running_mean"].view(1, 2, 1, 1)
    return edge, dp


def edge_displacement(x, sd, crop_size=512, stride=4):
    """EdgeDisplacement.forward (net/resnet50_irn.py:223-234): x [2,3,H,W] ->
    edge [1,h,w] in (0,1), dp [2,h,w]."""
    H, W = x.shape[2], x.shape[3]
    fh, fw = (H - 1) // stride + 1, (W - 1) // stride + 1
    xp = F.pad(x, [0, crop_size - W, 0, crop_size - H])
    e, d = irn_forward(xp, sd)
    e = e[..., :fh, :fw]
    d = d[..., :fh, :fw]
    return torch.sigmoid(e[0] / 2 + e[1].flip(-1) / 2), d[0]
