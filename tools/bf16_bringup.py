"""Bring-up aid for conv_bf16_kernel: a few layers in mode 2 against fp64 torch (max rel err, signed bias), with the env
switches IRN_BF_SWAP / IRN_BF_NACC / IRN_BF_WIDE_MINK applied by the caller."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from irn_b200.ops import Conv2d

dev = torch.device("cuda:0")
CASES = [(64, 64, 1, 1, 2, 32, 32), (64, 128, 3, 1, 2, 32, 48), (256, 256, 3, 1, 2, 32, 32), (512, 512, 3, 1, 2, 20, 24), (2048, 512, 1, 1, 2, 16, 16),
         (128, 128, 3, 2, 2, 64, 64)]
for cin, cout, k, stride, B, H, W in CASES:
    g = torch.Generator().manual_seed(cin + cout + k)
    w = torch.randn((cout, cin, k, k), generator=g) * (2.0 / (cin * k * k)) ** 0.5
    x = torch.randn((B, cin, H, W), generator=g).to(dev)
    conv = Conv2d(w.numpy(), None, stride, k // 2)
    ref = F.conv2d(x.double(), w.to(dev).double(), stride=stride, padding=k // 2).float().permute(0, 2, 3, 1).contiguous()
    xn = x.permute(0, 2, 3, 1).contiguous()
    out = {}
    for mode in (1, 2):
        try:
            y = conv(xn, None, relu=False, mode=mode)
            torch.cuda.synchronize()
            out["mode%d_err" % mode] = ((y - ref).abs().max() / ref.abs().max()).item()
            out["mode%d_bias" % mode] = ((y - ref).double().sum() / ref.double().abs().sum()).item()
        except Exception as e:
            out["mode%d_err" % mode] = repr(e)[:200]
    print(json.dumps({"case": [cin, cout, k, stride, B, H, W], "env": {k: v for k, v in os.environ.items() if k.startswith("IRN_BF")}, **out}), flush=True)
