"""Time individual convolution layers through the C ABI (development aid).  IRN_B200_LIB selects the build."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from irn_b200.ops import Conv2d

LAYERS = [  # name, cin, cout, k, stride, H(in), B, residual
    ("L1 c1 64->64", 64, 64, 1, 1, 256, 16, False),
    ("L1 c2 3x3 64", 64, 64, 3, 1, 256, 16, False),
    ("L1 c3 64->256 +res", 64, 256, 1, 1, 256, 16, True),
    ("L1 c1 256->64", 256, 64, 1, 1, 256, 16, False),
    ("L2 c3 128->512 +res", 128, 512, 1, 1, 128, 16, True),
    ("L3 c2 3x3 256", 256, 256, 3, 1, 64, 16, False),
    ("L3 c3 256->1024 +res", 256, 1024, 1, 1, 64, 16, True),
    ("L4 c2 3x3 512", 512, 512, 3, 1, 64, 16, False),
    ("L4 c3 512->2048 +res", 512, 2048, 1, 1, 64, 16, True),
    ("L4 c1 2048->512", 2048, 512, 1, 1, 64, 16, False),
    ("L2 c1 512->128", 512, 128, 1, 1, 128, 16, False),
    ("L2 ds 256->512 s2", 256, 512, 1, 2, 256, 16, False),
    ("L3 c1 512->256", 512, 256, 1, 1, 128, 16, False),
    ("L3 ds 512->1024 s2", 512, 1024, 1, 2, 128, 16, False),
]
dev = torch.device("cuda:0")
MODE = int(os.environ.get("CONV_MODE", "1"))
only = os.environ.get("CONV_ONLY")
if only:
    LAYERS = [l for l in LAYERS if any(o in l[0] for o in only.split(','))]
for name, cin, cout, k, s, H, B, res in LAYERS:
    w = (torch.randn(cout, cin, k, k) * (2.0 / (cin * k * k)) ** 0.5).numpy()
    bn = [np.ones(cout, np.float32), np.zeros(cout, np.float32), np.zeros(cout, np.float32), np.ones(cout, np.float32)]
    conv = Conv2d(w, bn, s, k // 2)
    x = torch.randn(B, H, H, cin, device=dev)
    Ho = (H + 2 * (k // 2) - k) // s + 1
    r = torch.randn(B, Ho, Ho, cout, device=dev) if res else None
    for _ in range(3):
        conv(x, r, relu=True, mode=MODE)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        conv(x, r, relu=True, mode=MODE)
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / n
    fl = 2.0 * B * Ho * Ho * cin * cout * k * k
    byt = 4.0 * (B * H * H * cin + B * Ho * Ho * cout * (2 if res else 1))
    print(json.dumps({"layer": name, "us": round(us, 1), "TFLOPs_alg": round(fl / us / 1e6, 1), "GBps": round(byt / us / 1e3, 1)}), flush=True)
