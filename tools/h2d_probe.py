"""Does a pinned H2D copy on a side stream overlap with kernels launched through libirn_b200 on torch's current stream?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irn_b200.ops import Conv2d
dev = torch.device("cuda:0")
x_host = torch.empty((256, 1024, 1024), dtype=torch.float32).pin_memory()   # 1 GiB
buf = torch.empty_like(x_host, device=dev)
w = (torch.randn(512, 512, 3, 3) * 0.02).numpy()
conv = Conv2d(w, None, 1, 1)
a = torch.randn(16, 64, 64, 512, device=dev)
side = torch.cuda.Stream()
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
def copy_only():
    with torch.cuda.stream(side): buf.copy_(x_host, non_blocking=True)
def conv_only():
    for _ in range(40): conv(a, None, relu=True, mode=1)
def both():
    copy_only(); conv_only()
print("copy 1 GiB: %.1f ms (%.1f GB/s)" % (t(copy_only), 1.0737 / t(copy_only) * 1e3))
print("40 convs  : %.1f ms" % t(conv_only))
print("both      : %.1f ms" % t(both))
