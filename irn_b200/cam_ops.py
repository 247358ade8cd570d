"""C4 host wrapper: multi-scale CAM merge through the C ABI (step/make_cam.py:38-52)."""
import ctypes

import numpy as np
import torch

from . import _lib


def merge_cams(outputs, size, label):
    """outputs: list of cuda fp32 [20,h_s,w_s] (one per scale); size=(H,W); label: fp32[20] multi-hot (any device).
    Returns (keys LongTensor[K] on cpu, strided_cam cuda [K,ceil(H/4),ceil(W/4)], highres_cam cuda [K,H,W])."""
    L = _lib.lib()
    _lib.require_cuda(*outputs)
    dev = outputs[0].device
    outs = [o.contiguous().float() for o in outputs]
    H, W = int(size[0]), int(size[1])
    keys = torch.nonzero(torch.as_tensor(label).cpu())[:, 0]          # step/make_cam.py:46
    K = int(keys.numel())
    h4, w4 = (H - 1) // 4 + 1, (W - 1) // 4 + 1
    strided = torch.empty((K, h4, w4), dtype=torch.float32, device=dev)
    highres = torch.empty((K, H, W), dtype=torch.float32, device=dev)
    if K == 0:
        return keys, strided, highres
    n = len(outs)
    ptrs = (ctypes.c_void_p * n)(*[o.data_ptr() for o in outs])
    hs = (ctypes.c_int * n)(*[int(o.shape[1]) for o in outs])
    ws = (ctypes.c_int * n)(*[int(o.shape[2]) for o in outs])
    keys_host = np.ascontiguousarray(keys.numpy().astype(np.int32))
    scratch = torch.empty(2 * K + 4, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = L.irn_cam_merge(ptrs, hs, ws, n, H, W, keys_host.ctypes.data, K, _lib.ptr(strided), _lib.ptr(highres), _lib.ptr(scratch),
                             _lib.stream_ptr())
    _lib.check(rc, "irn_cam_merge")
    return keys, strided, highres
