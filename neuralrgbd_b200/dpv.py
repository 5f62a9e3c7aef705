"""Thin host wrappers for the per-pixel DPV reductions (C ABI: nrgbd_dpv_normalize)."""
import ctypes

import numpy as np
import torch

from . import _lib
from ._devcache import planes_tensor
from ._lib import ptr, check

_F = ctypes.c_float


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def log_softmax_planes(x, sign=1.0, add=None, d_candi=None):
    """x [1,D,h,w] -> log_softmax(sign*(x+add), dim=1) [1,D,h,w]
    (models/basic.py:299-300 with sign=-1; models/KVNET.py:172-173 with add=BV_predict).
    With d_candi also returns (expected depth [1,h,w], confidence [1,h,w])."""
    L = _lib.lib()
    with torch.cuda.device(x.device):
        a = x.float().contiguous()
        b = add.float().contiguous() if add is not None else None
        _, D, H, W = a.shape
        out = torch.empty_like(a)
        dpl = depth = conf = None
        if d_candi is not None:
            dpl = planes_tensor(d_candi, a.device)
            depth = torch.empty((1, H, W), device=a.device, dtype=torch.float32)
            conf = torch.empty((1, H, W), device=a.device, dtype=torch.float32)
        check(L.nrgbd_dpv_normalize(ptr(a), H * W, 1, ptr(b), H * W, 1, _F(sign), H * W, D, ptr(out), H * W, 1, ptr(dpl),
                                    ptr(depth), ptr(conf), _stream()))
    return out if d_candi is None else (out, depth, conf)
