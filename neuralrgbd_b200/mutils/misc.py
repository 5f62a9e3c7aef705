"""Mirror of the hot-path helpers of the reference's `mutils.misc`
(/root/reference/code/mutils/misc.py:100-115, 509-517, 532-548)."""
import ctypes

import numpy as np
import torch

from .. import _lib
from .._devcache import planes_tensor
from .._lib import ptr, check


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def valid_dpv(dpv_in):
    """misc.py:100-115: a DPV whose first element is NaN marks an invalid batch slot."""
    if dpv_in is None:
        return False
    assert isinstance(dpv_in, torch.Tensor), 'input should a Tensor'
    if dpv_in.dim() not in (2, 3, 4, 5):
        raise Exception('wrong dimension for input dpv !')
    return not bool(torch.isnan(dpv_in.reshape(-1)[0]))


def split_frame_list(frame_list, t_win_r):
    """misc.py:509-517."""
    ref_frame = frame_list[t_win_r]
    src_frames = [frame_list[idx] for idx in range(len(frame_list)) if idx != t_win_r]
    return ref_frame, src_frames


def depth_val_regression(BV_measure, d_candi_cur, BV_log=True, return_conf=False):
    """misc.py:532-548: depth = sum_d exp(BV[0,d]) * d -> [1,h,w]. One kernel instead of D launches."""
    assert len(d_candi_cur) == BV_measure.shape[1], \
        'BV_measure should have the same # of slices as len(d_candi_cur) !'
    if not BV_measure.is_cuda:
        raise _lib.NrgbdError('neuralrgbd_b200 has no CPU path: expected a CUDA tensor')
    L = _lib.lib()
    with torch.cuda.device(BV_measure.device):
        bv = BV_measure[0].float().contiguous()
        D, H, W = bv.shape
        dpl = planes_tensor(d_candi_cur, bv.device)
        depth = torch.empty((1, H, W), device=bv.device, dtype=torch.float32)
        conf = torch.empty((1, H, W), device=bv.device, dtype=torch.float32) if return_conf else None
        check(L.nrgbd_depth_regression(ptr(bv), H * W, D, H * W, 1, ptr(dpl), 1 if BV_log else 0, ptr(depth),
                                       ptr(conf), _stream()))
    return (depth, conf) if return_conf else depth
