"""Device-side mirror of the reference's per-frame input stage.

Reference: mdataloader/scanNet.py:358-452 (`__getitem__`: PIL open -> resize(img_size, NEAREST) -> get_transform())
and mdataloader/m_preprocess.py:12-34 (ToTensor + Normalize with the ImageNet statistics). The reference converts
every frame on the host and re-reads each frame five times (once per window it appears in, test_KVNet.py:249-250).
Here the decoded uint8 frame is uploaded once, resized + normalised by one kernel (bit-exact with PIL + torchvision),
and `FrameWindow` keeps the sliding window resident, so a streamed frame costs one 0.9 MB upload instead of 18.4 MB.
"""
import ctypes
import functools

import numpy as np
import torch

from .. import _lib
from .._lib import check

__imagenet_stats = {'mean': [0.485, 0.456, 0.406], 'std': [0.229, 0.224, 0.225]}     # m_preprocess.py:12-13
IMAGENET_STATS = __imagenet_stats


@functools.lru_cache(maxsize=64)
def nearest_index(n_out, n_in):
    """Source index of every output sample of PIL's NEAREST resize (libImaging Geometry.c, ImagingScaleAffine:
    xo = 0.5 * scale, truncated, then xo += scale accumulated in double)."""
    a = n_in / n_out
    xo = a * 0.5
    out = np.empty(n_out, np.int32)
    for x in range(n_out):
        out[x] = min(int(xo), n_in - 1)
        xo += a
    return out


@functools.lru_cache(maxsize=64)
def _index_tensors(H, W, Hs, Ws, device):
    return (torch.from_numpy(nearest_index(H, Hs)).to(device), torch.from_numpy(nearest_index(W, Ws)).to(device))


def _as_u8_hwc(img):
    if isinstance(img, torch.Tensor):
        a = img
    else:
        a = torch.from_numpy(np.ascontiguousarray(np.asarray(img)))       # PIL.Image or ndarray
    assert a.dtype == torch.uint8 and a.dim() == 3 and a.shape[2] == 3, 'expected an H x W x 3 uint8 image'
    return a


def device_transform(img, img_size=None, device='cuda', stats=None):
    """PIL image / H x W x 3 uint8 array (host or device) -> normalised float tensor [1, 3, H, W] on `device`;
    img_size = (W, H) as in scanNet.py (`img.resize(self.img_size, PIL.Image.NEAREST)`), None = keep the size."""
    stats = stats or IMAGENET_STATS
    a = _as_u8_hwc(img)
    Hs, Ws = int(a.shape[0]), int(a.shape[1])
    W, H = (Ws, Hs) if img_size is None else (int(img_size[0]), int(img_size[1]))
    dev = torch.device(device)
    if dev.type != 'cuda' or not torch.cuda.is_available():
        raise _lib.NrgbdError('neuralrgbd_b200 has no CPU path: the input stage runs on a CUDA device')
    if not a.is_cuda:
        a = a.to(dev, non_blocking=True)          # pass a pinned tensor to make this upload asynchronous
    a = a.contiguous()
    ys, xs = _index_tensors(H, W, Hs, Ws, a.device)
    out = torch.empty((1, 3, H, W), device=a.device, dtype=torch.float32)
    mean = (ctypes.c_float * 3)(*stats['mean']); std = (ctypes.c_float * 3)(*stats['std'])
    with torch.cuda.device(a.device):
        check(_lib.lib().nrgbd_preprocess_rgb_u8(ctypes.c_void_p(a.data_ptr()), Hs, Ws, ctypes.c_void_p(ys.data_ptr()),
                                                 ctypes.c_void_p(xs.data_ptr()), H, W, mean, std, ctypes.c_void_p(out.data_ptr()),
                                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


def get_transform(img_size=None, device='cuda'):
    """m_preprocess.py:28-34 get_transform(): a callable image -> tensor (here: on the device, [1, 3, H, W])."""
    return functools.partial(device_transform, img_size=img_size, device=device)


class FrameWindow:
    """Device-resident sliding window over a video (test_KVNet.py:241-250 re-reads 2 * t_win_r + 1 frames per step).
    push() uploads and normalises ONE new frame; window() returns (ref_frame [1,3,H,W], src_frames [1,V,3,H,W]) and the
    frame dicts of the current window in the reference's order (src = all frames but the middle one)."""

    def __init__(self, t_win_r=2, img_size=None, device='cuda'):
        self.t_win_r, self.img_size, self.device = t_win_r, img_size, device
        self.n = 2 * t_win_r + 1
        self.frames = []            # [(tensor [1,3,H,W], extM)]

    def push(self, img, extM=None):
        self.frames.append((device_transform(img, self.img_size, self.device), extM))
        if len(self.frames) > self.n:
            self.frames.pop(0)
        return len(self.frames) == self.n

    def window(self):
        assert len(self.frames) == self.n, 'window not full yet'
        ref = self.frames[self.t_win_r][0]
        src = torch.stack([f[0][0] for i, f in enumerate(self.frames) if i != self.t_win_r], 0).unsqueeze(0)
        return ref, src

    def frame_dicts(self):
        """[{'img': ..., 'extM': ...}] in window order, the structure test_utils.test_KVNet.test consumes."""
        return [{'img': f[0], 'extM': f[1]} for f in self.frames]
