"""Camera intrinsics dict in the reference's format.

Mirrors mdataloader/scanNet.py:239-270 (read_IntM_from_txt with out_size) and
warping/View.py:16-62 (normalised_pixel_to_ray_array, normalize_z=True): the ray table is
built vectorised (the reference runs a Python double loop over H x W) with the same float64
arithmetic, so the values are bit-identical.
"""
import math

import numpy as np
import torch


def unit_ray_array(width, height, hfov, vfov):
    """View.py:32-62: (tan(hfov/2)(2(x+.5)/W-1), tan(vfov/2)(2(y+.5)/H-1), 1), float64 [H,W,3]."""
    th = math.tan(math.radians(hfov / 2.0))
    tv = math.tan(math.radians(vfov / 2.0))
    xs = th * ((2.0 * ((np.arange(width, dtype=np.float64) + 0.5) / width)) - 1.0)
    ys = tv * ((2.0 * ((np.arange(height, dtype=np.float64) + 0.5) / height)) - 1.0)
    out = np.empty((height, width, 3), dtype=np.float64)
    out[:, :, 0] = xs[None, :]
    out[:, :, 1] = ys[:, None]
    out[:, :, 2] = 1.0
    return out


def make_cam_intrinsics(fx, fy, cx, cy, out_size, full_width=None):
    """scanNet.py:239-270. fx, fy, cx, cy: calibration of the full-size image;
    out_size = [width, height] of the (quarter-resolution) maps the sweep runs on."""
    h_fov = math.degrees(math.atan(cx / fx) * 2)
    v_fov = math.degrees(math.atan(cy / fy) * 2)
    pw, ph = int(out_size[0]), int(out_size[1])
    K = np.zeros((3, 4))
    K[2, 2] = 1.
    K[0, 0] = (pw / 2.0) / math.tan(math.radians(h_fov / 2.0))
    K[0, 2] = pw / 2.0
    K[1, 1] = (ph / 2.0) / math.tan(math.radians(v_fov / 2.0))
    K[1, 2] = ph / 2.0
    width = full_width if full_width is not None else 2.0 * cx
    rays = unit_ray_array(pw, ph, h_fov, v_fov)
    rays2d = np.reshape(np.transpose(rays, axes=[2, 0, 1]), [3, -1])
    return {'hfov': h_fov, 'vfov': v_fov, 'unit_ray_array': rays,
            'unit_ray_array_2D': torch.from_numpy(rays2d.astype(np.float32)),
            'intrinsic_M_cuda': torch.from_numpy(K[:3, :3].astype(np.float32)),
            'focal_length': pw / width * float(np.mean([fx, fy])),
            'intrinsic_M': K}
