"""Golden vectors for the LBA depth-map back-warp (oracle pinned ahead of the device kernels).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_lba.py
Runs the UNMODIFIED reference back_warp_th_Rt / back_warp_th_Rt_msrc (warping/homography.py:479-574) on CPU and
differentiates the single-view warp w.r.t. R, t and the image with torch autograd - what ICP/opt_pose_numerical.py does.
Writes tests/golden/lba_outputs.npz and the oracle's deviations to PINNING_lba.json.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/code')
warnings.filterwarnings('ignore')

torch.Tensor.cuda = lambda s, *a, **k: s
torch.nn.Module.cuda = lambda s, *a, **k: s
torch.cuda.current_device = lambda: 0
torch.Tensor.get_device = lambda s: 0

import warping.homography as wh                      # noqa: E402  (reference)

from oracle import planesweep_oracle as O            # noqa: E402
from tests import cases                              # noqa: E402

T = torch.from_numpy


def main():
    out, pin = {}, {'torch': torch.__version__, 'cases': {}}
    for name in cases.LBA_CASES:
        c = cases.lba_case(name)
        cam = cases.cam_for(O.make_cam_intrinsics, c['w'], c['h'])
        camt = dict(cam); camt['unit_ray_array_2D'] = T(cam['unit_ray_array_2D']); camt['intrinsic_M_cuda'] = T(cam['intrinsic_M_cuda'])
        multi = wh.back_warp_th_Rt_msrc(T(c['imgs']), T(c['dmap']), T(c['Rs']), T(c['ts']), camt).numpy()
        R = T(c['Rs'][0]).clone().requires_grad_(True); t = T(c['ts'][0]).clone().requires_grad_(True)
        img = T(c['imgs'][:1]).clone().requires_grad_(True)
        single = wh.back_warp_th_Rt(img, T(c['dmap']), R, t, camt)
        g = np.random.RandomState(7).standard_normal(tuple(single.shape)).astype(np.float32)
        single.backward(T(g))
        out[name + '/warp_msrc'] = multi
        out[name + '/grad_out'] = g
        out[name + '/g_R'] = R.grad.numpy(); out[name + '/g_t'] = t.grad.numpy(); out[name + '/g_img'] = img.grad.numpy()
        o_multi = O.back_warp_th_Rt_msrc(c['imgs'], c['dmap'], c['Rs'], c['ts'], cam)
        gR, gt, gi = O.back_warp_th_Rt_backward(g, c['imgs'][:1], c['dmap'], c['Rs'][0], c['ts'][0], cam)
        pin['cases'][name] = {'warp_maxabs': float(np.abs(o_multi - multi).max()),
                              'single_equals_msrc_view0': float(np.abs(single.detach().numpy() - multi[:1]).max()),
                              'g_R_rel': float(np.abs(gR - out[name + '/g_R']).max() / np.abs(out[name + '/g_R']).max()),
                              'g_t_rel': float(np.abs(gt - out[name + '/g_t']).max() / np.abs(out[name + '/g_t']).max()),
                              'g_img_maxabs': float(np.abs(gi - out[name + '/g_img']).max())}
    np.savez_compressed(os.path.join(HERE, 'lba_outputs.npz'), **out)
    with open(os.path.join(HERE, 'PINNING_lba.json'), 'w') as f:
        json.dump(pin, f, indent=1, sort_keys=True)
    print(json.dumps(pin, indent=1))


if __name__ == '__main__':
    main()
