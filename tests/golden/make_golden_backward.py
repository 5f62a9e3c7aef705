"""Golden gradients of the plane sweep: torch autograd through the UNMODIFIED reference est_swp_volume_v4 (CPU).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_backward.py
This is what train_utils/train_KVNet.py:149-153 differentiates. Inputs: tests/cases.sweep_case (seeded) and a seeded
upstream gradient cases.sweep_grad(name). Stored: d loss / d feat_img_ref and d loss / d feat_img_src
(tests/golden/sweep_backward.npz), and the oracle's deviation from them (tests/golden/PINNING_backward.json).
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
    out, pin = {}, {'torch': torch.__version__, 'numpy': np.__version__, 'cases': {}}
    for name in cases.SWEEP_BACKWARD_CASES:
        c = cases.sweep_case(name)
        cam = cases.cam_for(O.make_cam_intrinsics, c['w'], c['h'])
        camt = dict(cam); camt['unit_ray_array_2D'] = T(cam['unit_ray_array_2D']); camt['intrinsic_M_cuda'] = T(cam['intrinsic_M_cuda'])
        ref = T(c['ref']).clone().requires_grad_(True)
        src = T(c['src']).clone().requires_grad_(True)
        cost = wh.est_swp_volume_v4(ref, src, c['d'], T(c['R']), T(c['t']), camt, c['sigma'], feat_dist=c['feat_dist'])
        g = cases.sweep_grad(name, cost.shape)
        cost.backward(T(g))
        g_ref, g_src = ref.grad.numpy(), src.grad.numpy()
        out[name + '/g_ref'] = g_ref
        out[name + '/g_src'] = g_src
        o_ref, o_src = O.est_swp_volume_v4_backward(g, c['ref'], c['src'], c['d'], c['R'], c['t'], cam, c['sigma'], c['feat_dist'])
        pin['cases'][name] = {
            'g_ref_maxabs': float(np.abs(o_ref - g_ref).max()), 'g_ref_max': float(np.abs(g_ref).max()),
            'g_src_maxabs': float(np.abs(o_src - g_src).max()), 'g_src_max': float(np.abs(g_src).max()),
        }
        print(name, pin['cases'][name])
    np.savez_compressed(os.path.join(HERE, 'sweep_backward.npz'), **out)
    with open(os.path.join(HERE, 'PINNING_backward.json'), 'w') as f:
        json.dump(pin, f, indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
