"""Print (and write to gpurun_out/parity_report.json) the measured deviations of the CUDA path from
the committed reference outputs for the full-pipeline cases: per output max-abs in log space and in
probability, expected-depth error in mm. Development / documentation aid (DESIGN.md quotes it)."""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import planesweep_oracle as O            # noqa: E402
from tests import cases                              # noqa: E402
from neuralrgbd_b200.models.KVNET import KVNET       # noqa: E402
from neuralrgbd_b200.test_utils.test_KVNet import test as kv_test   # noqa: E402

dev = 'cuda:0'
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)     # noqa: E731
gold = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'reference_outputs.npz'))


def main(mode=None):
    rep = {}
    for name in cases.KVNET_CASES:
        c = cases.kvnet_case(name)
        cam = cases.cam_for(O.make_cam_intrinsics, c['W'] // 4, c['H'] // 4)
        camt = dict(cam); camt['unit_ray_array_2D'] = torch.from_numpy(cam['unit_ray_array_2D']); camt['intrinsic_M_cuda'] = torch.from_numpy(cam['intrinsic_M_cuda'])
        with contextlib.redirect_stdout(io.StringIO()):
            m = KVNET(64, camt, c['d'], c['sigma'], 64, None, t_win_r=2)
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in c['sd'].items()})
        m = m.to(dev)
        if mode is not None:
            m.conv_math = mode
        bv_pred = None
        n_steps = len(c['frames']) - 4
        for step in range(n_steps):
            ref_f, src_f, poses = cases.window(c, 2 + step)
            key = 'kvnet/%s/step%d' % (name, step)
            with torch.no_grad():
                full = m(T(ref_f), T(src_f), T(poses), torch.zeros(1), cam_intrinsics=[camt], BV_predict=bv_pred)
            rec = {}
            for nm, a in zip(['dmap_cur_refined', 'dmap_refined', 'BV_cur', 'DPV'], full):
                a = cases.subsample(a.cpu().numpy()); r = gold['%s/%s' % (key, nm)]
                rec[nm + '_log'] = float(np.abs(a - r).max()); rec[nm + '_prob'] = float(np.abs(np.exp(a) - np.exp(r)).max())
                rec[nm + '_argmax_flips'] = int((a.argmax(1) != r.argmax(1)).sum())
            dep = O.depth_val_regression(cases.subsample(full[3].cpu().numpy()), c['d'])
            rec['depth_mm'] = float(1000 * np.abs(dep - O.depth_val_regression(gold[key + '/DPV'], c['d'])).max())
            nxt = m.propagate(full[3], T(gold[key + '/rel_inv']))
            rec['BV_predict_next_prob'] = float(np.abs(np.exp(cases.subsample(nxt.cpu().numpy())) - np.exp(gold[key + '/BV_predict_next'])).max())
            rep[key] = rec
            print(key, json.dumps(rec))
            if step < n_steps - 1:
                bv_pred = T(gold[key + '/BV_predict_next_full'])
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(rep, open('gpurun_out/parity_report%s.json' % ('' if mode is None else '_' + mode), 'w'), indent=1)


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else None)
