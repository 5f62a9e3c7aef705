"""How fast do two fp32 implementations of the K-Net recursion drift apart? (the floor for the free-running GPU gates)

    python tests/golden/make_golden_drift.py [case]        (CPU, ~25 min; no reference needed)

Runs the independent numpy oracle (oracle/kvnet_oracle.py) FREE-RUNNING over the whole stream of a BIG case - feeding
its own propagated prior, exactly like the engine in tests/test_gpu_configs.py - and compares every step with the
committed outputs of the live reference (tests/golden/configs_<case>.npz), which was also free-running. Both are fp32;
they differ only in summation order. The per-step deviations (probability of the filtered DPV / refined DPV / next
prior, expected depth in mm) are written to tests/golden/PINNING_drift_<case>.json: the measured growth of rounding
noise through the recursion, which no fp32 implementation can stay below and against which the GPU gates are set.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import planesweep_oracle as O            # noqa: E402
from oracle import kvnet_oracle as N                 # noqa: E402
from tests import cases                              # noqa: E402


def dev(a, b):
    return float(np.nanmax(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'c23_640x480_d64_v4_stream30'
    n_max = int(sys.argv[2]) if len(sys.argv) > 2 else 10 ** 9
    gold = np.load(os.path.join(HERE, 'configs_%s.npz' % name))
    c = cases.big_case(name)
    cam = cases.big_cam(O.make_cam_intrinsics, c)
    r = c['t_win_r']
    rows = []
    prior = None
    out_path = os.path.join(HERE, 'PINNING_drift_%s.json' % name)
    for step in range(min(c['n_steps'], n_max)):
        t0 = time.time()
        ref_f, src_f, poses = cases.window(c, r + step)
        o = N.kvnet_forward(c['sd'], ref_f, src_f, poses, cam, c['d'], c['sigma'], BV_predict=prior)
        import torch
        rel = torch.from_numpy(poses)[0, r].inverse().numpy()          # the same fp32 inverse the drivers use
        prior = O.propagate_dpv(o[3], rel, cam, c['d'])
        key = 'cfg/%s/step%d' % (name, step)
        thin = step >= cases.BIG_FULL_STEPS
        row = {'step': step}
        for nm, a in zip(['dmap_cur_refined', 'dmap_refined', 'BV_cur', 'DPV'], o):
            k = '%s/%s' % (key, nm)
            if k in gold.files:
                sub = cases.subsample_to(a, 30000) if thin else cases.subsample(a)
                row[nm + '_prob'] = dev(np.exp(sub), np.exp(gold[k]))
        row['depth_mm'] = 1000 * dev(cases.subsample_to(O.depth_val_regression(o[3], c['d']), 5000), gold[key + '/depth'])
        row['prior_next_prob'] = dev(np.exp(cases.subsample_to(prior, 30000)), np.exp(gold[key + '/BV_predict_next']))
        row['seconds'] = time.time() - t0
        rows.append(row)
        print(json.dumps(row), flush=True)
        with open(out_path, 'w') as f:
            json.dump({'case': name, 'what': 'numpy oracle free-running vs live reference free-running (both fp32)', 'steps': rows}, f, indent=1)


if __name__ == '__main__':
    main()
