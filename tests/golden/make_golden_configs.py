"""Golden fixtures for the BASELINE.json configurations at their real sizes, from the UNMODIFIED reference.

Run in the build container only (needs /root/reference; CPU torch, the 4-line .cuda() shim of SURVEY 8c):
    python tests/golden/make_golden_configs.py [case ...]

For every case of tests/cases.py:BIG_CASES the reference's own streaming step
(test_utils/test_KVNet.py:19-67, unmodified) is driven FREE-RUNNING over the whole frame list: step k receives
the prior the reference itself propagated at step k-1. A recording wrapper around the model keeps the four
outputs of KVNET.forward that test() discards. Stored per step (reference OUTPUTS only, inputs are re-generated
from seeds): strided samples + full-array statistics of the outputs, the expected depth of the filtered DPV, the
propagated prior (strided); the full prior after step 0 goes to configs_priors.npz so that one steady step can
also be tested re-seeded. The numpy oracle is run on step 0 and step 1 (fed the reference's prior) and its
deviation from the reference is written to PINNING_configs.json: that is the measured fp32-vs-fp32 noise floor
AT THESE SHAPES which the GPU gates are set against.
"""
import contextlib
import io
import json
import os
import sys
import time
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

import models.KVNET as m_kvnet                       # noqa: E402  (reference)
import test_utils.test_KVNet as ref_test             # noqa: E402  (reference)
import mutils.misc as m_misc                         # noqa: E402  (reference)

from oracle import planesweep_oracle as O            # noqa: E402
from oracle import kvnet_oracle as N                 # noqa: E402
from tests import cases                              # noqa: E402

T = torch.from_numpy
NAMES4 = ['dmap_cur_refined', 'dmap_refined', 'BV_cur', 'DPV']


def cam_torch(cam):
    c = dict(cam)
    c['unit_ray_array_2D'] = T(cam['unit_ray_array_2D'])
    c['intrinsic_M_cuda'] = T(cam['intrinsic_M_cuda'])
    return c


def dev(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.nanmax(np.abs(a - b))) if a.size else 0.0


class Recorder(torch.nn.Module):
    """Passes the call through to the reference model and keeps what it returned."""

    def __init__(self, model):
        super().__init__()
        self.model = model
        self.last = None

    def forward(self, **kw):
        self.last = self.model(**kw)
        return self.last


def run_case(name, oracle_steps=2):
    c = cases.big_case(name)
    cam = cases.big_cam(O.make_cam_intrinsics, c)
    camt = cam_torch(cam)
    r = c['t_win_r']
    with contextlib.redirect_stdout(io.StringIO()):
        model = m_kvnet.KVNET(feature_dim=64, cam_intrinsics=camt, d_candi=c['d'], sigma_soft_max=c['sigma'],
                              KVNet_feature_dim=64, d_upsample_ratio_KV_net=None, t_win_r=r, if_refined=True)
    model.load_state_dict({k: T(np.asarray(v)) for k, v in c['sd'].items()})
    rec_model = Recorder(torch.nn.DataParallel(model))      # DataParallel as test_KVNet.py:163 (no GPUs: falls through)
    out, priors, pin = {}, {}, {}
    bv_pred = None
    for step in range(c['n_steps']):
        t0 = time.time()
        ref_f, src_f, poses = cases.window(c, r + step)
        Ref_Dats = [{'img': T(ref_f)}]
        Src_Dats = [[{'img': T(src_f[0, v:v + 1])} for v in range(src_f.shape[1])]]
        kv_dpv, bv_next = ref_test.test(rec_model, c['d'], [camt], r, Ref_Dats, Src_Dats, T(poses), bv_pred, R_net=False)
        full = [a.numpy() for a in rec_model.last]
        key = 'cfg/%s/step%d' % (name, step)
        rec = {'seconds_reference': None}
        thin = step >= cases.BIG_FULL_STEPS
        for nm, a in zip(NAMES4, full):
            if thin and nm in ('dmap_cur_refined', 'BV_cur'):
                continue
            out['%s/%s' % (key, nm)] = cases.subsample_to(a, 30000) if thin else cases.subsample(a)
            out['%s/%s_stats' % (key, nm)] = cases.stats(np.exp(a.astype(np.float64)))
            rec[nm + '_min'] = float(a.min())
        dep = m_misc.depth_val_regression(T(full[3]), c['d'], BV_log=True).numpy()
        out[key + '/depth'] = cases.subsample_to(dep, 5000)
        out[key + '/BV_predict_next'] = cases.subsample_to(bv_next.numpy(), 30000)
        rec['kv_dpv_eq_forward'] = dev(kv_dpv.numpy(), full[3])
        if step == 0:
            priors['cfg/%s/step0/BV_predict_next_full' % name] = bv_next.numpy()
        if step < oracle_steps:
            to = time.time()
            o = N.kvnet_forward(c['sd'], ref_f, src_f, poses, cam, c['d'], c['sigma'],
                                BV_predict=None if bv_pred is None else bv_pred.numpy())
            for nm, a, b in zip(NAMES4, full, o):
                rec['oracle_' + nm + '_prob'] = dev(np.exp(a), np.exp(b))
                rec['oracle_' + nm + '_log'] = dev(a, b)
            rec['oracle_depth_mm'] = 1000 * dev(dep, O.depth_val_regression(o[3], c['d']))
            rec['oracle_argmax_flips'] = int((full[3].argmax(1) != o[3].argmax(1)).sum())
            rel_t = T(poses)[0, r].inverse().numpy()
            rec['oracle_BV_predict_next_prob'] = dev(np.exp(bv_next.numpy()), np.exp(O.propagate_dpv(o[3], rel_t, cam, c['d'])))
            rec['seconds_oracle'] = time.time() - to
            t0 += rec['seconds_oracle']
        rec['seconds_reference'] = time.time() - t0
        pin[key] = rec
        print(key, json.dumps(rec), flush=True)
        bv_pred = bv_next
    np.savez_compressed(os.path.join(HERE, 'configs_%s.npz' % name), **out)
    np.savez_compressed(os.path.join(HERE, 'configs_priors_%s.npz' % name), **priors)
    return pin


def main():
    names = sys.argv[1:] or cases.BIG_CASES
    path = os.path.join(HERE, 'PINNING_configs.json')
    pin = {'torch': torch.__version__, 'numpy': np.__version__, 'threads': torch.get_num_threads(), 'cases': {}}
    if os.path.exists(path):
        pin['cases'] = json.load(open(path)).get('cases', {})
    for name in names:
        pin['cases'].update(run_case(name))
        with open(path, 'w') as f:
            json.dump(pin, f, indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
