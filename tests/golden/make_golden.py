"""Generate the golden fixtures by running the UNMODIFIED reference (CPU torch).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference hard-codes .cuda(); on this GPU-less container the four-line shim of
SURVEY.md §8(c) turns those into no-ops. Inputs come from tests/cases.py (seeded);
only reference OUTPUTS are stored (large ones strided, plus full-array checksums).
Also runs the oracle on the same inputs and writes the measured oracle-vs-reference
deviations to tests/golden/PINNING.json (the record that pins the oracle).
"""
import contextlib
import io
import json
import math
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
import models.KVNET as m_kvnet                       # noqa: E402  (reference)
import test_utils.test_KVNet as ref_test             # noqa: E402  (reference)
import mutils.misc as m_misc                         # noqa: E402  (reference)

from oracle import planesweep_oracle as O            # noqa: E402
from oracle import kvnet_oracle as N                 # noqa: E402
from tests import cases                              # noqa: E402

T = torch.from_numpy


def cam_torch(cam):
    c = dict(cam)
    c['unit_ray_array_2D'] = T(cam['unit_ray_array_2D'])
    c['intrinsic_M_cuda'] = T(cam['intrinsic_M_cuda'])
    return c


def dev(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.nanmax(np.abs(a - b))) if a.size else 0.0


def main():
    pin = {'torch': torch.__version__, 'numpy': np.__version__, 'cases': {}}
    out = {}

    # ---- a1-a3 sweep ------------------------------------------------------
    for name in cases.SWEEP_CASES:
        c = cases.sweep_case(name)
        cam = cases.cam_for(O.make_cam_intrinsics, c['w'], c['h'])
        ref = wh.est_swp_volume_v4(T(c['ref']), T(c['src']), c['d'], T(c['R']), T(c['t']), cam_torch(cam),
                                   c['sigma'], feat_dist=c['feat_dist']).numpy()
        orc = O.est_swp_volume_v4(c['ref'], c['src'], c['d'], c['R'], c['t'], cam, c['sigma'], c['feat_dist'])
        bv_r = torch.log_softmax(-T(ref), 1).numpy(); bv_o = O.d_net_dpv_from_cost(orc)
        out['sweep/%s/cost' % name] = cases.subsample(ref)
        out['sweep/%s/cost_stats' % name] = cases.stats(ref)
        pin['cases']['sweep/' + name] = {'cost_maxabs': dev(ref, orc), 'cost_max': float(ref.max()),
                                         'logdpv_maxabs': dev(bv_r, bv_o),
                                         'prob_maxabs': dev(np.exp(bv_r), np.exp(bv_o))}
        print(name, pin['cases']['sweep/' + name])

    # ---- a7 image warp ----------------------------------------------------
    for name in cases.WARP_CASES:
        c = cases.warp_case(name)
        cam = cases.cam_for(O.make_cam_intrinsics, c['w'], c['h'])
        ref = wh.warp_img_feats_v3([T(i) for i in c['imgs']], c['d'], [T(r) for r in c['R']],
                                   [T(t) for t in c['t']], cam_torch(cam))
        ref = np.stack([r.numpy() for r in ref])
        orc = np.stack(O.warp_img_feats_v3(c['imgs'], c['d'], c['R'], c['t'], cam))
        # mgpu entry (homography.py:183-232) must give the same numbers
        mg = wh.warp_img_feats_mgpu([T(i) for i in c['imgs']], c['d'], [T(r) for r in c['R']],
                                    [T(t) for t in c['t']], T(cam['intrinsic_M_cuda'])[None],
                                    T(cam['unit_ray_array_2D'])[None])
        mg = np.stack([r.numpy() for r in mg])
        out['warp/%s/vol' % name] = cases.subsample(ref)
        out['warp/%s/vol_stats' % name] = cases.stats(ref)
        pin['cases']['warp/' + name] = {'maxabs': dev(ref, orc), 'mgpu_vs_v3': dev(ref, mg)}
        print(name, pin['cases']['warp/' + name])

    # ---- a12 resample -----------------------------------------------------
    for name in cases.RESAMPLE_CASES:
        c = cases.resample_case(name)
        cam = cases.cam_for(O.make_cam_intrinsics, c['w'], c['h'])
        ref = wh.resample_vol_cuda(T(c['vol']), T(c['rel']), cam_torch(cam), c['d'], d_candi_new=c['d_new'],
                                   padding_value=c['pad']).numpy()
        orc = O.resample_vol_cuda(c['vol'], c['rel'], cam, c['d'], d_candi_new=c['d_new'], padding_value=c['pad'])
        out['resample/%s/vol' % name] = cases.subsample(ref)
        out['resample/%s/vol_stats' % name] = cases.stats(ref)
        pin['cases']['resample/' + name] = {'maxabs': dev(ref, orc),
                                            'nonidentity_vs_input': dev(ref, c['vol'][0])}
        print(name, pin['cases']['resample/' + name])

    # ---- a11 depth regression ---------------------------------------------
    c = cases.resample_case('resample_pose_d32')
    ref = m_misc.depth_val_regression(T(c['vol']), c['d'], BV_log=True).numpy()
    orc = O.depth_val_regression(c['vol'], c['d'])
    out['regress/resample_pose_d32/depth'] = ref
    pin['cases']['regress/resample_pose_d32'] = {'maxabs_m': dev(ref, orc)}
    print('regress', pin['cases']['regress/resample_pose_d32'])

    # ---- a4-a10 full KVNET.forward + unmodified test() streaming ------------
    for name in cases.KVNET_CASES:
        c = cases.kvnet_case(name)
        cam = cases.cam_for(O.make_cam_intrinsics, c['W'] // 4, c['H'] // 4)
        camt = cam_torch(cam)
        with contextlib.redirect_stdout(io.StringIO()):
            model = m_kvnet.KVNET(feature_dim=64, cam_intrinsics=camt, d_candi=c['d'], sigma_soft_max=c['sigma'],
                                  KVNet_feature_dim=64, d_upsample_ratio_KV_net=None, t_win_r=2, if_refined=True)
        model.load_state_dict({k: T(np.asarray(v)) for k, v in c['sd'].items()})
        model = torch.nn.DataParallel(model)      # as test_KVNet.py:163 (no GPUs -> falls through)
        bv_pred = None; bv_pred_o = None
        n_steps = len(c['frames']) - 4
        for step in range(n_steps):
            ridx = 2 + step
            ref_f, src_f, poses = cases.window(c, ridx)
            Ref_Dats = [{'img': T(ref_f)}]
            Src_Dats = [[{'img': T(src_f[0, v:v + 1])} for v in range(src_f.shape[1])]]
            # R_net=False -> (kv_dpv, BVs_predict); also capture the refined output via a direct call
            with torch.no_grad():
                full = model(ref_frame=T(ref_f), src_frames=T(src_f), src_cam_poses=T(poses),
                             BatchIdx=torch.zeros(1), cam_intrinsics=[camt], BV_predict=bv_pred)
            kv_dpv, bv_next = ref_test.test(model, c['d'], [camt], 2, Ref_Dats, Src_Dats, T(poses), bv_pred,
                                            R_net=False)
            o = N.kvnet_forward(c['sd'], ref_f, src_f, poses, cam, c['d'], c['sigma'], BV_predict=bv_pred_o)
            rel = np.linalg.inv(poses[0, 2].astype(np.float64)).astype(np.float32)
            rel_t = T(poses)[0, 2].inverse().numpy()
            bv_next_o = O.propagate_dpv(o[3], rel_t, cam, c['d'])
            key = 'kvnet/%s/step%d' % (name, step)
            names4 = ['dmap_cur_refined', 'dmap_refined', 'BV_cur', 'DPV']
            rec = {}
            for nm, a, b in zip(names4, full, o):
                a = a.numpy()
                out['%s/%s' % (key, nm)] = cases.subsample(a)
                out['%s/%s_stats' % (key, nm)] = cases.stats(a)
                rec[nm + '_log'] = dev(a, b); rec[nm + '_prob'] = dev(np.exp(a), np.exp(b))
                rec[nm + '_min'] = float(a.min())
            dep_r = m_misc.depth_val_regression(full[3], c['d'], BV_log=True).numpy()
            dep_o = O.depth_val_regression(o[3], c['d'])
            rec['depth_mm'] = 1000 * dev(dep_r, dep_o)
            out[key + '/rel_inv'] = rel_t
            out[key + '/BV_predict_next'] = cases.subsample(bv_next.numpy())
            out[key + '/BV_predict_next_stats'] = cases.stats(bv_next.numpy())
            if step < n_steps - 1:      # the next step's prior, in full, so tests can feed the reference's prior
                out[key + '/BV_predict_next_full'] = bv_next.numpy()
            rec['kv_dpv_eq_full'] = dev(kv_dpv.numpy(), full[3].numpy())
            rec['BV_predict_next_log'] = dev(bv_next.numpy(), bv_next_o)
            pin['cases'][key] = rec
            print(key, rec)
            bv_pred = bv_next
            bv_pred_o = bv_next.numpy()      # feed the REFERENCE prior to the oracle: per-step deviation
        # NaN-prior branch (KVNET.py:142): must equal the first-window outputs
        ref_f, src_f, poses = cases.window(c, 2)
        nanp = torch.full((1, c['D'], c['H'] // 4, c['W'] // 4), float('nan'))
        with torch.no_grad():
            full = model(ref_frame=T(ref_f), src_frames=T(src_f), src_cam_poses=T(poses),
                         BatchIdx=torch.zeros(1), cam_intrinsics=[camt], BV_predict=nanp)
        pin['cases']['kvnet/%s/nan_prior_eq_first' % name] = dev(
            cases.subsample(full[3].numpy()), out['kvnet/%s/step0/DPV' % name])

    np.savez_compressed(os.path.join(HERE, 'reference_outputs.npz'), **out)
    with open(os.path.join(HERE, 'PINNING.json'), 'w') as f:
        json.dump(pin, f, indent=1, sort_keys=True)
    sz = os.path.getsize(os.path.join(HERE, 'reference_outputs.npz'))
    print('wrote reference_outputs.npz %.2f MB, %d arrays' % (sz / 1e6, len(out)))


if __name__ == '__main__':
    main()
