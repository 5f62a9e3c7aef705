"""Micro-benchmarks of the geometry kernels at the BASELINE metric shape (development aid; the
judged numbers come from bench.py). CUDA events on the current stream, L2 flushed between
iterations by writing a 256 MB buffer."""
import ctypes
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralrgbd_b200 import _lib, synth           # noqa: E402
from neuralrgbd_b200._lib import ptr, check       # noqa: E402
import neuralrgbd_b200.warping.homography as H    # noqa: E402

dev = torch.device('cuda:0')
L = _lib.lib()
F = ctypes.c_float
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)     # noqa: E731
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


def main():
    h, w, C, V, D = 120, 160, 67, 4, 64
    pos = [a for a in sys.argv[1:] if '=' not in a and a != 'bnsweep']
    if pos:
        h, w, C, V, D = [int(x) for x in pos[:5]]
    hw = h * w
    rng = np.random.RandomState(0)
    Cw, Cn = C - C % 4, C % 4
    ref_w = torch.randn(hw, Cw, device=dev); src_w = torch.randn(V, hw, Cw, device=dev)
    ref_n = torch.randn(hw, 4, device=dev); src_n = torch.randn(V, hw, 4, device=dev)
    exts = synth.camera_track(rng, V + 1)
    poses, _ = synth.window_rel_poses(exts, V // 2, V // 2)
    R = torch.from_numpy(np.ascontiguousarray(poses[:, :3, :3])).to(dev); t = torch.from_numpy(np.ascontiguousarray(poses[:, :3, 3])).to(dev)
    fx = (w / 2) / np.tan(np.arctan(320 / 585.)); fy = (h / 2) / np.tan(np.arctan(240 / 585.))
    K = torch.tensor([[fx, 0, w / 2], [0, fy, h / 2], [0, 0, 1]], dtype=torch.float32, device=dev)
    xs = (np.arange(w) + .5) / w * 2 - 1; ys = (np.arange(h) + .5) / h * 2 - 1
    rays = np.stack([np.tile(320 / 585. * xs[None], (h, 1)), np.tile(240 / 585. * ys[:, None], (1, w)), np.ones((h, w))]).reshape(3, -1)
    rays = torch.from_numpy(rays.astype(np.float32)).to(dev)
    dpl = torch.from_numpy(synth.d_candidates(D).astype(np.float32)).to(dev)
    ws = torch.empty(V * 12, device=dev); cost = torch.empty(hw, D, device=dev)

    def sweep():
        check(L.nrgbd_plane_sweep_cost_packed(ptr(ref_w) if Cw else None, ptr(ref_n) if Cn else None,
                                              ptr(src_w) if Cw else None, ptr(src_n) if Cn else None, Cw, Cn, V, D, h, w,
                                              ptr(K), ptr(R), ptr(t), ptr(rays), ptr(dpl), F(w / 2), F(h / 2), F(10.), 0,
                                              ptr(ws), ptr(cost), st()))
    res = {}
    med, mn = timeit(sweep)
    alg = (1 + V) * C * hw * 4 + D * hw * 4 + 3 * hw * 4
    res['sweep'] = dict(us_med=med, us_min=mn, alg_MB=alg / 1e6, GBps=alg / med / 1e3,
                        gflops=V * D * hw * (11 * C + 30) / med / 1e3)
    bv = torch.log_softmax(-cost, 1).contiguous(); out = torch.empty_like(bv)
    dep = torch.empty(hw, device=dev); conf = torch.empty(hw, device=dev)

    def norm():
        check(L.nrgbd_dpv_normalize(ptr(cost), 1, D, None, 0, 0, F(-1.), hw, D, ptr(out), 1, D, ptr(dpl), ptr(dep), ptr(conf), st()))
    med, mn = timeit(norm)
    res['dpv_normalize_hwd'] = dict(us_med=med, us_min=mn, GBps=2 * hw * D * 4 / med / 1e3)
    E = torch.from_numpy(np.linalg.inv(poses[V // 2].astype(np.float64)).astype(np.float32)).to(dev)
    vol = bv.t().contiguous(); outv = torch.empty_like(vol)

    def resample():
        check(L.nrgbd_resample_dpv(ptr(vol), hw, 1, ptr(E), ptr(rays), ptr(dpl), D, h, w, F(320 / 585.), F(240 / 585.),
                                   F(2.55), F(2.45), F(-4.16), 1, F(-1000.), F(0.), ptr(outv), hw, 1, st()))
    med, mn = timeit(resample)
    res['resample_dhw'] = dict(us_med=med, us_min=mn, GBps=2 * hw * D * 4 / med / 1e3)
    rgb = torch.randn(V, hw, 4, device=dev); outw = torch.empty(V, 3, D, hw, device=dev)

    def warp():
        check(L.nrgbd_warp_to_volume(ptr(rgb), 3, 0, 3, V, D, h, w, ptr(K), ptr(R), ptr(t), ptr(rays), ptr(dpl), F(w / 2),
                                     F(h / 2), ptr(ws), ptr(outw), st()))
    med, mn = timeit(warp)
    res['warp_to_volume'] = dict(us_med=med, us_min=mn, GBps=V * (3 * hw + 3 * D * hw) * 4 / med / 1e3)
    # f-1: sweep backward (scatter of the four bilinear weights with 16-byte vector atomics)
    gcost = torch.randn(hw, D, device=dev)
    g_rw = torch.empty_like(ref_w); g_rn = torch.empty_like(ref_n); g_sw = torch.empty_like(src_w); g_sn = torch.empty_like(src_n)

    def bwd():
        check(L.nrgbd_plane_sweep_backward_packed(ptr(ref_w), ptr(ref_n), ptr(src_w), ptr(src_n), Cw, Cn, V, D, h, w, ptr(K), ptr(R), ptr(t),
                                                  ptr(rays), ptr(dpl), F(w / 2), F(h / 2), F(10.), 0, ptr(ws), ptr(gcost), ptr(g_rw), ptr(g_rn),
                                                  ptr(g_sw), ptr(g_sn), st()))
    med, mn = timeit(bwd)
    n_red = V * D * hw * 4 * (Cw // 4 + (1 if Cn else 0))
    res['sweep_backward'] = dict(us_med=med, us_min=mn, vector_atomics=n_red, atomics_per_ns=n_red / med / 1e3,
                                 alg_MB=(2 * (1 + V) * C * hw * 4 + D * hw * 4) / 1e6)
    # f-2: output stage at full resolution (640x480 when the sweep shape is 120x160)
    HW = 16 * hw
    bvf = torch.log_softmax(torch.randn(D, HW, device=dev), 0).contiguous()
    dm = torch.empty(HW, device=dev); cf = torch.empty(HW, device=dev)
    d16 = torch.empty(HW, device=dev, dtype=torch.int16); c16 = torch.empty(HW, device=dev, dtype=torch.int16)

    def export():
        check(L.nrgbd_export_depth_conf(ptr(bvf), ptr(dpl), D, HW, F(1000.), F(1000.), ptr(dm), ptr(cf), ptr(d16), ptr(c16), st()))
    med, mn = timeit(export)
    res['export_depth_conf'] = dict(us_med=med, us_min=mn, alg_MB=(D * HW * 4 + 12 * HW) / 1e6, GBps=(D * HW * 4 + 12 * HW) / med / 1e3)
    # K-Net input volume rows (fp32 and operand-pair outputs), LBA back-warp (f-3) forward / backward at full resolution
    refq = torch.randn(hw, 4, device=dev); prior = torch.log_softmax(torch.randn(hw, D, device=dev), 1).contiguous()
    volf = torch.empty(D, hw, 32, device=dev); vh = torch.empty(D, hw, 32, device=dev, dtype=torch.float16); vl = torch.empty_like(vh)

    def knet_f32():
        check(L.nrgbd_knet_input_volume(ptr(rgb), ptr(refq), ptr(bv), ptr(prior), V, D, h, w, 32, ptr(K), ptr(R), ptr(t), ptr(rays), ptr(dpl),
                                        F(w / 2), F(h / 2), ptr(ws), ptr(volf), st()))

    def knet_pair():
        check(L.nrgbd_knet_input_volume_pair(ptr(rgb), ptr(refq), ptr(bv), ptr(prior), V, D, h, w, 32, ptr(K), ptr(R), ptr(t), ptr(rays), ptr(dpl),
                                             F(w / 2), F(h / 2), ptr(ws), None, ptr(vh), ptr(vl), st()))
    med, mn = timeit(knet_f32)
    res['knet_input_volume_f32'] = dict(us_med=med, us_min=mn, GBps=(D * hw * 32 * 4 + 2 * D * hw * 4) / med / 1e3)
    med, mn = timeit(knet_pair)
    res['knet_input_volume_pair'] = dict(us_med=med, us_min=mn, GBps=(D * hw * 32 * 4 + 2 * D * hw * 4) / med / 1e3)
    # BatchNorm pass over a K-Net volume [D][hw][64] (statistics -> scale / shift, ReLU, operand-pair output; with a pair residual)
    xk = torch.randn(D * hw, 64, device=dev)
    st64 = torch.stack([xk.double().sum(0), (xk.double() ** 2).sum(0)]).contiguous()
    gam = torch.rand(64, device=dev) + 0.5; bet = torch.randn(64, device=dev)
    yh = torch.empty(D * hw, 64, device=dev, dtype=torch.float16); yl = torch.empty_like(yh)
    rh = torch.randn(D * hw, 64, device=dev).half(); rl = torch.zeros_like(rh)

    def bn_plain():
        check(L.nrgbd_bn_apply_stats_pair(ptr(xk), ctypes.c_void_p(st64.data_ptr()), float(D * hw), ptr(gam), ptr(bet), F(1e-5), None, None, F(0.1), None, None,
                                          None, 1, D * hw, 64, 64, None, ptr(yh), ptr(yl), None, st()))

    def bn_res():
        check(L.nrgbd_bn_apply_stats_pair(ptr(xk), ctypes.c_void_p(st64.data_ptr()), float(D * hw), ptr(gam), ptr(bet), F(1e-5), None, None, F(0.1), None, ptr(rh),
                                          ptr(rl), 0, D * hw, 64, 64, None, ptr(yh), ptr(yl), None, st()))
    if 'bnsweep' in sys.argv:          # development: vectors in flight per thread x grid cap
        DL = _lib.dev_lib()
        for u in (1, 2, 4):
            for bps in (4, 8, 16, 32):
                DL.nrgbd_dev_set_bn_unroll(u); DL.nrgbd_dev_set_bn_blocks_per_sm(bps)
                a, _ = timeit(bn_plain); b, _ = timeit(bn_res)
                print(json.dumps(dict(unroll=u, blocks_per_sm=bps, plain_us=a, plain_GBps=D * hw * 64 * 8 / a / 1e3, res_us=b, res_GBps=D * hw * 64 * 12 / b / 1e3)), flush=True)
        return
    med, mn = timeit(bn_plain)
    res['bn_pass_knet_volume_relu_pair'] = dict(us_med=med, us_min=mn, alg_MB=D * hw * 64 * 8 / 1e6, GBps=D * hw * 64 * 8 / med / 1e3)
    med, mn = timeit(bn_res)
    res['bn_pass_knet_volume_pair_residual'] = dict(us_med=med, us_min=mn, alg_MB=D * hw * 64 * 12 / 1e6, GBps=D * hw * 64 * 12 / med / 1e3)
    Hf, Wf = 4 * h, 4 * w
    imgs = torch.randn(V, 3, Hf, Wf, device=dev); dm_full = (0.8 + 2.5 * torch.rand(Hf, Wf, device=dev))
    xsf = (np.arange(Wf) + .5) / Wf * 2 - 1; ysf = (np.arange(Hf) + .5) / Hf * 2 - 1
    raysf = np.stack([np.tile(320 / 585. * xsf[None], (Hf, 1)), np.tile(240 / 585. * ysf[:, None], (1, Wf)), np.ones((Hf, Wf))]).reshape(3, -1)
    raysf = torch.from_numpy(raysf.astype(np.float32)).to(dev)
    Kf = torch.tensor([[4 * fx, 0, Wf / 2], [0, 4 * fy, Hf / 2], [0, 0, 1]], dtype=torch.float32, device=dev)
    wout = torch.empty_like(imgs); gout = torch.randn_like(imgs); gimg = torch.empty_like(imgs)
    gR = torch.empty(V, 3, 3, device=dev); gt = torch.empty(V, 3, device=dev); wsd = torch.empty(V * 12, device=dev, dtype=torch.float64)

    def lba_f():
        check(L.nrgbd_lba_back_warp(ptr(imgs), ptr(dm_full), ptr(R), ptr(t), ptr(Kf), ptr(raysf), V, 3, Hf, Wf, ptr(wout), st()))

    def lba_b():
        check(L.nrgbd_lba_back_warp_backward(ptr(gout), ptr(imgs), ptr(dm_full), ptr(R), ptr(t), ptr(Kf), ptr(raysf), V, 3, Hf, Wf, None, ptr(gR), ptr(gt),
                                             ctypes.c_void_p(wsd.data_ptr()), st()))
    med, mn = timeit(lba_f)
    res['lba_back_warp_4views_640x480'] = dict(us_med=med, us_min=mn, GBps=(2 * V * 3 * Hf * Wf * 4 + 4 * Hf * Wf * 4) / med / 1e3)
    med, mn = timeit(lba_b)
    res['lba_back_warp_pose_gradients'] = dict(us_med=med, us_min=mn, GBps=(2 * V * 3 * Hf * Wf * 4 + 4 * Hf * Wf * 4) / med / 1e3)
    print(json.dumps(dict(shape=dict(h=h, w=w, C=C, V=V, D=D), kernels=res), indent=1))


if __name__ == '__main__':
    main()
