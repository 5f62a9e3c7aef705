"""CPU torch port of the reference's path, used ONLY as the timed CPU arm of bench.py
(`cpu_baseline`, `--impl reference`) and cross-checked against the numpy oracle in tests.

TEST / BASELINE INFRASTRUCTURE ONLY (see planesweep_oracle.py header for the rules).

The reference cannot travel to the GPU box and its source may not be copied, so this file
restates the same op sequence with the same ATen ops the reference calls (`matmul`, `repeat`,
`F.grid_sample`, `conv2d/3d`, `batch_norm(training=True)`, `log_softmax`), on CPU tensors, with
the state_dict names of the reference. It therefore costs what the reference costs on the host
cores (same kernels, same D x C x h x w intermediates), unlike the numpy oracle whose gather is
~4x slower. Citations are to /root/reference/code.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def _P(sd):
    out = {}
    for k, v in sd.items():
        k = k[7:] if k.startswith('module.') else k
        out[k] = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))
    return out


def back_warp_parallel(img_rep, d, term1, term2, cx, cy, H, W):
    """warping/homography.py:421-448."""
    n_d = d.shape[0]
    P = term1.unsqueeze(0) + term2.repeat(n_d, 1, 1) * d.reshape(n_d, 1, 1)
    P = P / (P[:, 2, :].unsqueeze(1) + 1e-10)
    grid = torch.empty(n_d, H, W, 2)
    grid[..., 0] = (P[:, 0, :].reshape(n_d, H, W) - cx) / cx
    grid[..., 1] = (P[:, 1, :].reshape(n_d, H, W) - cy) / cy
    return F.grid_sample(img_rep, grid, mode='bilinear', padding_mode='zeros', align_corners=False)


def est_swp_volume_v4(ref, src, d_candi, R, t, cam, sigma, feat_dist='L2'):
    """warping/homography.py:293-331."""
    H, W, D = ref.shape[2], ref.shape[3], len(d_candi)
    cost = torch.zeros(1, D, H, W)
    K = torch.as_tensor(cam['intrinsic_M_cuda']).float()
    rays = torch.as_tensor(cam['unit_ray_array_2D']).float()
    d = torch.from_numpy(np.asarray(d_candi).astype(np.float32))
    cx, cy = cam['intrinsic_M'][0, 2], cam['intrinsic_M'][1, 2]
    for v in range(src.shape[1]):
        term1 = K.matmul(t[v, :]).reshape(3, 1)
        term2 = K.matmul(R[v, :, :]).matmul(rays)
        warped = back_warp_parallel(src[:, v].repeat(D, 1, 1, 1), d, term1, term2, cx, cy, H, W)
        if feat_dist == 'L2':
            cost[0] = cost[0] + torch.sum((warped - ref) ** 2, 1) / sigma
        elif feat_dist == 'L1':
            cost[0] = cost[0] + torch.sum(torch.abs(warped - ref), 1) / sigma
        else:
            raise Exception('undefined metric for feature distance ...')
    return cost


def _convbn(P, pre, x, stride, pad, dil):
    y = F.conv2d(x, P[pre + '.0.weight'], None, stride, dil if dil > 1 else pad, dil)
    return F.batch_norm(y, None, None, P[pre + '.1.weight'], P[pre + '.1.bias'], True, 0.1, 1e-5)


def _block(P, pre, x, stride, dil, down):
    o = F.relu(_convbn(P, pre + '.conv1.0', x, stride, 1, dil))
    o = _convbn(P, pre + '.conv2', o, 1, 1, dil)
    if down:
        x = F.batch_norm(F.conv2d(x, P[pre + '.downsample.0.weight'], None, stride), None, None,
                         P[pre + '.downsample.1.weight'], P[pre + '.downsample.1.bias'], True, 0.1, 1e-5)
    return o + x


def _layer(P, pre, x, blocks, stride, dil, down):
    for i in range(blocks):
        x = _block(P, '%s.%d' % (pre, i), x, stride if i == 0 else 1, dil, down and i == 0)
    return x


def feature_extraction(P, x, pre='feature_extractor.feature_extraction'):
    """models/psm_submodule.py:141-167."""
    o = F.relu(_convbn(P, pre + '.firstconv.0', x, 2, 1, 1))
    o = F.relu(_convbn(P, pre + '.firstconv.2', o, 1, 1, 1))
    o = F.relu(_convbn(P, pre + '.firstconv.4', o, 1, 1, 1))
    l1 = _layer(P, pre + '.layer1', o, 3, 1, 1, False)
    raw = _layer(P, pre + '.layer2', l1, 16, 2, 1, True)
    o = _layer(P, pre + '.layer3', raw, 3, 1, 1, True)
    skip = _layer(P, pre + '.layer4', o, 3, 1, 2, False)
    br = []
    for name, k in (('branch1', 64), ('branch2', 32), ('branch3', 16), ('branch4', 8)):
        b = F.relu(_convbn(P, '%s.%s.1' % (pre, name), F.avg_pool2d(skip, k, k), 1, 0, 1))
        br.append(F.interpolate(b, skip.shape[2:], mode='bilinear', align_corners=True))
    cat = torch.cat((raw, skip, br[3], br[2], br[1], br[0]), 1)
    o = F.relu(_convbn(P, pre + '.lastconv.0', cat, 1, 1, 1))
    return l1, F.conv2d(o, P[pre + '.lastconv.2.weight'])


def r_net(P, dpv, feats, pre='r_net'):
    """models/Refine.py:79-107."""
    def cl(n, x):
        return F.leaky_relu(F.conv2d(x, P[n + '.0.weight'], P[n + '.0.bias'], 1, 1))

    def tl(n, x):
        return F.leaky_relu(F.conv_transpose2d(x, P[n + '.0.weight'], P[n + '.0.bias'], 2, 1))
    o = cl(pre + '.conv0_1', cl(pre + '.conv0', torch.cat([dpv, feats[0]], 1)))
    o = tl(pre + '.trans_conv0', o)
    o = cl(pre + '.conv1_1', cl(pre + '.conv1', torch.cat([o, feats[1]], 1)))
    o = tl(pre + '.trans_conv1', o)
    o = cl(pre + '.conv2_1', cl(pre + '.conv2', torch.cat([o, feats[2]], 1)))
    o = F.conv2d(o, P[pre + '.conv2_2.weight'], P[pre + '.conv2_2.bias'], 1, 1)
    return F.log_softmax(o, dim=1)


def kvnet_first_window(sd, ref_frame, src_frames, poses, cam, d_candi, sigma):
    """models/KVNET.py:93-143 first-window branch (D-Net + R-Net); numpy in, numpy out."""
    P = _P(sd)
    ref_frame = torch.as_tensor(ref_frame).float(); src_frames = torch.as_tensor(src_frames).float()
    poses = torch.as_tensor(poses).float()
    with torch.no_grad():
        l1, feats = feature_extraction(P, torch.cat((src_frames[0], ref_frame), 0))
        fr = torch.cat((feats[-1:], F.avg_pool2d(ref_frame, 4)), 1)
        fs = torch.cat((feats[:-1].unsqueeze(0), F.avg_pool2d(src_frames[0], 4).unsqueeze(0)), 2)
        cost = est_swp_volume_v4(fr, fs, d_candi, poses[0, :, :3, :3], poses[0, :, :3, 3], cam, sigma)
        BV = F.log_softmax(-cost, dim=1)
        ref = r_net(P, torch.exp(BV), [fr[:, :-3], l1[-1:], ref_frame])
        depth = torch.zeros(1, ref.shape[2], ref.shape[3])
        for i, dd in enumerate(d_candi):            # mutils/misc.py:541-546
            depth = depth + torch.exp(ref[0, i]) * dd
    return ref.numpy(), BV.numpy(), depth.numpy()
