"""CPU oracle for the plane-sweep DPV hot path of NVlabs/neuralrgbd.

TEST INFRASTRUCTURE ONLY.  Nothing under ``neuralrgbd_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker /
the timed CPU arm — never as the product.

What it is: a numpy (float32) restatement of the reference's algorithm for the
path named in BASELINE.json, function by function, each citing the reference
file:line it follows (paths relative to /root/reference/code).  The reference is
pure Python over PyTorch ATen ops; the arithmetic of ``F.grid_sample`` /
``conv`` / ``batch_norm`` / ``log_softmax`` lives in PyTorch (requirements.txt
pins no version; this container has torch 2.11.0, whose ``grid_sample`` default
is ``align_corners=False``).  The ATen semantics restated here are the published
ones of ``ATen/native/GridSampler.h``:
    unnormalise   ix = ((g + 1) * size - 1) / 2
    zeros padding an out-of-range corner contributes 0 (its weight is kept)
    border padding clip ix to [0, size-1] before interpolation

Pinning: the reference ships no tests/golden vectors for this path,
so the oracle is pinned against the LIVE reference imported in the build
container (tests/golden/make_golden.py, which also writes the committed
fixtures in tests/golden/*.npz that the GPU box uses).  tests/test_oracle_golden.py
re-checks the oracle against those fixtures on every run.

Layout conventions follow the reference: NCHW / NCDHW float32, N = 1.
All arithmetic that determines sampling coordinates is done op by op in
float32 in the order the reference's torch ops execute: the small matmuls
(K.t, K.R, (K.R).rays, E.[X;1]) as sgemm-style FMA chains over k, everything
else as separate rounded mul / add / div (torch eager never contracts across
ops).  The CUDA kernels use __fmaf_rn / __fmul_rn / __fadd_rn / __fdiv_rn in the
same order, so they agree with this file to the last bit on coordinates.
"""
from __future__ import annotations

import math
import numpy as np

f32 = np.float32


# --------------------------------------------------------------------------
# a14: intrinsics recipe + unit-ray table
# --------------------------------------------------------------------------
def pixel_to_ray_array(width, height, hfov, vfov):
    """warping/View.py:16-30,32-62 (normalize_z=True): float64 H x W x 3 table
    (tan(hfov/2)*(2(x+.5)/W-1), tan(vfov/2)*(2(y+.5)/H-1), 1)."""
    xs = np.array([math.tan(math.radians(hfov / 2.0)) * ((2.0 * ((x + 0.5) / width)) - 1.0)
                   for x in range(width)], dtype=np.float64)
    ys = np.array([math.tan(math.radians(vfov / 2.0)) * ((2.0 * ((y + 0.5) / height)) - 1.0)
                   for y in range(height)], dtype=np.float64)
    out = np.zeros((height, width, 3), dtype=np.float64)
    out[:, :, 0] = xs[None, :]
    out[:, :, 1] = ys[:, None]
    out[:, :, 2] = 1.0
    return out


def make_cam_intrinsics(fx, fy, cx, cy, out_size):
    """mdataloader/scanNet.py:239-270 (read_IntM_from_txt with out_size=[w,h]).
    Returns the reference's dict with numpy members (the host mirror converts the
    two tensor members to torch)."""
    h_fov = math.degrees(math.atan(cx / fx) * 2)
    v_fov = math.degrees(math.atan(cy / fy) * 2)
    pw, ph = out_size[0], out_size[1]
    K = np.zeros((3, 4))
    K[2, 2] = 1.
    K[0, 0] = (pw / 2.0) / math.tan(math.radians(h_fov / 2.0))
    K[0, 2] = pw / 2.0
    K[1, 1] = (ph / 2.0) / math.tan(math.radians(v_fov / 2.0))
    K[1, 2] = ph / 2.0
    rays = pixel_to_ray_array(pw, ph, h_fov, v_fov)
    rays2d = np.reshape(np.transpose(rays, axes=[2, 0, 1]), [3, -1]).astype(np.float32)
    return {'hfov': h_fov, 'vfov': v_fov, 'unit_ray_array': rays,
            'unit_ray_array_2D': rays2d,
            'intrinsic_M_cuda': K[:3, :3].astype(np.float32),
            'focal_length': float(np.mean([fx, fy])) * pw / (2.0 * cx),
            'intrinsic_M': K}


def get_rel_extrinsicM(ext_ref, ext_src):
    """warping/homography.py:904-906."""
    return ext_src.dot(np.linalg.inv(ext_ref))


# --------------------------------------------------------------------------
# fp32 helpers with a fixed, documented rounding order
# --------------------------------------------------------------------------
def _fma(a, b, c):
    """fp32 fused multiply-add emulated through float64 (the product of two
    fp32 values is exact in fp64; the one extra rounding fp64->fp32 differs from a
    true FMA only in vanishingly rare double-rounding cases)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def _dot3(a0, a1, a2, b0, b1, b2):
    """3-term dot product in the order an sgemm micro-kernel uses:
    acc = a0*b0; acc = fma(a1, b1, acc); acc = fma(a2, b2, acc).
    Probe (tests/golden/make_golden.py, recorded in DESIGN.md): this reproduces
    torch 2.11 CPU `matmul` for the 3x3.3xN products of homography.py:315-317
    bit for bit; the engine's setup kernel uses __fmaf_rn in the same order."""
    a0, a1, a2 = f32(a0), f32(a1), f32(a2)
    acc = (a0 * np.asarray(b0, f32)).astype(f32)
    acc = _fma(a1, b1, acc)
    return _fma(a2, b2, acc)


def homography_terms(K, R, t, rays2d):
    """warping/homography.py:315-317: term1 = K.t ; term2 = (K.R).rays."""
    K = np.asarray(K, f32); R = np.asarray(R, f32); t = np.asarray(t, f32)
    rays2d = np.asarray(rays2d, f32)
    term1 = np.array([_dot3(K[i, 0], K[i, 1], K[i, 2], t[0], t[1], t[2]) for i in range(3)], f32)
    KR = np.array([[_dot3(K[i, 0], K[i, 1], K[i, 2], R[0, j], R[1, j], R[2, j]) for j in range(3)]
                   for i in range(3)], f32)
    term2 = np.stack([_dot3(KR[i, 0], KR[i, 1], KR[i, 2], rays2d[0], rays2d[1], rays2d[2])
                      for i in range(3)]).astype(f32)
    return term1, KR, term2


def back_warp_grid(term1, term2, d, cx, cy):
    """warping/homography.py:434-446: P = term1 + term2*d; P /= (P_z + 1e-10);
    g = ((P_x - cx)/cx, (P_y - cy)/cy).  Returns gx, gy of shape [D, hw]."""
    d = np.asarray(d, f32).reshape(-1, 1)
    px = term1[0] + term2[0][None, :] * d
    py = term1[1] + term2[1][None, :] * d
    pz = term1[2] + term2[2][None, :] * d
    den = pz + f32(1e-10)
    px = px / den
    py = py / den
    gx = (px - f32(cx)) / f32(cx)
    gy = (py - f32(cy)) / f32(cy)
    return gx.astype(f32), gy.astype(f32)


def _unnormalize(g, size):
    """ATen GridSampler.h grid_sampler_unnormalize, align_corners=False."""
    return ((g + f32(1)) * f32(size) - f32(1)) / f32(2)


def grid_sample_2d_zeros(img, gx, gy):
    """F.grid_sample(mode='bilinear', padding_mode='zeros', align_corners=False)
    for one image. img [C,H,W]; gx, gy any shape S -> out [C, *S].
    Corner weights as in ATen's CUDA/generic kernel: nw=(ix_se-ix)*(iy_se-iy) ..."""
    C, H, W = img.shape
    ix = _unnormalize(gx.astype(f32), W)
    iy = _unnormalize(gy.astype(f32), H)
    x0f = np.floor(ix); y0f = np.floor(iy)
    x1f = x0f + f32(1); y1f = y0f + f32(1)
    wnw = (x1f - ix) * (y1f - iy)
    wne = (ix - x0f) * (y1f - iy)
    wsw = (x1f - ix) * (iy - y0f)
    wse = (ix - x0f) * (iy - y0f)
    # NaN / huge coordinates: every corner is out of range -> 0 (ATen: the
    # within_bounds test fails for all four corners).
    bad = ~np.isfinite(ix) | ~np.isfinite(iy) | (np.abs(ix) > 1e9) | (np.abs(iy) > 1e9)
    x0 = np.where(bad, -10, x0f).astype(np.int64); y0 = np.where(bad, -10, y0f).astype(np.int64)
    x1 = x0 + 1; y1 = y0 + 1
    flat = img.reshape(C, H * W)

    def tap(xx, yy, wgt):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        idx = np.where(ok, yy * W + xx, 0)
        v = flat[:, idx.reshape(-1)].reshape((C,) + idx.shape)
        wgt = np.where(ok, wgt, f32(0)).astype(f32)
        wgt = np.where(bad, f32(0), wgt)
        return v * wgt[None]

    out = tap(x0, y0, wnw)
    out = out + tap(x1, y0, wne)
    out = out + tap(x0, y1, wsw)
    out = out + tap(x1, y1, wse)
    return out.astype(f32)


# --------------------------------------------------------------------------
# a1-a3: plane-sweep cost volume
# --------------------------------------------------------------------------
def est_swp_volume_v4(feat_img_ref, feat_img_src, d_candi, R, t, cam_intrinsic,
                      costV_sigma, feat_dist='L2', d_chunk=8):
    """warping/homography.py:293-331 (with _back_warp_homo_parallel :421-448 and
    img_dis_L2_pard/L1_pard :81-87).
    feat_img_ref [1,C,h,w]; feat_img_src [1,V,C,h,w]; R [V,3,3]; t [V,3]."""
    if feat_dist not in ('L2', 'L1'):
        raise Exception('undefined metric for feature distance ...')
    ref = np.asarray(feat_img_ref, f32)[0]
    src = np.asarray(feat_img_src, f32)[0]
    C, h, w = ref.shape
    V = src.shape[0]
    d32 = np.asarray(d_candi).astype(f32)
    D = len(d32)
    K = cam_intrinsic['intrinsic_M_cuda']
    rays = cam_intrinsic['unit_ray_array_2D']
    cx, cy = cam_intrinsic['intrinsic_M'][0, 2], cam_intrinsic['intrinsic_M'][1, 2]
    cost = np.zeros((D, h, w), f32)
    sigma = f32(costV_sigma)
    for v in range(V):
        term1, _, term2 = homography_terms(K, R[v], t[v], rays)
        gx, gy = back_warp_grid(term1, term2, d32, cx, cy)
        for d0 in range(0, D, d_chunk):
            d1 = min(D, d0 + d_chunk)
            warped = grid_sample_2d_zeros(src[v], gx[d0:d1].reshape(-1, h, w), gy[d0:d1].reshape(-1, h, w))
            diff = warped - ref[:, None]
            if feat_dist == 'L2':
                dist = np.sum(diff * diff, axis=0, dtype=f32)
            else:
                dist = np.sum(np.abs(diff), axis=0, dtype=f32)
            cost[d0:d1] = cost[d0:d1] + dist / sigma
    return cost[None]


def log_softmax(x, axis):
    """F.log_softmax: x - max - log(sum(exp(x - max)))."""
    x = np.asarray(x, f32)
    m = np.max(x, axis=axis, keepdims=True)
    z = x - m
    s = np.sum(np.exp(z.astype(np.float64)), axis=axis, keepdims=True)
    return (z - np.log(s).astype(f32)).astype(f32)


def d_net_dpv_from_cost(costV):
    """models/basic.py:299-300 (BV_log=True): BV = log_softmax(-costV, dim=1)."""
    return log_softmax(-np.asarray(costV, f32), axis=1)


# --------------------------------------------------------------------------
# a7: image warp to volume
# --------------------------------------------------------------------------
def warp_img_feats_v3(feat_img_src, d_candi, R, t, cam_intrinsic):
    """warping/homography.py:234-280 (list branch :252-262): for each view returns
    [c, D, h, w] (transpose of the D x c x h x w parallel warp)."""
    d32 = np.asarray(d_candi).astype(f32)
    K = cam_intrinsic['intrinsic_M_cuda']
    rays = cam_intrinsic['unit_ray_array_2D']
    cx, cy = cam_intrinsic['intrinsic_M'][0, 2], cam_intrinsic['intrinsic_M'][1, 2]
    outs = []
    for v, img in enumerate(feat_img_src):
        img = np.asarray(img, f32)[0]
        c, h, w = img.shape
        term1, _, term2 = homography_terms(K, R[v], t[v], rays)
        gx, gy = back_warp_grid(term1, term2, d32, cx, cy)
        outs.append(grid_sample_2d_zeros(img, gx.reshape(-1, h, w), gy.reshape(-1, h, w)))
    return outs


def warp_img_feats_mgpu(feat_img_src, d_candi, R, t, IntM_tensors, unit_ray_arrays_2D):
    """warping/homography.py:183-232: same warp; intrinsics arrive as stacked
    tensors (1x3x3, 1x3xhw) and u/v centre come from IntM[0,2], IntM[1,2] (fp32)."""
    K = np.asarray(IntM_tensors, f32).reshape(3, 3)
    cam = {'intrinsic_M_cuda': K, 'unit_ray_array_2D': np.asarray(unit_ray_arrays_2D, f32).reshape(3, -1),
           'intrinsic_M': K}
    return warp_img_feats_v3(feat_img_src, d_candi, R, t, cam)


# --------------------------------------------------------------------------
# a12: DPV re-projection (3-D resample)
# --------------------------------------------------------------------------
def set_vol_border(vol, border_val):
    """warping/homography.py:873-887 on a [D,H,W] volume (clone, 6 face fills)."""
    v = np.array(vol, f32, copy=True)
    b = f32(border_val)
    v[0, :, :] = b; v[:, 0, :] = b; v[:, :, 0] = b
    v[-1, :, :] = b; v[:, -1, :] = b; v[:, :, -1] = b
    return v


def _clip(x, hi):
    """ATen clip_coordinates: min(hi, max(x, 0))."""
    return np.minimum(f32(hi), np.maximum(x, f32(0)))


def grid_sample_3d_border(vol, gx, gy, gz):
    """F.grid_sample 5-D, mode='bilinear' (trilinear), padding_mode='border',
    align_corners=False.  vol [D,H,W]; gx,gy,gz same shape S -> out S."""
    D, H, W = vol.shape
    ix = _clip(_unnormalize(gx, W), W - 1)
    iy = _clip(_unnormalize(gy, H), H - 1)
    iz = _clip(_unnormalize(gz, D), D - 1)
    # NaN coordinates (0/0 in the projection) propagate through min/max in ATen
    # as NaN -> floor(NaN) -> index garbage; ATen's within-bounds test then drops
    # every corner, giving 0.  Mirror that.
    bad = np.isnan(ix) | np.isnan(iy) | np.isnan(iz)
    ix = np.where(bad, f32(0), ix); iy = np.where(bad, f32(0), iy); iz = np.where(bad, f32(0), iz)
    x0 = np.floor(ix); y0 = np.floor(iy); z0 = np.floor(iz)
    x1 = x0 + f32(1); y1 = y0 + f32(1); z1 = z0 + f32(1)
    flat = vol.reshape(-1)
    out = np.zeros(ix.shape, f32)
    first = True
    # ATen order: tnw, tne, tsw, tse, bnw, bne, bsw, bse  (t = z0, b = z1)
    for (zz, wz) in ((z0, z1 - iz), (z1, iz - z0)):
        for (yy, wy) in ((y0, y1 - iy), (y1, iy - y0)):
            for (xx, wx) in ((x0, x1 - ix), (x1, ix - x0)):
                wgt = (wx * wy) * wz
                xi = xx.astype(np.int64); yi = yy.astype(np.int64); zi = zz.astype(np.int64)
                ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H) & (zi >= 0) & (zi < D)
                idx = np.where(ok, (zi * H + yi) * W + xi, 0)
                val = np.where(ok, flat[idx], f32(0))
                out = out + (val * wgt).astype(f32)
    return np.where(bad, f32(0), out).astype(f32)


def resample_vol_cuda(src_vol, rel_extM, cam_intrinsic=None, d_candi=None, d_candi_new=None,
                      padding_value=0.):
    """warping/homography.py:654-723.  src_vol [1,D,H,W]; rel_extM 4x4 (fp32).
    Returns [D,H,W].  Note the bug-for-bug details restated:
      * points = d * unit_ray (float32(d) * float32(ray)), homogeneous 1      :673-697
      * z range from the point cloud z (= d, since ray_z = 1) unless d_candi_new :686-693
      * x/(z+1e-10)/tan(hhfov), y likewise, (z - z_half)/z_radius               :703-705
      * then all four rows divided by (row3 + 1e-10)                              :708
      * faces overwritten with padding_value before border-clamped sampling       :713-716
    """
    assert d_candi is not None, 'd_candi should be some np.array object'
    vol = np.asarray(src_vol, f32)
    _, D, H, W = vol.shape
    hhfov = math.radians(cam_intrinsic['hfov']) * .5
    hvfov = math.radians(cam_intrinsic['vfov']) * .5
    d_ = d_candi_new if d_candi_new is not None else d_candi
    rays = np.asarray(cam_intrinsic['unit_ray_array']).astype(f32)            # FloatTensor(unit_ray_array)
    # `d * FloatTensor` : python/np float64 scalar times fp32 tensor -> fp32 multiply by f32(d)
    pts = np.stack([f32(d) * rays for d in d_]).astype(f32)                    # [D,H,W,3]
    if d_candi_new is not None:
        z_max, z_min = f32(np.max(d_candi)), f32(np.min(d_candi))
        z_half = f32((np.max(d_candi) + np.min(d_candi)) * .5)
        z_radius = f32((np.max(d_candi) - np.min(d_candi)) * .5)
    else:
        z_max = pts[..., 2].max(); z_min = pts[..., 2].min()
        z_half = (z_max + z_min) * f32(.5)
        z_radius = (z_max - z_min) * f32(.5)
    E = np.asarray(rel_extM, f32)
    X = pts[..., 0].reshape(-1); Y = pts[..., 1].reshape(-1); Z = pts[..., 2].reshape(-1)
    one = f32(1)

    def row(i):
        acc = (E[i, 0] * X).astype(f32)
        acc = _fma(E[i, 1], Y, acc)
        acc = _fma(E[i, 2], Z, acc)
        return _fma(E[i, 3], one, acc)
    xs, ys, zs, ws = row(0), row(1), row(2), row(3)
    den = zs + f32(1e-10)
    tx = f32(math.tan(hhfov)); ty = f32(math.tan(hvfov))
    gx = xs / den / tx
    gy = ys / den / ty
    gz = (zs - z_half) / z_radius
    wden = ws + f32(1e-10)
    gx = gx / wden; gy = gy / wden; gz = gz / wden
    volb = set_vol_border(vol[0], padding_value)
    out = grid_sample_3d_border(volb, gx.astype(f32), gy.astype(f32), gz.astype(f32))
    return out.reshape(D, H, W)


def propagate_dpv(kv_dpv, rel_Rt_inv, cam_intrinsic, d_candi):
    """test_utils/test_KVNet.py:54-59: resample with padding log(1/D), clamp to
    [-1000, 0], add the batch dim -> [1,D,h,w]."""
    D = len(d_candi)
    r = resample_vol_cuda(kv_dpv, rel_Rt_inv, cam_intrinsic, d_candi,
                          padding_value=math.log(1. / float(D)))
    return np.clip(r, f32(-1000.), f32(0))[None]


# --------------------------------------------------------------------------
# a9 / a11: Bayesian update, expected depth, NaN sentinel
# --------------------------------------------------------------------------
def bayes_update(BV_gain, BV_predict):
    """models/KVNET.py:172-173: DPV = log_softmax(squeeze(gain,1) + BV_predict, dim=1)."""
    g = np.asarray(BV_gain, f32)
    if g.ndim == 5:
        g = g[:, 0]
    return log_softmax(g + np.asarray(BV_predict, f32), axis=1)


def depth_val_regression(BV_measure, d_candi_cur, BV_log=True):
    """mutils/misc.py:532-548: sequential sum over d of exp(BV[0,d]) * d (fp32;
    `tensor * d` multiplies by float32(d))."""
    BV = np.asarray(BV_measure, f32)
    assert len(d_candi_cur) == BV.shape[1]
    acc = np.zeros((1,) + BV.shape[2:], f32)
    for i, d in enumerate(d_candi_cur):
        p = np.exp(BV[0, i]) if BV_log else BV[0, i]
        acc = acc + p.astype(f32) * f32(d)
    return acc


def valid_dpv(dpv_in):
    """mutils/misc.py:100-115: a DPV is invalid when its first element is NaN."""
    if dpv_in is None:
        return False
    a = np.asarray(dpv_in)
    if a.ndim not in (2, 3, 4, 5):
        raise Exception('wrong dimension for input dpv !')
    return not bool(np.isnan(a.reshape(-1)[0]))


# --------------------------------------------------------------------------
# f-1: backward of the plane-sweep cost volume (what autograd does to
# est_swp_volume_v4 in train_utils/train_KVNet.py:149-153: grid_sample backward =
# scatter-add of the bilinear weights, then the distance; R, t, d, K are constants)
# --------------------------------------------------------------------------
def _corners(gx, gy, W, H):
    """Corner indices / validity / weights of grid_sample_2d_zeros for grids of any shape."""
    ix = _unnormalize(gx.astype(f32), W)
    iy = _unnormalize(gy.astype(f32), H)
    x0f = np.floor(ix); y0f = np.floor(iy)
    x1f = x0f + f32(1); y1f = y0f + f32(1)
    ws = [(x1f - ix) * (y1f - iy), (ix - x0f) * (y1f - iy), (x1f - ix) * (iy - y0f), (ix - x0f) * (iy - y0f)]
    bad = ~np.isfinite(ix) | ~np.isfinite(iy) | (np.abs(ix) > 1e9) | (np.abs(iy) > 1e9)
    x0 = np.where(bad, -10, x0f).astype(np.int64); y0 = np.where(bad, -10, y0f).astype(np.int64)
    out = []
    for (xx, yy), wgt in zip(((x0, y0), (x0 + 1, y0), (x0, y0 + 1), (x0 + 1, y0 + 1)), ws):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H) & ~bad
        out.append((np.where(ok, yy * W + xx, 0), np.where(ok, wgt, f32(0)).astype(f32)))
    return out


def est_swp_volume_v4_backward(grad_cost, feat_img_ref, feat_img_src, d_candi, R, t, cam_intrinsic,
                               costV_sigma, feat_dist='L2'):
    """Gradients of est_swp_volume_v4 w.r.t. feat_img_ref [1,C,h,w] and feat_img_src [1,V,C,h,w]
    for grad_cost [1,D,h,w]. Accumulated in float64 (the device scatter has no defined order)."""
    if feat_dist not in ('L2', 'L1'):
        raise Exception('undefined metric for feature distance ...')
    ref = np.asarray(feat_img_ref, f32)[0]
    src = np.asarray(feat_img_src, f32)[0]
    g = np.asarray(grad_cost, np.float64)[0]
    C, h, w = ref.shape
    V = src.shape[0]
    d32 = np.asarray(d_candi).astype(f32)
    D = len(d32)
    K = cam_intrinsic['intrinsic_M_cuda']
    rays = cam_intrinsic['unit_ray_array_2D']
    cx, cy = cam_intrinsic['intrinsic_M'][0, 2], cam_intrinsic['intrinsic_M'][1, 2]
    gref = np.zeros((C, h * w), np.float64)
    gsrc = np.zeros((V, C, h * w), np.float64)
    sigma = float(f32(costV_sigma))
    for v in range(V):
        term1, _, term2 = homography_terms(K, R[v], t[v], rays)
        gx, gy = back_warp_grid(term1, term2, d32, cx, cy)
        flat = src[v].reshape(C, h * w)
        for d in range(D):
            cor = _corners(gx[d].reshape(h, w), gy[d].reshape(h, w), w, h)
            s = np.zeros((C, h, w), f32)
            for idx, wgt in cor:
                s = s + flat[:, idx.reshape(-1)].reshape(C, h, w) * wgt[None]
            diff = (s - ref).astype(np.float64)
            dphi = 2.0 * diff if feat_dist == 'L2' else np.sign(diff)
            gs = dphi * (g[d] / sigma)[None]                 # dLoss/dS  [C,h,w]
            gref -= gs.reshape(C, -1)
            for idx, wgt in cor:
                contrib = (gs * wgt[None].astype(np.float64)).reshape(C, -1)
                for c in range(C):
                    np.add.at(gsrc[v, c], idx.reshape(-1), contrib[c])
    return gref.reshape(1, C, h, w).astype(f32), gsrc.reshape(1, V, C, h, w).astype(f32)


# --------------------------------------------------------------------------
# f-3 (oracle first; the device kernels are the next round's work): depth-map back-warp used by the
# local bundle adjustment, warping/homography.py:479-574 (back_warp_th_Rt, back_warp_th_Rt_msrc), and
# its pose gradients (ICP/opt_pose_numerical.py differentiates the warped image w.r.t. R, t through the
# sampling grid: F.grid_sample backward w.r.t. the grid, the perspective division, the two matmuls).
# --------------------------------------------------------------------------
def _warp_points(dmap, R, t, cam_intrinsic):
    """Pixel grid of the source view for every reference pixel: X = d * ray, P = K (R X + t), u = P_x/P_z, v = P_y/P_z,
    normalised (u - cx)/cx, (v - cy)/cy (:546-567). Returns gx, gy [H*W] and the intermediates the backward needs."""
    K = np.asarray(cam_intrinsic['intrinsic_M_cuda'], f32)
    rays = np.asarray(cam_intrinsic['unit_ray_array_2D'], f32)
    d = np.asarray(dmap, f32).reshape(1, -1)
    X = (d * rays).astype(f32)                                        # [3, n]
    Xc = (np.asarray(R, f32) @ X + np.asarray(t, f32).reshape(3, 1)).astype(f32)
    P = (K @ Xc).astype(f32)
    u = (P[0] / P[2]).astype(f32); v = (P[1] / P[2]).astype(f32)
    cx, cy = K[0, 2], K[1, 2]
    return ((u - cx) / cx).astype(f32), ((v - cy) / cy).astype(f32), X, Xc, P


def back_warp_th_Rt(img_src, dmap, R, t, cam_intrinsic):
    """warping/homography.py:531-574: img_src [1,C,H,W] sampled at the projection of the reference depth map."""
    img = np.asarray(img_src, f32)[0]
    C, H, W = img.shape
    gx, gy, _, _, _ = _warp_points(dmap, R, t, cam_intrinsic)
    return grid_sample_2d_zeros(img, gx.reshape(H, W), gy.reshape(H, W))[None]


def back_warp_th_Rt_msrc(imgs_src, dmap, Rs, ts, cam_intrinsic):
    """warping/homography.py:479-529: the same warp for N source frames [N,C,H,W] with their own poses."""
    return np.concatenate([back_warp_th_Rt(imgs_src[i:i + 1], dmap, Rs[i], ts[i], cam_intrinsic) for i in range(len(imgs_src))], 0)


def back_warp_th_Rt_backward(grad_out, img_src, dmap, R, t, cam_intrinsic):
    """Gradients of back_warp_th_Rt w.r.t. R [3,3], t [3] and img_src [1,C,H,W] for grad_out [1,C,H,W] (float64 sums).
    grid_sample backward w.r.t. the grid (ATen grid_sampler_2d_backward, bilinear / zeros / align_corners=False):
    d out/d ix = sum over the 4 corners of value * d weight/d ix, times size/2 from the un-normalisation."""
    img = np.asarray(img_src, f32)[0]
    g = np.asarray(grad_out, np.float64)[0]
    C, H, W = img.shape
    K = np.asarray(cam_intrinsic['intrinsic_M_cuda'], np.float64)
    gx, gy, X, Xc, P = _warp_points(dmap, R, t, cam_intrinsic)
    ix = _unnormalize(gx, W).astype(np.float64); iy = _unnormalize(gy, H).astype(np.float64)
    x0 = np.floor(ix); y0 = np.floor(iy)
    flat = img.reshape(C, -1).astype(np.float64)
    g_flat = g.reshape(C, -1)
    g_img = np.zeros((C, H * W), np.float64)
    g_ix = np.zeros(H * W); g_iy = np.zeros(H * W)
    for dx, dy in ((0, 0), (1, 0), (0, 1), (1, 1)):
        xx = x0 + dx; yy = y0 + dy
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H) & np.isfinite(ix) & np.isfinite(iy)
        idx = np.where(ok, yy * W + xx, 0).astype(np.int64)
        wx = (1 - np.abs(ix - xx)); wy = (1 - np.abs(iy - yy))              # bilinear weights of this corner
        val = np.where(ok[None], flat[:, idx], 0.0)
        gv = (g_flat * val).sum(0)                                          # sum_c grad * value
        g_ix += np.where(ok, gv * wy * (1.0 if dx else -1.0), 0.0)
        g_iy += np.where(ok, gv * wx * (1.0 if dy else -1.0), 0.0)
        contrib = g_flat * np.where(ok, wx * wy, 0.0)[None]
        for c in range(C):
            np.add.at(g_img[c], idx, contrib[c])
    cx, cy = K[0, 2], K[1, 2]
    g_u = g_ix * (W / 2.0) / cx                                             # ix = ((u-cx)/cx + 1) * W/2 - 1/2
    g_v = g_iy * (H / 2.0) / cy
    Pd = P.astype(np.float64)
    g_P = np.stack([g_u / Pd[2], g_v / Pd[2], -(g_u * Pd[0] + g_v * Pd[1]) / (Pd[2] * Pd[2])])   # [3, n]
    g_Xc = K.T @ g_P
    g_R = g_Xc @ X.astype(np.float64).T
    g_t = g_Xc.sum(1)
    return g_R.astype(f32), g_t.astype(f32), g_img.reshape(1, C, H, W).astype(f32)
