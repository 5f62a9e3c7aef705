"""Seeded synthetic inputs for the plane-sweep DPV path (numpy only).

Shapes/statistics follow SURVEY.md §8(d): 7Scenes intrinsics (fx=fy=585, 640x480,
DSO/cam_info_7scenes.mat) pushed through the reference's ScanNet recipe
(mdataloader/scanNet.py:239-270) at quarter resolution, d_candi = linspace(0.1, 5, D),
V source views at [t-2, t-1, t+1, t+2] with ~3 cm / <=2 deg relative motion given as
E_src . E_ref^-1 (warping/homography.py:904-906), smooth textured images at the
post-ImageNet-normalisation scale (mdataloader/m_preprocess.py:12-21).
"""
import math
import numpy as np


def smooth_image(rng, c, h, w, n_waves=12, noise=0.05):
    """Sum of random low-frequency sinusoids + white noise, roughly unit variance."""
    ys = np.arange(h, dtype=np.float64)[:, None] / h
    xs = np.arange(w, dtype=np.float64)[None, :] / w
    img = np.zeros((c, h, w), np.float64)
    for ch in range(c):
        for _ in range(n_waves):
            fx, fy = rng.uniform(0.5, 9.0, 2)
            ph = rng.uniform(0, 2 * math.pi)
            amp = rng.uniform(0.3, 1.0)
            img[ch] += amp * np.sin(2 * math.pi * (fx * xs + fy * ys) + ph)
    img /= math.sqrt(n_waves * 0.25)
    img += noise * rng.standard_normal((c, h, w))
    return img.astype(np.float32)


def rot_from_axis_angle(a):
    th = float(np.linalg.norm(a))
    if th < 1e-12:
        return np.eye(3)
    k = a / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + math.sin(th) * Kx + (1 - math.cos(th)) * (Kx @ Kx)


def camera_track(rng, n_frames, step_t=0.03, step_rot_deg=0.7):
    """World->camera 4x4 extrinsics of a smooth hand-held-like trajectory (float64)."""
    exts = []
    E = np.eye(4)
    for _ in range(n_frames):
        exts.append(E.copy())
        dR = rot_from_axis_angle(np.radians(step_rot_deg) * rng.uniform(-1, 1, 3))
        dt = step_t * (np.array([1.0, 0.2, 0.1]) + 0.3 * rng.standard_normal(3))
        S = np.eye(4); S[:3, :3] = dR; S[:3, 3] = dt
        E = S @ E
    return exts


def window_rel_poses(exts, ref_idx, t_win_r):
    """Relative poses E_src . E_ref^-1 for src order [t-r..t-1, t+1..t+r]
    (mutils/misc.py:509-517, warping/homography.py:904-906)."""
    idxs = [i for i in range(ref_idx - t_win_r, ref_idx + t_win_r + 1) if i != ref_idx]
    inv_ref = np.linalg.inv(exts[ref_idx])
    return np.stack([exts[i].dot(inv_ref) for i in idxs]).astype(np.float32), idxs


def d_candidates(D, d_min=0.1, d_max=5.0):
    """test_KVNet.py:76: np.linspace(d_min, d_max, ndepth) (float64)."""
    return np.linspace(d_min, d_max, D)


def video(seed, n_frames, H, W, parallax=0.0):
    """n_frames images [3,H,W]; consecutive frames are shifted copies of one smooth
    texture plus per-frame noise so that costs have structure."""
    rng = np.random.RandomState(seed)
    pad = 64 if n_frames <= 20 else 8 + 6 * (n_frames // 2 + 1)      # long streams need a wider texture margin
    base = smooth_image(rng, 3, H + 2 * pad, W + 2 * pad)
    frames = []
    for i in range(n_frames):
        ox = pad + int(round(6 * (i - n_frames // 2)))
        oy = pad + int(round(2 * (i - n_frames // 2)))
        f = base[:, oy:oy + H, ox:ox + W] + 0.02 * rng.standard_normal((3, H, W)).astype(np.float32)
        frames.append(np.ascontiguousarray(f, dtype=np.float32))
    return frames, rng
