"""Steady-state streaming throughput (640x480, D=64, V=4, D-Net + K-Net + R-Net x2 + DPV propagation) through
the public mirror `test_utils.test_KVNet.test` - the loop of the reference's test_KVNet.py:241-262. Development aid;
the judged metric (first-window C2) is bench.py. Frames resident on the device, CUDA events, 1 video.
usage: bench_stream.py [n_frames]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralrgbd_b200 import arch, camera, synth                      # noqa: E402
from neuralrgbd_b200.models.KVNET import KVNET                       # noqa: E402
from neuralrgbd_b200.test_utils import test_KVNet as T               # noqa: E402

H, W, D, V = 480, 640, 64, 4
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16          # synth.video shifts 6 px per frame inside a 64 px margin: <= 17 frames + 4
dev = torch.device('cuda:0')
cam = camera.make_cam_intrinsics(585., 585., 320., 240., [W // 4, H // 4])
d_candi = synth.d_candidates(D)
model = KVNET(feature_dim=64, cam_intrinsics=cam, d_candi=d_candi, sigma_soft_max=10., KVNet_feature_dim=64,
              d_upsample_ratio_KV_net=None, t_win_r=2, if_refined=True)
model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in arch.synth_state_dict(5, 64, D, 2, 64).items()})
model = model.to(dev)
frames, rng = synth.video(7, n + 4, H, W)
exts = synth.camera_track(rng, n + 4)
fr = [{'img': torch.from_numpy(f[None]).to(dev), 'extM': e} for f, e in zip(frames, exts)]


def step(i, bv):
    poses, idx = synth.window_rel_poses(exts, 2 + i, 2)
    src_poses = torch.from_numpy(poses[None]).to(dev)
    nxt = torch.from_numpy((exts[3 + i] @ np.linalg.inv(exts[2 + i])).astype(np.float32)).to(dev)
    return T.test(model, d_candi, [cam], 2, [fr[2 + i]], [[fr[j] for j in idx]], src_poses, bv, cam_pose_next=nxt, R_net=True)


bv = None
for i in range(4):                      # first window + steady-state warm-up (graph capture)
    _, bv = step(i, bv)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(4, n):
    dmap, bv = step(i, bv)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / (n - 4)
print(json.dumps({'workload': 'C3 steady state 640x480 D=64 V=4 (D-Net + K-Net + 2x R-Net + propagation), 1 stream, resident frames',
                  'frames': n - 4, 'ms_per_frame': ms, 'frames_per_s': 1e3 / ms, 'algorithmic_tflop_per_frame': 3.73,
                  'algorithmic_tflops': 3.73 / (ms * 1e-3), 'finite': bool(torch.isfinite(dmap).all())}))
