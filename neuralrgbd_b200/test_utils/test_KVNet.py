"""Inference step of the streaming loop, behind the reference's name and signature:
`test_utils.test_KVNet.test` (/root/reference/code/test_utils/test_KVNet.py:19-67).

What one call does (reference line numbers):
  1. window tensors from the frame dicts: reference image [B,3,H,W], sources [B,V,3,H,W]   (:30-33)
  2. one KVNET forward under no_grad with the predicted prior                               (:35-40)
  3. first frame of a trajectory (prior None): the measured DPV stands in for the filtered  (:42-44)
  4. prior for the NEXT frame: the filtered low-resolution DPV resampled into the next camera
     (pose = inverse of the given next pose, or of the window's (t_win_r)-th source pose), faces padded
     with log(1/D), result clamped to [-1000, 0]                                            (:46-62)
  5. returns (refined full-resolution DPV if R_net else filtered DPV, stacked priors)       (:64-67)
Here step 4 is one fused device kernel per batch entry (resample + pad + clamp), the frames are moved with
`.cuda()` only if they are not on a device yet, and the model may be the engine-backed KVNET or that module
wrapped in nn.DataParallel - both expose the same keyword interface as the reference's.
"""
import math

import numpy as np
import torch

from ..warping import homography as warp_homo


def _on_device(t):
    return t if t.is_cuda else t.cuda()


def _stack_window(Ref_Dats, Src_Dats):
    """Frame dicts -> (ref [B,3,H,W], src [B,V,3,H,W])."""
    ref = torch.cat([_on_device(d['img']) for d in Ref_Dats], dim=0)
    per_traj = [torch.cat([_on_device(f['img']) for f in traj], dim=0) for traj in Src_Dats]
    return ref, torch.stack(per_traj, dim=0)


def _next_prior(dpv_lowres, pose_to_next, cam_intrinsic, d_candi):
    """Step 4 for one batch entry: [D,h,w] log-DPV -> [1,D,h,w] prior in the next camera."""
    uniform_log = math.log(1.0 / float(len(d_candi)))
    # Tensor.inverse() checks its LU status on the host, i.e. it waits for everything queued on the stream - the forward that
    # was just launched; inv_ex returns the same inverse without the host round trip, so the step stays asynchronous
    inv = torch.linalg.inv_ex(pose_to_next).inverse if pose_to_next.is_cuda else pose_to_next.inverse()
    moved = warp_homo.resample_vol_cuda(src_vol=dpv_lowres.unsqueeze(0), rel_extM=inv,
                                        cam_intrinsic=cam_intrinsic, d_candi=d_candi, padding_value=uniform_log,
                                        clamp=(-1000., 0.))
    return moved.unsqueeze(0)


def test(model_KV, d_candi, Cam_Intrinsics, t_win_r, Ref_Dats, Src_Dats, Src_CamPoses, BV_predict,
         cam_pose_next=None, R_net=False, Cam_Intrinsics_imgsize=None, ref_indx=None):
    """One depth frame of the stream. Arguments and return values as the reference (see module docstring);
    `Cam_Intrinsics_imgsize` and `ref_indx` are accepted and unused, as there."""
    ref_frame, src_frames = _stack_window(Ref_Dats, Src_Dats)
    batch_ids = torch.FloatTensor(np.arange(1))          # nGPU is fixed to 1 for testing (:27-28)
    with torch.no_grad():
        refined_cur, refined_filtered, dpv_measured, dpv_filtered = model_KV(
            ref_frame=ref_frame, src_frames=src_frames, src_cam_poses=Src_CamPoses, BatchIdx=batch_ids,
            cam_intrinsics=Cam_Intrinsics, BV_predict=BV_predict)
    first_of_trajectory = BV_predict is None
    if first_of_trajectory:
        dpv_filtered, refined_filtered = dpv_measured, refined_cur

    priors = []
    for b in range(dpv_measured.shape[0]):
        pose = cam_pose_next if cam_pose_next is not None else Src_CamPoses[b, t_win_r, :, :]
        priors.append(_next_prior(dpv_filtered[b], pose, Cam_Intrinsics[b], d_candi))
    BVs_predict = torch.cat(priors, dim=0)
    return (refined_filtered if R_net else dpv_filtered), BVs_predict
