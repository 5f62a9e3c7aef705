"""a14 fixture: the reference's OWN intrinsics recipe and ray table, run unmodified.

    python tests/golden/make_golden_camera.py      (build container only: needs /root/reference)

Writes a ScanNet-style `_info.txt` for each calibration, calls mdataloader/scanNet.py:read_IntM_from_txt
(:204-272, which calls warping/View.py:32-62 normalised_pixel_to_ray_array) and stores what it returned
(hfov, vfov, focal_length, intrinsic_M, intrinsic_M_cuda, unit_ray_array_2D) in camera_outputs.npz.
tests/test_oracle_camera.py holds neuralrgbd_b200.camera and oracle.make_cam_intrinsics - the function every
other fixture's inputs are built with - to these values bit for bit.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/code')

import mdataloader.scanNet as ref_scannet            # noqa: E402  (reference)
from tests import cases                              # noqa: E402


def info_txt(width, height, fx, fy, cx, cy):
    """The 8 lines read_IntM_from_txt parses (SensReader's _info.txt layout)."""
    m = [fx, 0, cx, 0, 0, fy, cy, 0, 0, 0, 1, 0, 0, 0, 0, 1]
    return '\n'.join([
        'm_versionNumber = 4', 'm_sensorName = synthetic', 'm_colorWidth = %d' % width, 'm_colorHeight = %d' % height,
        'm_depthWidth = %d' % width, 'm_depthHeight = %d' % height, 'm_depthShift = 1000',
        'm_calibrationColorIntrinsic = ' + ' '.join(repr(float(v)) for v in m)]) + '\n'


def main():
    out = {}
    for name, c in cases.CAMERA_CASES.items():
        with tempfile.NamedTemporaryFile('w', suffix='_info.txt', delete=False) as f:
            f.write(info_txt(c['width'], c['height'], c['fx'], c['fy'], c['cx'], c['cy']))
            path = f.name
        cam = ref_scannet.read_IntM_from_txt(path, out_size=c['out_size'])
        os.unlink(path)
        out[name + '/scalars'] = np.array([cam['hfov'], cam['vfov'], cam['focal_length']], np.float64)
        out[name + '/intrinsic_M'] = np.asarray(cam['intrinsic_M'], np.float64)
        out[name + '/intrinsic_M_cuda'] = cam['intrinsic_M_cuda'].numpy()
        out[name + '/unit_ray_array_2D'] = cam['unit_ray_array_2D'].numpy()
        ura = np.asarray(cam['unit_ray_array'], np.float64)
        out[name + '/unit_ray_array_stats'] = np.array([ura.sum(), np.square(ura).sum(), ura[0, 0, 0], ura[-1, -1, 1]])
        print(name, out[name + '/scalars'])
    np.savez_compressed(os.path.join(HERE, 'camera_outputs.npz'), **out)


if __name__ == '__main__':
    main()
