"""Golden vectors for the input stage: the reference's own transform (mdataloader.m_preprocess.get_transform, i.e.
torchvision ToTensor + Normalize) applied after PIL's NEAREST resize, exactly as mdataloader/scanNet.py:368-369,438 do.

Run in the build container only (needs /root/reference, Pillow, torchvision):
    python tests/golden/make_golden_preprocess.py
Writes tests/golden/preprocess_outputs.npz and the oracle's deviations (expected: 0) to PINNING_preprocess.json.
"""
import json
import os
import sys

import numpy as np
import PIL
import PIL.Image
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/code')

import mdataloader.m_preprocess as ref_pre           # noqa: E402  (reference)

from oracle import preprocess_oracle as P            # noqa: E402
from tests import cases                              # noqa: E402


def main():
    out, pin = {}, {'pillow': PIL.__version__, 'torchvision': torchvision.__version__, 'torch': torch.__version__, 'cases': {}}
    tf = ref_pre.get_transform()
    for name in cases.PREPROCESS_CASES:
        img, size = cases.preprocess_case(name)
        pil = PIL.Image.fromarray(img)
        if size is not None:
            pil = pil.resize(size, PIL.Image.NEAREST)              # scanNet.py:369
        ref = tf(pil).unsqueeze_(0).numpy()                        # scanNet.py:438, :446 ('img': img.unsqueeze_(0))
        out[name] = ref
        mine = P.preprocess(img, size)
        pin['cases'][name] = {'bit_identical': bool(np.array_equal(mine, ref)), 'maxabs': float(np.abs(mine - ref).max())}
    np.savez_compressed(os.path.join(HERE, 'preprocess_outputs.npz'), **out)
    with open(os.path.join(HERE, 'PINNING_preprocess.json'), 'w') as f:
        json.dump(pin, f, indent=1, sort_keys=True)
    print(json.dumps(pin, indent=1))


if __name__ == '__main__':
    main()
