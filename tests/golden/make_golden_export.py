"""Golden vectors for the output stage: runs the UNMODIFIED reference export_res_img on CPU.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_export.py
matplotlib is not installed here; export_res.py imports it only to write preview PNGs, so a stub module with a
no-op imsave stands in for it (the .pgm files, which are what is pinned, are written by PIL). The .cuda() shim is
the one of make_golden.py. Inputs come from tests/cases.export_case() (seeded); the reference's d_*.pgm / conf_*.pgm
bytes and its float maps are stored in tests/golden/export_outputs.npz, and the oracle's deviations from them in
tests/golden/PINNING_export.json.
"""
import json
import os
import sys
import tempfile
import types
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
_mpl = types.ModuleType('matplotlib'); _mpl.use = lambda *a, **k: None
_plt = types.ModuleType('matplotlib.pyplot'); _plt.imsave = lambda *a, **k: None
_mpl.pyplot = _plt
sys.modules.setdefault('matplotlib', _mpl); sys.modules.setdefault('matplotlib.pyplot', _plt)

import test_utils.export_res as ref_export           # noqa: E402  (reference)

from oracle import export_oracle as E                # noqa: E402
from tests import cases                              # noqa: E402


def main():
    out, pin = {}, {'torch': torch.__version__, 'numpy': np.__version__, 'cases': {}}
    for name in cases.EXPORT_CASES:
        bv, d_candi, img = cases.export_case(name)
        with tempfile.TemporaryDirectory() as tmp:
            ref_export.export_res_img({'img': torch.from_numpy(img)}, torch.from_numpy(bv), d_candi, tmp, 7)
            d_bytes = open(os.path.join(tmp, 'd_00007.pgm'), 'rb').read()
            c_bytes = open(os.path.join(tmp, 'conf_00007.pgm'), 'rb').read()
        D, H, W = bv.shape[1:]
        vol = torch.ones(1, D, H, W)
        for i in range(D):
            vol[0, i] = vol[0, i] * d_candi[i]
        dmap = ref_export.depth_regression(vol, torch.from_numpy(bv)).astype(np.float32)
        conf = torch.exp(torch.max(torch.from_numpy(bv), dim=1)[0]).squeeze().numpy()
        out[name + '/d_pgm'] = np.frombuffer(d_bytes, np.uint8)
        out[name + '/conf_pgm'] = np.frombuffer(c_bytes, np.uint8)
        out[name + '/dmap'] = dmap
        out[name + '/conf'] = conf
        # oracle against the reference
        od, oc, od16, oc16 = E.export_maps(bv[0], d_candi)
        hdr = b'P5\n%d %d\n65535\n' % (W, H)
        assert d_bytes.startswith(hdr) and c_bytes.startswith(hdr), 'unexpected PGM header from PIL'
        rd16 = np.frombuffer(d_bytes[len(hdr):], '>u2').reshape(H, W).astype(np.int64)
        rc16 = np.frombuffer(c_bytes[len(hdr):], '>u2').reshape(H, W).astype(np.int64)
        pin['cases'][name] = {
            'dmap_max_rel': float(np.max(np.abs(od - dmap) / np.maximum(np.abs(dmap), 1e-6))),
            'conf_max_rel': float(np.max(np.abs(oc - conf) / np.maximum(conf, 1e-6))),
            'dmap_u16_max_lsb': int(np.max(np.abs(od16.astype(np.int64) - rd16))),
            'conf_u16_max_lsb': int(np.max(np.abs(oc16.astype(np.int64) - rc16))),
            'dmap_u16_mismatch_frac': float(np.mean(od16.astype(np.int64) != rd16)),
            'pgm_bytes_equal_given_same_u16': bool(E.pgm16_bytes(rd16.astype(np.uint16)) == d_bytes),
        }
    np.savez_compressed(os.path.join(HERE, 'export_outputs.npz'), **out)
    with open(os.path.join(HERE, 'PINNING_export.json'), 'w') as f:
        json.dump(pin, f, indent=1, sort_keys=True)
    print(json.dumps(pin, indent=1))


if __name__ == '__main__':
    main()
