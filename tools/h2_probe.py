"""Which shared-memory descriptor semantics does the halo tile need? (development; run on the GPU box)

Runs a few convolutions of tests/test_gpu_conv_h2.py under the variant flags of csrc/conv_f16.cu
(include/nrgbd_dev.h: nrgbd_dev_conv_h2_set_flags) and prints the relative error of each against torch fp32:
  0  halo, pitch = TW + 2 pad (SBO not a multiple of the swizzle atom), no base offset     <- the design
  1  same + base-offset field     2  halo pitch 16      3  pitch 16 + base offset      4  one box per tap (no halo)
  8  64-channel chunks / 128-byte rows (with 0, 2, 4)
"""
import json
import math
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, '.')
from neuralrgbd_b200 import _lib, convops       # noqa: E402

import subprocess

if len(sys.argv) < 2:          # driver: one subprocess per variant (a trapped kernel poisons its CUDA context)
    res = {}
    for flags in (4, 0, 1, 2, 3, 12, 8, 10):
        r = subprocess.run([sys.executable, __file__, str(flags)], capture_output=True, text=True, timeout=300)
        line = [l for l in r.stdout.splitlines() if l.startswith('{')]
        res[str(flags)] = json.loads(line[-1]) if line else {'error': (r.stderr or r.stdout)[-300:]}
        print('flags', flags, res[str(flags)], flush=True)
    json.dump(res, open('gpurun_out/h2_probe.json', 'w'), indent=1)
    sys.exit(0)

flags = int(sys.argv[1])
L = _lib.dev_lib()
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator(device='cuda').manual_seed(3)
cases = [dict(N=1, Cin=32, Cout=64, H=16, W=32, k=3, p=1, d=1), dict(N=2, Cin=64, Cout=64, H=30, W=40, k=3, p=1, d=1),
         dict(N=1, Cin=128, Cout=128, H=20, W=28, k=3, p=2, d=2), dict(N=1, Cin=64, Cout=64, H=16, W=8, k=1, p=0, d=1)]
row = []
for c in cases:
    x = torch.randn((c['N'], c['Cin'], c['H'], c['W']), device='cuda', generator=g)
    w = torch.randn((c['Cout'], c['Cin'], c['k'], c['k']), device='cuda', generator=g) / math.sqrt(c['Cin'] * c['k'] ** 2)
    ref = F.conv2d(x, w, None, 1, c['p'], c['d'])
    L.nrgbd_dev_conv_h2_set_flags(flags)
    try:
        y = convops.conv_h2(x, w, None, 1, c['p'], c['d'])
        torch.cuda.synchronize()
        row.append(float((y - ref).abs().max() / ref.abs().max()))
    except Exception as e:          # noqa: BLE001
        row.append(str(e)[:80])
        break
    L.nrgbd_dev_conv_h2_set_flags(0)
print(json.dumps({'rel_err': row}))
