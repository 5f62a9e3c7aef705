"""Per-CTA timeline of conv_h2_kernel (development; GPU box): clock64 stamps of every CTA of one launch.
usage: h2_timeline.py N D H W Cin Cout [k=3] [pad=1] [dil=1] [flags=0]
Prints the mean / p90 of: setup, first operands landed, main loop (issue), drain, epilogue, teardown, lifetime, and the
cycles the producer / issuer spent waiting on each barrier class; plus tiles per SM and the launch's wall cycles."""
import ctypes
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from neuralrgbd_b200 import _lib, convops       # noqa: E402
from neuralrgbd_b200._lib import ptr, check     # noqa: E402

a = [int(v) for v in sys.argv[1:]]
N, D, H, W, Cin, Cout = a[:6]
k = a[6] if len(a) > 6 else 3
pad = a[7] if len(a) > 7 else 1
dil = a[8] if len(a) > 8 else 1
flags = a[9] if len(a) > 9 else 0
dev = torch.device('cuda:0'); L = _lib.dev_lib()
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)     # noqa: E731
kd = 3 if D > 1 else 1
Cs = convops.pad_to(Cin, 32)
x = torch.randn((N, D, H, W, Cs), device=dev)
w = torch.randn((Cout, Cin) + ((kd,) if D > 1 else ()) + (k, k), device=dev) / np.sqrt(Cin * k * k * kd)
wp, cin_p, cout_p, bn = convops.pack_weight_h2(w)
hi, lo = convops.split_f16_pair(x)
Ho = (H + 2 * pad - dil * (k - 1) - 1) + 1; Wo = (W + 2 * pad - dil * (k - 1) - 1) + 1
y = torch.zeros((N, D, Ho, Wo, convops.pad4(Cout)), device=dev)
stats = torch.zeros((2, Cout), device=dev, dtype=torch.float64)
tiles = N * D * ((Ho + 15) // 16) * ((Wo + 7) // 8)
ny = (cout_p + bn - 1) // bn
L.nrgbd_dev_conv_h2_set_flags(flags)


def run():
    check(L.nrgbd_conv_nhwc_h2(ptr(hi), ptr(lo), N, D, H, W, cin_p, Cs, ptr(wp), None, Cout, cout_p, bn, kd, k, k, 1, pad, dil, ptr(y), Ho, Wo,
                               y.shape[-1], 0, 0, ctypes.c_void_p(stats.data_ptr()), st()))


for _ in range(3):
    run()
n_cta = min(ny * tiles, torch.cuda.get_device_properties(0).multi_processor_count)
dbg = torch.zeros((n_cta, 16), device=dev, dtype=torch.int64)
L.nrgbd_dev_conv_h2_set_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
run(); torch.cuda.synchronize()
L.nrgbd_dev_conv_h2_set_debug_buffer(None)
t = dbg.cpu().numpy().astype(np.float64)
msub = 4 if bn <= 32 else 2 if bn <= 64 else 1
G = 3 if (k * k) % 3 == 0 else 1
items = ny * N * D * ((Ho + 16 * msub - 1) // (16 * msub)) * ((Wo + 7) // 8) / n_cta
steps = kd * (k * k // G) * (cin_p // 32)
seg = {'setup': t[:, 1] - t[:, 0], 'lifetime': t[:, 6] - t[:, 0], 'issuer_done_at': t[:, 3] - t[:, 0], 'producer_done_at': t[:, 8] - t[:, 0],
       'epilogue_done_at': t[:, 5] - t[:, 0],
       'issuer_main_loops_total': t[:, 14], 'issuer_wait_a_full': t[:, 11], 'issuer_wait_b_full': t[:, 12], 'issuer_wait_acc_free': t[:, 13],
       'producer_wait_a_empty': t[:, 9], 'producer_wait_b_empty': t[:, 10], 'epilogue_wait_acc_full': t[:, 4], 'epilogue_busy': t[:, 15]}
out = {kk: {'mean': float(v.mean()), 'p90': float(np.percentile(v, 90))} for kk, v in seg.items()}
out['items_per_cta'] = items; out['steps_per_item'] = steps; out['BN'] = bn; out['mma_per_step'] = 4 * msub * G
out['cycles_per_step_in_main_loop'] = float(t[:, 14].mean() / (items * steps))
out['lifetime_per_item'] = float((t[:, 6] - t[:, 0]).mean() / items)
print(json.dumps({'shape': a, 'segments': out}, indent=1))
