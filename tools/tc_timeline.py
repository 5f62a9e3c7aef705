"""Per-CTA clock64 timeline of the tcgen05 conv kernels (development aid).

usage: tc_timeline.py [Cin] [Cout] [v1|v2] [dev_flags]
"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralrgbd_b200 import _lib, convops
from neuralrgbd_b200._lib import ptr, check
dev = torch.device('cuda:0'); L = _lib.dev_lib()
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
N, H, W, Cin, Cout, k = 5, 120, 160, int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 64, 3
x = torch.randn((N, 1, H, W, Cin), device=dev); xh = torch.empty_like(x); xl = torch.empty_like(x)
w = torch.randn((Cout, Cin, k, k), device=dev) / np.sqrt(Cin * 9)
wh, wl = convops.pack_weight_tc(w)
y = torch.zeros((N, 1, H, W, Cout), device=dev); stats = torch.zeros((2, Cout), device=dev, dtype=torch.float64)
check(L.nrgbd_split_tf32(ptr(x), x.numel(), ptr(xh), ptr(xl), st()))
grid = N * (H // 8) * (W // 16)
dbg = torch.zeros((grid, 64), device=dev, dtype=torch.int64)
impl = sys.argv[3] if len(sys.argv) > 3 else 'v1'
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
L.nrgbd_conv_tc_set_dev(int(sys.argv[5]) if len(sys.argv) > 5 else 0, flags)
def run():
    if impl == 'v2':
        check(L.nrgbd_conv_nhwc_tc2(ptr(x), N, 1, H, W, Cin, Cin, ptr(wh), ptr(wl), None, Cout, Cout, 1, k, k, 1, 1, 1, ptr(y), H, W, Cout, 0, 0,
                                    ctypes.c_void_p(stats.data_ptr()), st()))
        return
    check(L.nrgbd_conv_nhwc_tc(ptr(xh), ptr(xl), N, 1, H, W, Cin, Cin, ptr(wh), ptr(wl), None, Cout, Cout, 1, k, k, 1, 1, 1, ptr(y), H, W, Cout, 0, 0,
                               ctypes.c_void_p(stats.data_ptr()), st()))
for _ in range(3): run()
L.nrgbd_conv_tc_set_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
torch.cuda.synchronize(); run(); torch.cuda.synchronize()
L.nrgbd_conv_tc_set_debug_buffer(None)
d = dbg.cpu().numpy()
if impl == 'v2':
    names = ['setup(alloc+sync)', 'converter loop', 'mma drain', 'epilogue ld/st', 'stats+teardown',
             '  chunk1: tmem loads', '  chunk1: math', '  chunk1: global store', '  chunk1: smem stage']
    seg = np.stack([d[:, 1] - d[:, 0], d[:, 3] - d[:, 1], d[:, 4] - d[:, 3], d[:, 5] - d[:, 4], d[:, 6] - d[:, 5],
                    d[:, 9] - d[:, 8], d[:, 10] - d[:, 9], d[:, 11] - d[:, 10], d[:, 12] - d[:, 11]], 1)
else:
    names = ['setup(alloc+sync)', 'first operands', 'mainloop issue', 'mma drain', 'epilogue ld/st', 'stats+teardown']
    seg = np.stack([d[:, 1] - d[:, 0], d[:, 2] - d[:, 1], d[:, 3] - d[:, 2], d[:, 4] - d[:, 3], d[:, 5] - d[:, 4], d[:, 6] - d[:, 5]], 1)
print('grid', grid, 'median cycles per segment:')
for n_, v in zip(names, np.median(seg, 0)): print('  %-20s %8.0f' % (n_, v))
print('  total per CTA        %8.0f' % np.median(d[:, 6] - d[:, 0]))
# per-SM schedule: gap between consecutive CTAs on one SM
sm = d[:, 7]; gaps = []
for s in np.unique(sm):
    idx = np.where(sm == s)[0]; o = idx[np.argsort(d[idx, 0])]
    gaps += list(d[o[1:], 0] - d[o[:-1], 6])
print('  CTAs per SM', np.bincount(sm.astype(int)).max(), 'median gap between CTAs on an SM', np.median(gaps))

if impl == 'v2':
    # K-step trace of the TMA lane, one converter thread (group 0) and the MMA issue lane, K-steps 8..13,
    # relative to the converter's "raw tile landed" stamp of K-step 8; median over CTAs
    base = d[:, 16:17]
    tr = np.median(d[:, 16:64].reshape(-1, 6, 8) - base[:, None, :], 0)
    print('  ks | tma:A-issue tma:B-issue | cvt:landed  cvt:math  cvt:aempty  cvt:sttm+arrive | mma:start  mma:issued')
    for i in range(6):
        print('  %2d | %10.0f %11.0f | %10.0f %9.0f %11.0f %16.0f | %9.0f %11.0f' % ((8 + i, tr[i][6], tr[i][7]) + tuple(tr[i][:6])))
