"""Where does conv_tc2 differ from the fp32 SIMT conv? (development aid)
usage: tc2_diag.py N D H W Cin Cout [flags] [stages]"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralrgbd_b200 import _lib, convops
from neuralrgbd_b200._lib import ptr, check
dev = torch.device('cuda:0'); L = _lib.dev_lib()
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
N, D, H, W, Cin, Cout = [int(a) for a in sys.argv[1:7]]
flags = int(sys.argv[7]) if len(sys.argv) > 7 else 0
stages = int(sys.argv[8]) if len(sys.argv) > 8 else 0
L.nrgbd_conv_tc_set_dev(stages, flags)
k, kd = 3, (3 if D > 1 else 1)
torch.manual_seed(0)
Cs = convops.pad_to(Cin, 32)
x = torch.randn((N, D, H, W, Cs), device=dev)
w = torch.randn((Cout, Cin) + ((kd,) if D > 1 else ()) + (k, k), device=dev) / np.sqrt(Cin * k * k * kd)
wp = convops.pack_weight(w); wh, wl = convops.pack_weight_tc(w)
Cso = convops.pad4(Cout)
y = torch.zeros((N, D, H, W, Cso), device=dev); y3 = torch.full_like(y, 7.0)
check(L.nrgbd_conv_nhwc(ptr(x), N, D, H, W, convops.pad4(Cin), Cs, ptr(wp), None, Cout, convops.pad4(Cout), kd, k, k, 1, 1, 1,
                        ptr(y), H, W, Cso, 0, 0, None, st()))
for it in range(3):
    y3.fill_(7.0)
    check(L.nrgbd_conv_nhwc_tc2(ptr(x), N, D, H, W, Cs, Cs, ptr(wh), ptr(wl), None, Cout, convops.pad_to(Cout, 16), kd, k, k,
                                1, 1, 1, ptr(y3), H, W, Cso, 0, 0, None, st()))
    torch.cuda.synchronize()
    e = (y - y3).abs()[..., :Cout]
    bad = e > 1e-3 * float(y.abs().max())
    print('iter', it, 'max err', float(e.max()), 'bad elements', int(bad.sum()), 'of', bad.numel(), 'untouched(7.0)', int((y3[..., :Cout] == 7.0).sum()))
    if bad.any():
        idx = bad.nonzero().cpu().numpy()
        n_, z_, yy, xx, cc = idx.T
        print('  bad n', np.unique(n_)[:10], 'z', np.unique(z_)[:10])
        print('  bad rows (y) count by y%8:', np.bincount(yy % 8, minlength=8), ' by x%16:', np.bincount(xx % 16, minlength=16))
        print('  bad channels by c//16:', np.bincount(cc // 16))
        ty, tx = yy // 8, xx // 16
        tiles = np.unique(np.stack([n_, z_, ty, tx], 1), axis=0)
        print('  bad tiles:', len(tiles), 'of', N * D * ((H + 7) // 8) * ((W + 15) // 16), 'first', tiles[:8].tolist())
        lin = ((tiles[:, 0] * D + tiles[:, 1]) * ((H + 7) // 8) + tiles[:, 2]) * ((W + 15) // 16) + tiles[:, 3]
        print('  bad tile linear ids (first 20):', lin[:20].tolist(), ' min', lin.min(), 'max', lin.max())
        # sample values
        j = idx[0]; print('  sample', j.tolist(), 'ref', float(y[tuple(j)]), 'got', float(y3[tuple(j)]))
