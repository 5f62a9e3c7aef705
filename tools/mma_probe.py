"""Raw tcgen05.mma (kind::tf32, M=128) rate probe: cycles per MMA for several N and destination patterns."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralrgbd_b200 import _lib
from neuralrgbd_b200._lib import check
L = _lib.dev_lib(); dev = torch.device('cuda:0')
out = torch.zeros(2, dtype=torch.int64, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
n = 1200
print('N   pattern                 ctas  cycles/MMA(total) cycles/MMA(issue)  tensor floor')
for BN in (32, 64, 128, 256):
    for (pat, nd, grp, two, name) in ((0, 1, 1, 0, 'same D'), (1, 2, 1, 0, 'alternate 2 D'), (2, 2, 4, 0, 'groups of 4, 2 D'),
                                      (2, 2, 12, 0, 'groups of 12, 2 D'), (1, 4, 1, 0, 'rotate 4 D'), (0, 1, 1, 1, 'same D, 2 warps'),
                                      (3, 1, 1, 0, 'lean, same D'), (3, 2, 1, 0, 'lean, lo/lo/main 2 D'), (3, 2, 1, 1, 'lean, 2 D, 2 warps')):
        if (2 if two else 1) * nd * BN > 512:
            continue
        for ctas in (1, 148):
            for _ in range(2):
                check(L.nrgbd_mma_probe(BN, n, pat, nd, grp, two, ctas, ctypes.c_void_p(out.data_ptr()), st))
            torch.cuda.synchronize()
            t = out.cpu().numpy()
            print('%-3d %-22s %4d  %8.1f %17.1f %12.1f' % (BN, name, ctas, t[0] / n, t[1] / n, 128 * BN * 8 / 2048.0))

print('TMEM read-back: cycles per load (4 warps, 128 lanes)')
for pat, name in ((10, 'x16, wait each'), (13, 'x16, one wait'), (11, 'x32, wait each'), (14, 'x32, one wait')):
    for _ in range(2):
        check(L.nrgbd_mma_probe(64, 256, pat, 1, 1, 0, 1, ctypes.c_void_p(out.data_ptr()), st))
    torch.cuda.synchronize()
    print('  %-16s %8.1f' % (name, out.cpu().numpy()[0] / 256.0))

print('queue depth: issue vs completion cycles of n back-to-back MMAs, lean stream, 1 CTA')
print('BN  form  n_mma   issue_cycles  total_cycles')
for BN in (64, 128):
    for pat, name in ((3, 'SS'), (4, 'TS')):
        for n in (12, 24, 48, 96, 384):
            for _ in range(2):
                check(L.nrgbd_mma_probe(BN, n, pat, 2, 1, 0, 1, ctypes.c_void_p(out.data_ptr()), st))
            torch.cuda.synchronize()
            t = out.cpu().numpy()
            print('%-3d %-4s %5d %12d %12d' % (BN, name, n, t[1], t[0]))
