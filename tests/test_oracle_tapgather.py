"""CPU check of the re-association behind the engine's single-output-channel convolution: for K-Net's last
layer, Conv3d(f -> 1, k3, pad 1) (/root/reference/code/models/basic.py:136-137),
    out[p] = sum_t sum_c x[p + off_t][c] w[0][c][t]  ==  sum_t Q[p + off_t][t],   Q = pointwise conv of x with w viewed as [t][c].
The oracle's conv3d is the reference statement; Q and the shifted tap sum are restated here in numpy exactly as
nrgbd_conv_nhwc_h2 (1x1x1, one output channel per tap) + nrgbd_tap_gather_sum compute them (zero padding = skipped taps).
The GPU counterpart is tests/test_gpu_conv_h2.py::test_conv3d_single_output_channel_tap_gather_vs_oracle."""
import numpy as np

from oracle import kvnet_oracle as N


def tap_gather(Q, kd):
    """Q [D][H][W][taps] -> out [D][H][W]: out[d,h,w] = sum_t Q[d+tz-kd//2, h+ty-1, w+tx-1, t], zero outside."""
    D, H, W, _ = Q.shape
    out = np.zeros((D, H, W), np.float64)
    pd = kd // 2
    for tz in range(kd):
        for ty in range(3):
            for tx in range(3):
                t = (tz * 3 + ty) * 3 + tx
                d0, d1 = max(0, pd - tz), min(D, D + pd - tz)
                h0, h1 = max(0, 1 - ty), min(H, H + 1 - ty)
                w0, w1 = max(0, 1 - tx), min(W, W + 1 - tx)
                out[d0:d1, h0:h1, w0:w1] += Q[d0 + tz - pd:d1 + tz - pd, h0 + ty - 1:h1 + ty - 1, w0 + tx - 1:w1 + tx - 1, t]
    return out


def test_single_channel_conv3d_equals_pointwise_plus_tap_gather():
    rng = np.random.RandomState(3)
    for cin, shape in ((8, (4, 6, 7)), (5, (1, 5, 9)), (16, (3, 3, 3))):
        x = rng.standard_normal((1, cin) + shape).astype(np.float32)
        w = rng.standard_normal((1, cin, 3, 3, 3)).astype(np.float32)
        ref = N.conv3d(x, w)[0, 0]
        wt = w[0].reshape(cin, 27).astype(np.float64)                        # [c][t]: the transposed-kind source layout
        Q = np.einsum('cdhw,ct->dhwt', x[0].astype(np.float64), wt)
        out = tap_gather(Q, 3)
        assert out.shape == ref.shape
        assert np.abs(out - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
