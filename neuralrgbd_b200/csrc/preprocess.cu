// Input stage: PIL nearest resize + torchvision ToTensor + Normalize of one RGB frame, on the device.
//
// Replaces, per frame, mdataloader/scanNet.py:368-369 (PIL.Image.resize(img_size, NEAREST)) and
// mdataloader/m_preprocess.py:15-21 (transforms.ToTensor(), transforms.Normalize(imagenet stats)), which run on
// the host and ship a float NCHW tensor (3.7 MB at 640x480) through pageable memory; here the decoded uint8 HWC
// frame is uploaded once (0.9 MB, or 3.8 MB for a raw 1296x968 ScanNet frame) and converted in one pass:
//   out[c][y][x] = ((float)src[ys[y]][xs[x]][c] / 255 - mean[c]) / std[c]
// with every operation rounded separately in fp32 (ToTensor's .div(255), Normalize's .sub_(mean).div_(std)): bit-exact.
// ys / xs are PIL's nearest-neighbour source indices (Geometry.c ImagingScaleAffine: xo = 0.5 * scale, then
// xo += scale accumulated in double, truncated) computed on the host once per (source, target) size.
#include "common.cuh"
#include "../../include/nrgbd.h"

namespace {

__global__ void __launch_bounds__(256)
preprocess_rgb_u8_kernel(const unsigned char* __restrict__ src, int Ws, const int* __restrict__ ys, const int* __restrict__ xs, int H, int W,
                         float m0, float m1, float m2, float s0, float s1, float s2, float* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W) return;
  const unsigned char* p = src + ((size_t)__ldg(ys + y) * Ws + __ldg(xs + x)) * 3;
  const size_t hw = (size_t)H * W, o = (size_t)y * W + x;
  dst[o] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)p[0], 255.f), m0), s0);
  dst[hw + o] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)p[1], 255.f), m1), s1);
  dst[2 * hw + o] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)p[2], 255.f), m2), s2);
}

}  // namespace

extern "C" int nrgbd_preprocess_rgb_u8(const unsigned char* src_hwc, int Hs, int Ws, const int* ys, const int* xs, int H, int W,
                                       const float* mean3, const float* std3, float* dst_chw, nrgbd_stream_t st) {
  NRGBD_REQUIRE(src_hwc && ys && xs && mean3 && std3 && dst_chw, "null pointer");
  NRGBD_REQUIRE(Hs >= 1 && Ws >= 1 && H >= 1 && W >= 1 && H <= 65535, "bad image extent");
  NRGBD_REQUIRE(std3[0] != 0.f && std3[1] != 0.f && std3[2] != 0.f, "zero std");
  dim3 grid(ceil_div(W, 256), H);
  preprocess_rgb_u8_kernel<<<grid, 256, 0, (cudaStream_t)st>>>(src_hwc, Ws, ys, xs, H, W, mean3[0], mean3[1], mean3[2], std3[0], std3[1],
                                                               std3[2], dst_chw);
  NRGBD_COUNT(1);
  NRGBD_LAUNCH_CHECK();
  return NRGBD_OK;
}
