// Output stage: expected depth + confidence of a log-DPV, scaled and truncated to
// uint16 on the device, replacing test_utils/export_res.py:43-75 which copies the whole D x H x W volume
// to the host to do the same (78.6 MB per 640x480x64 frame against 1.2 MB of maps).
//
//   dmap[p]  = sum_d exp(BV[d][p]) * d_candi[d]        (export_res.py:37-41, :49-53; mutils/misc.py:532-548)
//   conf[p]  = exp(max_d BV[d][p])                     (export_res.py:56-59, :88-90)
//   u16      = (uint16)(map * scale)                   (export_res.py:74-75, numpy astype: truncation)
//
// Layout: BV is the reference's [D][H*W] plane-major volume, read once, coalesced across pixels
// (algorithmic bytes = D*HW*4 in, <= 12*HW out); one thread per pixel, two independent partial sums.
#include <cuda_runtime.h>

#include "common.cuh"
#include "../../include/nrgbd.h"

namespace {

__device__ __forceinline__ unsigned short to_u16(float x) {
  // numpy float32 -> uint16 cast truncates toward zero; out-of-range input is undefined there, clamped here
  x = fminf(fmaxf(x, 0.f), 65535.f);
  return (unsigned short)(int)x;
}

__global__ void __launch_bounds__(256)
export_depth_conf_kernel(const float* __restrict__ bv, const float* __restrict__ d_candi, int D, long long HW,
                         float depth_scale, float conf_scale, float* __restrict__ dmap, float* __restrict__ conf,
                         unsigned short* __restrict__ dmap_u16, unsigned short* __restrict__ conf_u16) {
  extern __shared__ float dc[];
  for (int i = threadIdx.x; i < D; i += blockDim.x) dc[i] = d_candi[i];
  __syncthreads();
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  // products are rounded before they are added (exp(BV) * Depth_val_vol is materialised in the reference)
  float acc = 0.f, m = -INFINITY;
  int d = 0;
  for (; d + 8 <= D; d += 8) {                 // eight independent loads in flight per thread (HBM latency)
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldg(bv + (long long)(d + j) * HW + p);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc = __fadd_rn(acc, __fmul_rn(expf(v[j]), dc[d + j]));
      m = fmaxf(m, v[j]);
    }
  }
  for (; d < D; ++d) {
    const float v = __ldg(bv + (long long)d * HW + p);
    acc = __fadd_rn(acc, __fmul_rn(expf(v), dc[d]));
    m = fmaxf(m, v);
  }
  const float c = expf(m);
  if (dmap) dmap[p] = acc;
  if (conf) conf[p] = c;
  if (dmap_u16) dmap_u16[p] = to_u16(__fmul_rn(acc, depth_scale));
  if (conf_u16) conf_u16[p] = to_u16(__fmul_rn(c, conf_scale));
}

}  // namespace

extern "C" int nrgbd_export_depth_conf(const float* log_dpv, const float* d_candi, int D, long long HW, float depth_scale,
                                       float conf_scale, float* dmap, float* conf, unsigned short* dmap_u16,
                                       unsigned short* conf_u16, nrgbd_stream_t st) {
  NRGBD_REQUIRE(log_dpv && d_candi, "null pointer");
  NRGBD_REQUIRE(D >= 1 && D <= 4096 && HW >= 1, "bad volume extent");
  NRGBD_REQUIRE(dmap || conf || dmap_u16 || conf_u16, "no output requested");
  const long long blocks = (HW + 255) / 256;
  export_depth_conf_kernel<<<(unsigned)blocks, 256, D * sizeof(float), (cudaStream_t)st>>>(log_dpv, d_candi, D, HW, depth_scale, conf_scale,
                                                                                        dmap, conf, dmap_u16, conf_u16);
  NRGBD_COUNT(1);
  NRGBD_LAUNCH_CHECK();
  return NRGBD_OK;
}
