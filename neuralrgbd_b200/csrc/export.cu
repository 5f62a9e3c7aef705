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

// Four consecutive pixels per thread (one 16-byte load per plane, eight planes in flight: 128 bytes per thread outstanding,
// ~64 KB per SM); the per-pixel sums stay sequential over the planes, as in the reference's loop. VEC = 1 handles a ragged tail /
// unaligned volumes.
template <int VEC>
__global__ void __launch_bounds__(256)
export_depth_conf_kernel(const float* __restrict__ bv, const float* __restrict__ d_candi, int D, long long HW,
                         float depth_scale, float conf_scale, float* __restrict__ dmap, float* __restrict__ conf,
                         unsigned short* __restrict__ dmap_u16, unsigned short* __restrict__ conf_u16, long long p_begin) {
  extern __shared__ float dc[];
  for (int i = threadIdx.x; i < D; i += blockDim.x) dc[i] = d_candi[i];
  __syncthreads();
  const long long p = p_begin + ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (p + VEC > HW) return;
  // products are rounded before they are added (exp(BV) * Depth_val_vol is materialised in the reference)
  float acc[VEC], m[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) { acc[k] = 0.f; m[k] = -INFINITY; }
  // software pipeline: the loads of plane batch b + 1 are in flight while batch b is reduced (the block is a single wave:
  // its run time is the sum of its batches' latencies, so they must overlap); sums stay sequential over the planes
  constexpr int B = 8;
  float cur[B][VEC], nxt[B][VEC];
  auto load_batch = [&](float (&dst)[B][VEC], int d0) {
#pragma unroll
    for (int j = 0; j < B; ++j) {
      if (d0 + j < D) {
        if (VEC == 4) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(bv + (long long)(d0 + j) * HW + p));
          dst[j][0] = q.x; dst[j][1 % VEC] = q.y; dst[j][2 % VEC] = q.z; dst[j][3 % VEC] = q.w;
        } else {
          dst[j][0] = __ldg(bv + (long long)(d0 + j) * HW + p);
        }
      }
    }
  };
  load_batch(cur, 0);
  for (int d = 0; d < D; d += B) {
    if (d + B < D) load_batch(nxt, d + B);
#pragma unroll
    for (int j = 0; j < B; ++j)
      if (d + j < D) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          acc[k] = __fadd_rn(acc[k], __fmul_rn(expf(cur[j][k]), dc[d + j]));
          m[k] = fmaxf(m[k], cur[j][k]);
        }
      }
#pragma unroll
    for (int j = 0; j < B; ++j)
#pragma unroll
      for (int k = 0; k < VEC; ++k) cur[j][k] = nxt[j][k];
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const float c = expf(m[k]);
    if (dmap) dmap[p + k] = acc[k];
    if (conf) conf[p + k] = c;
    if (dmap_u16) dmap_u16[p + k] = to_u16(__fmul_rn(acc[k], depth_scale));
    if (conf_u16) conf_u16[p + k] = to_u16(__fmul_rn(c, conf_scale));
  }
}

}  // namespace

extern "C" int nrgbd_export_depth_conf(const float* log_dpv, const float* d_candi, int D, long long HW, float depth_scale,
                                       float conf_scale, float* dmap, float* conf, unsigned short* dmap_u16,
                                       unsigned short* conf_u16, nrgbd_stream_t st) {
  NRGBD_REQUIRE(log_dpv && d_candi, "null pointer");
  NRGBD_REQUIRE(D >= 1 && D <= 4096 && HW >= 1, "bad volume extent");
  NRGBD_REQUIRE(dmap || conf || dmap_u16 || conf_u16, "no output requested");
  // vector body (4 pixels per thread) when the planes are 16-byte aligned, scalar kernel for the tail / unaligned volumes
  long long body = 0;
  if (HW % 4 == 0 && ((uintptr_t)log_dpv & 15) == 0) body = HW;
  int launches = 0;
  if (body > 0) {
    const long long blocks = (body / 4 + 255) / 256;
    export_depth_conf_kernel<4><<<(unsigned)blocks, 256, D * sizeof(float), (cudaStream_t)st>>>(log_dpv, d_candi, D, HW, depth_scale, conf_scale,
                                                                                             dmap, conf, dmap_u16, conf_u16, 0);
    ++launches;
  }
  if (body < HW) {
    const long long blocks = (HW - body + 255) / 256;
    export_depth_conf_kernel<1><<<(unsigned)blocks, 256, D * sizeof(float), (cudaStream_t)st>>>(log_dpv, d_candi, D, HW, depth_scale, conf_scale,
                                                                                             dmap, conf, dmap_u16, conf_u16, body);
    ++launches;
  }
  NRGBD_COUNT(launches);
  NRGBD_LAUNCH_CHECK();
  return NRGBD_OK;
}
