// Backward of the fused plane-sweep cost volume.
//
// The reference trains through est_swp_volume_v4 (train_utils/train_KVNet.py:149-153 backpropagates the
// D-Net loss through warping/homography.py:293-331) with autograd over the materialised D x C x h x w
// intermediates: grid_sample backward (a scatter-add of the four bilinear weights) and the distance.
// Poses, intrinsics and plane depths are constants there (no gradient), so only
//     cost[d,p] = sum_v sum_c phi(S_v(d,p)[c] - ref[c,p]) / sigma,   S_v = bilinear(src_v) at the homography,
// needs its two feature gradients, with g = dLoss/dcost:
//     dref[c,p]   = - sum_d g[d,p] sum_v phi'(S_v[c] - ref[c,p]) / sigma
//     dsrc_v[c,q] =   sum_{d,p} g[d,p] phi'(S_v(d,p)[c] - ref[c,p]) / sigma * w_q(d,p)   (q: the 4 corners)
// phi' = 2x (L2, homography.py:81-83) or sign(x) (L1, :85-87; torch.abs backward: sign, 0 at 0).
//
// Same data layout and coordinate arithmetic as the forward kernel (sweep.cu): packed wide [hw][Cw] + narrow
// [hw][4] features, the homography re-evaluated in registers with the pinned rounding order, so the corners and
// weights are the forward's bit for bit. A lane-group owns a reference pixel: dref accumulates in registers and
// is written once; dsrc is scattered with 16-byte vector atomics (red.global.add.v4.f32, sm_90+), one per corner
// and channel quad, skipped when the corner weight is zero (zeros padding). Like ATen's grid_sampler backward
// the accumulation order into dsrc is not deterministic.
#include "common.cuh"

namespace {

template <bool L1>
__device__ __forceinline__ float dphi(float x) {
  if (L1) return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f);
  return 2.f * x;
}

__device__ __forceinline__ float4 bilerp4b(float4 a, float4 b, float4 c, float4 d, const float* w) {
  float4 r;
  r.x = fmaf(d.x, w[3], fmaf(c.x, w[2], fmaf(b.x, w[1], a.x * w[0])));
  r.y = fmaf(d.y, w[3], fmaf(c.y, w[2], fmaf(b.y, w[1], a.y * w[0])));
  r.z = fmaf(d.z, w[3], fmaf(c.z, w[2], fmaf(b.z, w[1], a.z * w[0])));
  r.w = fmaf(d.w, w[3], fmaf(c.w, w[2], fmaf(b.w, w[1], a.w * w[0])));
  return r;
}

__device__ __forceinline__ void red_add4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Quad index q in [0, Q): q < G -> wide channels 4q..4q+3, q == G -> the narrow quad. LANES lanes share a pixel and
// take quads lane, lane + LANES, ... (at most QPL each).
template <int LANES, int QPL, bool L1>
__global__ void __launch_bounds__(256)
plane_sweep_backward_kernel(const float4* __restrict__ ref_w, const float4* __restrict__ src_w, int G,
                            const float4* __restrict__ ref_n, const float4* __restrict__ src_n,
                            const float* __restrict__ t1, const float* __restrict__ KR, const float* __restrict__ rays,
                            const float* __restrict__ dpl, int V, int D, int w, int h, float cx, float cy, float sigma,
                            const float* __restrict__ gcost /*[hw][D]*/, float4* __restrict__ gref_w, float4* __restrict__ gref_n,
                            float* __restrict__ gsrc_w, float* __restrict__ gsrc_n) {
  constexpr int GROUPS_PER_BLOCK = 256 / LANES;
  const int hw = w * h;
  const int lane = threadIdx.x % LANES;
  const int pix = blockIdx.x * GROUPS_PER_BLOCK + threadIdx.x / LANES;
  if (pix >= hw) return;
  const int Q = G + (ref_n ? 1 : 0);
  const float Wf = (float)w, Hf = (float)h;
  const float r0 = rays[pix], r1 = rays[hw + pix], r2 = rays[2 * hw + pix];
  float4 refq[QPL], gr[QPL];
#pragma unroll
  for (int i = 0; i < QPL; ++i) {
    const int q = lane + i * LANES;
    refq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    gr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < G) refq[i] = __ldg(ref_w + (size_t)pix * G + q);
    else if (q == G && ref_n) refq[i] = __ldg(ref_n + pix);
  }
  for (int v = 0; v < V; ++v) {
    const float* kr = KR + v * 9;
    const float t2x = dot3_chain(kr[0], kr[1], kr[2], r0, r1, r2);
    const float t2y = dot3_chain(kr[3], kr[4], kr[5], r0, r1, r2);
    const float t2z = dot3_chain(kr[6], kr[7], kr[8], r0, r1, r2);
    const float4* sw = src_w ? src_w + (size_t)v * hw * G : nullptr;
    const float4* sn = src_n ? src_n + (size_t)v * hw : nullptr;
    float* gw = gsrc_w ? gsrc_w + (size_t)v * hw * G * 4 : nullptr;
    float* gn = gsrc_n ? gsrc_n + (size_t)v * hw * 4 : nullptr;
    for (int d = 0; d < D; ++d) {
      const float g = __ldg(gcost + (size_t)pix * D + d);
      if (g == 0.f) continue;                               // uniform across the lane-group
      float ix, iy;
      plane_project(t1[v * 3], t1[v * 3 + 1], t1[v * 3 + 2], t2x, t2y, t2z, __ldg(dpl + d), cx, cy, Wf, Hf, ix, iy);
      const Tap2D tp = make_tap2d(ix, iy, w, h);
      const float wt[4] = {tp.w_nw, tp.w_ne, tp.w_sw, tp.w_se};
      const int off[4] = {tp.o_nw, tp.o_ne, tp.o_sw, tp.o_se};
      const float coef = g / sigma;
#pragma unroll
      for (int i = 0; i < QPL; ++i) {
        const int q = lane + i * LANES;
        if (q >= Q) continue;
        const bool wide = q < G;
        const float4* base = wide ? sw + q : sn;
        const size_t stride = wide ? (size_t)G : 1;
        const float4 a = __ldg(base + off[0] * stride), b = __ldg(base + off[1] * stride);
        const float4 c = __ldg(base + off[2] * stride), e = __ldg(base + off[3] * stride);
        const float4 s = bilerp4b(a, b, c, e, wt);
        float4 gs;                                          // dLoss / dS
        gs.x = coef * dphi<L1>(s.x - refq[i].x); gs.y = coef * dphi<L1>(s.y - refq[i].y);
        gs.z = coef * dphi<L1>(s.z - refq[i].z); gs.w = coef * dphi<L1>(s.w - refq[i].w);
        gr[i].x -= gs.x; gr[i].y -= gs.y; gr[i].z -= gs.z; gr[i].w -= gs.w;
        float* gb = wide ? gw + (size_t)q * 4 : gn;
        const size_t gstride = wide ? (size_t)G * 4 : 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (wt[k] != 0.f) red_add4(gb + off[k] * gstride, make_float4(gs.x * wt[k], gs.y * wt[k], gs.z * wt[k], gs.w * wt[k]));
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < QPL; ++i) {
    const int q = lane + i * LANES;
    if (q < G) gref_w[(size_t)pix * G + q] = gr[i];
    else if (q == G && gref_n) gref_n[pix] = gr[i];
  }
}

// wide [hw][Cw] + narrow [hw][4] -> NCHW [C][hw]   (inverse of pack_features_kernel)
__global__ void unpack_features_kernel(const float* __restrict__ wide, const float* __restrict__ narrow, int C, int hw, int Cw,
                                       float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (p < hw && c < C) v = c < Cw ? wide[(size_t)p * Cw + c] : narrow[(size_t)p * 4 + (c - Cw)];
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (p < hw && c < C) out[(size_t)c * hw + p] = tile[threadIdx.x][i];
  }
}

__global__ void sweep_setup_bwd_kernel(const float* __restrict__ K, const float* __restrict__ R, const float* __restrict__ t, int V,
                                       float* __restrict__ t1, float* __restrict__ KR) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const float* Rv = R + v * 9;
  const float* tv = t + v * 3;
  for (int i = 0; i < 3; ++i) {
    t1[v * 3 + i] = dot3_chain(K[i * 3 + 0], K[i * 3 + 1], K[i * 3 + 2], tv[0], tv[1], tv[2]);
    for (int j = 0; j < 3; ++j)
      KR[v * 9 + i * 3 + j] = dot3_chain(K[i * 3 + 0], K[i * 3 + 1], K[i * 3 + 2], Rv[j], Rv[3 + j], Rv[6 + j]);
  }
}

}  // namespace

extern "C" {

// Inverse of nrgbd_pack_features: n_img images of wide [hw][Cw] (+ narrow [hw][4]) -> NCHW.
int nrgbd_unpack_features(const float* wide, const float* narrow, int C, int hw, int n_img, float* nchw, cudaStream_t st) {
  NRGBD_REQUIRE(nchw && C > 0 && hw > 0 && n_img > 0, "bad arguments");
  const int Cn = C % 4, Cw = C - Cn;
  NRGBD_REQUIRE((Cw == 0 || wide) && (Cn == 0 || narrow), "missing input buffer");
  dim3 blk(32, 8), grid(ceil_div(hw, 32), ceil_div(C, 32));
  for (int n = 0; n < n_img; ++n)
    unpack_features_kernel<<<grid, blk, 0, st>>>(Cw ? wide + (size_t)n * hw * Cw : nullptr, Cn ? narrow + (size_t)n * hw * 4 : nullptr, C, hw,
                                                 Cw, nchw + (size_t)n * C * hw);
  NRGBD_COUNT(n_img);
  NRGBD_LAUNCH_CHECK();
  return NRGBD_OK;
}

// Gradients of nrgbd_plane_sweep_cost_packed with respect to the packed reference and source features.
// grad_cost_hwd: [h*w][D] (pixel-major, like the forward's cost). The four gradient buffers have the shapes of the
// corresponding inputs; g_src_* are zeroed here before the scatter. ws: V*12 floats (as the forward).
int nrgbd_plane_sweep_backward_packed(const float* ref_wide, const float* ref_narrow, const float* src_wide, const float* src_narrow,
                                      int Cw, int Cn, int V, int D, int h, int w, const float* K, const float* R, const float* t,
                                      const float* rays, const float* d_planes, float cx, float cy, float sigma, int metric,
                                      float* ws, const float* grad_cost_hwd, float* g_ref_wide, float* g_ref_narrow,
                                      float* g_src_wide, float* g_src_narrow, cudaStream_t st) {
  NRGBD_REQUIRE(V > 0 && D > 0 && h > 0 && w > 0, "empty problem");
  NRGBD_REQUIRE(Cw % 4 == 0 && Cn >= 0 && Cn <= 4 && Cw + Cn > 0, "bad channel split");
  NRGBD_REQUIRE((Cw == 0 || (ref_wide && src_wide && g_ref_wide && g_src_wide)) &&
                    (Cn == 0 || (ref_narrow && src_narrow && g_ref_narrow && g_src_narrow)), "null features / gradients");
  NRGBD_REQUIRE(K && R && t && rays && d_planes && ws && grad_cost_hwd, "null pointer");
  if (metric != 0 && metric != 1) {
    nrgbd_set_error("undefined metric for feature distance ...");   // homography.py:329
    return NRGBD_ERR_BAD_ARG;
  }
  const int hw = h * w, G = Cw / 4, Q = G + (Cn ? 1 : 0);
  if (Q > 64) { nrgbd_set_error("plane sweep backward supports at most 252 wide channels per call"); return NRGBD_ERR_UNSUPPORTED; }
  float* t1 = ws;
  float* KR = ws + 3 * V;
  sweep_setup_bwd_kernel<<<ceil_div(V, 32), 32, 0, st>>>(K, R, t, V, t1, KR);
  if (Cw) NRGBD_CUDA_CHECK(cudaMemsetAsync(g_src_wide, 0, sizeof(float) * (size_t)V * hw * Cw, st));
  if (Cn) NRGBD_CUDA_CHECK(cudaMemsetAsync(g_src_narrow, 0, sizeof(float) * (size_t)V * hw * 4, st));
  const float4* rw = reinterpret_cast<const float4*>(ref_wide);
  const float4* sw = reinterpret_cast<const float4*>(src_wide);
  const float4* rn = Cn ? reinterpret_cast<const float4*>(ref_narrow) : nullptr;
  const float4* sn = Cn ? reinterpret_cast<const float4*>(src_narrow) : nullptr;
  float4* grw = reinterpret_cast<float4*>(g_ref_wide);
  float4* grn = Cn ? reinterpret_cast<float4*>(g_ref_narrow) : nullptr;
#define NRGBD_BWD(L, QPL)                                                                                                      \
  do {                                                                                                                         \
    dim3 grid(ceil_div((long long)hw * L, 256));                                                                               \
    if (metric == 0)                                                                                                           \
      plane_sweep_backward_kernel<L, QPL, false><<<grid, 256, 0, st>>>(rw, sw, G, rn, sn, t1, KR, rays, d_planes, V, D, w, h, cx, cy, \
                                                                       sigma, grad_cost_hwd, grw, grn, g_src_wide, Cn ? g_src_narrow : nullptr); \
    else                                                                                                                       \
      plane_sweep_backward_kernel<L, QPL, true><<<grid, 256, 0, st>>>(rw, sw, G, rn, sn, t1, KR, rays, d_planes, V, D, w, h, cx, cy,  \
                                                                      sigma, grad_cost_hwd, grw, grn, g_src_wide, Cn ? g_src_narrow : nullptr);  \
  } while (0)
  if (Q <= 1) NRGBD_BWD(1, 1);
  else if (Q <= 4) NRGBD_BWD(4, 1);
  else if (Q <= 8) NRGBD_BWD(8, 1);
  else if (Q <= 16) NRGBD_BWD(16, 1);
  else if (Q <= 32) NRGBD_BWD(16, 2);
  else NRGBD_BWD(16, 4);
#undef NRGBD_BWD
  NRGBD_COUNT(2);
  NRGBD_LAUNCH_CHECK();
  return NRGBD_OK;
}

}  // extern "C"
