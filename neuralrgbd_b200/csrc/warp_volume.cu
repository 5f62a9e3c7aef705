// Image warp to a plane-sweep volume and K-Net input assembly (a8 head).
//
// Replaces warping/homography.py:234-280 (warp_img_feats_v3) / :183-232 (_mgpu) and the
// cat/repeat/transpose assembly of models/KVNET.py:147-166. Pure write-bound gather:
// one thread per (pixel, plane) evaluates the homography per view and bilinearly samples
// a <=4-channel image stored pixel-interleaved ([hw][4], one LDG.128 per corner).
#include "common.cuh"

namespace {

__device__ __forceinline__ float4 sample4(const float4* __restrict__ img, const Tap2D& tp) {
  float4 a = __ldg(img + tp.o_nw), b = __ldg(img + tp.o_ne), c = __ldg(img + tp.o_sw), e = __ldg(img + tp.o_se);
  float4 r;
  r.x = fmaf(e.x, tp.w_se, fmaf(c.x, tp.w_sw, fmaf(b.x, tp.w_ne, a.x * tp.w_nw)));
  r.y = fmaf(e.y, tp.w_se, fmaf(c.y, tp.w_sw, fmaf(b.y, tp.w_ne, a.y * tp.w_nw)));
  r.z = fmaf(e.z, tp.w_se, fmaf(c.z, tp.w_sw, fmaf(b.z, tp.w_ne, a.z * tp.w_nw)));
  r.w = fmaf(e.w, tp.w_se, fmaf(c.w, tp.w_sw, fmaf(b.w, tp.w_ne, a.w * tp.w_nw)));
  return r;
}

// MODE 0: out[v][c_off + c][d][pix]  (reference layout: list of V tensors C x D x h x w)
// MODE 1: out[d][pix][CK] with channels [v*3+c | ref rgb | bv_cur - bv_pred]  (K-Net input, NDHWC)
template <int MODE>
__global__ void __launch_bounds__(256)
warp_volume_kernel(const float4* __restrict__ imgs, int c_cnt, int c_off, int C_total,
                   const float* __restrict__ t1, const float* __restrict__ KR,
                   const float* __restrict__ rays, const float* __restrict__ dpl, int V, int D, int w, int h,
                   float cx, float cy, float* __restrict__ out, const float4* __restrict__ ref_img,
                   const float* __restrict__ bv_cur_hwd, const float* __restrict__ bv_pred_hwd, int CK) {
  const int hw = w * h;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)hw * D) return;
  const int pix = (int)(idx % hw);
  const int d = (int)(idx / hw);
  const float Wf = (float)w, Hf = (float)h;
  const float r0 = rays[pix], r1 = rays[hw + pix], r2 = rays[2 * hw + pix];
  const float dval = __ldg(dpl + d);
  float* o1 = nullptr;
  if (MODE == 1) o1 = out + ((size_t)d * hw + pix) * CK;
  for (int v = 0; v < V; ++v) {
    const float* kr = KR + v * 9;
    const float t2x = dot3_chain(kr[0], kr[1], kr[2], r0, r1, r2);
    const float t2y = dot3_chain(kr[3], kr[4], kr[5], r0, r1, r2);
    const float t2z = dot3_chain(kr[6], kr[7], kr[8], r0, r1, r2);
    float ix, iy;
    plane_project(t1[v * 3], t1[v * 3 + 1], t1[v * 3 + 2], t2x, t2y, t2z, dval, cx, cy, Wf, Hf, ix, iy);
    Tap2D tp = make_tap2d(ix, iy, w, h);
    float4 s = sample4(imgs + (size_t)v * hw, tp);
    const float sv[4] = {s.x, s.y, s.z, s.w};
    if (MODE == 0) {
      for (int c = 0; c < c_cnt; ++c)
        out[(((size_t)v * C_total + c_off + c) * D + d) * hw + pix] = sv[c];
    } else {
      o1[v * 3 + 0] = s.x; o1[v * 3 + 1] = s.y; o1[v * 3 + 2] = s.z;
    }
  }
  if (MODE == 1) {
    float4 r = __ldg(ref_img + pix);
    o1[3 * V + 0] = r.x; o1[3 * V + 1] = r.y; o1[3 * V + 2] = r.z;
    o1[3 * V + 3] = bv_cur_hwd[(size_t)pix * D + d] - bv_pred_hwd[(size_t)pix * D + d];
    for (int c = 3 * V + 4; c < CK; ++c) o1[c] = 0.f;
  }
}

__global__ void warp_setup_kernel(const float* __restrict__ K, const float* __restrict__ R,
                                  const float* __restrict__ t, int V, float* __restrict__ t1,
                                  float* __restrict__ KR) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
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

// imgs_packed: [V][hw][4] (channels c_off..c_off+c_cnt of each view, zero padded to 4)
// out: [V][C_total][D][hw]; ws: V*12 floats.
int nrgbd_warp_to_volume(const float* imgs_packed, int c_cnt, int c_off, int C_total, int V, int D, int h, int w,
                         const float* K, const float* R, const float* t, const float* rays,
                         const float* d_planes, float cx, float cy, float* ws, float* out, cudaStream_t st) {
  NRGBD_REQUIRE(imgs_packed && K && R && t && rays && d_planes && ws && out, "null pointer");
  NRGBD_REQUIRE(V > 0 && D > 0 && h > 0 && w > 0, "empty problem");
  NRGBD_REQUIRE(c_cnt >= 1 && c_cnt <= 4 && c_off >= 0 && c_off + c_cnt <= C_total, "bad channel window");
  float* t1 = ws; float* KR = ws + 3 * V;
  warp_setup_kernel<<<ceil_div(V, 32), 32, 0, st>>>(K, R, t, V, t1, KR);
  long long n = (long long)h * w * D;
  warp_volume_kernel<0><<<ceil_div(n, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(imgs_packed), c_cnt,
                                                          c_off, C_total, t1, KR, rays, d_planes, V, D, w, h, cx,
                                                          cy, out, nullptr, nullptr, nullptr, 0);
  NRGBD_COUNT(2);
  NRGBD_LAUNCH_CHECK();
  return NRGBD_OK;
}

// K-Net input volume [D][hw][CK], CK >= 3V+4 (models/KVNET.py:163-166), from the
// quarter-resolution RGB of the V sources ([V][hw][4]) and the reference ([hw][4]) and the two
// pixel-major log-DPVs.
int nrgbd_knet_input_volume(const float* src_rgb_packed, const float* ref_rgb_packed, const float* bv_cur_hwd,
                            const float* bv_pred_hwd, int V, int D, int h, int w, int CK, const float* K,
                            const float* R, const float* t, const float* rays, const float* d_planes, float cx,
                            float cy, float* ws, float* out, cudaStream_t st) {
  NRGBD_REQUIRE(src_rgb_packed && ref_rgb_packed && bv_cur_hwd && bv_pred_hwd && K && R && t && rays &&
                    d_planes && ws && out, "null pointer");
  NRGBD_REQUIRE(V > 0 && D > 0 && h > 0 && w > 0 && CK >= 3 * V + 4, "bad shape");
  float* t1 = ws; float* KR = ws + 3 * V;
  warp_setup_kernel<<<ceil_div(V, 32), 32, 0, st>>>(K, R, t, V, t1, KR);
  long long n = (long long)h * w * D;
  warp_volume_kernel<1><<<ceil_div(n, 256), 256, 0, st>>>(
      reinterpret_cast<const float4*>(src_rgb_packed), 3, 0, 3, t1, KR, rays, d_planes, V, D, w, h, cx, cy, out,
      reinterpret_cast<const float4*>(ref_rgb_packed), bv_cur_hwd, bv_pred_hwd, CK);
  NRGBD_COUNT(2);
  NRGBD_LAUNCH_CHECK();
  return NRGBD_OK;
}

}  // extern "C"
