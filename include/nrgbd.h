/* nrgbd.h - C ABI of the B200-native plane-sweep DPV engine (libnrgbd.so).
 *
 * The reference (NVlabs/neuralrgbd) has no FFI / plugin ABI of its own: its boundary is the
 * Python import surface `warping.homography` (free functions) and `models.KVNET.KVNET`
 * (an nn.Module). Each entry point below names the reference call it replaces (paths relative to
 * /root/reference/code); `neuralrgbd_b200/warping/homography.py` and
 * `neuralrgbd_b200/models/KVNET.py` bind them through ctypes with the reference's names and
 * argument conventions.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer to float32 unless the name
 *    says `host`; `stream` is a cudaStream_t (0 = legacy default stream).
 *  - return value: 0 on success, negative NRGBD_ERR_* otherwise; nothing throws across the
 *    boundary. nrgbd_last_error() returns a thread-local message for the last failure.
 *  - the caller owns every buffer; the library keeps no hidden global state besides the launch
 *    counter. All entry points are re-entrant and launch on the stream they are given.
 */
#ifndef NRGBD_H_
#define NRGBD_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* nrgbd_stream_t;

#define NRGBD_OK 0
#define NRGBD_ERR_BAD_ARG (-1)
#define NRGBD_ERR_CUDA (-2)
#define NRGBD_ERR_UNSUPPORTED (-3)
#define NRGBD_ERR_NOMEM (-4)

#define NRGBD_METRIC_L2 0 /* img_dis_L2_pard, warping/homography.py:81-83 */
#define NRGBD_METRIC_L1 1 /* img_dis_L1_pard, warping/homography.py:85-87 */

/* ---- runtime -------------------------------------------------------------------------- */
int nrgbd_abi_version(void);
const char* nrgbd_last_error(void);
long long nrgbd_launch_count(void);      /* kernels launched by this library since the last reset */
void nrgbd_reset_launch_count(void);

/* ---- layout helpers --------------------------------------------------------------------- */
/* Channel split used by the sweep: C = Cw + Cn, Cw % 4 == 0 ("wide", [hw][Cw]), Cn = C % 4
 * ("narrow", [hw][4] zero padded). */
void nrgbd_sweep_channel_split(int C, int* Cw, int* Cn);
/* n_img images [C][hw] (NCHW planes) -> wide [n_img][hw][Cw] and narrow [n_img][hw][4]. */
int nrgbd_pack_features(const float* nchw, int C, int hw, int n_img, float* wide, float* narrow,
                        nrgbd_stream_t stream);
/* out[b][a] = in[a][b] (in has A rows of B). Used for [hw][D] <-> [D][hw]. */
int nrgbd_transpose2d(const float* in, int A, int B, float* out, nrgbd_stream_t stream);
int nrgbd_sweep_workspace_floats(int V); /* scratch floats needed by the sweep / warp entries */

/* ---- a1-a3: fused plane-sweep cost volume ------------------------------------------------
 * replaces warping/homography.py:293-331 est_swp_volume_v4 (+ :421-448 _back_warp_homo_parallel,
 * :81-87 img_dis_L{2,1}_pard).  cost_hwd[pix][d] = sum_v dist(warp_{v,d}(src_v)[pix], ref[pix]) / sigma.
 * K 3x3, R [V][3][3], t [V][3], rays [3][hw] (cam_intrinsic['unit_ray_array_2D']), d_planes [D];
 * cx, cy = cam_intrinsic['intrinsic_M'][0,2], [1,2]. ws: nrgbd_sweep_workspace_floats(V) floats.
 * metric other than L2/L1 -> NRGBD_ERR_BAD_ARG with message
 * "undefined metric for feature distance ..." (homography.py:329). */
int nrgbd_plane_sweep_cost_packed(const float* ref_wide, const float* ref_narrow, const float* src_wide,
                                  const float* src_narrow, int Cw, int Cn, int V, int D, int h, int w,
                                  const float* K, const float* R, const float* t, const float* rays,
                                  const float* d_planes, float cx, float cy, float sigma, int metric,
                                  float* ws, float* cost_hwd, nrgbd_stream_t stream);

/* ---- a7: image warp to volume --------------------------------------------------------------
 * replaces warping/homography.py:234-280 warp_img_feats_v3 and :183-232 warp_img_feats_mgpu.
 * imgs_packed [V][hw][4] holds channels [c_off, c_off+c_cnt) of each source view;
 * out [V][C_total][D][hw] (the reference returns V tensors C x D x h x w). */
int nrgbd_warp_to_volume(const float* imgs_packed, int c_cnt, int c_off, int C_total, int V, int D, int h,
                         int w, const float* K, const float* R, const float* t, const float* rays,
                         const float* d_planes, float cx, float cy, float* ws, float* out,
                         nrgbd_stream_t stream);
/* K-Net input volume, channels-last [D][hw][CK]: [v*3+c warped src RGB | ref RGB | BV_cur - BV_predict]
 * replaces models/KVNET.py:147-166 (avg-pooled RGB is passed in packed [.][hw][4]). */
int nrgbd_knet_input_volume(const float* src_rgb_packed, const float* ref_rgb_packed,
                            const float* bv_cur_hwd, const float* bv_pred_hwd, int V, int D, int h, int w,
                            int CK, const float* K, const float* R, const float* t, const float* rays,
                            const float* d_planes, float cx, float cy, float* ws, float* out,
                            nrgbd_stream_t stream);

/* ---- a12: DPV re-projection -----------------------------------------------------------------
 * replaces warping/homography.py:654-723 resample_vol_cuda + :873-887 _set_vol_border (+ the
 * clamp of test_utils/test_KVNet.py:54-59 when do_clamp != 0). Element (d,pix) of vol/out lives at
 * d*stride_d + pix*stride_pix. E: 4x4 row-major rel_extM. d_pts: plane depths of the sampled
 * points (d_candi_new if given else d_candi). tan_hh/tan_hv = tan(radians(hfov)/2), tan(radians(vfov)/2);
 * z_half/z_radius as in :695-696. */
int nrgbd_resample_dpv(const float* vol, long long in_stride_d, long long in_stride_pix, const float* E,
                       const float* rays, const float* d_pts, int D, int H, int W, float tan_hh,
                       float tan_hv, float z_half, float z_radius, float pad_value, int do_clamp,
                       float clamp_lo, float clamp_hi, float* out, long long out_stride_d,
                       long long out_stride_pix, nrgbd_stream_t stream);

/* ---- a4 tail, a9, a11: reductions over the D planes -------------------------------------------
 * out = log_softmax_d(sign * (a + b)) (b may be NULL):
 *   sign=-1, b=NULL : BV = log_softmax(-costV)            models/basic.py:299-300
 *   sign=+1, b=prior: DPV = log_softmax(gain + BV_predict) models/KVNET.py:172-173
 * with d_planes also depth[pix] = sum_d exp(out)*d (mutils/misc.py:532-548) and
 * conf[pix] = max_d exp(out) (test_utils/export_res.py:55-62). out may be NULL. */
int nrgbd_dpv_normalize(const float* a, const float* b, float sign, int n_pix, int D, long long in_sd,
                        long long in_sp, float* out, long long out_sd, long long out_sp,
                        const float* d_planes, float* depth, float* conf, nrgbd_stream_t stream);
/* depth[pix] = sum_d (bv_log ? exp(bv) : bv)(d,pix) * d_planes[d]  -- mutils/misc.py:532-548
 * depth_val_regression (no normalisation); conf[pix] = max_d of the same probability. */
int nrgbd_depth_regression(const float* bv, int n_pix, int D, long long in_sd, long long in_sp,
                           const float* d_planes, int bv_log, float* depth, float* conf,
                           nrgbd_stream_t stream);
int nrgbd_exp(const float* x, long long n, float* y, nrgbd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NRGBD_H_ */
