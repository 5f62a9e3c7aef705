// Layout / pooling / resampling helpers of the conv stacks (channels-last, fp32).
//   NCHW <-> NHWC(Cs)            boundary conversion of the reference-facing tensors
//   avg_pool2d(k)                F.avg_pool2d (models/basic.py:259-262, KVNET.py:149-151) and the SPP
//                                AvgPool2d branches (models/psm_submodule.py:103-117)
//   bilinear upsample            F.upsample(mode='bilinear', align_corners=True) (psm_submodule.py:148-159)
//   channel copy / exp           torch.cat(...) and torch.exp(BV) feeding R-Net (KVNET.py:134, Refine.py:91)
#include "common.cuh"

namespace {

// in [C][P] -> out[p*Cs + c_off + c]
__global__ void planes_to_interleaved_kernel(const float* __restrict__ in, int C, long long P, int Cs, int c_off,
                                             float* __restrict__ out) {
  __shared__ float tile[32][33];
  long long p0 = (long long)blockIdx.x * 32;
  int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i; long long p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < P) ? in[(long long)c * P + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    long long p = p0 + i; int c = c0 + threadIdx.x;
    if (p < P && c < C) out[p * Cs + c_off + c] = tile[threadIdx.x][i];
  }
}

// small C (<= 4): one thread per pixel, coalesced plane reads, one 16-byte store (Cs == 4, c_off == 0)
__global__ void __launch_bounds__(256)
planes_to_interleaved4_kernel(const float* __restrict__ in, int C, long long P, int N, float* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * N) return;
  long long n = i / P, p = i - n * P;
  const float* b = in + n * C * P + p;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = 0; c < C; ++c) v[c] = b[(long long)c * P];
  reinterpret_cast<float4*>(out)[i] = make_float4(v[0], v[1], v[2], v[3]);
}

// in[p*Cs + c_off + c] -> out [C][P]
__global__ void interleaved_to_planes_kernel(const float* __restrict__ in, int C, long long P, int Cs, int c_off,
                                             float* __restrict__ out) {
  __shared__ float tile[32][33];
  long long p0 = (long long)blockIdx.x * 32;
  int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    long long p = p0 + i; int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < P && c < C) ? in[p * Cs + c_off + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i; long long p = p0 + threadIdx.x;
    if (c < C && p < P) out[(long long)c * P + p] = tile[threadIdx.x][i];
  }
}

// one block per output position; threads = 32 channels x 8 row lanes
__global__ void __launch_bounds__(256)
avgpool_nhwc_kernel(const float* __restrict__ x, int N, int H, int W, int Cs_in, int C, int k, float* __restrict__ y,
                    int Ho, int Wo, int Cs_out, int c_off) {
  __shared__ float red[8][33];
  int pos = blockIdx.x;
  int ox = pos % Wo; int oy = (pos / Wo) % Ho; int n = pos / (Wo * Ho);
  int cl = threadIdx.x % 32, rl = threadIdx.x / 32;
  for (int c0 = 0; c0 < C; c0 += 32) {
    int c = c0 + cl;
    float s = 0.f;
    if (c < C) {
      for (int r = rl; r < k; r += 8) {
        const float* row = x + (((long long)n * H + oy * k + r) * W + ox * k) * Cs_in + c;
        for (int q = 0; q < k; ++q) s += row[(long long)q * Cs_in];
      }
    }
    red[rl][cl] = s;
    __syncthreads();
    if (rl == 0 && c < C) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) t += red[r][cl];
      y[(((long long)n * Ho + oy) * Wo + ox) * Cs_out + c_off + c] = t / (float)(k * k);
    }
    __syncthreads();
  }
}

// small windows (k <= 8): one thread per (output position, group of 4 channels), float4 loads
__global__ void __launch_bounds__(256)
avgpool_small_nhwc_kernel(const float* __restrict__ x, int N, int H, int W, int Cs_in, int C, int k, float* __restrict__ y,
                          int Ho, int Wo, int Cs_out, int c_off) {
  const int G = (C + 3) / 4;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * Ho * Wo * G) return;
  int g = (int)(i % G); long long pos = i / G;
  int ox = (int)(pos % Wo); int oy = (int)((pos / Wo) % Ho); int n = (int)(pos / ((long long)Wo * Ho));
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = 0; r < k; ++r) {
    const float* row = x + (((long long)n * H + oy * k + r) * W + (long long)ox * k) * Cs_in + g * 4;
    for (int q = 0; q < k; ++q) {
      float4 v = __ldg(reinterpret_cast<const float4*>(row + (long long)q * Cs_in));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  const float inv = (float)(k * k);
  float o[4] = {s.x / inv, s.y / inv, s.z / inv, s.w / inv};
  float* dst = y + (((long long)n * Ho + oy) * Wo + ox) * Cs_out + c_off + g * 4;
  for (int c = 0; c < 4; ++c) if (g * 4 + c < C) dst[c] = o[c];
}

__global__ void __launch_bounds__(256)
upsample_bilinear_ac_kernel(const float* __restrict__ x, int N, int Hi, int Wi, int Cs_in, int C, float* __restrict__ y,
                            int Ho, int Wo, int Cs_out, int c_off, float sy, float sx) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)N * Ho * Wo * C;
  if (i >= total) return;
  int c = (int)(i % C); long long r = i / C;
  int ox = (int)(r % Wo); r /= Wo;
  int oy = (int)(r % Ho); int n = (int)(r / Ho);
  float fy = sy * (float)oy, fx = sx * (float)ox;       // area_pixel_compute_source_index, align_corners=True
  int y0 = min((int)fy, Hi - 1), x0 = min((int)fx, Wi - 1);
  int y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
  float ly = fy - (float)y0, lx = fx - (float)x0;
  float hy = 1.f - ly, hx = 1.f - lx;
  const float* b = x + (long long)n * Hi * Wi * Cs_in + c;
  float v00 = b[((long long)y0 * Wi + x0) * Cs_in], v01 = b[((long long)y0 * Wi + x1) * Cs_in];
  float v10 = b[((long long)y1 * Wi + x0) * Cs_in], v11 = b[((long long)y1 * Wi + x1) * Cs_in];
  float v = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
  y[(((long long)n * Ho + oy) * Wo + ox) * Cs_out + c_off + c] = v;
}

// y[p*Cs_out + c_off_out + c] = op(x[p*Cs_in + c_off_in + c]); op 0 copy, 1 exp
__global__ void __launch_bounds__(256)
copy_channels_kernel(const float* __restrict__ x, long long P, int Cs_in, int c_off_in, int C, int op,
                     float* __restrict__ y, int Cs_out, int c_off_out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * C) return;
  int c = (int)(i % C); long long p = i / C;
  float v = x[p * Cs_in + c_off_in + c];
  if (op == 1) v = expf(v);
  y[p * Cs_out + c_off_out + c] = v;
}

}  // namespace

extern "C" {

int nrgbd_nchw_to_nhwc(const float* x, int N, int C, long long P, float* y, int Cs, int c_off, cudaStream_t st) {
  NRGBD_REQUIRE(x && y && N > 0 && C > 0 && P > 0 && c_off + C <= Cs, "bad arguments");
  if (C <= 4 && Cs == 4 && c_off == 0) {
    planes_to_interleaved4_kernel<<<ceil_div(P * N, 256), 256, 0, st>>>(x, C, P, N, y);
    NRGBD_COUNT(1);
    NRGBD_LAUNCH_CHECK();
    return NRGBD_OK;
  }
  dim3 blk(32, 8), grid(ceil_div(P, 32), ceil_div(C, 32));
  for (int n = 0; n < N; ++n)
    planes_to_interleaved_kernel<<<grid, blk, 0, st>>>(x + (long long)n * C * P, C, P, Cs, c_off, y + (long long)n * P * Cs);
  NRGBD_COUNT(N);
  NRGBD_LAUNCH_CHECK();
  return NRGBD_OK;
}

int nrgbd_nhwc_to_nchw(const float* x, int N, int C, long long P, int Cs, int c_off, float* y, cudaStream_t st) {
  NRGBD_REQUIRE(x && y && N > 0 && C > 0 && P > 0 && c_off + C <= Cs, "bad arguments");
  dim3 blk(32, 8), grid(ceil_div(P, 32), ceil_div(C, 32));
  for (int n = 0; n < N; ++n)
    interleaved_to_planes_kernel<<<grid, blk, 0, st>>>(x + (long long)n * P * Cs, C, P, Cs, c_off, y + (long long)n * C * P);
  NRGBD_COUNT(N);
  NRGBD_LAUNCH_CHECK();
  return NRGBD_OK;
}

int nrgbd_avgpool_nhwc(const float* x, int N, int H, int W, int Cs_in, int C, int k, float* y, int Cs_out, int c_off,
                       cudaStream_t st) {
  NRGBD_REQUIRE(x && y && k >= 1 && H / k >= 1 && W / k >= 1 && C <= Cs_in && c_off + C <= Cs_out, "bad arguments");
  int Ho = H / k, Wo = W / k;
  if (k <= 8 && Cs_in % 4 == 0) {
    long long n = (long long)N * Ho * Wo * ((C + 3) / 4);
    avgpool_small_nhwc_kernel<<<ceil_div(n, 256), 256, 0, st>>>(x, N, H, W, Cs_in, C, k, y, Ho, Wo, Cs_out, c_off);
  } else {
    avgpool_nhwc_kernel<<<N * Ho * Wo, 256, 0, st>>>(x, N, H, W, Cs_in, C, k, y, Ho, Wo, Cs_out, c_off);
  }
  NRGBD_COUNT(1);
  NRGBD_LAUNCH_CHECK();
  return NRGBD_OK;
}

int nrgbd_upsample_bilinear_ac_nhwc(const float* x, int N, int Hi, int Wi, int Cs_in, int C, float* y, int Ho, int Wo,
                                    int Cs_out, int c_off, cudaStream_t st) {
  NRGBD_REQUIRE(x && y && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && C <= Cs_in && c_off + C <= Cs_out, "bad arguments");
  float sy = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
  float sx = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
  long long total = (long long)N * Ho * Wo * C;
  upsample_bilinear_ac_kernel<<<ceil_div(total, 256), 256, 0, st>>>(x, N, Hi, Wi, Cs_in, C, y, Ho, Wo, Cs_out, c_off, sy, sx);
  NRGBD_COUNT(1);
  NRGBD_LAUNCH_CHECK();
  return NRGBD_OK;
}

int nrgbd_copy_channels(const float* x, long long P, int Cs_in, int c_off_in, int C, int op, float* y, int Cs_out,
                        int c_off_out, cudaStream_t st) {
  NRGBD_REQUIRE(x && y && P > 0 && C > 0 && c_off_in + C <= Cs_in && c_off_out + C <= Cs_out && (op == 0 || op == 1),
                "bad arguments");
  copy_channels_kernel<<<ceil_div(P * C, 256), 256, 0, st>>>(x, P, Cs_in, c_off_in, C, op, y, Cs_out, c_off_out);
  NRGBD_COUNT(1);
  NRGBD_LAUNCH_CHECK();
  return NRGBD_OK;
}

}  // extern "C"
