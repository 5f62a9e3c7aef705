// Tensor-core convolution for the KVNET conv stacks: tcgen05 / TMEM / TMA implicit GEMM with
// error-compensated 3xTF32 products.
//
// Why 3xTF32: the reference's results must be matched to 1e-4 on the DPV through 61 (2-D) / 12 (3-D)
// convolutions separated by batch-statistics BatchNorm; single-pass TF32 operands give 1e-2-level
// DPV errors. Each fp32 operand is split a = a_hi + a_lo with a_hi = RN_tf32(a),
// a_lo = RN_tf32(a - a_hi) and the product is accumulated as a_hi*b_hi + a_lo*b_hi + a_hi*b_lo in
// the fp32 TMEM accumulator (dropped term <= 2^-22 |a b|).
//
// Structure (one 128-pixel x Cout output tile per CTA):
//   warp 0    TMA producer: per K-step (one filter tap x 32 input channels) four bulk-tensor loads -
//             an 8x16-pixel x 32-channel box of the hi and lo activation tensors (5-D tensor map over
//             [N][D][H][W][C]; padding = TMA out-of-bounds zero fill, conv stride = element stride,
//             dilation / transposed-conv parity = box origin) and the hi / lo weight slices, all in the
//             128-byte-swizzled K-major layout tcgen05 consumes; mbarrier expect-tx pipeline.
//   warp 1    MMA issuer: 4 K-slices x 3 tcgen05.mma.kind::tf32 (M=128, N=Cout_pad, K=8) per step into
//             a TMEM accumulator; tcgen05.commit releases the smem stage / signals the epilogue.
//   warps 2-5 epilogue: tcgen05.ld 32x32b (one output pixel per thread), bias / LeakyReLU, vector
//             stores to the channels-last output, BatchNorm sum / sum-of-squares via a smem
//             transpose and one double atomicAdd per channel per CTA.
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int TH = 8, TW = 16;       // spatial tile: 8 rows x 16 columns = 128 GEMM rows
constexpr int BK = 32;               // input channels per K-step (128 bytes = one swizzle row)
constexpr int A_TILE_BYTES = 128 * BK * 4;
constexpr int MAX_TAPS_TC = 27;
constexpr int NUM_THREADS = 192;

struct TcParams {
  float* y; const float* bias; double* stats;
  int N, Dz, Hy, Wx;                 // iteration space (output positions before out_stride/off)
  int tiles_x, tiles_y;
  int cin_chunks, n_taps, in_stride;
  int Cout, Cout_pad;
  int Dout, Hout, Wout, Cs_out, c_off, out_stride, out_off_y, out_off_x;
  int leaky, stages, tmem_cols, nacc;
  signed char dz[MAX_TAPS_TC], dy[MAX_TAPS_TC], dx[MAX_TAPS_TC];
  unsigned char wsel[MAX_TAPS_TC];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a protocol bug becomes a trap (error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 28); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100):
// start address >> 4 [0,14), LBO >> 4 [16,30) (ignored for swizzled K-major, set to 1),
// SBO >> 4 [32,46) = 1024 B between 8-row groups, version 1 [46,48), layout SWIZZLE_128B (2) [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
               const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int BN = p.Cout_pad;
  const uint32_t b_tile_bytes = (uint32_t)BN * BK * 4;
  const uint32_t stage_bytes = 2 * A_TILE_BYTES + 2 * b_tile_bytes;
  const uint32_t bars = smem_base + p.stages * stage_bytes;       // full[stages], empty[stages], tmem_full, tmem_ptr
  const uint32_t bar_full = bars, bar_empty = bars + 8 * p.stages, bar_tmem = bars + 16 * p.stages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_gen + p.stages * stage_bytes + 16 * p.stages + 8);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  // tile coordinates
  int t = blockIdx.x;
  const int tx = t % p.tiles_x; t /= p.tiles_x;
  const int ty = t % p.tiles_y; t /= p.tiles_y;
  const int z0 = t % p.Dz;
  const int n0 = t / p.Dz;
  const int ox0 = tx * TW, oy0 = ty * TH;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_tmem, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_a_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_a_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_b_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_b_lo) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int nk = p.n_taps * p.cin_chunks;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int ks = 0; ks < nk; ++ks) {
        const int s = ks % p.stages;
        const uint32_t ph = (uint32_t)(ks / p.stages) & 1u;
        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        mbar_expect_tx(bar_full + 8 * s, stage_bytes);
        const int tap = ks / p.cin_chunks, cc = ks - tap * p.cin_chunks;
        const uint32_t sa = smem_base + s * stage_bytes;
        const int cx = ox0 * p.in_stride + p.dx[tap], cy = oy0 * p.in_stride + p.dy[tap], cz = z0 + p.dz[tap];
        tma_load_5d(sa, &tm_a_hi, bar_full + 8 * s, cc * BK, cx, cy, cz, n0);
        tma_load_5d(sa + A_TILE_BYTES, &tm_a_lo, bar_full + 8 * s, cc * BK, cx, cy, cz, n0);
        tma_load_3d(sa + 2 * A_TILE_BYTES, &tm_b_hi, bar_full + 8 * s, cc * BK, 0, p.wsel[tap]);
        tma_load_3d(sa + 2 * A_TILE_BYTES + b_tile_bytes, &tm_b_lo, bar_full + 8 * s, cc * BK, 0, p.wsel[tap]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A/B=TF32 [7,10)/[10,13)=2,
      // K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
      for (int ks = 0; ks < nk; ++ks) {
        const int s = ks % p.stages;
        const uint32_t ph = (uint32_t)(ks / p.stages) & 1u;
        mbar_wait(bar_full + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_base + s * stage_bytes;
        // Accumulator plan (tensor-core fp32 accumulation truncates, so long chains drift):
        //   columns [i*BN, (i+1)*BN), i < nacc : a_hi*b_hi of the K-steps with ks % nacc == i
        //   columns [nacc*BN, (nacc+1)*BN)     : the two small cross terms of every K-step
        // The epilogue adds them in fp32 with round-to-nearest.
        const int ai = ks % p.nacc;
        const uint32_t d_main = tmem_base + (uint32_t)(ai * BN);
        const uint32_t d_lo = tmem_base + (uint32_t)(p.nacc * BN);
#pragma unroll
        for (int k4 = 0; k4 < BK / 8; ++k4) {
          const uint64_t a_hi = umma_desc_sw128(sa + k4 * 32);
          const uint64_t a_lo = umma_desc_sw128(sa + A_TILE_BYTES + k4 * 32);
          const uint64_t b_hi = umma_desc_sw128(sa + 2 * A_TILE_BYTES + k4 * 32);
          const uint64_t b_lo = umma_desc_sw128(sa + 2 * A_TILE_BYTES + b_tile_bytes + k4 * 32);
          umma_tf32(d_lo, a_lo, b_hi, idesc, (ks > 0 || k4 > 0) ? 1u : 0u);
          umma_tf32(d_lo, a_hi, b_lo, idesc, 1u);
          umma_tf32(d_main, a_hi, b_hi, idesc, (ks >= p.nacc || k4 > 0) ? 1u : 0u);
        }
        umma_commit(bar_empty + 8 * s);           // frees the smem stage once these MMAs have read it
      }
      umma_commit(bar_tmem);                      // accumulator complete
    }
  } else {
    // ===== epilogue (warps 2..5): TMEM lanes [32*(warp%4), +32) =====
    const int q = warp & 3;
    const int r = q * 32 + lane;                  // GEMM row = TMEM lane = pixel within the tile
    const int py = r / TW, px = r % TW;
    const int iy = oy0 + py, ix = ox0 + px;
    const bool valid = iy < p.Hy && ix < p.Wx;
    const int oy = iy * p.out_stride + p.out_off_y, ox = ix * p.out_stride + p.out_off_x;
    float* dst = p.y + ((((long long)n0 * p.Dout + z0) * p.Hout + oy) * p.Wout + ox) * (long long)p.Cs_out + p.c_off;
    float* ep = reinterpret_cast<float*>(smem_gen);          // [128][BN+1] staging, reuses the pipeline stages
    const int EPS = BN + 1;
    mbar_wait(bar_tmem, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const bool vec_ok = ((p.Cs_out | p.c_off) & 3) == 0;
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float accv[16];
      const int n_used = nk < p.nacc ? nk : p.nacc;
#pragma unroll 1
      for (int ai = -1; ai < n_used; ++ai) {            // -1: the cross-term accumulator first (small terms)
        uint32_t v[16];
        const uint32_t col = (uint32_t)((ai < 0 ? p.nacc : ai) * BN + c0);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + col;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) accv[j] = ai < 0 ? __uint_as_float(v[j]) : accv[j] + __uint_as_float(v[j]);
      }
      float f[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float x = accv[j];
        const int co = c0 + j;
        if (p.bias && co < p.Cout) x += __ldg(p.bias + co);
        if (p.leaky) x = x >= 0.f ? x : x * 0.01f;
        f[j] = (valid && co < p.Cout) ? x : 0.f;
      }
      if (valid) {
        if (vec_ok && c0 + 16 <= p.Cout) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) if (c0 + j < p.Cout) dst[c0 + j] = f[j];
        }
      }
      if (p.stats) {
#pragma unroll
        for (int j = 0; j < 16; ++j) ep[r * EPS + c0 + j] = f[j];
      }
    }
    if (p.stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");         // the four epilogue warps only
      const int e = threadIdx.x - 64;                        // 0..127
      for (int co = e; co < p.Cout; co += 128) {
        float s1 = 0.f, s2 = 0.f;
        for (int rr = 0; rr < 128; ++rr) { float x = ep[rr * EPS + co]; s1 += x; s2 = fmaf(x, x, s2); }
        atomicAdd(p.stats + co, (double)s1);
        atomicAdd(p.stats + p.Cout + co, (double)s2);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// x -> hi = RN_tf32(x), lo = RN_tf32(x - hi)
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float4* __restrict__ x, long long n4, float4* __restrict__ hi, float4* __restrict__ lo) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = x[i];
  float a[4] = {v.x, v.y, v.z, v.w}, h[4], l[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t hb, lb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(a[k]));
    h[k] = __uint_as_float(hb);
    float d = a[k] - h[k];
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(d));
    l[k] = __uint_as_float(lb);
  }
  hi[i] = make_float4(h[0], h[1], h[2], h[3]);
  lo[i] = make_float4(l[0], l[1], l[2], l[3]);
}

// PyTorch weight -> K-major packed hi / lo [tap][Cout_pad][Cin_pad]
__global__ void pack_weight_tc_kernel(const float* __restrict__ w, int kind, int Cout, int Cin, int taps, int Cin_pad,
                                      int Cout_pad, float* __restrict__ hi, float* __restrict__ lo) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long n = (long long)taps * Cin_pad * Cout_pad;
  if (i >= n) return;
  int ci = (int)(i % Cin_pad);
  int co = (int)((i / Cin_pad) % Cout_pad);
  int t = (int)(i / ((long long)Cout_pad * Cin_pad));
  float v = 0.f;
  if (co < Cout && ci < Cin) v = kind == 0 ? w[((long long)co * Cin + ci) * taps + t] : w[((long long)ci * Cout + co) * taps + t];
  uint32_t hb, lb;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
  float h = __uint_as_float(hb);
  float d = v - h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(d));
  hi[i] = h; lo[i] = __uint_as_float(lb);
}

int g_force_nacc = 0;      // development knob (nrgbd_conv_tc_set_nacc): cap on the main accumulators

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int encode_act_map(CUtensorMap* tm, const float* x, int N, int D, int H, int W, int Cin_pad, int Cs, int stride) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { nrgbd_set_error("cuTensorMapEncodeTiled unavailable"); return NRGBD_ERR_CUDA; }
  cuuint64_t dims[5] = {(cuuint64_t)Cin_pad, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)Cs * 4, (cuuint64_t)W * Cs * 4, (cuuint64_t)H * W * Cs * 4, (cuuint64_t)D * H * W * Cs * 4};
  cuuint32_t box[5] = {(cuuint32_t)BK, (cuuint32_t)(TW * stride), (cuuint32_t)(TH * stride), 1, 1};
  cuuint32_t estr[5] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { nrgbd_set_error("cuTensorMapEncodeTiled(activation) failed: %d", (int)r); return NRGBD_ERR_CUDA; }
  return NRGBD_OK;
}

int encode_w_map(CUtensorMap* tm, const float* w, int taps, int Cout_pad, int Cin_pad) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { nrgbd_set_error("cuTensorMapEncodeTiled unavailable"); return NRGBD_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)Cin_pad, (cuuint64_t)Cout_pad, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)Cin_pad * 4, (cuuint64_t)Cout_pad * Cin_pad * 4};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)Cout_pad, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { nrgbd_set_error("cuTensorMapEncodeTiled(weights) failed: %d", (int)r); return NRGBD_ERR_CUDA; }
  return NRGBD_OK;
}

int launch_tc(const float* x_hi, const float* x_lo, int N, int Din, int Hin, int Win, int Cin_pad, int Cs_in, const float* w_hi,
              const float* w_lo, int n_wslices, TcParams& p, cudaStream_t st) {
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  int rc = encode_act_map(&ta_hi, x_hi, N, Din, Hin, Win, Cin_pad, Cs_in, p.in_stride);
  if (rc == NRGBD_OK) rc = encode_act_map(&ta_lo, x_lo, N, Din, Hin, Win, Cin_pad, Cs_in, p.in_stride);
  if (rc == NRGBD_OK) rc = encode_w_map(&tb_hi, w_hi, n_wslices, p.Cout_pad, Cin_pad);
  if (rc == NRGBD_OK) rc = encode_w_map(&tb_lo, w_lo, n_wslices, p.Cout_pad, Cin_pad);
  if (rc != NRGBD_OK) return rc;
  p.cin_chunks = Cin_pad / BK;
  p.tiles_x = ceil_div(p.Wx, TW); p.tiles_y = ceil_div(p.Hy, TH);
  int nacc = 512 / p.Cout_pad - 1;
  if (nacc > 4) nacc = 4;
  if (nacc < 1) nacc = 1;
  if (g_force_nacc > 0 && g_force_nacc < nacc) nacc = g_force_nacc;
  p.nacc = nacc;
  int cols = 32; while (cols < (nacc + 1) * p.Cout_pad) cols <<= 1;
  if (cols > 512) { nrgbd_set_error("conv_tc: accumulators do not fit TMEM"); return NRGBD_ERR_UNSUPPORTED; }
  p.tmem_cols = cols;
  const size_t stage = 2 * (size_t)A_TILE_BYTES + 2 * (size_t)p.Cout_pad * BK * 4;
  int stages = (int)((220 * 1024 - 2048) / stage);
  if (stages > 4) stages = 4;
  if (stages < 2) { nrgbd_set_error("conv_tc: Cout too large for the shared-memory pipeline"); return NRGBD_ERR_UNSUPPORTED; }
  size_t ep_bytes = (size_t)128 * (p.Cout_pad + 1) * 4;
  p.stages = stages;
  size_t smem = (size_t)stages * stage + 1024 /*align*/ + 256 /*barriers*/;
  if (stages * stage < ep_bytes) smem = ep_bytes + 1024 + 256;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { nrgbd_set_error("conv_tc: cannot opt in to %zu bytes of shared memory: %s", smem, cudaGetErrorString(e)); return NRGBD_ERR_CUDA; }
    configured = smem;
  }
  const long long tiles = (long long)N * p.Dz * p.tiles_x * p.tiles_y;
  conv_tc_kernel<<<(unsigned)tiles, NUM_THREADS, smem, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, p);
  return NRGBD_OK;
}

}  // namespace

extern "C" {

void nrgbd_conv_tc_set_nacc(int n) { g_force_nacc = n; }

// Whether the tensor-core path can run a convolution with these channel counts.
int nrgbd_conv_tc_supported(int Cin_pad, int Cout_pad) {
  return (Cin_pad % 32 == 0 && Cin_pad >= 32 && Cout_pad % 16 == 0 && Cout_pad >= 16 && Cout_pad <= 256) ? 1 : 0;
}

int nrgbd_split_tf32(const float* x, long long n, float* hi, float* lo, cudaStream_t st) {
  NRGBD_REQUIRE(x && hi && lo && n > 0 && n % 4 == 0, "bad arguments");
  split_tf32_kernel<<<ceil_div(n / 4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(x), n / 4, reinterpret_cast<float4*>(hi),
                                                         reinterpret_cast<float4*>(lo));
  NRGBD_COUNT(1);
  NRGBD_LAUNCH_CHECK();
  return NRGBD_OK;
}

// PyTorch weight [Cout][Cin][taps] (transposed=0) or [Cin][Cout][taps] (1) -> hi / lo, each
// [taps][Cout_pad][Cin_pad] (K-major), TF32-split.
int nrgbd_pack_conv_weight_tc(const float* w, int transposed, int Cout, int Cin, int taps, int Cin_pad, int Cout_pad,
                              float* hi, float* lo, cudaStream_t st) {
  NRGBD_REQUIRE(w && hi && lo && Cout > 0 && Cin > 0 && taps > 0 && Cin_pad >= Cin && Cout_pad >= Cout, "bad arguments");
  long long n = (long long)taps * Cin_pad * Cout_pad;
  pack_weight_tc_kernel<<<ceil_div(n, 256), 256, 0, st>>>(w, transposed ? 1 : 0, Cout, Cin, taps, Cin_pad, Cout_pad, hi, lo);
  NRGBD_COUNT(1);
  NRGBD_LAUNCH_CHECK();
  return NRGBD_OK;
}

// Tensor-core counterpart of nrgbd_conv_nhwc: same semantics, inputs given as the TF32 hi / lo split
// of the activations and of the (K-major packed) weights. Requires nrgbd_conv_tc_supported().
int nrgbd_conv_nhwc_tc(const float* x_hi, const float* x_lo, int N, int Din, int Hin, int Win, int Cin_pad, int Cs_in,
                       const float* w_hi, const float* w_lo, const float* bias, int Cout, int Cout_pad, int kd, int kh, int kw,
                       int stride, int pad, int dilation, float* y, int Hout, int Wout, int Cs_out, int c_off, int leaky,
                       double* stats, cudaStream_t st) {
  NRGBD_REQUIRE(x_hi && x_lo && w_hi && w_lo && y, "null pointer");
  NRGBD_REQUIRE(nrgbd_conv_tc_supported(Cin_pad, Cout_pad) && Cin_pad <= Cs_in && Cs_in % 4 == 0 && Cout <= Cout_pad,
                "channel counts not supported by the tensor-core path");
  NRGBD_REQUIRE(kd * kh * kw <= MAX_TAPS_TC && stride >= 1 && stride <= 8, "unsupported filter");
  NRGBD_REQUIRE(Hout == (Hin + 2 * pad - dilation * (kh - 1) - 1) / stride + 1 &&
                    Wout == (Win + 2 * pad - dilation * (kw - 1) - 1) / stride + 1, "output extent mismatch");
  TcParams p;
  p.y = y; p.bias = bias; p.stats = stats;
  p.N = N; p.Dz = Din; p.Hy = Hout; p.Wx = Wout;
  p.in_stride = stride; p.Cout = Cout; p.Cout_pad = Cout_pad;
  p.Dout = Din; p.Hout = Hout; p.Wout = Wout; p.Cs_out = Cs_out; p.c_off = c_off;
  p.out_stride = 1; p.out_off_y = 0; p.out_off_x = 0; p.leaky = leaky;
  int t = 0;
  for (int a = 0; a < kd; ++a)
    for (int b = 0; b < kh; ++b)
      for (int c = 0; c < kw; ++c) {
        p.dz[t] = (signed char)(a - kd / 2); p.dy[t] = (signed char)(b * dilation - pad); p.dx[t] = (signed char)(c * dilation - pad);
        p.wsel[t] = (unsigned char)t; ++t;
      }
  p.n_taps = t;
  int rc = launch_tc(x_hi, x_lo, N, Din, Hin, Win, Cin_pad, Cs_in, w_hi, w_lo, t, p, st);
  if (rc != NRGBD_OK) return rc;
  NRGBD_COUNT(1);
  NRGBD_LAUNCH_CHECK();
  return NRGBD_OK;
}

// Tensor-core counterpart of nrgbd_conv_transpose2d_k4s2_nhwc (four parity-class launches).
int nrgbd_conv_transpose2d_k4s2_nhwc_tc(const float* x_hi, const float* x_lo, int N, int Hin, int Win, int Cin_pad, int Cs_in,
                                        const float* w_hi, const float* w_lo, const float* bias, int Cout, int Cout_pad, float* y,
                                        int Cs_out, int c_off, int leaky, cudaStream_t st) {
  NRGBD_REQUIRE(x_hi && x_lo && w_hi && w_lo && y, "null pointer");
  NRGBD_REQUIRE(nrgbd_conv_tc_supported(Cin_pad, Cout_pad) && Cin_pad <= Cs_in && Cs_in % 4 == 0 && Cout <= Cout_pad,
                "channel counts not supported by the tensor-core path");
  const int kys[2][2] = {{1, 3}, {0, 2}};
  const int dys[2][2] = {{0, -1}, {1, 0}};
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      TcParams p;
      p.y = y; p.bias = bias; p.stats = nullptr;
      p.N = N; p.Dz = 1; p.Hy = Hin; p.Wx = Win;
      p.in_stride = 1; p.Cout = Cout; p.Cout_pad = Cout_pad;
      p.Dout = 1; p.Hout = 2 * Hin; p.Wout = 2 * Win; p.Cs_out = Cs_out; p.c_off = c_off;
      p.out_stride = 2; p.out_off_y = py; p.out_off_x = px; p.leaky = leaky;
      int t = 0;
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          p.dz[t] = 0; p.dy[t] = (signed char)dys[py][a]; p.dx[t] = (signed char)dys[px][b];
          p.wsel[t] = (unsigned char)(kys[py][a] * 4 + kys[px][b]); ++t;
        }
      p.n_taps = 4;
      int rc = launch_tc(x_hi, x_lo, N, 1, Hin, Win, Cin_pad, Cs_in, w_hi, w_lo, 16, p, st);
      if (rc != NRGBD_OK) return rc;
    }
  NRGBD_COUNT(4);
  NRGBD_LAUNCH_CHECK();
  return NRGBD_OK;
}

}  // extern "C"
