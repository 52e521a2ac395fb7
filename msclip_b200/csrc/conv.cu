// Data-movement kernels that turn the convolutions of the early-conv stem (EarlyconvRes, M.py:1993-2000),
// the parallel branch (Resnet_Stage / ConvResBlock, M.py:1842-1861) and the lateral adapters
// (M.py:1752-1759) into GEMM operands for gemm.cu.  Activations are NHWC op16 so that 8 channels of one
// tap are one 16-byte vector; BatchNorm (eval) is folded into the packed weights.  All kernels are
// HBM-bound: 16-byte loads/stores, adjacent threads on adjacent vectors.
#include "common.cuh"
#include "kernels.h"

#include <cuda_fp16.h>

namespace msclip {

namespace {

__device__ __forceinline__ float load_pixel(const void* img, int dtype, long long idx) {
  if (dtype == 0) return reinterpret_cast<const float*>(img)[idx];
  if (dtype == 1) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(img)[idx]);
  return __half2float(reinterpret_cast<const __half*>(img)[idx]);
}

// One thread per output pixel: 27 taps (c, ky, kx) of a 3x3 / stride 2 / pad 1 window + 5 zero columns.
__global__ void __launch_bounds__(256)
im2col_first_kernel(const void* __restrict__ img, int dtype, op16* __restrict__ out, long long total, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < total; idx += gridDim.x * 256ll) {
    const int ox = static_cast<int>(idx % Wo);
    const int oy = static_cast<int>((idx / Wo) % Ho);
    const long long bi = idx / (static_cast<long long>(Wo) * Ho);
    float v[32];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = 2 * oy - 1 + ky;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = 2 * ox - 1 + kx;
          const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
          v[c * 9 + ky * 3 + kx] = ok ? load_pixel(img, dtype, ((bi * 3 + c) * H + iy) * W + ix) : 0.f;
        }
      }
    }
#pragma unroll
    for (int k = 27; k < 32; ++k) v[k] = 0.f;
    uint4* o4 = reinterpret_cast<uint4*>(out + idx * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o4[j] = make_uint4(pack16(v[8 * j], v[8 * j + 1]), pack16(v[8 * j + 2], v[8 * j + 3]),
                         pack16(v[8 * j + 4], v[8 * j + 5]), pack16(v[8 * j + 6], v[8 * j + 7]));
  }
}

// One thread per (output pixel, tap, 8-channel vector).
__global__ void __launch_bounds__(256)
im2col_nhwc_kernel(const op16* __restrict__ in, int H, int W, int cpix, int c_off, int C, int ksize, int stride,
                   int pad, int Ho, int Wo, op16* __restrict__ out, long long out_ld, int out_off, long long total) {
  const int cv = C / 8;
  const int per_pixel = ksize * ksize * cv;
  for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < total; idx += gridDim.x * 256ll) {
    const int e = static_cast<int>(idx % per_pixel);
    const long long pix = idx / per_pixel;
    const int c8 = e % cv;
    const int tap = e / cv;
    const int ky = tap / ksize, kx = tap % ksize;
    const int ox = static_cast<int>(pix % Wo);
    const int oy = static_cast<int>((pix / Wo) % Ho);
    const long long bi = pix / (static_cast<long long>(Wo) * Ho);
    const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = *reinterpret_cast<const uint4*>(in + ((bi * H + iy) * W + ix) * cpix + c_off + c8 * 8);
    *reinterpret_cast<uint4*>(out + pix * out_ld + out_off + tap * C + c8 * 8) = v;
  }
}

// One thread per (output cell, 8-channel vector): fp32 accumulation over the k x k patch.
__global__ void __launch_bounds__(256)
patch_pool_kernel(const op16* __restrict__ in, int H, int W, int cpix, int c_off, int C, int k,
                  const float* __restrict__ w, const float* __restrict__ bias, op16* __restrict__ out, long long total) {
  const int cv = C / 8;
  const int Ho = H / k, Wo = W / k;
  for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < total; idx += gridDim.x * 256ll) {
    const int c8 = static_cast<int>(idx % cv);
    const long long cell = idx / cv;
    const int ox = static_cast<int>(cell % Wo);
    const int oy = static_cast<int>((cell / Wo) % Ho);
    const long long bi = cell / (static_cast<long long>(Wo) * Ho);
    float acc[8];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(bias + c8 * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(bias + c8 * 8 + 4);
      acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
      acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
    }
    for (int ky = 0; ky < k; ++ky) {
      const op16* row = in + ((bi * H + oy * k + ky) * W + ox * k) * cpix + c_off + c8 * 8;
      for (int kx = 0; kx < k; ++kx) {
        const uint4 raw = *reinterpret_cast<const uint4*>(row + static_cast<long long>(kx) * cpix);
        const float* wp = w + static_cast<long long>(ky * k + kx) * C + c8 * 8;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
        const op162* h = reinterpret_cast<const op162*>(&raw);
        const float2 f0 = op162_to_float2(h[0]), f1 = op162_to_float2(h[1]);
        const float2 f2 = op162_to_float2(h[2]), f3 = op162_to_float2(h[3]);
        acc[0] = fmaf(f0.x, w0.x, acc[0]); acc[1] = fmaf(f0.y, w0.y, acc[1]);
        acc[2] = fmaf(f1.x, w0.z, acc[2]); acc[3] = fmaf(f1.y, w0.w, acc[3]);
        acc[4] = fmaf(f2.x, w1.x, acc[4]); acc[5] = fmaf(f2.y, w1.y, acc[5]);
        acc[6] = fmaf(f3.x, w1.z, acc[6]); acc[7] = fmaf(f3.y, w1.w, acc[7]);
      }
    }
    *reinterpret_cast<uint4*>(out + cell * C + c8 * 8) =
        make_uint4(pack16(acc[0], acc[1]), pack16(acc[2], acc[3]), pack16(acc[4], acc[5]),
                   pack16(acc[6], acc[7]));
  }
}

__global__ void __launch_bounds__(256)
pack_op16_kernel(const float* __restrict__ src, long long sn, long long sk, const float* __restrict__ row_scale,
                 op16* __restrict__ dst, long long ldd, int N, int K) {
  const long long total = static_cast<long long>(N) * K;
  for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < total; idx += gridDim.x * 256ll) {
    const int k = static_cast<int>(idx % K);
    const long long n = idx / K;
    float v = src[n * sn + k * sk];
    if (row_scale) v *= row_scale[n];
    dst[n * ldd + k] = to_op16(v);
  }
}

// LN fold: one warp per weight row n.  dst[n][k] = op16(src[n][k] * rs[n] * gamma[k]); colsum[n] = sum_k dst[n][k] (of
// the rounded values, which is what the MMA will see); bias_out[n] = rs[n] * (bias[n] + sum_k src[n][k] * beta[k])
__global__ void __launch_bounds__(256)
pack_ln_fold_kernel(const float* __restrict__ src, const float* __restrict__ row_scale, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ bias, op16* __restrict__ dst,
                    float* __restrict__ colsum, float* __restrict__ bias_out, int N, int K) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const float rs = row_scale ? row_scale[n] : 1.0f;
  float cs = 0.f, bb = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = src[static_cast<long long>(n) * K + k];
    const op16 q = to_op16(w * rs * gamma[k]);
    dst[static_cast<long long>(n) * K + k] = q;
    cs += op16_to_float(q);
    bb = fmaf(w, beta[k], bb);
  }
  cs = warp_sum(cs);
  bb = warp_sum(bb);
  if (lane == 0) {
    colsum[n] = cs;
    bias_out[n] = rs * ((bias ? bias[n] : 0.f) + bb);
  }
}

inline int flat_grid(long long total) {
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 32;
  return static_cast<int>(blocks < cap ? blocks : cap);
}

}  // namespace

int launch_im2col_first(const void* img, int img_dtype, op16* out, int batch, int H, int W, cudaStream_t stream) {
  if (batch <= 0) return 0;
  MSCLIP_REQUIRE(img_dtype >= 0 && img_dtype <= 2, "image dtype must be 0 (f32), 1 (bf16) or 2 (f16)");
  MSCLIP_REQUIRE(H % 2 == 0 && W % 2 == 0, "image height/width must be even");
  const long long total = static_cast<long long>(batch) * (H / 2) * (W / 2);
  im2col_first_kernel<<<flat_grid(total), 256, 0, stream>>>(img, img_dtype, out, total, H, W);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_im2col_nhwc(const op16* in, int batch, int H, int W, int cpix, int c_off, int C, int ksize, int stride,
                       int pad, op16* out, int64_t out_ld, int out_off, cudaStream_t stream) {
  if (batch <= 0) return 0;
  MSCLIP_REQUIRE(C % 8 == 0 && cpix % 8 == 0 && c_off % 8 == 0 && out_ld % 8 == 0 && out_off % 8 == 0,
                 "im2col: channel counts / pitches must be multiples of 8");
  const int Ho = (H + 2 * pad - ksize) / stride + 1, Wo = (W + 2 * pad - ksize) / stride + 1;
  const long long total = static_cast<long long>(batch) * Ho * Wo * ksize * ksize * (C / 8);
  im2col_nhwc_kernel<<<flat_grid(total), 256, 0, stream>>>(in, H, W, cpix, c_off, C, ksize, stride, pad, Ho, Wo, out,
                                                          out_ld, out_off, total);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_patch_pool(const op16* in, int batch, int H, int W, int cpix, int c_off, int C, int k, const float* w,
                      const float* bias, op16* out, cudaStream_t stream) {
  if (batch <= 0) return 0;
  MSCLIP_REQUIRE(C % 8 == 0 && cpix % 8 == 0 && c_off % 8 == 0, "patch_pool: channels must be multiples of 8");
  MSCLIP_REQUIRE(H % k == 0 && W % k == 0, "patch_pool: kernel must tile the feature map");
  const long long total = static_cast<long long>(batch) * (H / k) * (W / k) * (C / 8);
  patch_pool_kernel<<<flat_grid(total), 256, 0, stream>>>(in, H, W, cpix, c_off, C, k, w, bias, out, total);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_pack_ln_fold(const float* src, const float* row_scale, const float* gamma, const float* beta, const float* bias,
                        op16* dst, float* colsum, float* bias_out, int N, int K, cudaStream_t stream) {
  if (N <= 0 || K <= 0) return 0;
  pack_ln_fold_kernel<<<(N + 7) / 8, 256, 0, stream>>>(src, row_scale, gamma, beta, bias, dst, colsum, bias_out, N, K);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// dense K = (source, tap, c) -> padded K = (source, tap, 64-channel block, 64): one thread per padded element
struct KpadDesc {
  int nsrc, N;
  int C[2], taps[2], cblk[2], dense_begin[2], pad_begin[2];
  int kpad;
};
__global__ void __launch_bounds__(256)
pack_conv_kpad_kernel(const op16* __restrict__ dense, long long ldd, op16* __restrict__ padded, KpadDesc d) {
  const long long total = static_cast<long long>(d.N) * d.kpad;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int n = static_cast<int>(i / d.kpad), kp = static_cast<int>(i % d.kpad);
    const int s = (d.nsrc > 1 && kp >= d.pad_begin[1]) ? 1 : 0;
    const int r = kp - d.pad_begin[s];
    const int per_tap = d.cblk[s] * 64;
    const int tap = r / per_tap, c = r - tap * per_tap;
    op16 v = to_op16(0.f);
    if (c < d.C[s]) v = dense[n * ldd + d.dense_begin[s] + tap * d.C[s] + c];
    padded[i] = v;
  }
}

int launch_pack_conv_kpad(const op16* dense, int64_t ldd, const ConvSource* src, int nsrc, int N, op16* padded,
                          cudaStream_t stream) {
  KpadDesc d = {};
  d.nsrc = nsrc;
  d.N = N;
  int kd = 0, kp = 0;
  for (int i = 0; i < nsrc; ++i) {
    d.C[i] = src[i].C;
    d.taps[i] = src[i].ksize * src[i].ksize;
    d.cblk[i] = (src[i].C + 63) / 64;
    d.dense_begin[i] = kd;
    d.pad_begin[i] = kp;
    kd += d.taps[i] * d.C[i];
    kp += d.taps[i] * d.cblk[i] * 64;
  }
  d.kpad = kp;
  MSCLIP_REQUIRE(kp == conv_tma_kpad(src, nsrc) && ldd >= kd, "pack_conv_kpad: layout mismatch");
  pack_conv_kpad_kernel<<<flat_grid(static_cast<long long>(N) * kp), 256, 0, stream>>>(dense, ldd, padded, d);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_pack_op16(const float* src, int64_t sn, int64_t sk, const float* row_scale, op16* dst, int64_t ldd, int N,
                     int K, cudaStream_t stream) {
  if (N <= 0 || K <= 0) return 0;
  pack_op16_kernel<<<flat_grid(static_cast<long long>(N) * K), 256, 0, stream>>>(src, sn, sk, row_scale, dst, ldd, N, K);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace msclip
