// Row primitives of the 768-wide LayerNorm (M.py:204-219), shared by the row kernels (elementwise.cu) and by the
// LayerNorm warps of the residual GEMMs (gemm.cu) so that both produce bit-identical rows: one warp owns one row, 6 float4
// per lane, statistics by warp shuffle in fp32, two-pass (mean, then centred variance) like the reference's
// (x-u).pow(2).mean().
#pragma once

#include "common.cuh"
#include "kernels.h"

namespace msclip {
namespace rowops {

constexpr int kD = 768;
constexpr int kVec = kD / 128;  // float4 per lane
constexpr float kLnEps = 1e-12f;  // M.py:205

__device__ __forceinline__ void load_row(const float* __restrict__ src, int lane, float4 (&v)[kVec]) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int i = 0; i < kVec; ++i) v[i] = s4[lane + 32 * i];
}

// in-register LayerNorm of one row held as 6 float4 per lane; returns normalised*w+b in place
__device__ __forceinline__ void layer_norm_row(float4 (&v)[kVec], const float* __restrict__ w,
                                               const float* __restrict__ b, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / kD);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    v[i].x -= mean;
    v[i].y -= mean;
    v[i].z -= mean;
    v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / kD) + kLnEps);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const float4 ww = __ldg(w4 + lane + 32 * i);
    const float4 bb = __ldg(b4 + lane + 32 * i);
    v[i].x = ww.x * (v[i].x * rstd) + bb.x;
    v[i].y = ww.y * (v[i].y * rstd) + bb.y;
    v[i].z = ww.z * (v[i].z * rstd) + bb.z;
    v[i].w = ww.w * (v[i].w * rstd) + bb.w;
  }
}

__device__ __forceinline__ void store_row_bf16(op16* __restrict__ dst, int lane, const float4 (&v)[kVec]) {
  uint2* d2 = reinterpret_cast<uint2*>(dst);
#pragma unroll
  for (int i = 0; i < kVec; ++i) d2[lane + 32 * i] = make_uint2(pack16(v[i].x, v[i].y), pack16(v[i].z, v[i].w));
}
}  // namespace rowops
}  // namespace msclip
