// Pieces shared by the tcgen05 GEMM kernels (gemm.cu: TMA-fed operands; conv_gemm.cu: implicit-GEMM gather of
// the A operand): tile constants, parameters and the fused epilogue.
#pragma once

#include "common.cuh"
#include "kernels.h"

namespace msclip {
namespace gemm_detail {

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 op16 = 128 B = one swizzle row
constexpr int kNumEpilogueWarps = 8;
constexpr int kGemmThreads = 128 + 32 * kNumEpilogueWarps;  // TMA, MMA, TMEM-alloc, spare + epilogue warps
constexpr int kAccStride = 256;  // TMEM columns between the two accumulator stages

struct GemmParams {
  int M, N, K;
  int tiles_n, total_tiles;
  const float* bias;
  void* out;
  const float* resid;
  long long ldo, ldr;
  float alpha;  // acc is scaled by alpha before the bias (similarity logits: exp(logit_scale))
  int vec_ok;   // rows of out / resid keep 16-byte alignment -> vector stores
};

__device__ __forceinline__ float quick_gelu(float x) {
  // x * sigmoid(1.702 x)   (M.py:224)
  // 1 / (1 + 2^(-1.702 log2(e) x)) with MUFU.EX2 + MUFU.RCP (2^-22 relative error each)
  return x * fast_rcp(1.0f + fast_ex2(-2.4554669595930157f * x));
}

// Operands the epilogue needs from global memory for one chunk (bias slice, residual row segment); fetched
// before the accumulator chunk is waited for so that their latency overlaps the TMEM load.
template <int EPI, int CH>
struct EpiOperands {
  float4 bias[CH / 4];
  float4 resid[(EPI == EPI_RESID_F32) ? CH / 4 : 1];
};

template <int EPI, int CH>
__device__ __forceinline__ void epilogue_prefetch(EpiOperands<EPI, CH>& o, const GemmParams& p, int row, int col0,
                                                  bool fast) {
  if (!fast) return;
  if (p.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
    for (int j = 0; j < CH / 4; ++j) o.bias[j] = __ldg(b4 + j);
  } else {
#pragma unroll
    for (int j = 0; j < CH / 4; ++j) o.bias[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (EPI == EPI_RESID_F32) {
    const float4* x4 = reinterpret_cast<const float4*>(p.resid + static_cast<long long>(row) * p.ldr + col0);
#pragma unroll
    for (int j = 0; j < CH / 4; ++j) o.resid[j] = x4[j];
  }
}

template <int EPI, int CH>
__device__ __forceinline__ void epilogue_store(const uint32_t (&r)[CH], const EpiOperands<EPI, CH>& o,
                                               const GemmParams& p, int row, int col0, bool fast) {
  float v[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
  if (!fast) {
    // ragged edge (N not a multiple of the tile) or unaligned rows: scalar, bounds-checked
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int col = col0 + j;
      if (col < p.N) {
        float x = v[j] + (p.bias ? p.bias[col] : 0.f);
        if (EPI == EPI_QGELU_BF16) x = quick_gelu(x);
        if (EPI == EPI_RELU_BF16) x = fmaxf(x, 0.f);
        if (EPI == EPI_RESID_F32) x += p.resid[static_cast<long long>(row) * p.ldr + col];
        if (EPI == EPI_RESID_F32 || EPI == EPI_F32)
          reinterpret_cast<float*>(p.out)[static_cast<long long>(row) * p.ldo + col] = x;
        else
          reinterpret_cast<op16*>(p.out)[static_cast<long long>(row) * p.ldo + col] = to_op16(x);
      }
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < CH / 4; ++j) {
    v[4 * j + 0] += o.bias[j].x;
    v[4 * j + 1] += o.bias[j].y;
    v[4 * j + 2] += o.bias[j].z;
    v[4 * j + 3] += o.bias[j].w;
  }
  if (EPI == EPI_QGELU_BF16) {
#pragma unroll
    for (int j = 0; j < CH; ++j) v[j] = quick_gelu(v[j]);
  } else if (EPI == EPI_RELU_BF16) {
#pragma unroll
    for (int j = 0; j < CH; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
  if (EPI == EPI_RESID_F32 || EPI == EPI_F32) {
    if (EPI == EPI_RESID_F32) {
#pragma unroll
      for (int j = 0; j < CH / 4; ++j) {
        v[4 * j + 0] += o.resid[j].x;
        v[4 * j + 1] += o.resid[j].y;
        v[4 * j + 2] += o.resid[j].z;
        v[4 * j + 3] += o.resid[j].w;
      }
    }
    float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + col0);
#pragma unroll
    for (int j = 0; j < CH / 4; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
    uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<op16*>(p.out) + static_cast<long long>(row) * p.ldo + col0);
#pragma unroll
    for (int j = 0; j < CH / 8; ++j)
      o4[j] = make_uint4(pack16(v[8 * j], v[8 * j + 1]), pack16(v[8 * j + 2], v[8 * j + 3]),
                         pack16(v[8 * j + 4], v[8 * j + 5]), pack16(v[8 * j + 6], v[8 * j + 7]));
  }
}

template <int CH>
__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, uint32_t (&r)[CH]) {
  if constexpr (CH == 32) {
    tmem_ld_32x32(taddr, r);
  } else {
    tmem_ld_32x16(taddr, r);
  }
}

}  // namespace gemm_detail
}  // namespace msclip
