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

// A operand = implicit im2col of up to two NHWC sources onto one Ho x Wo output grid (conv_tma: the producer issues
// one im2col-mode TMA per k-block = one filter tap x 64 channels).  K index = (source, tap, 64-channel block, channel).
struct ConvGeom {
  int Ho, Wo;
  uint32_t hw_mul, hw_shr, wo_mul, wo_shr;  // n / (Ho*Wo), n / Wo as multiply-high + shift
  int kb_begin1;                            // first k-block of source 1 (== number of k-blocks when there is one source)
  int ksize[2], stride[2], pad[2], cblk[2]; // cblk = 64-channel blocks per tap
};

// q = n / d for 0 <= n < 2^31 as umulhi(n, mul) >> shr (d == 1: mul = 0 selects the identity)
inline void find_divisor(uint32_t d, uint32_t* mul, uint32_t* shr) {
  if (d == 1) {
    *mul = 0;
    *shr = 0;
    return;
  }
  uint32_t lg = 0;
  while ((1ull << lg) < d) ++lg;
  const uint32_t p = 31 + lg;
  *mul = static_cast<uint32_t>(((1ull << p) + d - 1) / d);
  *shr = p - 32;
}
#ifdef __CUDACC__
__device__ __forceinline__ int fast_div(int n, uint32_t mul, uint32_t shr) {
  return mul != 0 ? static_cast<int>(__umulhi(static_cast<uint32_t>(n), mul) >> shr) : n;
}
#endif

struct GemmParams {
  int M, N, K;
  int tiles_n, total_tiles;
  const float* bias;
  void* out;
  const float* resid;
  long long ldo, ldr;
  float alpha;  // acc is scaled by alpha before the bias (similarity logits: exp(logit_scale))
  int vec_ok;   // rows of out / resid keep 16-byte alignment -> vector stores
  const op16* aux16;     // EPI_DGELU_BF16: u (fc1 pre-activation), pitch ldo
  op16* out2;            // EPI_DGELU_BF16: quickgelu(u), pitch ldo
  uint32_t operand_fmt;  // kind::f16 A/B format field (0 = F16, 1 = BF16): the loss backward runs fp16 operands in both builds
  long long split_stride;  // != 0: column tile j writes a separate [M, BN] matrix at out + j * split_stride (elements)
  // LayerNorm folded into the GEMMs around it (kernel template parameter LN, see below)
  const float* ln_in;    // LN = 1: row records of A's rows; LN = 2: records of the residual input (its mean = new shift)
  float* ln_out;         // LN = 2: row records of the new residual stream
  op16* out16;           // LN = 2: op16(x_new - shift), the next GEMM's A operand, pitch ldo16
  const float* colsum;   // LN = 1: sum_k W'[n][k] of the packed (gamma-folded) weight
  long long ldo16;
  ConvGeom conv;         // CONV kernels only
  // LN = 3 (out-proj, fc2; N = 768): four extra warps per CTA re-read every finished 128 x 768 row block of `out`
  // from L2 and write its LayerNorm (the next GEMM's A operand) - the separate LayerNorm pass over HBM disappears
  const float* lnw_gamma;
  const float* lnw_beta;
  op16* lnw_out;
  long long lnw_ld;
  uint32_t* lnw_counters;  // [row blocks][2]: column tiles stored so far (self-cleaning)
};

// ---- LayerNorm folding --------------------------------------------------------------------------------------
// LN(x) . W^T + b  ==  rstd * (xc . W'^T - mean_c * colsum(W')) + b'   with  xc = x - shift (any per-row shift),
// mean_c = mean(xc), rstd = 1 / sqrt(var(xc) + eps), W' = W * diag(gamma), b' = b + W . beta.
// The GEMM that produces the residual stream (LN = 2: out-proj, fc2) therefore also emits op16(x - shift) and a
// 64-byte record per row - the shift (mean of the row before this update, so the centred values stay small and
// the 16-bit rounding error does not grow with a common-mode offset) and (sum, sum of squares) of the centred
// values per 128-column slice - and the GEMM that consumes LN(x) (LN = 1: QKV, fc1) reads the centred rows as its A
// operand and finishes the normalisation in its epilogue.  The separate LayerNorm pass over the fp32 stream
// (6 bytes per element, 47 launches per step) disappears.
constexpr int kLnRec = 16;    // floats per row record: [0] shift, [4 + 2*slot] sum, [5 + 2*slot] sum of squares
constexpr int kLnSlots = 6;   // slot = 256-column tile * 2 + epilogue half (N = 768)
constexpr float kLnEps = 1e-12f;  // inside the sqrt (M.py:217)

struct LnRow {   // LN = 1: statistics of this thread's row
  float mean_c, rstd;
};
struct LnEmit {  // LN = 2: running sums of the two rows of the lane pair (pieces held after the swap)
  float shift_e, shift_o, s1e, s2e, s1o, s2o;
};

__device__ __forceinline__ void ln_load_record(const float* rec, float inv_width, float& shift, float& mean_c, float& rstd) {
  const float4* r4 = reinterpret_cast<const float4*>(rec);
  const float4 a = __ldg(r4), b = __ldg(r4 + 1), c = __ldg(r4 + 2), d = __ldg(r4 + 3);
  shift = a.x;
  const float s1 = (b.x + b.z) + (c.x + c.z) + (d.x + d.z);
  const float s2 = (b.y + b.w) + (c.y + c.w) + (d.y + d.w);
  mean_c = s1 * inv_width;
  const float var = fmaxf(s2 * inv_width - mean_c * mean_c, 0.f);
  rstd = 1.0f / sqrtf(var + kLnEps);
}

__device__ __forceinline__ float quick_gelu(float x) {
  // x * sigmoid(1.702 x)   (M.py:224)
#ifndef MSCLIP_QGELU_EXP
  // sigmoid(z) = 0.5 + 0.5 tanh(z / 2): ONE MUFU op per element (tanh.approx, 2^-11 relative) instead of two - the fc1
  // epilogue is MUFU-co-limited (65 536 MUFU ops per 128 x 256 tile at 16 / clk against ~6 100 clk of MMA): text fc1
  // 1.203 -> 1.156 ms, image fc1 0.820 -> 0.729 ms, step 132.3 -> 130.5 ms (round 2).  The absolute error, <= 2.5e-4 |x|,
  // stays below the 16-bit rounding of the output and every model-level parity gate holds unchanged
  // (tests/test_model_gpu.py); -DMSCLIP_QGELU_EXP (MSCLIP_QGELU_EXP=1 python -m msclip_b200.build) restores ex2 + rcp.
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
#else
  // 1 / (1 + 2^(-1.702 log2(e) x)) with MUFU.EX2 + MUFU.RCP (2^-22 relative error each)
  return x * fast_rcp(1.0f + fast_ex2(-2.4554669595930157f * x));
#endif
}

// ---- fused epilogue ---------------------------------------------------------------------------------------
// TMEM hands every thread one accumulator row (32 lanes = 32 rows, CH consecutive columns each).  Storing that
// layout directly makes each warp-wide 16-byte access touch 32 different sectors and use half of each, which
// doubles the L2 sector traffic of the epilogue.  Lane pairs therefore swap half of their values first: after the
// swap lanes 2i / 2i+1 hold the even / odd 16-byte pieces of BOTH rows of the pair, so each warp-wide access
// covers whole 32-byte sectors (rows are processed as "even row" then "odd row" of every pair).

// Operands the epilogue needs from global memory for one chunk (bias slice, residual pieces in the swapped
// layout); fetched before the accumulator chunk is waited for so their latency overlaps the TMEM load.
template <int EPI, int CH, int LN = 0>
struct EpiOperands {
  float4 bias[CH / 4];
  float4 resid_even[(EPI == EPI_RESID_F32) ? CH / 8 : 1];  // even row of the lane pair, pieces 2j + (lane & 1)
  float4 resid_odd[(EPI == EPI_RESID_F32) ? CH / 8 : 1];   // odd row of the lane pair
  float4 colsum[(LN == 1) ? CH / 4 : 1];
  uint4 aux[(EPI == EPI_DGELU_BF16) ? CH / 8 : 1];  // this thread's own row of u, CH consecutive columns
};

template <int EPI, int CH, int LN = 0>
__device__ __forceinline__ void epilogue_prefetch(EpiOperands<EPI, CH, LN>& o, const GemmParams& p, int row, int col0,
                                                  bool fast) {
  if (!fast) return;
  if (LN == 1) {
    const float4* c4 = reinterpret_cast<const float4*>(p.colsum + col0);
#pragma unroll
    for (int j = 0; j < CH / 4; ++j) o.colsum[j] = __ldg(c4 + j);
  }
  if (p.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
    for (int j = 0; j < CH / 4; ++j) o.bias[j] = __ldg(b4 + j);
  } else {
#pragma unroll
    for (int j = 0; j < CH / 4; ++j) o.bias[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (EPI == EPI_DGELU_BF16) {
    if (row < p.M) {
      const uint4* up = reinterpret_cast<const uint4*>(p.aux16 + static_cast<long long>(row) * p.ldo + col0);
#pragma unroll
      for (int j = 0; j < CH / 8; ++j) o.aux[j] = __ldg(up + j);
    } else {
#pragma unroll
      for (int j = 0; j < CH / 8; ++j) o.aux[j] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (EPI == EPI_RESID_F32) {
    const int par = threadIdx.x & 1;
    const int row_e = row & ~1, row_o = row | 1;
    const float4* xe = reinterpret_cast<const float4*>(p.resid + static_cast<long long>(row_e) * p.ldr + col0);
    const float4* xo = reinterpret_cast<const float4*>(p.resid + static_cast<long long>(row_o) * p.ldr + col0);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < CH / 8; ++j) {
      o.resid_even[j] = row_e < p.M ? xe[2 * j + par] : z;
      o.resid_odd[j] = row_o < p.M ? xo[2 * j + par] : z;
    }
  }
}

// 16-bit results of a lane's own row (CH columns as CH / 2 packed words) -> global memory through the lane-pair swap
template <int CH>
__device__ __forceinline__ void store16_swapped(const uint32_t (&u)[CH / 2], op16* base, long long ld, int row_e, int row_o, int col0,
                                                bool ok_e, bool ok_o, bool par) {
  uint4* oe = reinterpret_cast<uint4*>(base + static_cast<long long>(row_e) * ld + col0);
  uint4* oo = reinterpret_cast<uint4*>(base + static_cast<long long>(row_o) * ld + col0);
#pragma unroll
  for (int j = 0; j < CH / 16; ++j) {
    uint32_t e[4], d[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const uint32_t lo = u[8 * j + t], hi = u[8 * j + 4 + t];
      const uint32_t keep = par ? hi : lo;
      const uint32_t recv = __shfl_xor_sync(0xffffffffu, par ? lo : hi, 1);
      e[t] = par ? recv : keep;
      d[t] = par ? keep : recv;
    }
    if (ok_e) oe[2 * j + par] = make_uint4(e[0], e[1], e[2], e[3]);
    if (ok_o) oo[2 * j + par] = make_uint4(d[0], d[1], d[2], d[3]);
  }
}

// Must be called by all 32 lanes of the warp (it shuffles); rows >= M are masked inside.
template <int EPI, int CH, int LN = 0>
__device__ __forceinline__ void epilogue_store(const uint32_t (&r)[CH], const EpiOperands<EPI, CH, LN>& o,
                                               const GemmParams& p, int row, int col0, bool fast,
                                               const LnRow& lnr = LnRow(), LnEmit* lne = nullptr) {
  float v[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
  if (!fast) {
    // ragged edge (N not a multiple of the tile) or unaligned rows: scalar, bounds-checked (never taken with LN != 0:
    // the launcher requires aligned, tile-sized problems there)
    if (row >= p.M) return;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int col = col0 + j;
      if (col < p.N) {
        float x = v[j] + (p.bias ? p.bias[col] : 0.f);
        if (EPI == EPI_QGELU_BF16 || EPI == EPI_QGELU_DUAL_BF16) x = quick_gelu(x);
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
  if (LN == 1) {
    // finish the LayerNorm of this row: rstd * (xc . W'^T - mean_c * colsum)
    const float nm = -lnr.mean_c;
#pragma unroll
    for (int j = 0; j < CH / 4; ++j) {
      v[4 * j + 0] = lnr.rstd * fmaf(nm, o.colsum[j].x, v[4 * j + 0]);
      v[4 * j + 1] = lnr.rstd * fmaf(nm, o.colsum[j].y, v[4 * j + 1]);
      v[4 * j + 2] = lnr.rstd * fmaf(nm, o.colsum[j].z, v[4 * j + 2]);
      v[4 * j + 3] = lnr.rstd * fmaf(nm, o.colsum[j].w, v[4 * j + 3]);
    }
  }
#pragma unroll
  for (int j = 0; j < CH / 4; ++j) {
    v[4 * j + 0] += o.bias[j].x;
    v[4 * j + 1] += o.bias[j].y;
    v[4 * j + 2] += o.bias[j].z;
    v[4 * j + 3] += o.bias[j].w;
  }
  const bool par = (threadIdx.x & 1) != 0;
  const int row_e = row & ~1, row_o = row | 1;
  const bool ok_e = row_e < p.M, ok_o = row_o < p.M;
  if (EPI == EPI_QGELU_DUAL_BF16) {
    uint32_t pre[CH / 2];
#pragma unroll
    for (int j = 0; j < CH / 2; ++j) pre[j] = pack16(v[2 * j], v[2 * j + 1]);
    store16_swapped<CH>(pre, p.out2, p.ldo, row_e, row_o, col0, ok_e, ok_o, par);
  }
  if (EPI == EPI_QGELU_BF16 || EPI == EPI_QGELU_DUAL_BF16) {
#pragma unroll
    for (int j = 0; j < CH; ++j) v[j] = quick_gelu(v[j]);
  } else if (EPI == EPI_RELU_BF16) {
#pragma unroll
    for (int j = 0; j < CH; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
  if (EPI == EPI_RESID_F32 || EPI == EPI_F32) {
    // pieces = float4 (4 columns); lane parity selects pieces 2j + par of both rows
    float4* oe = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<long long>(row_e) * p.ldo + col0);
    float4* oo = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<long long>(row_o) * p.ldo + col0);
#pragma unroll
    for (int j = 0; j < CH / 8; ++j) {
      float e[4], d[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float lo = v[8 * j + t], hi = v[8 * j + 4 + t];  // pieces 2j and 2j+1 of this lane's own row
        const float keep = par ? hi : lo;
        const float recv = __shfl_xor_sync(0xffffffffu, par ? lo : hi, 1);
        e[t] = par ? recv : keep;  // even row of the pair
        d[t] = par ? keep : recv;  // odd row of the pair
      }
      if (EPI == EPI_RESID_F32) {
        e[0] += o.resid_even[j].x; e[1] += o.resid_even[j].y; e[2] += o.resid_even[j].z; e[3] += o.resid_even[j].w;
        d[0] += o.resid_odd[j].x;  d[1] += o.resid_odd[j].y;  d[2] += o.resid_odd[j].z;  d[3] += o.resid_odd[j].w;
      }
      if (ok_e) oe[2 * j + par] = make_float4(e[0], e[1], e[2], e[3]);
      if (ok_o) oo[2 * j + par] = make_float4(d[0], d[1], d[2], d[3]);
      if (LN == 2) {
        // centred 16-bit copy for the next GEMM + this lane's share of the row sums
        uint2* be = reinterpret_cast<uint2*>(p.out16 + static_cast<long long>(row_e) * p.ldo16 + col0);
        uint2* bo = reinterpret_cast<uint2*>(p.out16 + static_cast<long long>(row_o) * p.ldo16 + col0);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          e[t] -= lne->shift_e;
          d[t] -= lne->shift_o;
          lne->s1e += e[t];
          lne->s2e = fmaf(e[t], e[t], lne->s2e);
          lne->s1o += d[t];
          lne->s2o = fmaf(d[t], d[t], lne->s2o);
        }
        if (ok_e) be[2 * j + par] = make_uint2(pack16(e[0], e[1]), pack16(e[2], e[3]));
        if (ok_o) bo[2 * j + par] = make_uint2(pack16(d[0], d[1]), pack16(d[2], d[3]));
      }
    }
  } else {
    // pieces = uint4 (8 columns of 16-bit values)
    uint32_t u[CH / 2];
    if (EPI == EPI_DGELU_BF16) {
      // v = d a (gradient of the fc2 input); aux = u.  sigmoid with one MUFU op as in quick_gelu()
      uint32_t act[CH / 2];
      const op162* uu = reinterpret_cast<const op162*>(o.aux);
#pragma unroll
      for (int j = 0; j < CH / 2; ++j) {
        // with t = tanh(0.851 x): sigmoid s = (1 + t) / 2, quickgelu(x) = x s = hx + hx t (hx = x / 2, as in quick_gelu()),
        // quickgelu'(x) = s + 1.702 x s (1 - s) = (1 + t) / 2 + (0.851 x / 2) (1 - t^2)
        const float2 x = op162_to_float2(uu[j]);
        const float q0 = 0.851f * x.x, q1 = 0.851f * x.y;
        float t0, t1;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(q0));
        asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(q1));
        const float h0 = 0.5f * x.x, h1 = 0.5f * x.y;
        act[j] = pack16(fmaf(h0, t0, h0), fmaf(h1, t1, h1));
        const float g0 = fmaf(0.5f * q0, fmaf(-t0, t0, 1.0f), fmaf(0.5f, t0, 0.5f));
        const float g1 = fmaf(0.5f * q1, fmaf(-t1, t1, 1.0f), fmaf(0.5f, t1, 0.5f));
        u[j] = pack16(v[2 * j] * g0, v[2 * j + 1] * g1);
      }
      store16_swapped<CH>(act, p.out2, p.ldo, row_e, row_o, col0, ok_e, ok_o, par);
    } else {
#pragma unroll
      for (int j = 0; j < CH / 2; ++j) u[j] = pack16(v[2 * j], v[2 * j + 1]);
    }
    store16_swapped<CH>(u, reinterpret_cast<op16*>(p.out), p.ldo, row_e, row_o, col0, ok_e, ok_o, par);
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
