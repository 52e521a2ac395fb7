// Backward of the fused self-attention (attention.cu; Attention_CUST.forward M.py:707-738) for the short sequences of
// MS-CLIP-S B/32 (L = 50, L = 77 causal; head_dim 64) - SURVEY.md section 8f-1.  The reference ships no backward: the
// oracle is torch.autograd through softmax(q k^T / 8 + mask) v.
//
// One CTA per (batch, head).  Q / K / V head slices of the packed QKV tensor and the dO slice of the context gradient are
// staged by TMA (four cp.async.bulk.tensor.3d boxes, 128-B swizzle, rows >= L zero-filled) exactly like the forward.
// Nothing of the forward is stored: the probabilities are recomputed (the whole L x L score block of a head fits in the
// registers of L / 16 warps).  Per warp = 16 query rows:
//   S = Q K^T, P = softmax(S)          (fp32, quad shuffles)
//   dP = dO V^T, D = rowsum(P * dP), dS = P * (dP - D)
//   dQ = dS K                          (written with the 1/8 of the folded q scaling: the gradient of the UNSCALED q)
// P and dS go to shared memory as 16-bit tiles; after one barrier every warp owns 16 key rows:
//   dV = P^T dO,  dK = dS^T Q          (ldmatrix.trans turns the stored [query][key] tiles into A fragments)
// Contractions are warp-level mma.sync m16n8k16 (attention is 1-2 % of the FLOPs; see attention.cu).
#include "common.cuh"
#include "kernels.h"

namespace msclip {

namespace {

constexpr int kHeadDim = 64;
constexpr int kRowBytes = kHeadDim * 2;

__device__ __forceinline__ uint32_t sw128(uint32_t base, int row, int chunk) {
  return base + static_cast<uint32_t>(row) * kRowBytes + (static_cast<uint32_t>(chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32." MSCLIP_MMA_OPERANDS ".f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// PAD = padded sequence length (64 or 80): PAD / 16 warps
template <int PAD, bool CAUSAL>
__global__ void __launch_bounds__(PAD * 2)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                     op16* __restrict__ dqkv, int L, int heads) {
  constexpr int NT = PAD / 8;            // 8-wide key tiles
  constexpr uint32_t kTile = PAD * kRowBytes;
  constexpr int kPitch = PAD * 2 + 16;   // bytes per row of the P / dS tiles: 8 consecutive rows hit 8 different 16-byte slots
  static_assert(PAD % 16 == 0 && (kPitch / 16) % 2 == 1, "tile shapes");
  extern __shared__ uint8_t att_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(att_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * kTile);
  uint8_t* sp_raw = smem + 4 * kTile + 64;
  const int b = blockIdx.x / heads;
  const int h = blockIdx.x % heads;
  const int width = heads * kHeadDim;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    mbar_init(bar, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(bar, 4 * kTile);
#pragma unroll
    for (int which = 0; which < 3; ++which) tma_load_3d(smem + which * kTile, &tmap_qkv, bar, which * width + h * kHeadDim, 0, b);
    tma_load_3d(smem + 3 * kTile, &tmap_do, bar, h * kHeadDim, 0, b);
  }
  __syncthreads();
  mbar_wait(bar, 0, 51);
  const uint32_t sq = smem_u32(smem), sk = sq + kTile, sv = sk + kTile, sdo = sv + kTile;
  const uint32_t sp = smem_u32(sp_raw), sds = sp + PAD * kPitch;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = warp * 16;
  const int g = lane >> 2, t = lane & 3;
  const int mi = lane >> 3, r8 = lane & 7;
  const int row_lo = m0 + g, row_hi = m0 + g + 8;
  constexpr float kLog2e = 1.4426950408889634f;

  // ---------------------------------------------------------------- phase 1: this warp's 16 query rows
  {
    uint32_t qf[4][4], dof[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      ldmatrix_x4(qf[kk], sw128(sq, m0 + r8 + 8 * (mi & 1), 2 * kk + (mi >> 1)));
      ldmatrix_x4(dof[kk], sw128(sdo, m0 + r8 + 8 * (mi & 1), 2 * kk + (mi >> 1)));
    }
    float s[NT][4], dp[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
    }
    // causal: key groups entirely above this warp's last query row are fully masked - neither their scores nor their dP are
    // needed (P = dS = 0 there, and phase 2 never reads a [query tile][key tile] block with key tile > query tile)
    const int jp_end = CAUSAL ? min(NT / 2, m0 / 16 + 1) : NT / 2;
#pragma unroll
    for (int jp = 0; jp < NT / 2; ++jp) {
      if (jp >= jp_end) continue;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t kf[4], vf[4];
        ldmatrix_x4(kf, sw128(sk, 16 * jp + r8 + 8 * (mi >> 1), 2 * kk + (mi & 1)));
        mma_16816(s[2 * jp], qf[kk], kf[0], kf[1]);
        mma_16816(s[2 * jp + 1], qf[kk], kf[2], kf[3]);
        ldmatrix_x4(vf, sw128(sv, 16 * jp + r8 + 8 * (mi >> 1), 2 * kk + (mi & 1)));
        mma_16816(dp[2 * jp], dof[kk], vf[0], vf[1]);
        mma_16816(dp[2 * jp + 1], dof[kk], vf[2], vf[3]);
      }
    }
    // mask + softmax: key column c = 8j + 2t + (e & 1) is live for a row iff c <= min(L - 1, row) (causal) or c <= L - 1
    const int lim_lo = (CAUSAL ? min(L - 1, row_lo) : L - 1) - 2 * t;
    const int lim_hi = (CAUSAL ? min(L - 1, row_hi) : L - 1) - 2 * t;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = 8 * j + (e & 1) <= ((e < 2) ? lim_lo : lim_hi);
        if (!ok) s[j][e] = -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], s[j][e]);
      }
    float ls[2] = {0.f, 0.f};
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
      mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
      if (mx[hh] == -INFINITY) mx[hh] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = fast_ex2(fmaf(s[j][e], kLog2e, -mx[e >> 1] * kLog2e));
        s[j][e] = pv;
        ls[e >> 1] += pv;
      }
    float inv[2];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      ls[hh] += __shfl_xor_sync(0xffffffffu, ls[hh], 1);
      ls[hh] += __shfl_xor_sync(0xffffffffu, ls[hh], 2);
      inv[hh] = 1.0f / ls[hh];
    }
    // query rows >= L are padding: zero probabilities (their dO rows are zero anyway)
    if (row_lo >= L) inv[0] = 0.f;
    if (row_hi >= L) inv[1] = 0.f;
    float dsum[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s[j][e] *= inv[e >> 1];
        dsum[e >> 1] = fmaf(s[j][e], dp[j][e], dsum[e >> 1]);
      }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      dsum[hh] += __shfl_xor_sync(0xffffffffu, dsum[hh], 1);
      dsum[hh] += __shfl_xor_sync(0xffffffffu, dsum[hh], 2);
    }
    // dS = P * (dP - D); P and dS -> shared memory [query][key] for phase 2
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      if (j >= 2 * jp_end) continue;   // masked key groups: nothing to store (never read)
#pragma unroll
      for (int e = 0; e < 4; ++e) dp[j][e] = s[j][e] * (dp[j][e] - dsum[e >> 1]);
      const uint32_t off_lo = static_cast<uint32_t>(row_lo) * kPitch + (8 * j + 2 * t) * 2;
      const uint32_t off_hi = static_cast<uint32_t>(row_hi) * kPitch + (8 * j + 2 * t) * 2;
      *reinterpret_cast<uint32_t*>(sp_raw + off_lo) = pack16(s[j][0], s[j][1]);
      *reinterpret_cast<uint32_t*>(sp_raw + off_hi) = pack16(s[j][2], s[j][3]);
      *reinterpret_cast<uint32_t*>(sp_raw + PAD * kPitch + off_lo) = pack16(dp[j][0], dp[j][1]);
      *reinterpret_cast<uint32_t*>(sp_raw + PAD * kPitch + off_hi) = pack16(dp[j][2], dp[j][3]);
    }
    // dQ = dS K: A fragments straight from the accumulator layout (as P.V in the forward), B = K via ldmatrix.trans
    float dq[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
#pragma unroll
    for (int kk2 = 0; kk2 < NT / 2; ++kk2) {
      if (kk2 >= jp_end) continue;
      uint32_t pa[4];
      pa[0] = pack16(dp[2 * kk2][0], dp[2 * kk2][1]);
      pa[1] = pack16(dp[2 * kk2][2], dp[2 * kk2][3]);
      pa[2] = pack16(dp[2 * kk2 + 1][0], dp[2 * kk2 + 1][1]);
      pa[3] = pack16(dp[2 * kk2 + 1][2], dp[2 * kk2 + 1][3]);
#pragma unroll
      for (int dpi = 0; dpi < 4; ++dpi) {
        uint32_t kf[4];
        ldmatrix_x4_trans(kf, sw128(sk, 16 * kk2 + r8 + 8 * (mi & 1), 2 * dpi + (mi >> 1)));
        mma_16816(dq[2 * dpi], pa, kf[0], kf[1]);
        mma_16816(dq[2 * dpi + 1], pa, kf[2], kf[3]);
      }
    }
    op16* qbase = dqkv + static_cast<long long>(b) * L * (3 * width) + h * kHeadDim;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = 8 * j + 2 * t;
      if (row_lo < L)
        *reinterpret_cast<uint32_t*>(qbase + static_cast<long long>(row_lo) * (3 * width) + col) =
            pack16(dq[j][0] * 0.125f, dq[j][1] * 0.125f);
      if (row_hi < L)
        *reinterpret_cast<uint32_t*>(qbase + static_cast<long long>(row_hi) * (3 * width) + col) =
            pack16(dq[j][2] * 0.125f, dq[j][3] * 0.125f);
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase 2: this warp's 16 key rows
  {
    float dv[8][4], dk[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
      dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
    }
    // contraction over the queries, 16 at a time.  A[m = key][k = query] = tile[query][key]: ldmatrix.trans of the 8x8
    // blocks (queries q0 .. q0+7 | q0+8 .., keys m0 .. m0+7 | m0+8 ..) in A-fragment order
    const int q_first = CAUSAL ? m0 / 16 : 0;  // causal: queries before this key tile see none of its keys
#pragma unroll
    for (int qq = 0; qq < PAD / 16; ++qq) {
      if (qq < q_first) continue;
      const int q0 = qq * 16;
      const uint32_t arow = static_cast<uint32_t>(q0 + r8 + 8 * (mi >> 1)) * kPitch + (m0 + 8 * (mi & 1)) * 2;
      uint32_t pa[4], da[4];
      ldmatrix_x4_trans(pa, sp + arow);
      ldmatrix_x4_trans(da, sds + arow);
#pragma unroll
      for (int dpi = 0; dpi < 4; ++dpi) {
        uint32_t of[4], qf[4];
        ldmatrix_x4_trans(of, sw128(sdo, q0 + r8 + 8 * (mi & 1), 2 * dpi + (mi >> 1)));
        mma_16816(dv[2 * dpi], pa, of[0], of[1]);
        mma_16816(dv[2 * dpi + 1], pa, of[2], of[3]);
        ldmatrix_x4_trans(qf, sw128(sq, q0 + r8 + 8 * (mi & 1), 2 * dpi + (mi >> 1)));
        mma_16816(dk[2 * dpi], da, qf[0], qf[1]);
        mma_16816(dk[2 * dpi + 1], da, qf[2], qf[3]);
      }
    }
    op16* kbase = dqkv + static_cast<long long>(b) * L * (3 * width) + width + h * kHeadDim;
    op16* vbase = kbase + width;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = 8 * j + 2 * t;
      if (row_lo < L) {
        *reinterpret_cast<uint32_t*>(kbase + static_cast<long long>(row_lo) * (3 * width) + col) = pack16(dk[j][0], dk[j][1]);
        *reinterpret_cast<uint32_t*>(vbase + static_cast<long long>(row_lo) * (3 * width) + col) = pack16(dv[j][0], dv[j][1]);
      }
      if (row_hi < L) {
        *reinterpret_cast<uint32_t*>(kbase + static_cast<long long>(row_hi) * (3 * width) + col) = pack16(dk[j][2], dk[j][3]);
        *reinterpret_cast<uint32_t*>(vbase + static_cast<long long>(row_hi) * (3 * width) + col) = pack16(dv[j][2], dv[j][3]);
      }
    }
  }
}

// ---- long sequences (B/16 image tower: L = 197) -------------------------------------------------------------------------
// A head's L x L block no longer fits in the registers of L / 16 warps, and [L][L] tiles of P and dS no longer fit in shared
// memory next to Q / K / V / dO.  Same staging, three passes without any P / dS round trip through shared memory:
//   A  warp = 16 query rows, walks the key groups: S and dP blocks (16 x 16) -> online (max, sum exp, sum exp * dP) -> per row
//      the log-sum-exp (log2 domain) and D = rowsum(P * dP); both go to shared memory for pass B
//   C  the same warp walks the key groups again: P = exp2(S - lse), dS = P * (dP - D), dQ += dS K
//   B  warp = 16 KEY rows, walks the query tiles with the TRANSPOSED blocks S^T = K Q^T, dP^T = V dO^T (rows = keys), so that
//      P^T and dS^T come out of the tensor cores already in A-fragment layout: dV += P^T dO, dK += dS^T Q
// 9 block products per (query tile, key tile) pair instead of 5 - attention is ~1 % of the step's FLOPs.
template <int PAD, bool CAUSAL>
__global__ void __launch_bounds__(PAD * 2, 1)
attention_bwd_long_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                          op16* __restrict__ dqkv, int L, int heads) {
  constexpr int NG = PAD / 16;  // 16-row groups (query tiles = key tiles = warps)
  constexpr uint32_t kTile = PAD * kRowBytes;
  extern __shared__ uint8_t att_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(att_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * kTile);
  float* s_lse = reinterpret_cast<float*>(smem + 4 * kTile + 64);
  float* s_d = s_lse + PAD;
  const int b = blockIdx.x / heads;
  const int h = blockIdx.x % heads;
  const int width = heads * kHeadDim;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    mbar_init(bar, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(bar, 4 * kTile);
#pragma unroll
    for (int which = 0; which < 3; ++which) tma_load_3d(smem + which * kTile, &tmap_qkv, bar, which * width + h * kHeadDim, 0, b);
    tma_load_3d(smem + 3 * kTile, &tmap_do, bar, h * kHeadDim, 0, b);
  }
  __syncthreads();
  mbar_wait(bar, 0, 52);
  const uint32_t sq = smem_u32(smem), sk = sq + kTile, sv = sk + kTile, sdo = sv + kTile;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = warp * 16;
  const int g = lane >> 2, t = lane & 3;
  const int mi = lane >> 3, r8 = lane & 7;
  constexpr float kLog2e = 1.4426950408889634f;
  const int n_live_groups = (L + 15) / 16;

  // ------------------------------------------------------------ passes A and C: this warp's 16 query rows
  {
    const int row_lo = m0 + g, row_hi = m0 + g + 8;
    uint32_t qf[4][4], dof[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      ldmatrix_x4(qf[kk], sw128(sq, m0 + r8 + 8 * (mi & 1), 2 * kk + (mi >> 1)));
      ldmatrix_x4(dof[kk], sw128(sdo, m0 + r8 + 8 * (mi & 1), 2 * kk + (mi >> 1)));
    }
    const int jp_end = CAUSAL ? min(n_live_groups, m0 / 16 + 1) : n_live_groups;
    const int lim_lo = CAUSAL ? min(L - 1, row_lo) : L - 1, lim_hi = CAUSAL ? min(L - 1, row_hi) : L - 1;
    // one 16-key group: scores and dP of the 16 x 16 block, masked scores = -inf
    auto block = [&](int jp, float (&s)[2][4], float (&dp)[2][4]) {
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[n][e] = dp[n][e] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t kf[4], vf[4];
        ldmatrix_x4(kf, sw128(sk, 16 * jp + r8 + 8 * (mi >> 1), 2 * kk + (mi & 1)));
        mma_16816(s[0], qf[kk], kf[0], kf[1]);
        mma_16816(s[1], qf[kk], kf[2], kf[3]);
        ldmatrix_x4(vf, sw128(sv, 16 * jp + r8 + 8 * (mi >> 1), 2 * kk + (mi & 1)));
        mma_16816(dp[0], dof[kk], vf[0], vf[1]);
        mma_16816(dp[1], dof[kk], vf[2], vf[3]);
      }
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = 16 * jp + 8 * n + 2 * t + (e & 1);
          if (col > ((e < 2) ? lim_lo : lim_hi)) s[n][e] = -INFINITY;
        }
    };
    // ---- pass A: online statistics
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f}, d_run[2] = {0.f, 0.f};
    for (int jp = 0; jp < jp_end; ++jp) {
      float s[2][4], dp[2][4];
      block(jp, s, dp);
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) mx[e >> 1] = fmaxf(mx[e >> 1], s[n][e]);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
        const float mnew = fmaxf(m_run[hh], mx[hh]);
        const float msafe = (mnew == -INFINITY) ? 0.f : mnew;
        const float alpha = fast_ex2((m_run[hh] - msafe) * kLog2e);
        l_run[hh] *= alpha;
        d_run[hh] *= alpha;
        m_run[hh] = mnew;
        mx[hh] = msafe;
      }
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = fast_ex2(fmaf(s[n][e], kLog2e, -mx[e >> 1] * kLog2e));
          l_run[e >> 1] += pv;
          d_run[e >> 1] = fmaf(pv, dp[n][e], d_run[e >> 1]);
        }
    }
    float lse2[2], dd[2];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 1);
      l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 2);
      d_run[hh] += __shfl_xor_sync(0xffffffffu, d_run[hh], 1);
      d_run[hh] += __shfl_xor_sync(0xffffffffu, d_run[hh], 2);
      const float msafe = (m_run[hh] == -INFINITY) ? 0.f : m_run[hh];
      lse2[hh] = fmaf(msafe, kLog2e, __log2f(l_run[hh]));
      dd[hh] = d_run[hh] / l_run[hh];
    }
    if (t == 0) {
      s_lse[row_lo] = lse2[0];
      s_lse[row_hi] = lse2[1];
      s_d[row_lo] = dd[0];
      s_d[row_hi] = dd[1];
    }
    // ---- pass C: dQ = sum over key groups of dS K
    float dq[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
    for (int jp = 0; jp < jp_end; ++jp) {
      float s[2][4], dp[2][4];
      block(jp, s, dp);
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = fast_ex2(fmaf(s[n][e], kLog2e, -lse2[e >> 1]));
          dp[n][e] = pv * (dp[n][e] - dd[e >> 1]);
        }
      uint32_t pa[4];
      pa[0] = pack16(dp[0][0], dp[0][1]);
      pa[1] = pack16(dp[0][2], dp[0][3]);
      pa[2] = pack16(dp[1][0], dp[1][1]);
      pa[3] = pack16(dp[1][2], dp[1][3]);
#pragma unroll
      for (int dpi = 0; dpi < 4; ++dpi) {
        uint32_t kf[4];
        ldmatrix_x4_trans(kf, sw128(sk, 16 * jp + r8 + 8 * (mi & 1), 2 * dpi + (mi >> 1)));
        mma_16816(dq[2 * dpi], pa, kf[0], kf[1]);
        mma_16816(dq[2 * dpi + 1], pa, kf[2], kf[3]);
      }
    }
    op16* qbase = dqkv + static_cast<long long>(b) * L * (3 * width) + h * kHeadDim;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = 8 * j + 2 * t;
      if (row_lo < L)
        *reinterpret_cast<uint32_t*>(qbase + static_cast<long long>(row_lo) * (3 * width) + col) =
            pack16(dq[j][0] * 0.125f, dq[j][1] * 0.125f);
      if (row_hi < L)
        *reinterpret_cast<uint32_t*>(qbase + static_cast<long long>(row_hi) * (3 * width) + col) =
            pack16(dq[j][2] * 0.125f, dq[j][3] * 0.125f);
    }
  }
  __syncthreads();

  // ------------------------------------------------------------ pass B: this warp's 16 key rows
  if (m0 < L) {
    const int key_lo = m0 + g, key_hi = m0 + g + 8;
    uint32_t ka[4][4], va[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      ldmatrix_x4(ka[kk], sw128(sk, m0 + r8 + 8 * (mi & 1), 2 * kk + (mi >> 1)));
      ldmatrix_x4(va[kk], sw128(sv, m0 + r8 + 8 * (mi & 1), 2 * kk + (mi >> 1)));
    }
    float dv[8][4], dk[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
      dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
    }
    const int qt_begin = CAUSAL ? m0 / 16 : 0;
    for (int qt = qt_begin; qt < n_live_groups; ++qt) {
      const int q0 = qt * 16;
      float st[2][4], dpt[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) st[n][e] = dpt[n][e] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t qb[4], ob[4];
        ldmatrix_x4(qb, sw128(sq, q0 + r8 + 8 * (mi >> 1), 2 * kk + (mi & 1)));
        mma_16816(st[0], ka[kk], qb[0], qb[1]);
        mma_16816(st[1], ka[kk], qb[2], qb[3]);
        ldmatrix_x4(ob, sw128(sdo, q0 + r8 + 8 * (mi >> 1), 2 * kk + (mi & 1)));
        mma_16816(dpt[0], va[kk], ob[0], ob[1]);
        mma_16816(dpt[1], va[kk], ob[2], ob[3]);
      }
      // element (n, e): key = m0 + g + 8 (e >> 1), query = q0 + 8 n + 2 t + (e & 1)
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = (e < 2) ? key_lo : key_hi;
          const int qr = q0 + 8 * n + 2 * t + (e & 1);
          const bool live = key < L && qr < L && (!CAUSAL || key <= qr);
          const float pv = live ? fast_ex2(fmaf(st[n][e], kLog2e, -s_lse[qr])) : 0.f;
          st[n][e] = pv;
          dpt[n][e] = pv * (dpt[n][e] - s_d[qr]);
        }
      uint32_t pa[4], da[4];
      pa[0] = pack16(st[0][0], st[0][1]);
      pa[1] = pack16(st[0][2], st[0][3]);
      pa[2] = pack16(st[1][0], st[1][1]);
      pa[3] = pack16(st[1][2], st[1][3]);
      da[0] = pack16(dpt[0][0], dpt[0][1]);
      da[1] = pack16(dpt[0][2], dpt[0][3]);
      da[2] = pack16(dpt[1][0], dpt[1][1]);
      da[3] = pack16(dpt[1][2], dpt[1][3]);
#pragma unroll
      for (int dpi = 0; dpi < 4; ++dpi) {
        uint32_t of[4], qf2[4];
        ldmatrix_x4_trans(of, sw128(sdo, q0 + r8 + 8 * (mi & 1), 2 * dpi + (mi >> 1)));
        mma_16816(dv[2 * dpi], pa, of[0], of[1]);
        mma_16816(dv[2 * dpi + 1], pa, of[2], of[3]);
        ldmatrix_x4_trans(qf2, sw128(sq, q0 + r8 + 8 * (mi & 1), 2 * dpi + (mi >> 1)));
        mma_16816(dk[2 * dpi], da, qf2[0], qf2[1]);
        mma_16816(dk[2 * dpi + 1], da, qf2[2], qf2[3]);
      }
    }
    op16* kbase = dqkv + static_cast<long long>(b) * L * (3 * width) + width + h * kHeadDim;
    op16* vbase = kbase + width;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = 8 * j + 2 * t;
      if (key_lo < L) {
        *reinterpret_cast<uint32_t*>(kbase + static_cast<long long>(key_lo) * (3 * width) + col) = pack16(dk[j][0], dk[j][1]);
        *reinterpret_cast<uint32_t*>(vbase + static_cast<long long>(key_lo) * (3 * width) + col) = pack16(dv[j][0], dv[j][1]);
      }
      if (key_hi < L) {
        *reinterpret_cast<uint32_t*>(kbase + static_cast<long long>(key_hi) * (3 * width) + col) = pack16(dk[j][2], dk[j][3]);
        *reinterpret_cast<uint32_t*>(vbase + static_cast<long long>(key_hi) * (3 * width) + col) = pack16(dv[j][2], dv[j][3]);
      }
    }
  }
}

template <int PAD, bool CAUSAL>
int launch_bwd_long(const op16* qkv, const op16* dctx, op16* dqkv, int batch, int L, int heads, cudaStream_t stream) {
  constexpr int smem = 4 * PAD * kRowBytes + 64 + 2 * PAD * 4 + 1024;
  CUtensorMap tq, td;
  MSCLIP_TRY(make_tmap_op16_3d(&tq, qkv, batch, L, 3ull * heads * kHeadDim, 3ull * heads * kHeadDim, PAD));
  MSCLIP_TRY(make_tmap_op16_3d(&td, dctx, batch, L, 1ull * heads * kHeadDim, 1ull * heads * kHeadDim, PAD));
  static bool configured = false;
  if (!configured) {
    MSCLIP_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_long_kernel<PAD, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  attention_bwd_long_kernel<PAD, CAUSAL><<<batch * heads, PAD * 2, smem, stream>>>(tq, td, dqkv, L, heads);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int PAD, bool CAUSAL>
int launch_bwd_variant(const op16* qkv, const op16* dctx, op16* dqkv, int batch, int L, int heads, cudaStream_t stream) {
  constexpr int kPitch = PAD * 2 + 16;
  constexpr int smem = 4 * PAD * kRowBytes + 64 + 2 * PAD * kPitch + 1024;
  CUtensorMap tq, td;
  MSCLIP_TRY(make_tmap_op16_3d(&tq, qkv, batch, L, 3ull * heads * kHeadDim, 3ull * heads * kHeadDim, PAD));
  MSCLIP_TRY(make_tmap_op16_3d(&td, dctx, batch, L, 1ull * heads * kHeadDim, 1ull * heads * kHeadDim, PAD));
  static bool configured = false;
  if (!configured) {
    MSCLIP_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<PAD, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  attention_bwd_kernel<PAD, CAUSAL><<<batch * heads, PAD * 2, smem, stream>>>(tq, td, dqkv, L, heads);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

// qkv op16 [B*L, 3*64*heads] (q pre-scaled by 1/8, as the forward consumes it), dctx op16 [B*L, 64*heads] = gradient of the
// attention output -> dqkv op16 [B*L, 3*64*heads] = gradient of the UNSCALED (q | k | v).  L <= 80: one-pass kernel with P / dS
// tiles in shared memory; 80 < L <= 208 (B/16 image tower): three-pass kernel
int launch_attention_bwd(const op16* qkv, const op16* dctx, op16* dqkv, int batch, int L, int heads, int causal,
                         cudaStream_t stream) {
  if (batch <= 0) return 0;
  MSCLIP_REQUIRE(L >= 1 && L <= 208, "attention backward: sequence length must be in [1, 208]");
  MSCLIP_REQUIRE(heads >= 1, "attention backward: heads must be positive");
  if (L <= 64)
    return causal ? launch_bwd_variant<64, true>(qkv, dctx, dqkv, batch, L, heads, stream)
                  : launch_bwd_variant<64, false>(qkv, dctx, dqkv, batch, L, heads, stream);
  if (L <= 80)
    return causal ? launch_bwd_variant<80, true>(qkv, dctx, dqkv, batch, L, heads, stream)
                  : launch_bwd_variant<80, false>(qkv, dctx, dqkv, batch, L, heads, stream);
  return causal ? launch_bwd_long<208, true>(qkv, dctx, dqkv, batch, L, heads, stream)
                : launch_bwd_long<208, false>(qkv, dctx, dqkv, batch, L, heads, stream);
}

}  // namespace msclip
