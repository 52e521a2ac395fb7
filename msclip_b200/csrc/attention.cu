// Fused self-attention for the short sequences of MS-CLIP-S (L = 50 / 77 causal / 197, head_dim 64):
// softmax(q k^T [+ causal mask]) v per (batch, head), replacing the two bmm + softmax + five copies of
// Attention_CUST.forward (M.py:707-738).  The 1/sqrt(64) scaling of q (M.py:707) is folded into the packed
// QKV weights, heads are addressed by pointer arithmetic on the [B*L, 3*768] QKV matrix (M.py:709-711),
// scores never leave registers and the softmax is fp32 with quad shuffles.
//
// The Q / K / V head slices of a (batch, head) item ([L, 64] windows of the [B, L, 3*768] QKV tensor) are staged by
// TMA: one thread issues three cp.async.bulk.tensor boxes of 64 columns x KVPAD rows (128-byte rows, 128-B swizzle,
// rows >= L arrive zero-filled) that complete on an mbarrier; no thread spends instructions on copies.  Every warp owns
// 16 query rows and walks the keys in chunks (online softmax across chunks for L = 197); ldmatrix addresses apply the
// same XOR swizzle (conflict-free without padding).
// Attention is 1-2 % of the path's FLOPs and the per-head problems are far
// smaller than one 128-row tcgen05 tile, so the contractions use warp-level mma.sync m16n8k16 (op16 in,
// fp32 accumulate); the GEMMs that carry the other 98 % are tcgen05 (gemm.cu).
#include "common.cuh"
#include "kernels.h"

namespace msclip {

namespace {

constexpr int kHeadDim = 64;
constexpr int kRowBytes = kHeadDim * 2;  // one head-slice row = 128 B = one swizzle row

// shared-memory address of the 16-byte chunk `chunk` (0..7) of row `row` in a 128-B-swizzled tile (1024-B aligned base)
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

// scores, softmax and P.V for the 16 query rows of this warp; Q/K/V head slices are in (swizzled) shared memory
template <int QPAD, int KC, int NCHUNK, bool CAUSAL>
__device__ __forceinline__ void attention_compute(uint32_t sq, uint32_t sk, uint32_t sv, op16* out_base, int L, int width) {
  constexpr int NT = KC / 8;  // score n-tiles per chunk
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = warp * 16;
  if (m0 >= L) return;  // this warp has no query rows (it still took part in the staging)
  const int g = lane >> 2, t = lane & 3;
  const int mi = lane >> 3, r8 = lane & 7;

  uint32_t qf[4][4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
    ldmatrix_x4(qf[kk], sw128(sq, m0 + r8 + 8 * (mi & 1), 2 * kk + (mi >> 1)));

  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  const int row_lo = m0 + g, row_hi = m0 + g + 8;
  constexpr float kLog2e = 1.4426950408889634f;

#pragma unroll
  for (int ch = 0; ch < NCHUNK; ++ch) {
    const int kv0 = ch * KC;
    if (kv0 >= L) break;
    if (CAUSAL && kv0 > m0 + 15) break;
    // number of 16-wide key groups this warp needs in this chunk
    int ng = KC / 16;
    {
      int last = L - 1;
      if (CAUSAL && m0 + 15 < last) last = m0 + 15;
      const int need = (last - kv0) / 16 + 1;
      if (need < ng) ng = need;
    }
    float s[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int jp = 0; jp < NT / 2; ++jp) {
      if (jp < ng) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint32_t kf[4];
          ldmatrix_x4(kf, sw128(sk, kv0 + 16 * jp + r8 + 8 * (mi >> 1), 2 * kk + (mi & 1)));
          mma_16816(s[2 * jp], qf[kk], kf[0], kf[1]);
          mma_16816(s[2 * jp + 1], qf[kk], kf[2], kf[3]);
        }
      }
    }
    // mask + chunk max.  Key column c = kv0 + 8j + 2t + (e & 1) is live for a row iff c <= min(L - 1, row) (causal) or
    // c <= L - 1; columns of key groups beyond `ng` lie past that bound as well, so one compare per score suffices:
    // 8j + (e & 1) <= lim - kv0 - 2t, with the left side a compile-time constant.
    const int lim_lo = (CAUSAL ? min(L - 1, row_lo) : L - 1) - kv0 - 2 * t;
    const int lim_hi = (CAUSAL ? min(L - 1, row_hi) : L - 1) - kv0 - 2 * t;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = 8 * j + (e & 1) <= ((e < 2) ? lim_lo : lim_hi);
        if (!ok) s[j][e] = -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], s[j][e]);
      }
    }
    float alpha[2], mnew[2];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
      mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
      mnew[hh] = fmaxf(m_run[hh], mx[hh]);
      // rows of the padding region may be fully masked in a chunk: keep the maths finite
      const float msafe = (mnew[hh] == -INFINITY) ? 0.f : mnew[hh];
      alpha[hh] = fast_ex2((m_run[hh] - msafe) * kLog2e);
      m_run[hh] = mnew[hh];
      mnew[hh] = msafe;
    }
    float ls[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = fast_ex2(fmaf(s[j][e], kLog2e, -mnew[e >> 1] * kLog2e));  // one FFMA + MUFU.EX2
        s[j][e] = pv;
        ls[e >> 1] += pv;
      }
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) l_run[hh] = l_run[hh] * alpha[hh] + ls[hh];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= alpha[0];
      o[j][1] *= alpha[0];
      o[j][2] *= alpha[1];
      o[j][3] *= alpha[1];
    }
    // O += P V
#pragma unroll
    for (int kk2 = 0; kk2 < NT / 2; ++kk2) {
      if (kk2 < ng) {
        uint32_t pa[4];
        pa[0] = pack16(s[2 * kk2][0], s[2 * kk2][1]);
        pa[1] = pack16(s[2 * kk2][2], s[2 * kk2][3]);
        pa[2] = pack16(s[2 * kk2 + 1][0], s[2 * kk2 + 1][1]);
        pa[3] = pack16(s[2 * kk2 + 1][2], s[2 * kk2 + 1][3]);
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t vf[4];
          ldmatrix_x4_trans(vf, sw128(sv, kv0 + 16 * kk2 + r8 + 8 * (mi & 1), 2 * dp + (mi >> 1)));
          mma_16816(o[2 * dp], pa, vf[0], vf[1]);
          mma_16816(o[2 * dp + 1], pa, vf[2], vf[3]);
        }
      }
    }
  }

  // quad-reduce the row sums, normalise, store
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 1);
    l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 2);
  }
  const float inv_lo = 1.0f / l_run[0], inv_hi = 1.0f / l_run[1];
  op16* obase = out_base;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = 8 * j + 2 * t;
    if (row_lo < L)
      *reinterpret_cast<uint32_t*>(obase + static_cast<long long>(row_lo) * width + col) =
          pack16(o[j][0] * inv_lo, o[j][1] * inv_lo);
    if (row_hi < L)
      *reinterpret_cast<uint32_t*>(obase + static_cast<long long>(row_hi) * width + col) =
          pack16(o[j][2] * inv_hi, o[j][3] * inv_hi);
  }
}

// one (batch, head) item = three boxes [KVPAD rows][64 columns] at columns which * width + h * 64 of batch entry b
template <int KVPAD>
__device__ __forceinline__ void tma_stage_item(const CUtensorMap* tmap, uint8_t* buf, uint64_t* bar, int b, int h, int width) {
  constexpr uint32_t kTile = KVPAD * kRowBytes;
  mbar_arrive_expect_tx(bar, 3 * kTile);
#pragma unroll
  for (int which = 0; which < 3; ++which) tma_load_3d(buf + which * kTile, tmap, bar, which * width + h * kHeadDim, 0, b);
}

// one CTA per (batch, head): used for L = 197, where two staging buffers would not fit next to each other
template <int QPAD, int KC, int NCHUNK, bool CAUSAL>
__global__ void __launch_bounds__(QPAD * 2)
attention_kernel(const __grid_constant__ CUtensorMap tmap, op16* __restrict__ out, int L, int heads) {
  constexpr int KVPAD = KC * NCHUNK;
  constexpr uint32_t kTile = KVPAD * kRowBytes;
  static_assert(QPAD % 16 == 0 && KC % 16 == 0 && KVPAD >= QPAD && KVPAD % 8 == 0, "tile shapes");
  extern __shared__ uint8_t att_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(att_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 3 * kTile);
  const int b = blockIdx.x / heads;
  const int h = blockIdx.x % heads;
  const int width = heads * kHeadDim;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(bar, 1);
    fence_mbar_init();
    tma_stage_item<KVPAD>(&tmap, smem, bar, b, h, width);
  }
  __syncthreads();
  mbar_wait(bar, 0, 31);
  const uint32_t s0 = smem_u32(smem);
  attention_compute<QPAD, KC, NCHUNK, CAUSAL>(s0, s0 + kTile, s0 + 2 * kTile, out + static_cast<long long>(b) * L * width + h * kHeadDim,
                                               L, width);
}

// persistent variant for the short sequences (L <= 80): every CTA walks over (batch, head) items; the next item's Q / K / V
// boxes are in flight into the second staging buffer while the current item is computed, so the HBM stream never waits
// for the tensor-core / softmax phase
template <int QPAD, int KC, int NCHUNK, bool CAUSAL>
__global__ void __launch_bounds__(QPAD * 2)
attention_persistent_kernel(const __grid_constant__ CUtensorMap tmap, op16* __restrict__ out, int L, int heads, int items) {
  constexpr int KVPAD = KC * NCHUNK;
  constexpr uint32_t kTile = KVPAD * kRowBytes;
  constexpr uint32_t kBuf = 3 * kTile;
  static_assert(KVPAD >= QPAD && KVPAD % 8 == 0, "Q shares the K / V box height");
  extern __shared__ uint8_t att_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(att_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * kBuf);
  const int width = heads * kHeadDim;
  int item = blockIdx.x;
  if (item >= items) return;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
    tma_stage_item<KVPAD>(&tmap, smem, &bar[0], item / heads, item % heads, width);
  }
  __syncthreads();
  for (int it = 0; item < items; item += gridDim.x, ++it) {
    const int cur = it & 1;
    const int next_item = item + gridDim.x;
    // the other buffer was released by the __syncthreads that ended the previous iteration
    if (threadIdx.x == 0 && next_item < items)
      tma_stage_item<KVPAD>(&tmap, smem + (cur ^ 1) * kBuf, &bar[cur ^ 1], next_item / heads, next_item % heads, width);
    mbar_wait(&bar[cur], (it >> 1) & 1, 32);
    const int b = item / heads, h = item % heads;
    const uint32_t s0 = smem_u32(smem + cur * kBuf);
    attention_compute<QPAD, KC, NCHUNK, CAUSAL>(s0, s0 + kTile, s0 + 2 * kTile,
                                                 out + static_cast<long long>(b) * L * width + h * kHeadDim, L, width);
    __syncthreads();  // all warps are done with `cur` before the next iteration's boxes may land in it
  }
}

template <int QPAD, int KC, int NCHUNK, bool CAUSAL>
int launch_persistent(const op16* qkv, op16* out, int batch, int L, int heads, cudaStream_t stream) {
  constexpr int smem = 2 * 3 * KC * NCHUNK * kRowBytes + 64 + 1024;  // two staging buffers, two mbarriers, alignment slack
  CUtensorMap tmap;
  MSCLIP_TRY(make_tmap_op16_3d(&tmap, qkv, batch, L, 3ull * heads * kHeadDim, 3ull * heads * kHeadDim, KC * NCHUNK));
  static int ctas_per_sm = 0;
  if (ctas_per_sm == 0) {
    MSCLIP_CHECK_CUDA(cudaFuncSetAttribute(attention_persistent_kernel<QPAD, KC, NCHUNK, CAUSAL>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    MSCLIP_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &ctas_per_sm, attention_persistent_kernel<QPAD, KC, NCHUNK, CAUSAL>, QPAD * 2, smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
  }
  const int items = batch * heads;
  const int grid = items < num_sms() * ctas_per_sm ? items : num_sms() * ctas_per_sm;
  attention_persistent_kernel<QPAD, KC, NCHUNK, CAUSAL><<<grid, QPAD * 2, smem, stream>>>(tmap, out, L, heads, items);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int QPAD, int KC, int NCHUNK, bool CAUSAL>
int launch_variant(const op16* qkv, op16* out, int batch, int L, int heads, cudaStream_t stream) {
  constexpr int smem = 3 * KC * NCHUNK * kRowBytes + 64 + 1024;
  CUtensorMap tmap;
  MSCLIP_TRY(make_tmap_op16_3d(&tmap, qkv, batch, L, 3ull * heads * kHeadDim, 3ull * heads * kHeadDim, KC * NCHUNK));
  static bool configured = false;
  if (!configured) {
    MSCLIP_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<QPAD, KC, NCHUNK, CAUSAL>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  attention_kernel<QPAD, KC, NCHUNK, CAUSAL><<<batch * heads, QPAD * 2, smem, stream>>>(tmap, out, L, heads);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int launch_attention(const op16* qkv, op16* out, int batch, int L, int heads, int causal, cudaStream_t stream) {
  if (batch <= 0) return 0;
  MSCLIP_REQUIRE(L >= 1 && L <= 208, "attention: sequence length must be in [1, 208]");
  MSCLIP_REQUIRE(heads >= 1, "attention: heads must be positive");
  if (L <= 64)
    return causal ? launch_persistent<64, 64, 1, true>(qkv, out, batch, L, heads, stream)
                  : launch_persistent<64, 64, 1, false>(qkv, out, batch, L, heads, stream);
  if (L <= 80)
    return causal ? launch_persistent<80, 80, 1, true>(qkv, out, batch, L, heads, stream)
                  : launch_persistent<80, 80, 1, false>(qkv, out, batch, L, heads, stream);
  return causal ? launch_variant<208, 112, 2, true>(qkv, out, batch, L, heads, stream)
                : launch_variant<208, 112, 2, false>(qkv, out, batch, L, heads, stream);
}

}  // namespace msclip
