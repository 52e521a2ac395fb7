// Weight-gradient GEMM for sm_100a:  dW[n, k] = sum_m dY[m, n] * X[m, k]   (the wgrad of every
// F.linear of the shared block, M.py:612 / 747 / 794-798, and of the two projections, M.py:2690 / 3074).
//
// The contraction runs over the TOKEN dimension, which is the slow dimension of both operands in memory (dY [M, N] and
// X [M, K] are row-major activations).  Instead of transposing them, both operands are fed to tcgen05.mma as **MN-major**
// tiles: a TMA box of 64 tokens x 64 columns lands as 64 rows of 128 bytes (128-B swizzle) - exactly the canonical
// MN-major SWIZZLE_128B layout ((8 x 16 B contiguous along M/N, 8 token rows per swizzle atom): the shared-memory
// descriptor's leading byte offset is the distance between 64-column chunks (8 KB), its stride byte offset the distance
// between 8-token groups (1 KB), and the instruction descriptor sets the a_major / b_major bits.  One MMA consumes 16
// tokens (2 KB of every chunk).
//
// CTA = 128 x 256 output tile (fp32 accumulators double-buffered in TMEM), persistent over (token split, tile) work
// units; the token range is split S ways so that tiles x S fills the 148 SMs, every split writes its own fp32 partial
// matrix and reduce_partials sums them in a fixed order (bit-reproducible, no atomics).  Warp roles as in gemm.cu.
#include "gemm_common.cuh"

namespace msclip {

namespace {

using namespace gemm_detail;

constexpr int kWM = 128;  // rows of dW per tile (columns of dY)
constexpr int kWN = 256;  // columns of dW per tile (columns of X)
constexpr int kWK = 64;   // tokens per pipeline stage
constexpr int kChunkBytes = kWK * 128;           // one TMA box: 64 tokens x 64 columns
constexpr int kStageA = (kWM / 64) * kChunkBytes;
constexpr int kStageB = (kWN / 64) * kChunkBytes;
constexpr int kStage = kStageA + kStageB;        // 48 KB
constexpr int kStages = 4;
constexpr int kSmemBytes = kStages * kStage + 256 + 1024;
constexpr int kThreads = 128 + 32 * kNumEpilogueWarps;

struct WgradParams {
  int tokens, N, K;
  int tiles_m, tiles_n, splits, kb_per_split, num_kb;
  float* part;  // [splits][N][K]
  uint32_t lbo, sbo;
};

// MN-major SWIZZLE_128B operand (see the header comment); same bit layout as umma_desc_sw128
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStage);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles = p.tiles_m * p.tiles_n;
  const int units = tiles * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_dy);
    tma_prefetch_desc(&tmap_x);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kNumEpilogueWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // unit u = split * tiles + tile: units that run at the same time share their token range (L2 reuse of dY / X rows)
  auto kb_range = [&](int u, int& kb0, int& kb1) {
    const int s = u / tiles;
    kb0 = s * p.kb_per_split;
    kb1 = kb0 + p.kb_per_split < p.num_kb ? kb0 + p.kb_per_split : p.num_kb;
  };

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int tile = u % tiles;
        const int n0 = (tile / p.tiles_n) * kWM, k0 = (tile % p.tiles_n) * kWN;
        int kb0, kb1;
        kb_range(u, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1, 41);
          uint8_t* sa = smem + s * kStage;
          mbar_arrive_expect_tx(&full_bar[s], kStage);
#pragma unroll
          for (int c = 0; c < kWM / 64; ++c) tma_load_2d(sa + c * kChunkBytes, &tmap_dy, &full_bar[s], n0 + 64 * c, kb * kWK);
#pragma unroll
          for (int c = 0; c < kWN / 64; ++c)
            tma_load_2d(sa + kStageA + c * kChunkBytes, &tmap_x, &full_bar[s], k0 + 64 * c, kb * kWK);
          if (++s == kStages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // both operands MN-major: bits 15 (A) and 16 (B) of the instruction descriptor
      constexpr uint32_t idesc = umma_idesc_f32acc(kWM, kWN) | (1u << 15) | (1u << 16);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int kb0, kb1;
        kb_range(u, kb0, kb1);
        mbar_wait(&tempty_bar[as], aph ^ 1, 42);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kAccStride;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[s], ph, 43);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * kStage);
          const uint32_t b_addr = a_addr + kStageA;
#pragma unroll
          for (int k = 0; k < kWK / 16; ++k) {
            // 16 tokens = two 8-row swizzle atoms = 2 KB further into every chunk
            const uint64_t da = umma_desc_mn_sw128(a_addr + k * 2048, p.lbo, p.sbo);
            const uint64_t db = umma_desc_mn_sw128(b_addr + k * 2048, p.lbo, p.sbo);
            umma_f16(d_tmem, da, db, idesc, (kb > kb0 || k != 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
          if (++s == kStages) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit(&tfull_bar[as]);
        as ^= 1;
        if (as == 0) aph ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    constexpr int kCh = 32;
    constexpr int kPerHalf = kWN / kCh / 2;
    int as = 0;
    uint32_t aph = 0;
    GemmParams gp = {};
    gp.M = p.N;
    gp.N = p.K;
    gp.alpha = 1.0f;
    gp.vec_ok = 1;
    gp.ldo = p.K;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int tile = u % tiles, split = u / tiles;
      const int n0 = (tile / p.tiles_n) * kWM, k0 = (tile % p.tiles_n) * kWN;
      gp.out = p.part + static_cast<long long>(split) * p.N * p.K;
      const int row = n0 + q * 32 + lane;
      mbar_wait(&tfull_bar[as], aph, 44);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccStride;
      uint32_t acc[2][kCh];
      const int c_begin = half * kPerHalf;
      tmem_ld_chunk<kCh>(taddr + c_begin * kCh, acc[0]);
#pragma unroll
      for (int i = 0; i < kPerHalf; ++i) {
        const int c = c_begin + i;
        EpiOperands<EPI_F32, kCh> ops;
        epilogue_prefetch<EPI_F32, kCh>(ops, gp, row, k0 + c * kCh, true);
        tmem_ld_wait();
        if (i + 1 < kPerHalf) tmem_ld_chunk<kCh>(taddr + (c + 1) * kCh, acc[(i + 1) & 1]);
        epilogue_store<EPI_F32, kCh>(acc[i & 1], ops, gp, row, k0 + c * kCh, true);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      as ^= 1;
      if (as == 0) aph ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// debug / bring-up knob: descriptor offsets of the MN-major operands (bytes)
uint32_t g_wgrad_lbo = kChunkBytes, g_wgrad_sbo = 1024;

}  // namespace

void wgrad_set_desc(uint32_t lbo, uint32_t sbo) {
  g_wgrad_lbo = lbo ? lbo : kChunkBytes;
  g_wgrad_sbo = sbo ? sbo : 1024;
}

// number of token splits for a [N, K] weight gradient over `tokens` rows: fill the SMs, keep >= 4 stages of work per unit
int wgrad_pick_splits(int tokens, int N, int K) {
  const int tiles = (N / kWM) * (K / kWN);
  const int num_kb = (tokens + kWK - 1) / kWK;
  const int sms = num_sms();
  // wave efficiency of every candidate; take the SMALLEST split count within 3 % of the best (fewer partial matrices to write
  // and reduce: 30 splits of the QKV gradient wrote 200 MB of partials for a 1 % better tail, profiles/r02_backward_ncu.md)
  double eff[33] = {0.0};
  double best_eff = 0.0;
  int smax = 1;
  for (int s = 1; s <= 32; ++s) {
    if (s > 1 && num_kb / s < 4) break;
    const int units = tiles * s;
    const int waves = (units + sms - 1) / sms;
    eff[s] = static_cast<double>(units) / (static_cast<double>(waves) * sms);
    if (eff[s] > best_eff) best_eff = eff[s];
    smax = s;
  }
  for (int s = 1; s <= smax; ++s)
    if (eff[s] >= best_eff - 0.03) return s;
  return 1;
}

size_t wgrad_workspace_bytes(int tokens, int N, int K) {
  return static_cast<size_t>(wgrad_pick_splits(tokens, N, K)) * N * K * sizeof(float);
}

// dW [N, K] (f32, pitch K) (+)= dY[tokens, N]^T . X[tokens, K]; workspace: wgrad_workspace_bytes(tokens, N, K)
int launch_wgrad(const op16* dY, int64_t ldy, const op16* X, int64_t ldx, int tokens, int N, int K, float* dW, int accumulate,
                 void* workspace, cudaStream_t stream) {
  MSCLIP_REQUIRE(tokens > 0 && N > 0 && K > 0, "wgrad: empty problem");
  MSCLIP_REQUIRE(N % kWM == 0 && K % kWN == 0, "wgrad: N must be a multiple of 128 and K a multiple of 256");
  MSCLIP_REQUIRE(ldy % 8 == 0 && ldx % 8 == 0 && workspace != nullptr, "wgrad: operand pitches must be multiples of 8");
  static bool configured = false;
  if (!configured) {
    MSCLIP_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  CUtensorMap ta, tb;
  MSCLIP_TRY(make_tmap_op16_2d(&ta, dY, static_cast<uint64_t>(tokens), static_cast<uint64_t>(N), static_cast<uint64_t>(ldy), kWK));
  MSCLIP_TRY(make_tmap_op16_2d(&tb, X, static_cast<uint64_t>(tokens), static_cast<uint64_t>(K), static_cast<uint64_t>(ldx), kWK));
  WgradParams p = {};
  p.tokens = tokens;
  p.N = N;
  p.K = K;
  p.tiles_m = N / kWM;
  p.tiles_n = K / kWN;
  p.splits = wgrad_pick_splits(tokens, N, K);
  p.num_kb = (tokens + kWK - 1) / kWK;
  p.kb_per_split = (p.num_kb + p.splits - 1) / p.splits;
  // every split must own at least one k-block (the epilogue stores whatever the accumulator holds)
  while (p.splits > 1 && (p.splits - 1) * p.kb_per_split >= p.num_kb) --p.splits;
  p.part = static_cast<float*>(workspace);
  p.lbo = g_wgrad_lbo;
  p.sbo = g_wgrad_sbo;
  const int units = p.tiles_m * p.tiles_n * p.splits;
  const int grid = units < num_sms() ? units : num_sms();
  wgrad_tcgen05_kernel<<<grid, kThreads, kSmemBytes, stream>>>(ta, tb, p);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return launch_reduce_partials(p.part, p.splits, static_cast<long long>(N) * K, dW, static_cast<long long>(N) * K, accumulate, 0,
                                1.0f, stream);
}

}  // namespace msclip
