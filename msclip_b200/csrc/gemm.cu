// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  out[M,N] = epi(A[M,K] . W[N,K]^T + bias).
//
// Replaces every F.linear of the shared block (QKV M.py:612, out-proj M.py:747, MLP M.py:794-798), the
// projections (M.py:2690, 3074) and - through im2col - the convolutions of the stem / parallel branch
// (M.py:1993-2000, 1842-1861).  Both operands are K-major op16 (A = activations [M,K], W = nn.Linear
// weight [N,K]), accumulation is fp32 in TMEM, and the epilogue (bias, QuickGELU / ReLU, residual add
// on the fp32 stream) is fused so no GEMM output ever makes an extra HBM round trip.
//
// CTA layout (384 threads, 1 CTA / SM, persistent, static round-robin tile schedule):
//   warp 0    : TMA producer   (one lane)  global -> 128B-swizzled smem ring, 4-8 stages deep
//   warp 1    : MMA issuer     (one lane)  tcgen05.mma (128|256) x BN x 16, accumulators double-buffered in TMEM
//   warp 2    : TMEM allocator (512 columns = 2 accumulator stages of up to 256 columns)
//   warps 4-11: epilogue       tcgen05.ld -> registers -> fused math -> global stores (two warps per TMEM
//                              lane quarter, each draining half of the tile's columns)
// Three mbarrier pipelines: smem full/empty (TMA <-> MMA) and TMEM full/empty (MMA <-> epilogue), so the
// epilogue of tile i overlaps the main loop of tile i+1.
#include "gemm_common.cuh"
#include "rowops.cuh"

namespace msclip {

namespace {

using namespace gemm_detail;

constexpr int kLnWarps = 4;        // LN = 3: LayerNorm warps per CTA
constexpr int kLnSlotsQ = 4;       // tiles the epilogue may run ahead of the LayerNorm warps (short: the rows must still be in L2)
constexpr int kLnRowSlots = 4;     // fp32 rows each LayerNorm warp keeps in flight (bulk copies into its shared-memory ring)
constexpr int kLnRowBytes = 768 * 4;

template <int BN, int CG, int LN = 0>
struct GemmCfg {
  static constexpr int kBRows = BN / CG;               // rows of W this CTA stages per k-block
  static constexpr int kStageA = kBM * kBK * 2;
  static constexpr int kStageB = kBRows * kBK * 2;
  static constexpr int kStage = kStageA + kStageB;
  // LN = 3 gives one 32 KB stage to the LayerNorm warps (row staging rings + gamma / beta)
  static constexpr int kRingBytes = (LN == 3 ? 160 : 192) * 1024;
  static constexpr int kStagesRaw = kRingBytes / kStage;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kBarrierBytes = 512;
  static constexpr int kLnParamBytes = (LN == 3) ? 2 * kLnRowBytes : 0;                    // gamma | beta
  static constexpr int kLnStageBytes = (LN == 3) ? kLnWarps * kLnRowSlots * kLnRowBytes : 0;  // row rings
  static constexpr int kSmemBytes = kStages * kStage + kBarrierBytes + kLnParamBytes + kLnStageBytes + 1024;  // + alignment slack
  static constexpr int kChunk = (BN % 32 == 0) ? 32 : 16;                     // columns per tcgen05.ld
  static_assert(kStageB % 1024 == 0, "B stage must keep 1024-byte alignment");
  static_assert(BN % 16 == 0 && BN <= 256, "UMMA N constraint");
  static_assert(CG == 1 || BN % 16 == 0, "pair tiles: each CTA stages BN / 2 rows of W (whole 8-row swizzle groups)");
};

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, cta_group::2) per 256 x BN tile -
// each CTA stages its own 128 rows of A and one half of the W rows, the leader issues M = 256 MMAs that read
// both CTAs' shared memory, and each CTA's TMEM holds its 128 rows of the accumulator.  Halving the W bytes
// per CTA cuts the L2 -> SM operand traffic per flop by a third and deepens the smem ring from 4 to 6 stages.
// NP > 1 (CG = 2 only): NP CTA pairs form one cluster and work on NP vertically adjacent 256-row tiles of the
// same column tile; every W slice is fetched from L2 once and TMA-multicast to the NP CTAs that need it.
// LN = 1 / 2: LayerNorm folding, see gemm_common.cuh (consume in the epilogue of QKV / fc1, emit from out-proj / fc2)
// NE = number of epilogue warps (a multiple of 4: NE / 4 warps share each TMEM lane quarter and split the tile's column
// chunks between them); 8 by default, 12 / 16 selectable for the CTA-pair tiles (see g_epi_warps below).
// CONV = 1: implicit-GEMM convolution - the A operand is the im2col view of one or two NHWC activations, fetched by
// im2col-mode TMA (one instruction per k-block = filter tap x 64 channels: the hardware walks the 128 output pixels of
// the tile with the convolution stride, applies the tap offset and zero-fills the padding), tmap_a / tmap_a2 = the
// sources' im2col maps.  Everything downstream of the smem ring is the plain GEMM.
// LN = 3 (EPI_RESID_F32, N = 768, CTA pairs): four extra "LayerNorm warps" per CTA (512 threads, 128 registers each).
// The tile schedule is the plain one - the three column tiles of a 256-row block run on three neighbouring CTA pairs at
// about the same time, which is what keeps the A rows and the W tiles L2-resident - so a row block is complete only once
// three different CTAs have stored their tiles: every CTA bumps a global per-(block, CTA rank) counter after each tile.
// Block mb is normalised by the CTA that computed its column tile mb % 3 (a fixed owner, so every CTA normalises a third
// of the blocks it touches; "last arriver does it" piles the work on whichever pairs run late and makes them later
// still): its LayerNorm warps wait until the counter shows all three tiles, re-read the 128 x 768 fresh rows (L2 hits)
// and write LayerNorm(x_new) as the next GEMM's 16-bit A operand - with the row arithmetic of layernorm_kernel
// (rowops.cuh), so the result is bit-identical to the separate pass.  The wait cannot deadlock: all CTAs are resident, a
// CTA's epilogue never waits for another CTA, and its own LayerNorm warps lag by at most kLnSlotsQ tiles.
template <int BN, int EPI, int CG, int NP, int LN = 0, int NE = 8, int CONV = 0>
__global__ void __launch_bounds__(128 + 32 * NE + (LN == 3 ? 32 * kLnWarps : 0), 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2,
                    const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  using Cfg = GemmCfg<BN, CG, LN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStage);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* rows_done_bar = tempty_bar + 2;          // LN = 3: epilogue warps -> LayerNorm warps (tile stored), kLnSlotsQ slots
  uint64_t* ln_free_bar = rows_done_bar + kLnSlotsQ; // LN = 3: LayerNorm warps -> epilogue warps (slot consumed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ln_free_bar + kLnSlotsQ);
  uint64_t* ln_row_bar = reinterpret_cast<uint64_t*>(tmem_slot + 2);  // LN = 3: [kLnWarps][kLnRowSlots] row landed
  float4* ln_gamma4 = reinterpret_cast<float4*>(smem + Cfg::kStages * Cfg::kStage + Cfg::kBarrierBytes);
  float4* ln_beta4 = ln_gamma4 + 768 / 4;
  uint8_t* ln_rows = reinterpret_cast<uint8_t*>(ln_beta4 + 768 / 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  static_assert(NP == 1 || CG == 2, "multicast clusters are built from CTA pairs");
  const uint32_t cluster_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const uint32_t cta_rank = cluster_rank & 1u;   // rank inside the CTA pair
  const uint32_t pair_idx = cluster_rank >> 1;   // which pair of the cluster
  const uint32_t leader_rank = cluster_rank & ~1u;
  const bool leader = cta_rank == 0;
  const int first_tile = (CG == 2) ? static_cast<int>(cluster_id_x()) : static_cast<int>(blockIdx.x);
  const int tile_step = (CG == 2) ? static_cast<int>(num_clusters_x()) : static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    if (CONV) tma_prefetch_desc(&tmap_a2);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], NP);  // every pair that multicasts into this CTA must have consumed the slot
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], NE * CG);  // one arrive per epilogue warp (of both CTAs)
    }
    for (int i = 0; i < kLnSlotsQ; ++i) {
      mbar_init(&rows_done_bar[i], NE);
      mbar_init(&ln_free_bar[i], kLnWarps);
    }
    if (LN == 3)
      for (int i = 0; i < kLnWarps * kLnRowSlots; ++i) mbar_init(&ln_row_bar[i], 1);
    fence_mbar_init();
  }
  if (LN == 3 && warp >= 4 + NE) {
    // LayerNorm parameters -> shared memory (read once per row by the LayerNorm warps)
    for (int i = threadIdx.x - 32 * (4 + NE); i < 768 / 4; i += 32 * kLnWarps) {
      ln_gamma4[i] = __ldg(reinterpret_cast<const float4*>(p.lnw_gamma) + i);
      ln_beta4[i] = __ldg(reinterpret_cast<const float4*>(p.lnw_beta) + i);
    }
  }
  if (warp == 2) {
    if (CG == 2) {
      tmem_alloc_pair(tmem_slot, 512);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_kb = (p.K + kBK - 1) / kBK;
  static_assert(LN != 3 || (CG == 2 && NP == 1 && NE == 8 && EPI == EPI_RESID_F32 && BN == 256), "LayerNorm warps: pair tiles of the residual GEMMs");
  // LN = 3: 512 threads -> 128 registers each (ptxas takes the ceiling from the launch bounds, setmaxnreg does not raise
  // it for a region): the epilogue warps drain the accumulator in 16-column pieces there, like the 12 / 16-warp variants
  // it-th tile of this CTA (pair / cluster), -1 past the end
  auto tile_at = [&](int it) -> int {
    const int t = first_tile + it * tile_step;
    return t < p.total_tiles ? t : -1;
  };

  if (warp < 4) {
  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0, tile; (tile = tile_at(it)) >= 0; ++it) {
        const int m0 = (tile / p.tiles_n) * (kBM * CG * NP) + static_cast<int>(pair_idx) * (kBM * CG) +
                       static_cast<int>(cta_rank) * kBM;
        const int n0 = (tile % p.tiles_n) * BN + static_cast<int>(cta_rank) * Cfg::kBRows;
        // CONV: first output pixel of this CTA's 128 rows -> (image, oy, ox); filter position of the k-block
        int img = 0, oy = 0, ox = 0, src = 0, ky = 0, kx = 0, cb = 0;
        if (CONV) {
          const ConvGeom& g = p.conv;
          img = fast_div(m0, g.hw_mul, g.hw_shr);
          const int rem = m0 - img * (g.Ho * g.Wo);
          oy = fast_div(rem, g.wo_mul, g.wo_shr);
          ox = rem - oy * g.Wo;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1, 1);
          uint8_t* sa = smem + s * Cfg::kStage;
          int cw = 0, chh = 0, cc = 0;
          uint16_t offw = 0, offh = 0;
          const CUtensorMap* ta = &tmap_a;
          if (CONV) {
            const ConvGeom& g = p.conv;
            if (kb == g.kb_begin1) {
              src = 1;
              ky = kx = cb = 0;
            }
            ta = src ? &tmap_a2 : &tmap_a;
            cw = ox * g.stride[src] - g.pad[src];
            chh = oy * g.stride[src] - g.pad[src];
            cc = cb * kBK;
            offw = static_cast<uint16_t>(kx);
            offh = static_cast<uint16_t>(ky);
            if (++cb == g.cblk[src]) {
              cb = 0;
              if (++kx == g.ksize[src]) {
                kx = 0;
                ++ky;
              }
            }
          }
          if (CG == 2) {
            // both CTAs' bytes are counted on the leader's barrier, which the leader arms for the pair
            const uint32_t bar = mapa_shared(smem_u32(&full_bar[s]), leader_rank);
            if (leader) mbar_arrive_expect_tx(&full_bar[s], Cfg::kStage * 2);
            if (CONV) tma_load_im2col_4d_pair(sa, ta, bar, cc, cw, chh, img, offw, offh);
            else tma_load_2d_pair(sa, &tmap_a, bar, kb * kBK, m0);
            if (NP == 1) {
              tma_load_2d_pair(sa + Cfg::kStageA, &tmap_b, bar, kb * kBK, n0);
            } else {
              // this CTA fetches slice `pair_idx` of its W half and multicasts it to the same-rank CTA of every pair
              constexpr int kSliceRows = Cfg::kBRows / NP;
              uint16_t mask = 0;
#pragma unroll
              for (int q = 0; q < NP; ++q) mask |= static_cast<uint16_t>(1u << (2 * q + cta_rank));
              tma_load_2d_pair_mcast(sa + Cfg::kStageA + pair_idx * (kSliceRows * 128), &tmap_b, &full_bar[s], mask,
                                     kb * kBK, n0 + static_cast<int>(pair_idx) * kSliceRows);
            }
          } else {
            mbar_arrive_expect_tx(&full_bar[s], Cfg::kStage);
            if (CONV) tma_load_im2col_4d(sa, ta, &full_bar[s], cc, cw, chh, img, offw, offh);
            else tma_load_2d(sa, &tmap_a, &full_bar[s], kb * kBK, m0);
            tma_load_2d(sa + Cfg::kStageA, &tmap_b, &full_bar[s], kb * kBK, n0);
          }
          if (++s == Cfg::kStages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      const uint32_t idesc = umma_idesc_f32acc_fmt(kBM * CG, BN, p.operand_fmt);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int it = 0, tile; (tile = tile_at(it)) >= 0; ++it) {
        mbar_wait(&tempty_bar[as], aph ^ 1, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph, 3);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * Cfg::kStage);
          const uint32_t b_addr = a_addr + Cfg::kStageA;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t da = umma_desc_sw128(a_addr + k * 32), db = umma_desc_sw128(b_addr + k * 32);
            if (CG == 2) umma_f16_pair(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_f16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs) once these MMAs have read it
          if (CG == 2) umma_commit_pair(&empty_bar[s], static_cast<uint16_t>((1u << (2 * NP)) - 1u));
          else umma_commit(&empty_bar[s]);
          if (++s == Cfg::kStages) {
            s = 0;
            ph ^= 1;
          }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if (CG == 2) umma_commit_pair(&tfull_bar[as], static_cast<uint16_t>(3u << (2 * pair_idx)));
        else umma_commit(&tfull_bar[as]);
        as ^= 1;
        if (as == 0) aph ^= 1;
      }
    }
  }
  } else if (warp < 4 + NE) {
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;    // which share of the tile's column chunks this warp drains
    constexpr int kParts = NE / 4;
    // more epilogue warps run under a tighter register cap: they drain the tile in 16-column pieces
    constexpr int kCh = ((NE > 8 || LN == 3 || EPI == EPI_DGELU_BF16 || EPI == EPI_QGELU_DUAL_BF16) && Cfg::kChunk == 32) ? 16 : Cfg::kChunk;
    constexpr int kNumCh = BN / kCh;
    static_assert(NE % 4 == 0 && kParts >= 1 && (LN != 2 || kParts == 2), "LN emit slots assume two column halves per tile");
    constexpr int kPerHalf = (kNumCh + kParts - 1) / kParts;
    const int c_begin = half * kPerHalf;
    const int c_end = (c_begin + kPerHalf < kNumCh) ? c_begin + kPerHalf : kNumCh;
    int as = 0;
    uint32_t aph = 0;
    for (int it = 0, tile; (tile = tile_at(it)) >= 0; ++it) {
      const int m0 = (tile / p.tiles_n) * (kBM * CG * NP) + static_cast<int>(pair_idx) * (kBM * CG) +
                     static_cast<int>(cta_rank) * kBM;
      const int n0 = (tile % p.tiles_n) * BN;
      if (EPI == EPI_RESID_F32 && p.vec_ok) {
        // the residual rows of this tile are known long before its accumulator is complete: pull them into L2
        // while the MMAs run, so the epilogue's loads see L2 latency instead of DRAM latency
        const int prow = m0 + q * 32 + lane;
        if (prow < p.M) {
          const float* rp = p.resid + static_cast<long long>(prow) * p.ldr + n0 + c_begin * kCh;
#pragma unroll
          for (int i = 0; i < kPerHalf; ++i)
            if (c_begin + i < c_end && n0 + (c_begin + i + 1) * kCh <= p.N)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + i * kCh));
        }
      }
      const int row = m0 + q * 32 + lane;
      // LN fold: the row records are fetched while the MMAs of this tile still run (an L2 round trip on the
      // epilogue's critical path costs the short-K GEMMs ~15 %)
      LnRow lnr = {0.f, 1.f};
      LnEmit lne = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (LN == 1 && row < p.M) {
        float sh;
        ln_load_record(p.ln_in + static_cast<long long>(row) * kLnRec, 1.0f / static_cast<float>(p.K), sh, lnr.mean_c, lnr.rstd);
      }
      if (LN == 2) {
        // new shift of a row = its mean before this update (from the record of the residual input)
        const int row_e = row & ~1, row_o = row | 1;
        float mc, rs;
        if (row_e < p.M) {
          ln_load_record(p.ln_in + static_cast<long long>(row_e) * kLnRec, 1.0f / static_cast<float>(p.N), lne.shift_e, mc, rs);
          lne.shift_e += mc;
        }
        if (row_o < p.M) {
          ln_load_record(p.ln_in + static_cast<long long>(row_o) * kLnRec, 1.0f / static_cast<float>(p.N), lne.shift_o, mc, rs);
          lne.shift_o += mc;
        }
      }
      mbar_wait(&tfull_bar[as], aph, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccStride;
      GemmParams pt = p;  // per-tile view: with split outputs every column tile owns its own [M, BN] matrix
      if (p.split_stride != 0) {
        const long long shift = static_cast<long long>(tile % p.tiles_n) * p.split_stride - n0;
        pt.out = (EPI == EPI_RESID_F32 || EPI == EPI_F32) ? static_cast<void*>(reinterpret_cast<float*>(p.out) + shift)
                                                          : static_cast<void*>(reinterpret_cast<op16*>(p.out) + shift);
      }
      // software pipeline: the TMEM load of chunk c+1 and the global operands of chunk c are in flight while
      // chunk c is converted and stored (tcgen05.wait::ld waits for every outstanding load, so the next load
      // is issued right after the wait)
      uint32_t acc[2][kCh];
      tmem_ld_chunk<kCh>(taddr + c_begin * kCh, acc[0]);
#pragma unroll
      for (int i = 0; i < kPerHalf; ++i) {
        const int c = c_begin + i;
        if (c < c_end) {
          const int col0 = n0 + c * kCh;
          const bool fast = p.vec_ok && (col0 + kCh <= p.N);
          EpiOperands<EPI, kCh, LN> ops;
          epilogue_prefetch<EPI, kCh, LN>(ops, pt, row, col0, fast);
          tmem_ld_wait();
          if (c + 1 < c_end) tmem_ld_chunk<kCh>(taddr + (c + 1) * kCh, acc[(i + 1) & 1]);
          epilogue_store<EPI, kCh, LN>(acc[i & 1], ops, pt, row, col0, fast, lnr, &lne);
        }
      }
      if (LN == 2) {
        // row sums of this 128-column slice: the two lanes of a pair hold complementary pieces of both rows
        lne.s1e += __shfl_xor_sync(0xffffffffu, lne.s1e, 1);
        lne.s2e += __shfl_xor_sync(0xffffffffu, lne.s2e, 1);
        lne.s1o += __shfl_xor_sync(0xffffffffu, lne.s1o, 1);
        lne.s2o += __shfl_xor_sync(0xffffffffu, lne.s2o, 1);
        const bool odd = (lane & 1) != 0;
        const int r = odd ? (row | 1) : (row & ~1);
        if (r < p.M) {
          const int slot = (tile % p.tiles_n) * 2 + half;
          float* rec = p.ln_out + static_cast<long long>(r) * kLnRec;
          *reinterpret_cast<float2*>(rec + 4 + 2 * slot) = odd ? make_float2(lne.s1o, lne.s2o) : make_float2(lne.s1e, lne.s2e);
          if (slot == 0) rec[0] = odd ? lne.shift_o : lne.shift_e;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[as]), leader_rank));
        else mbar_arrive(&tempty_bar[as]);
        if (LN == 3) {
          // this warp's part of the tile is stored (__syncwarp ordered the other lanes' stores before this arrive):
          // tell the LayerNorm warps, once they have consumed the slot's previous tile
          const int slot = it % kLnSlotsQ;
          mbar_wait(&ln_free_bar[slot], ((it / kLnSlotsQ) & 1) ^ 1, 5);
          mbar_arrive(&rows_done_bar[slot]);
        }
      }
      as ^= 1;
      if (as == 0) aph ^= 1;
    }
  } else if (LN == 3) {
    // ------------------------------------------------------------------ LayerNorm warps
    // Every warp runs on its own (no barrier between them) with two cursors over this CTA's tile sequence:
    //   it_pub  - next tile to acknowledge: once the epilogue warps have stored it, warp 0 adds 1 to the row block's
    //             global counter (release at GPU scope) and every warp frees the queue slot;
    //   it_proc - next OWNED tile (column tile == block % 3) whose row block this CTA normalises: when the block's
    //             counter shows all three tiles, the warp streams its 32 rows through its shared-memory ring.
    // Waiting for sibling CTAs therefore never holds back the acknowledgements (and through them the epilogue), and the
    // poll is a relaxed load (an acquire load costs an L1 invalidation per iteration) followed by one fence.
    const int lw = warp - (4 + NE);
    const float* xsrc = reinterpret_cast<const float*>(p.out);
    constexpr int kV = rowops::kVec;
    constexpr int kRows = kBM / kLnWarps;
    constexpr uint32_t kTileMask = 15u, kWarpDone = 16u;  // counter = tiles stored (low bits) + 16 per LayerNorm warp done
    uint8_t* ring = ln_rows + lw * (kLnRowSlots * kLnRowBytes);
    uint64_t* rbar = ln_row_bar + lw * kLnRowSlots;
    uint32_t ln_row_phase = 0u;  // bit s: parity the next wait on row slot s expects
    auto owned = [&](int tile) { return (tile % p.tiles_n) == ((tile / p.tiles_n) % p.tiles_n); };
    int it_pub = 0, it_proc = 0;
    while (it_proc < it_pub && !owned(tile_at(it_proc))) ++it_proc;
    for (;;) {
      const int t_pub = tile_at(it_pub);
      const bool have_owned = it_proc < it_pub;  // it_proc always rests on an owned, acknowledged tile (or == it_pub)
      if (t_pub < 0 && !have_owned) break;
      bool progressed = false;
      // ---- normalise the oldest owned block if its three tiles are visible
      if (have_owned) {
        const int tile = tile_at(it_proc);
        const int mb = tile / p.tiles_n;
        uint32_t* cnt = p.lnw_counters + 2 * mb + cta_rank;
        uint32_t seen = 0;
        if (lane == 0) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
        seen = __shfl_sync(0xffffffffu, seen, 0);
        if ((seen & kTileMask) >= static_cast<uint32_t>(p.tiles_n)) {
          const int r0 = mb * (kBM * CG) + static_cast<int>(cta_rank) * kBM + lw * kRows;
          const int nrows = p.M - r0 < kRows ? (p.M - r0 > 0 ? p.M - r0 : 0) : kRows;
          auto fetch = [&](int i) {  // lane 0: row r0 + i -> slot i % kLnRowSlots (one 3 KB bulk copy, no registers held)
            const int sl = i % kLnRowSlots;
            mbar_arrive_expect_tx(&rbar[sl], kLnRowBytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(ring + sl * kLnRowBytes)),
                         "l"(xsrc + static_cast<long long>(r0 + i) * p.ldo), "n"(kLnRowBytes), "r"(smem_u32(&rbar[sl]))
                         : "memory");
          };
          if (lane == 0) {
            // acquire the three tiles (written with generic-proxy stores by up to three SMs, released through the
            // counter), then hand over to the async proxy the bulk copies read with
            __threadfence();
            asm volatile("fence.proxy.async.global;" ::: "memory");
            for (int i = 0; i < kLnRowSlots && i < nrows; ++i) fetch(i);
          }
#pragma unroll 1
          for (int i = 0; i < nrows; ++i) {
            const int sl = i % kLnRowSlots;
            mbar_wait(&rbar[sl], (ln_row_phase >> sl) & 1u, 7);
            ln_row_phase ^= 1u << sl;
            float4 v[kV];
            const float4* src = reinterpret_cast<const float4*>(ring + sl * kLnRowBytes);
#pragma unroll
            for (int k = 0; k < kV; ++k) v[k] = src[lane + 32 * k];
            __syncwarp();  // every lane has read the slot: refill it
            if (lane == 0 && i + kLnRowSlots < nrows) fetch(i + kLnRowSlots);
            // per row exactly rowops::layer_norm_row's arithmetic, gamma / beta from shared memory
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < kV; ++k) sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
            const float mean = warp_sum(sum) * (1.0f / rowops::kD);
            float q = 0.f;
#pragma unroll
            for (int k = 0; k < kV; ++k) {
              v[k].x -= mean;
              v[k].y -= mean;
              v[k].z -= mean;
              v[k].w -= mean;
              q += (v[k].x * v[k].x + v[k].y * v[k].y) + (v[k].z * v[k].z + v[k].w * v[k].w);
            }
            const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / rowops::kD) + rowops::kLnEps);
            uint2* dst = reinterpret_cast<uint2*>(p.lnw_out + static_cast<long long>(r0 + i) * p.lnw_ld);
#pragma unroll
            for (int k = 0; k < kV; ++k) {
              const float4 ww = ln_gamma4[lane + 32 * k];
              const float4 bb = ln_beta4[lane + 32 * k];
              const float ox = ww.x * (v[k].x * rstd) + bb.x;
              const float oy = ww.y * (v[k].y * rstd) + bb.y;
              const float oz = ww.z * (v[k].z * rstd) + bb.z;
              const float ow = ww.w * (v[k].w * rstd) + bb.w;
              dst[lane + 32 * k] = make_uint2(pack16(ox, oy), pack16(oz, ow));
            }
          }
          if (lane == 0) {
            // the last of the four warps to finish returns the counter to zero for the next launch
            const uint32_t old = atomicAdd(cnt, kWarpDone);
            if (old == static_cast<uint32_t>(p.tiles_n) + (kLnWarps - 1) * kWarpDone) *cnt = 0u;
          }
          ++it_proc;
          while (it_proc < it_pub && !owned(tile_at(it_proc))) ++it_proc;
          progressed = true;
        }
      }
      // ---- acknowledge the next stored tile (blocking only when there is nothing to normalise)
      if (t_pub >= 0) {
        const int slot = it_pub % kLnSlotsQ;
        const uint32_t par = (it_pub / kLnSlotsQ) & 1;
        bool stored;
        if (it_proc < it_pub) {
          stored = mbar_try_wait(&rows_done_bar[slot], par);
        } else {
          mbar_wait(&rows_done_bar[slot], par, 6);
          stored = true;
        }
        if (stored) {
          if (lw == 0 && lane == 0) {
            __threadfence();  // the tile (stored by the epilogue warps, observed through the mbarrier) -> GPU scope
            atomicAdd(p.lnw_counters + 2 * (t_pub / p.tiles_n) + cta_rank, 1u);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&ln_free_bar[slot]);
          ++it_pub;
          while (it_proc < it_pub && !owned(tile_at(it_proc))) ++it_proc;
          progressed = true;
        }
      }
      if (!progressed) __nanosleep(200);
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    if (CG == 2) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

template <int BN, int EPI, int CG, int NP, int LN = 0, int NE = 8, int CONV = 0>
int launch_variant(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream,
                   const CUtensorMap* ta2 = nullptr) {
  using Cfg = GemmCfg<BN, CG, LN>;
  static bool configured = false;
  if (!configured) {
    MSCLIP_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, EPI, CG, NP, LN, NE, CONV>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int units = num_sms() / (CG * NP);  // CTAs, CTA pairs or clusters that can be resident
  const int n = p.total_tiles < units ? p.total_tiles : units;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n * CG * NP);
  cfg.blockDim = dim3(128 + 32 * NE + (LN == 3 ? 32 * kLnWarps : 0));
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG * NP;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MSCLIP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BN, EPI, CG, NP, LN, NE, CONV>, ta, ta2 ? *ta2 : ta, tb, p));
  return 0;
}

// MSCLIP_GEMM_EPI_WARPS = 8 | 12 | 16: epilogue warps of the CTA-pair tiles (16 = 16 for the 16-bit epilogues and 12 for
// the fp32 residual epilogue).  Measured on the text-tower shapes (profiles/r01_kernel_bench.md): 8 warps draining
// 32-column pieces are as fast or faster (QKV 1330 / 1314 / 1292, fc1 1244 / 1224 / 1267, fc2 1365 / 1277 / 1278 TFLOP/s
// with 8 / 12 / 16 warps), so 8 stays the default and the wider variants are kept for A/B runs only.
static const int g_epi_warps = [] {
  const char* e = getenv("MSCLIP_GEMM_EPI_WARPS");
  const int v = e ? atoi(e) : 8;
  return (v == 8 || v == 12 || v == 16) ? v : 8;
}();

template <int BN, int CG, int NP>
int launch_bn(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int epi, cudaStream_t stream) {
  if constexpr (BN == 256 && CG == 2 && NP == 1) {
    if (g_epi_warps == 16) {
      switch (epi) {
        case EPI_BF16: return launch_variant<BN, EPI_BF16, CG, NP, 0, 16>(ta, tb, p, stream);
        case EPI_QGELU_BF16: return launch_variant<BN, EPI_QGELU_BF16, CG, NP, 0, 16>(ta, tb, p, stream);
        case EPI_RESID_F32: return launch_variant<BN, EPI_RESID_F32, CG, NP, 0, 12>(ta, tb, p, stream);
      }
    } else if (g_epi_warps == 12) {
      switch (epi) {
        case EPI_BF16: return launch_variant<BN, EPI_BF16, CG, NP, 0, 12>(ta, tb, p, stream);
        case EPI_QGELU_BF16: return launch_variant<BN, EPI_QGELU_BF16, CG, NP, 0, 12>(ta, tb, p, stream);
        case EPI_RESID_F32: return launch_variant<BN, EPI_RESID_F32, CG, NP, 0, 12>(ta, tb, p, stream);
      }
    }
  }
  switch (epi) {
    case EPI_BF16: return launch_variant<BN, EPI_BF16, CG, NP>(ta, tb, p, stream);
    case EPI_QGELU_BF16: return launch_variant<BN, EPI_QGELU_BF16, CG, NP>(ta, tb, p, stream);
    case EPI_RELU_BF16: return launch_variant<BN, EPI_RELU_BF16, CG, NP>(ta, tb, p, stream);
    case EPI_RESID_F32: return launch_variant<BN, EPI_RESID_F32, CG, NP>(ta, tb, p, stream);
    case EPI_F32: return launch_variant<BN, EPI_F32, CG, NP>(ta, tb, p, stream);
  }
  set_last_error("launch_gemm: unknown epilogue " + std::to_string(epi));
  return 2;
}

}  // namespace

// 0: one CTA per tile; 1: CTA pairs; 2 / 4: clusters of 2 / 4 pairs with TMA multicast of the W tile
constexpr int kDefaultPairMode = 1;
static int g_pair_mode = kDefaultPairMode;
void gemm_set_pair_mode(int mode) {
  g_pair_mode = (mode == 0 || mode == 1 || mode == 2 || mode == 4) ? mode : kDefaultPairMode;  // anything else: default
}

int gemm_pick_bn(int N) {
  const int cands[6] = {256, 192, 128, 96, 64, 48};
  for (int i = 0; i < 6; ++i)
    if (N % cands[i] == 0) return cands[i];
  if (N >= 256) return 256;  // ragged N: last tile is masked in the epilogue
  for (int i = 5; i >= 0; --i)
    if (cands[i] >= N) return cands[i];
  return 256;
}

int launch_gemm(const op16* A, int64_t lda, const op16* W, int64_t ldw, int M, int N, int K, const float* bias,
                void* out, int64_t ldo, const float* resid, int64_t ldr, int epi, cudaStream_t stream) {
  return launch_gemm_scaled(A, lda, W, ldw, M, N, K, 1.0f, bias, out, ldo, resid, ldr, epi, stream);
}

static int launch_gemm_impl(const op16* A, int64_t lda, const op16* W, int64_t ldw, int M, int N, int K, float alpha,
                            const float* bias, void* out, int64_t ldo, const float* resid, int64_t ldr, int epi,
                            int split_cols, cudaStream_t stream, uint32_t operand_fmt = kUmmaOperandFormat);

// the same GEMM on IEEE fp16 operands whatever the build's operand type (16-bit payloads are moved as opaque words)
int launch_gemm_f16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, float alpha, void* out, int64_t ldo,
                    int epi, cudaStream_t stream) {
  return launch_gemm_impl(static_cast<const op16*>(A), lda, static_cast<const op16*>(W), ldw, M, N, K, alpha, nullptr, out, ldo, nullptr,
                          0, epi, 0, stream, kUmmaFormatF16);
}

int launch_gemm_scaled(const op16* A, int64_t lda, const op16* W, int64_t ldw, int M, int N, int K, float alpha,
                       const float* bias, void* out, int64_t ldo, const float* resid, int64_t ldr, int epi,
                       cudaStream_t stream) {
  return launch_gemm_impl(A, lda, W, ldw, M, N, K, alpha, bias, out, ldo, resid, ldr, epi, 0, stream);
}

// Column tile j (split_cols wide) is written as its own dense [M, split_cols] matrix at out + j * M * split_cols:
// one GEMM producing several tensors (the shared first convolution of the stem and of the parallel branch).
int launch_gemm_split(const op16* A, int64_t lda, const op16* W, int64_t ldw, int M, int N, int K, const float* bias,
                      void* out, int split_cols, int epi, cudaStream_t stream) {
  MSCLIP_REQUIRE(split_cols == 48 || split_cols == 64 || split_cols == 96 || split_cols == 128 || split_cols == 192 ||
                     split_cols == 256,
                 "launch_gemm_split: split width must be a supported tile width");
  MSCLIP_REQUIRE(N % split_cols == 0 && epi != EPI_RESID_F32, "launch_gemm_split: N must be a multiple of the split width");
  return launch_gemm_impl(A, lda, W, ldw, M, N, K, 1.0f, bias, out, split_cols, nullptr, 0, epi, split_cols, stream);
}

int launch_gemm_ln(const op16* A, int64_t lda, const op16* W, int64_t ldw, int M, int N, int K, const float* bias, void* out,
                   int64_t ldo, const float* resid, int64_t ldr, int epi, int ln_mode, const float* ln_in, float* ln_out,
                   op16* out16, int64_t ldo16, const float* colsum, cudaStream_t stream) {
  MSCLIP_REQUIRE(M > 0 && K > 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "launch_gemm_ln: bad operand shapes");
  MSCLIP_REQUIRE(ln_mode == 1 || ln_mode == 2, "launch_gemm_ln: mode must be 1 (consume) or 2 (emit)");
  MSCLIP_REQUIRE(N % 256 == 0 && ln_in != nullptr && (reinterpret_cast<uintptr_t>(ln_in) & 15) == 0,
                 "launch_gemm_ln: N must be a multiple of 256 and the row records 16-byte aligned");
  const bool f32_out = epi == EPI_RESID_F32;
  if (ln_mode == 1) {
    MSCLIP_REQUIRE((epi == EPI_BF16 || epi == EPI_QGELU_BF16) && colsum != nullptr && (reinterpret_cast<uintptr_t>(colsum) & 15) == 0,
                   "launch_gemm_ln: consume mode needs a 16-bit epilogue and the column sums of the folded weight");
  } else {
    MSCLIP_REQUIRE(epi == EPI_RESID_F32 && N == 128 * kLnSlots && resid != nullptr && ln_out != nullptr && out16 != nullptr &&
                       ln_out != ln_in && ldr % 4 == 0 && ldo16 % 4 == 0 && (reinterpret_cast<uintptr_t>(resid) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(out16) & 7) == 0 && (reinterpret_cast<uintptr_t>(ln_out) & 15) == 0,
                   "launch_gemm_ln: emit mode needs the residual epilogue, N = 768 and separate in / out row records");
  }
  MSCLIP_REQUIRE(ldo % (f32_out ? 4 : 8) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                     (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0),
                 "launch_gemm_ln: output rows and bias must be 16-byte aligned");
  const int cg = (g_pair_mode >= 1 && M >= 256) ? 2 : 1;
  CUtensorMap ta, tb;
  MSCLIP_TRY(make_tmap_op16_2d(&ta, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(lda), kBM));
  MSCLIP_TRY(make_tmap_op16_2d(&tb, W, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldw),
                               static_cast<uint32_t>(256 / cg)));
  GemmParams p = {};
  p.operand_fmt = kUmmaOperandFormat;
  p.M = M;
  p.N = N;
  p.K = K;
  p.tiles_n = N / 256;
  p.alpha = 1.0f;
  p.vec_ok = 1;
  p.split_stride = 0;
  p.total_tiles = ((M + kBM * cg - 1) / (kBM * cg)) * p.tiles_n;
  p.bias = bias;
  p.out = out;
  p.resid = resid;
  p.ldo = ldo;
  p.ldr = ldr;
  p.ln_in = ln_in;
  p.ln_out = ln_out;
  p.out16 = out16;
  p.ldo16 = ldo16;
  p.colsum = colsum;
  if (ln_mode == 2) return cg == 2 ? launch_variant<256, EPI_RESID_F32, 2, 1, 2>(ta, tb, p, stream)
                                   : launch_variant<256, EPI_RESID_F32, 1, 1, 2>(ta, tb, p, stream);
  if (epi == EPI_BF16) return cg == 2 ? launch_variant<256, EPI_BF16, 2, 1, 1>(ta, tb, p, stream)
                                      : launch_variant<256, EPI_BF16, 1, 1, 1>(ta, tb, p, stream);
  return cg == 2 ? launch_variant<256, EPI_QGELU_BF16, 2, 1, 1>(ta, tb, p, stream)
                 : launch_variant<256, EPI_QGELU_BF16, 1, 1, 1>(ta, tb, p, stream);
}

// x (fp32, in place) += A . W^T + bias, and h = LayerNorm(x) * gamma + beta as op16 - the residual GEMMs of the shared
// block (out-proj M.py:747 + ln_2, fc2 M.py:798 + the next block's ln_1, M.py:1027-1028) with the LayerNorm done by extra
// warps of the same kernel (LN = 3 above).  N must be 768, M >= 256.
int launch_gemm_resid_ln(const op16* A, int64_t lda, const op16* W, int64_t ldw, int M, int N, int K, const float* bias, float* x,
                         int64_t ldx, const float* gamma, const float* beta, op16* h, int64_t ldh, uint32_t* counters,
                         cudaStream_t stream) {
  MSCLIP_REQUIRE(counters != nullptr, "launch_gemm_resid_ln: needs the zero-initialised block counters (gemm_resid_ln_counters(M) words)");
  MSCLIP_REQUIRE(M >= 256 && N == rowops::kD && K > 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0,
                 "launch_gemm_resid_ln: needs M >= 256, N = 768 and 16-byte aligned operand rows");
  MSCLIP_REQUIRE(ldx % 4 == 0 && ldh % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(h) & 7) == 0 &&
                     (reinterpret_cast<uintptr_t>(bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(gamma) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(beta) & 15) == 0 && bias && gamma && beta,
                 "launch_gemm_resid_ln: x, h, bias, gamma and beta must be aligned and non-null");
  CUtensorMap ta, tb;
  MSCLIP_TRY(make_tmap_op16_2d(&ta, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(lda), kBM));
  MSCLIP_TRY(make_tmap_op16_2d(&tb, W, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldw), 128));
  GemmParams p = {};
  p.operand_fmt = kUmmaOperandFormat;
  p.M = M;
  p.N = N;
  p.K = K;
  p.tiles_n = N / 256;
  p.alpha = 1.0f;
  p.vec_ok = 1;
  p.total_tiles = ((M + 2 * kBM - 1) / (2 * kBM)) * p.tiles_n;
  p.bias = bias;
  p.out = x;
  p.resid = x;
  p.ldo = ldx;
  p.ldr = ldx;
  p.lnw_gamma = gamma;
  p.lnw_beta = beta;
  p.lnw_out = h;
  p.lnw_ld = ldh;
  p.lnw_counters = counters;
  return launch_variant<256, EPI_RESID_F32, 2, 1, 3>(ta, tb, p, stream);
}

// one counter per (256-row block, CTA rank): zero before the first launch, left zero by every launch
size_t gemm_resid_ln_counters(int M) { return 2 * static_cast<size_t>((M + 2 * kBM - 1) / (2 * kBM)); }

static int launch_gemm_impl(const op16* A, int64_t lda, const op16* W, int64_t ldw, int M, int N, int K, float alpha,
                            const float* bias, void* out, int64_t ldo, const float* resid, int64_t ldr, int epi,
                            int split_cols, cudaStream_t stream, uint32_t operand_fmt) {
  MSCLIP_REQUIRE(M > 0 && N > 0 && K > 0, "launch_gemm: empty problem");
  MSCLIP_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "launch_gemm: K and operand pitches must be multiples of 8");
  const int bn = split_cols ? split_cols : gemm_pick_bn(N);
  const bool f32_out = (epi == EPI_RESID_F32 || epi == EPI_F32);
  if (epi == EPI_RESID_F32) MSCLIP_REQUIRE(resid != nullptr, "launch_gemm: residual operand missing");
  bool vec_ok = ldo % (f32_out ? 4 : 8) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0);
  if (epi == EPI_RESID_F32) vec_ok = vec_ok && ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(resid) & 15) == 0;
  // CTA pairs for the large transformer GEMMs (N a multiple of 256, at least one full pair tile of rows)
  const int cg = (g_pair_mode >= 1 && bn == 256 && N % 256 == 0 && M >= 256 && split_cols == 0) ? 2 : 1;
  int np = 1;
  if (cg == 2 && g_pair_mode >= 2) np = (g_pair_mode == 4 && M >= 4096) ? 4 : (M >= 1024 ? 2 : 1);
  CUtensorMap ta, tb;
  MSCLIP_TRY(make_tmap_op16_2d(&ta, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(lda),
                               kBM));
  MSCLIP_TRY(make_tmap_op16_2d(&tb, W, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldw),
                               static_cast<uint32_t>(bn / cg / np)));
  GemmParams p = {};
  p.operand_fmt = kUmmaOperandFormat;
  p.M = M;
  p.N = N;
  p.K = K;
  p.tiles_n = (N + bn - 1) / bn;
  p.alpha = alpha;
  p.vec_ok = vec_ok ? 1 : 0;
  p.split_stride = split_cols ? static_cast<long long>(M) * split_cols : 0;
  p.operand_fmt = operand_fmt;
  p.total_tiles = ((M + kBM * cg * np - 1) / (kBM * cg * np)) * p.tiles_n;
  p.bias = bias;
  p.out = out;
  p.resid = resid;
  p.ldo = ldo;
  p.ldr = ldr;
  if (cg == 2 && np == 4) return launch_bn<256, 2, 4>(ta, tb, p, epi, stream);
  if (cg == 2 && np == 2) return launch_bn<256, 2, 2>(ta, tb, p, epi, stream);
  if (cg == 2) return launch_bn<256, 2, 1>(ta, tb, p, epi, stream);
  switch (bn) {
    case 256: return launch_bn<256, 1, 1>(ta, tb, p, epi, stream);
    case 192: return launch_bn<192, 1, 1>(ta, tb, p, epi, stream);
    case 128: return launch_bn<128, 1, 1>(ta, tb, p, epi, stream);
    case 96: return launch_bn<96, 1, 1>(ta, tb, p, epi, stream);
    case 64: return launch_bn<64, 1, 1>(ta, tb, p, epi, stream);
    case 48: return launch_bn<48, 1, 1>(ta, tb, p, epi, stream);
  }
  return 2;
}

// Training forward of fc1: activation AND pre-activation from one epilogue (EPI_QGELU_DUAL_BF16, kernels.h)
int launch_gemm_qgelu_dual(const op16* A, int64_t lda, const op16* W, int64_t ldw, int M, int N, int K, const float* bias, op16* a,
                           op16* u, cudaStream_t stream) {
  MSCLIP_REQUIRE(M > 0 && N % 256 == 0 && K > 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "launch_gemm_qgelu_dual: bad operand shapes");
  MSCLIP_REQUIRE(a && u && bias && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0,
                 "launch_gemm_qgelu_dual: a, u and bias must be 16-byte aligned");
  const int cg = (g_pair_mode >= 1 && M >= 256) ? 2 : 1;
  CUtensorMap ta, tb;
  MSCLIP_TRY(make_tmap_op16_2d(&ta, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(lda), kBM));
  MSCLIP_TRY(make_tmap_op16_2d(&tb, W, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldw),
                               static_cast<uint32_t>(256 / cg)));
  GemmParams p = {};
  p.operand_fmt = kUmmaOperandFormat;
  p.M = M;
  p.N = N;
  p.K = K;
  p.tiles_n = N / 256;
  p.alpha = 1.0f;
  p.vec_ok = 1;
  p.total_tiles = ((M + kBM * cg - 1) / (kBM * cg)) * p.tiles_n;
  p.bias = bias;
  p.out = a;
  p.out2 = u;
  p.ldo = N;
  return cg == 2 ? launch_variant<256, EPI_QGELU_DUAL_BF16, 2, 1>(ta, tb, p, stream)
                 : launch_variant<256, EPI_QGELU_DUAL_BF16, 1, 1>(ta, tb, p, stream);
}

// Backward of the MLP's activation on the dgrad GEMM's epilogue (EPI_DGELU_BF16, kernels.h)
int launch_gemm_dgelu(const op16* A, int64_t lda, const op16* W, int64_t ldw, int M, int N, int K, const op16* u, op16* du, op16* a,
                      cudaStream_t stream) {
  MSCLIP_REQUIRE(M > 0 && N % 256 == 0 && K > 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "launch_gemm_dgelu: bad operand shapes");
  MSCLIP_REQUIRE(u && du && a && ((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(du) | reinterpret_cast<uintptr_t>(a)) & 15) == 0,
                 "launch_gemm_dgelu: u, du and a must be 16-byte aligned");
  const int cg = (g_pair_mode >= 1 && M >= 256) ? 2 : 1;
  CUtensorMap ta, tb;
  MSCLIP_TRY(make_tmap_op16_2d(&ta, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(lda), kBM));
  MSCLIP_TRY(make_tmap_op16_2d(&tb, W, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldw),
                               static_cast<uint32_t>(256 / cg)));
  GemmParams p = {};
  p.operand_fmt = kUmmaOperandFormat;
  p.M = M;
  p.N = N;
  p.K = K;
  p.tiles_n = N / 256;
  p.alpha = 1.0f;
  p.vec_ok = 1;
  p.total_tiles = ((M + kBM * cg - 1) / (kBM * cg)) * p.tiles_n;
  p.out = du;
  p.out2 = a;
  p.aux16 = u;
  p.ldo = N;
  return cg == 2 ? launch_variant<256, EPI_DGELU_BF16, 2, 1>(ta, tb, p, stream) : launch_variant<256, EPI_DGELU_BF16, 1, 1>(ta, tb, p, stream);
}

// ---- implicit-GEMM convolution fed by im2col-mode TMA ---------------------------------------------------------
// K layout of the packed weight: for every source, every filter tap (ky, kx) and every 64-channel block: 64 columns
// (channels beyond C are zero) - conv_tma_kpad() columns in total.
int conv_tma_kpad(const ConvSource* src, int nsrc) {
  int k = 0;
  for (int i = 0; i < nsrc; ++i) k += src[i].ksize * src[i].ksize * ((src[i].C + kBK - 1) / kBK) * kBK;
  return k;
}

namespace {
template <int BN, int CG>
int launch_conv_tma_bn(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tb, const GemmParams& p, int epi,
                       cudaStream_t stream) {
  switch (epi) {
    case EPI_RELU_BF16: return launch_variant<BN, EPI_RELU_BF16, CG, 1, 0, 8, 1>(ta, tb, p, stream, &ta2);
    case EPI_F32: return launch_variant<BN, EPI_F32, CG, 1, 0, 8, 1>(ta, tb, p, stream, &ta2);
  }
  set_last_error("launch_conv_tma: unsupported epilogue " + std::to_string(epi));
  return 2;
}
}  // namespace

// MSCLIP_CONV_PAIR=0: one CTA per 128-pixel tile instead of CTA pairs (A/B timing)
static const bool g_conv_pair = [] {
  const char* e = getenv("MSCLIP_CONV_PAIR");
  return e == nullptr || e[0] != '0';
}();

int launch_conv_tma(const ConvSource* src, int nsrc, int batch, int Ho, int Wo, const op16* Wp, int64_t ldw, int N,
                    const float* bias, void* out, int64_t ldo, int epi, cudaStream_t stream) {
  MSCLIP_REQUIRE(nsrc == 1 || nsrc == 2, "conv_tma: one or two sources");
  MSCLIP_REQUIRE(batch > 0 && Ho > 0 && Wo > 0 && N > 0, "conv_tma: empty problem");
  const int K = conv_tma_kpad(src, nsrc);
  MSCLIP_REQUIRE(ldw % 8 == 0 && ldw >= K, "conv_tma: weight pitch (padded K layout, see conv_tma_kpad)");
  int bn = 0;
  const int cands[4] = {256, 192, 96, 48};
  for (int i = 0; i < 4 && bn == 0; ++i)
    if (N % cands[i] == 0) bn = cands[i];
  MSCLIP_REQUIRE(bn != 0, "conv_tma: N must be a multiple of 48");
  const long long M = static_cast<long long>(batch) * Ho * Wo;
  MSCLIP_REQUIRE(M < (1ll << 31), "conv_tma: too many output pixels for one launch");
  const bool f32_out = (epi == EPI_F32);
  const bool vec_ok = ldo % (f32_out ? 4 : 8) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                      (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0);
  MSCLIP_REQUIRE(vec_ok, "conv_tma: output rows and bias must be 16-byte aligned");
  GemmParams p = {};
  p.operand_fmt = kUmmaOperandFormat;
  CUtensorMap ta[2], tb;
  int kb = 0;
  for (int i = 0; i < nsrc; ++i) {
    const ConvSource& s = src[i];
    MSCLIP_REQUIRE(s.C % 8 == 0 && s.cpix % 8 == 0 && s.c_off % 8 == 0, "conv_tma: channels must be multiples of 8");
    MSCLIP_REQUIRE((s.H + 2 * s.pad - s.ksize) / s.stride + 1 == Ho && (s.W + 2 * s.pad - s.ksize) / s.stride + 1 == Wo,
                   "conv_tma: source geometry does not produce the output grid");
    MSCLIP_TRY(make_tmap_im2col_nhwc(&ta[i], s.in, batch, s.H, s.W, s.cpix, s.c_off, s.C, s.ksize, s.stride, s.pad));
    p.conv.ksize[i] = s.ksize;
    p.conv.stride[i] = s.stride;
    p.conv.pad[i] = s.pad;
    p.conv.cblk[i] = (s.C + kBK - 1) / kBK;
    if (i == 1) p.conv.kb_begin1 = kb;
    kb += s.ksize * s.ksize * p.conv.cblk[i];
  }
  if (nsrc == 1) {
    ta[1] = ta[0];
    p.conv.kb_begin1 = kb;  // never reached
    p.conv.ksize[1] = p.conv.stride[1] = p.conv.cblk[1] = 1;
  }
  p.conv.Ho = Ho;
  p.conv.Wo = Wo;
  find_divisor(static_cast<uint32_t>(Ho) * static_cast<uint32_t>(Wo), &p.conv.hw_mul, &p.conv.hw_shr);
  find_divisor(static_cast<uint32_t>(Wo), &p.conv.wo_mul, &p.conv.wo_shr);
  const int cg = (g_conv_pair && M >= 256) ? 2 : 1;
  MSCLIP_TRY(make_tmap_op16_2d(&tb, Wp, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldw),
                               static_cast<uint32_t>(bn / cg)));
  p.M = static_cast<int>(M);
  p.N = N;
  p.K = K;
  p.tiles_n = N / bn;
  p.alpha = 1.0f;
  p.vec_ok = 1;
  p.total_tiles = static_cast<int>((M + kBM * cg - 1) / (kBM * cg)) * p.tiles_n;
  p.bias = bias;
  p.out = out;
  p.ldo = ldo;
  if (cg == 2) {
    switch (bn) {
      case 256: return launch_conv_tma_bn<256, 2>(ta[0], ta[1], tb, p, epi, stream);
      case 192: return launch_conv_tma_bn<192, 2>(ta[0], ta[1], tb, p, epi, stream);
      case 96: return launch_conv_tma_bn<96, 2>(ta[0], ta[1], tb, p, epi, stream);
      case 48: return launch_conv_tma_bn<48, 2>(ta[0], ta[1], tb, p, epi, stream);
    }
  }
  switch (bn) {
    case 256: return launch_conv_tma_bn<256, 1>(ta[0], ta[1], tb, p, epi, stream);
    case 192: return launch_conv_tma_bn<192, 1>(ta[0], ta[1], tb, p, epi, stream);
    case 96: return launch_conv_tma_bn<96, 1>(ta[0], ta[1], tb, p, epi, stream);
    case 48: return launch_conv_tma_bn<48, 1>(ta[0], ta[1], tb, p, epi, stream);
  }
  return 2;
}

}  // namespace msclip
