// Implicit-GEMM convolution on tcgen05: out[M = B*Ho*Wo, N] = relu(patches(in)[M, K] . W[N, K]^T + bias).
//
// Replaces im2col + GEMM for the 3x3 (and strided 1x1) convolutions of the early-conv stem
// (ResBasicBlock_v0, M.py:1920-1936) and of the parallel branch (ConvResBlock, M.py:1842-1861): the A operand
// is never materialised in HBM.  128 gather threads (one output pixel each) copy the 16-byte channel vectors
// of their patch straight from the NHWC activation into the 128B-swizzled smem stage with cp.async (zero
// fill for padding), so every activation byte is read from HBM once and re-used out of L2 by the 2.25
// (stride 2) to 9 (stride 1) windows that contain it.  The K axis may be the concatenation of two sources
// (ConvResBlock: [conv3 input | strided shortcut input], one GEMM).  W arrives by TMA; MMA issue, TMEM double
// buffering and the fused epilogue are those of gemm.cu.
//
// Warp roles (512 threads, 1 CTA / SM, persistent):
//   warps 0-7  : A gather (8 lanes per tile row, 4 rows per lane and k-block)
//   warp 8     : W TMA producer (one lane)      warp 9 : MMA issuer (one lane)     warp 10 : TMEM allocator
//   warps 12-15: epilogue (one per TMEM lane quarter; the tiles are at most 192 columns wide)
// The gather warps are the critical resource (profiles/r01d_conv_ncu.md: they were busy ~80 % of the time on
// address arithmetic while every other unit idled), so their inner loop is stripped to a table lookup, a bit
// test and the copy: per tile row the thread keeps the patch origin as an element offset plus a bit mask of the
// taps that fall inside the image, per k-chunk a table gives the tap's offset and bit.
#include "gemm_common.cuh"

namespace msclip {

namespace {

using namespace gemm_detail;

constexpr int kConvThreads = 512;
constexpr int kGatherWarps = 8;
constexpr int kGatherThreads = 32 * kGatherWarps;
constexpr int kRowsPerLane = kBM / (4 * kGatherWarps);  // 4
constexpr int kTmaWarp = kGatherWarps, kMmaWarp = kGatherWarps + 1, kAllocWarp = kGatherWarps + 2;
constexpr int kFirstEpiWarp = 12;
constexpr int kConvEpiWarps = 4;
constexpr int kMaxKChunks = 512;  // K <= 4096

struct ConvSeg {
  const op16* in;
  int H, W, cpix, c_off, C, ksize, stride, pad;
  int k_begin;  // first K index of this source
};

struct ConvParams {
  ConvSeg seg[2];
  int nseg;
  int Ho, Wo;
  uint32_t hw_mul, hw_shr, wo_mul, wo_shr;  // n / (Ho*Wo) and n / Wo as multiply-high + shift (n < 2^31)
  GemmParams g;
};


template <int BN>
struct ConvCfg {
  static constexpr int kStageA = kBM * kBK * 2;
  static constexpr int kStageB = BN * kBK * 2;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStagesRaw = (184 * 1024) / kStage;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTableBytes = kMaxKChunks * 8;
  static constexpr int kSmemBytes = kStages * kStage + kTableBytes + 256 + 1024;
  static constexpr int kChunk = (BN % 32 == 0) ? 32 : 16;
  static constexpr int kNumChunks = BN / kChunk;
  // cp.async groups a gather thread keeps in flight before it hands the oldest one to the MMA warp.  The ring is
  // split between bytes in flight (LAG stages landing) and slack (stages already handed over that the MMA warp can
  // run through while the gather threads sleep on a free-slot barrier): with LAG = kStages - 1 every k-block paid
  // the full MMA-done -> gather-wakes -> issue -> arrive -> MMA-wakes round trip (profiles/r01d_conv_ncu.md)
  static constexpr int kLagHalf = kStages / 2;
  static constexpr int kLagDeep = kStages - 2 > 0 ? kStages - 2 : 1;
  static_assert(kStages >= 3, "ring too shallow");
};

__device__ __forceinline__ void cp_async_16_zfill(uint32_t smem_dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int BN, int EPI, int LAG>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmap_b, const ConvParams cp) {
  using Cfg = ConvCfg<BN>;
  const GemmParams& p = cp.g;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  int2* ktab = reinterpret_cast<int2*>(smem + Cfg::kStages * Cfg::kStage);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(ktab) + Cfg::kTableBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.K + kBK - 1) / kBK;

  // K-chunk table: chunk q (K indices 8q .. 8q+7) -> x: element offset of the tap inside the source
  // ((ky*W + kx)*cpix + c), y: tap index ky*ksize + kx | source << 16; y = -1 marks chunks beyond K (zero filled)
  for (int q = threadIdx.x; q < num_kb * 8; q += kConvThreads) {
    const int k = q * 8;
    int2 e = make_int2(0, -1);
    if (k < p.K) {
      const int sidx = (cp.nseg > 1 && k >= cp.seg[1].k_begin) ? 1 : 0;
      const ConvSeg& sg = cp.seg[sidx];
      const int kk = k - sg.k_begin;
      const int tap = kk / sg.C, c = kk - tap * sg.C;
      const int ky = tap / sg.ksize, kx = tap - ky * sg.ksize;
      e.x = (ky * sg.W + kx) * sg.cpix + c;
      e.y = tap | (sidx << 16);
    }
    ktab[q] = e;
  }
  if (warp == kTmaWarp && lane == 0) tma_prefetch_desc(&tmap_b);
  if (warp == kMmaWarp && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], kGatherThreads + 1);  // every gather thread + the W producer's expect_tx arrive
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kConvEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == kAllocWarp) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kGatherWarps) {
    // ------------------------------------------------------------------ A gather
    // Lane l of warp w copies chunk j = l & 7 of tile rows w*16 + 4*i + (l >> 3), i = 0..3: the eight lanes that
    // share a row fetch its 128 contiguous-in-K bytes (one or two contiguous NHWC segments) and write one full,
    // conflict-free swizzled smem row, so each warp-wide cp.async touches 4 rows instead of 32.
    const int j = lane & 7;
    const int rsub = lane >> 3;
    const int row0 = warp * (4 * kRowsPerLane) + rsub;
    const op16* const in0 = cp.seg[0].in;
    const op16* const in1 = cp.seg[1].in;
    const int hw = cp.Ho * cp.Wo;
    uint32_t dst_off[kRowsPerLane];
#pragma unroll
    for (int i = 0; i < kRowsPerLane; ++i) {
      const uint32_t row = static_cast<uint32_t>(row0 + 4 * i);
      dst_off[i] = row * 128u + ((static_cast<uint32_t>(j) ^ (row & 7u)) << 4);
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      // per row: element offset of the patch origin in each source and the bit mask of the taps inside the image
      int pix_off[2][kRowsPerLane];
      uint32_t tapmask[2][kRowsPerLane];
      {
        const int m0 = (tile / p.tiles_n) * kBM + row0;
        int b = fast_div(m0, cp.hw_mul, cp.hw_shr);
        const int rem = m0 - b * hw;
        int oy = fast_div(rem, cp.wo_mul, cp.wo_shr), ox = rem - oy * cp.Wo;
#pragma unroll
        for (int i = 0; i < kRowsPerLane; ++i) {
          const bool row_ok = m0 + 4 * i < p.M;
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            if (s2 >= cp.nseg) {  // single source: the second slot is never selected
              pix_off[s2][i] = 0;
              tapmask[s2][i] = 0;
              continue;
            }
            const ConvSeg& sg = cp.seg[s2];
            const int iy0 = oy * sg.stride - sg.pad, ix0 = ox * sg.stride - sg.pad;
            pix_off[s2][i] = ((b * sg.H + iy0) * sg.W + ix0) * sg.cpix + sg.c_off;
            uint32_t mk = 0;
            if (row_ok) {
              if (iy0 >= 0 && ix0 >= 0 && iy0 + sg.ksize <= sg.H && ix0 + sg.ksize <= sg.W) {
                mk = (2u << (sg.ksize * sg.ksize - 1)) - 1u;  // interior pixel: every tap reads inside the image
              } else {
                // taps ky in [ky_lo, ky_hi] and kx in [kx_lo, kx_hi] read inside the image
                const int ky_lo = iy0 < 0 ? -iy0 : 0, kx_lo = ix0 < 0 ? -ix0 : 0;
                const int ky_hi = min(sg.ksize - 1, sg.H - 1 - iy0), kx_hi = min(sg.ksize - 1, sg.W - 1 - ix0);
                if (kx_hi >= kx_lo) {
                  const uint32_t cols = ((2u << kx_hi) - 1u) & ~((1u << kx_lo) - 1u);
                  for (int ky = ky_lo; ky <= ky_hi; ++ky) mk |= cols << (ky * sg.ksize);
                }
              }
            }
            tapmask[s2][i] = mk;
          }
          ox += 4;
          while (ox >= cp.Wo) {
            ox -= cp.Wo;
            ++oy;
          }
          while (oy >= cp.Ho) {
            oy -= cp.Ho;
            ++b;
          }
        }
      }
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % Cfg::kStages;
        const uint32_t ph = static_cast<uint32_t>(it / Cfg::kStages) & 1u;
        const int2 e = ktab[kb * 8 + j];
        const bool second = ((e.y >> 16) & 1) != 0;
        const uint32_t tap = static_cast<uint32_t>(e.y) & 0xFFu;
        const bool chunk_ok = e.y != -1;
        const op16* const base = second ? in1 : in0;
        if (it >= LAG) {
          // hand over the block issued LAG iterations ago BEFORE sleeping on a free slot: the MMA warp must never
          // wait for data that has already landed
          cp_async_wait<LAG - 1>();
          fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async proxy
          mbar_arrive(&full_bar[(it - LAG) % Cfg::kStages]);
        }
        mbar_wait(&empty_bar[s], ph ^ 1u, 21);
        const uint32_t stage_base = smem_u32(smem + s * Cfg::kStage);
#pragma unroll
        for (int i = 0; i < kRowsPerLane; ++i) {
          const uint32_t mk = second ? tapmask[1][i] : tapmask[0][i];
          const bool ok = chunk_ok && ((mk >> tap) & 1u) != 0;
          const int off = (second ? pix_off[1][i] : pix_off[0][i]) + e.x;
          cp_async_16_zfill(stage_base + dst_off[i], ok ? base + off : in0, ok ? 16u : 0u);
        }
        cp_async_commit();
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    for (int d = (it < LAG ? it : LAG); d > 0; --d) mbar_arrive(&full_bar[(it - d) % Cfg::kStages]);
  } else if (warp == kTmaWarp) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n0 = (tile % p.tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1, 22);
          mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageB);
          tma_load_2d(smem + s * Cfg::kStage + Cfg::kStageA, &tmap_b, &full_bar[s], kb * kBK, n0);
          if (++s == Cfg::kStages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f32acc(kBM, BN);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], aph ^ 1, 23);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph, 24);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * Cfg::kStage);
          const uint32_t b_addr = a_addr + Cfg::kStageA;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_f16(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          if (++s == Cfg::kStages) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit(&tfull_bar[as]);
        as ^= 1;
        if (as == 0) aph ^= 1;
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    constexpr int kPerHalf = Cfg::kNumChunks;
    constexpr int c_begin = 0, c_end = Cfg::kNumChunks;
    int as = 0;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.tiles_n) * kBM;
      const int n0 = (tile % p.tiles_n) * BN;
      mbar_wait(&tfull_bar[as], aph, 25);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccStride;
      const int row = m0 + q * 32 + lane;
      if (c_begin < c_end) {
        uint32_t acc[2][Cfg::kChunk];
        tmem_ld_chunk<Cfg::kChunk>(taddr + c_begin * Cfg::kChunk, acc[0]);
#pragma unroll
        for (int i = 0; i < kPerHalf; ++i) {
          const int c = c_begin + i;
          if (c < c_end) {
            const int col0 = n0 + c * Cfg::kChunk;
            const bool fast = p.vec_ok && (col0 + Cfg::kChunk <= p.N);
            EpiOperands<EPI, Cfg::kChunk> ops;
            epilogue_prefetch<EPI, Cfg::kChunk>(ops, p, row, col0, fast);
            tmem_ld_wait();
            if (c + 1 < c_end) tmem_ld_chunk<Cfg::kChunk>(taddr + (c + 1) * Cfg::kChunk, acc[(i + 1) & 1]);
            epilogue_store<EPI, Cfg::kChunk>(acc[i & 1], ops, p, row, col0, fast);
          }
        }
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
  if (warp == kAllocWarp) tmem_dealloc(tmem_base, 512);
}

template <int BN, int EPI, int LAG>
int launch_conv_lag(const CUtensorMap& tb, const ConvParams& cp, cudaStream_t stream) {
  using Cfg = ConvCfg<BN>;
  static bool configured = false;
  if (!configured) {
    MSCLIP_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN, EPI, LAG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg::kSmemBytes));
    configured = true;
  }
  const int grid = cp.g.total_tiles < num_sms() ? cp.g.total_tiles : num_sms();
  conv_gemm_kernel<BN, EPI, LAG><<<grid, kConvThreads, Cfg::kSmemBytes, stream>>>(tb, cp);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// MSCLIP_CONV_LAG=deep keeps kStages - 2 gather groups in flight instead of kStages / 2 (A/B timing)
static const bool g_lag_deep = [] {
  const char* e = getenv("MSCLIP_CONV_LAG");
  return e != nullptr && e[0] == 'd';
}();

template <int BN, int EPI>
int launch_conv_variant(const CUtensorMap& tb, const ConvParams& cp, cudaStream_t stream) {
  using Cfg = ConvCfg<BN>;
  if (g_lag_deep && Cfg::kLagDeep != Cfg::kLagHalf) return launch_conv_lag<BN, EPI, Cfg::kLagDeep>(tb, cp, stream);
  return launch_conv_lag<BN, EPI, Cfg::kLagHalf>(tb, cp, stream);
}

template <int BN>
int launch_conv_bn(const CUtensorMap& tb, const ConvParams& cp, int epi, cudaStream_t stream) {
  switch (epi) {
    case EPI_BF16: return launch_conv_variant<BN, EPI_BF16>(tb, cp, stream);
    case EPI_RELU_BF16: return launch_conv_variant<BN, EPI_RELU_BF16>(tb, cp, stream);
    case EPI_F32: return launch_conv_variant<BN, EPI_F32>(tb, cp, stream);
  }
  set_last_error("launch_conv_gemm: unsupported epilogue " + std::to_string(epi));
  return 2;
}

}  // namespace

static int conv_pick_bn(int N) {  // 256-wide tiles would leave only 3 smem stages next to the gather's cp.async lag
  const int cands[5] = {192, 128, 96, 64, 48};
  for (int i = 0; i < 5; ++i)
    if (N % cands[i] == 0) return cands[i];
  if (N >= 192) return 192;
  for (int i = 4; i >= 0; --i)
    if (cands[i] >= N) return cands[i];
  return 192;
}

int launch_conv_gemm(const ConvSource* src, int nsrc, int batch, int Ho, int Wo, const op16* W, int64_t ldw, int N,
                     const float* bias, void* out, int64_t ldo, int epi, cudaStream_t stream) {
  MSCLIP_REQUIRE(nsrc == 1 || nsrc == 2, "conv_gemm: one or two sources");
  MSCLIP_REQUIRE(batch > 0 && Ho > 0 && Wo > 0 && N > 0, "conv_gemm: empty problem");
  ConvParams cp = {};
  cp.nseg = nsrc;
  cp.Ho = Ho;
  cp.Wo = Wo;
  int K = 0;
  for (int i = 0; i < nsrc; ++i) {
    const ConvSource& s = src[i];
    MSCLIP_REQUIRE(s.C % 8 == 0 && s.cpix % 8 == 0 && s.c_off % 8 == 0, "conv_gemm: channels must be multiples of 8");
    MSCLIP_REQUIRE(s.ksize >= 1 && s.ksize <= 5 && s.stride >= 1, "conv_gemm: kernel size must be in [1, 5]");
    MSCLIP_REQUIRE((s.H + 2 * s.pad - s.ksize) / s.stride + 1 == Ho && (s.W + 2 * s.pad - s.ksize) / s.stride + 1 == Wo,
                   "conv_gemm: source geometry does not produce the output grid");
    MSCLIP_REQUIRE((reinterpret_cast<uintptr_t>(s.in) & 15) == 0, "conv_gemm: input must be 16-byte aligned");
    cp.seg[i] = ConvSeg{static_cast<const op16*>(s.in), s.H, s.W, s.cpix, s.c_off, s.C, s.ksize, s.stride, s.pad, K};
    K += s.ksize * s.ksize * s.C;
  }
  if (nsrc == 1) cp.seg[1] = cp.seg[0];
  find_divisor(static_cast<uint32_t>(Ho) * static_cast<uint32_t>(Wo), &cp.hw_mul, &cp.hw_shr);
  find_divisor(static_cast<uint32_t>(Wo), &cp.wo_mul, &cp.wo_shr);
  MSCLIP_REQUIRE(K <= kMaxKChunks * 8, "conv_gemm: K too large");
  MSCLIP_REQUIRE(ldw % 8 == 0 && ldw >= K, "conv_gemm: weight pitch");
  const int bn = conv_pick_bn(N);
  const bool f32_out = (epi == EPI_F32);
  const bool vec_ok = ldo % (f32_out ? 4 : 8) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                      (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0);
  CUtensorMap tb;
  MSCLIP_TRY(make_tmap_op16_2d(&tb, W, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldw),
                               static_cast<uint32_t>(bn)));
  const long long M = static_cast<long long>(batch) * Ho * Wo;
  MSCLIP_REQUIRE(M < (1ll << 31), "conv_gemm: too many output pixels for one launch");
  for (int i = 0; i < nsrc; ++i)
    MSCLIP_REQUIRE(static_cast<long long>(batch) * src[i].H * src[i].W * src[i].cpix < (1ll << 31) && src[i].pad <= 32,
                   "conv_gemm: source too large for 32-bit element offsets");
  GemmParams& p = cp.g;
  p.M = static_cast<int>(M);
  p.N = N;
  p.K = K;
  p.tiles_n = (N + bn - 1) / bn;
  p.total_tiles = static_cast<int>((M + kBM - 1) / kBM) * p.tiles_n;
  p.bias = bias;
  p.out = out;
  p.resid = nullptr;
  p.ldo = ldo;
  p.ldr = 0;
  p.alpha = 1.0f;
  p.vec_ok = vec_ok ? 1 : 0;
  p.split_stride = 0;
  switch (bn) {
    case 192: return launch_conv_bn<192>(tb, cp, epi, stream);
    case 128: return launch_conv_bn<128>(tb, cp, epi, stream);
    case 96: return launch_conv_bn<96>(tb, cp, epi, stream);
    case 64: return launch_conv_bn<64>(tb, cp, epi, stream);
    case 48: return launch_conv_bn<48>(tb, cp, epi, stream);
  }
  return 2;
}

}  // namespace msclip
