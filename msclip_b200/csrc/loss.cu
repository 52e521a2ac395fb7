// Global-batch contrastive loss without materialising the G x G logits (reference: CLIP.forward builds
// logits = exp(logit_scale) * I_all @ T_all^T on every rank after two NCCL all-gathers, M.py:3136-3141,
// lib/utils/comm.py:140-154; the symmetric cross-entropy itself is the north star's).
//
// Formulation: rank r owns rows [r*B, (r+1)*B).  For both directions (image rows vs all text columns,
// text rows vs all image columns) it needs  lse_i = log sum_j exp(s_ij)  over all G columns and the
// diagonal s_ii.  One CTA owns one 128-column tile of the *gathered* matrix: it pulls that tile straight
// from the owning rank's shard (a local or an NVLink peer-mapped pointer - this is the all-gather, issued
// as plain loads from inside the kernel, each peer byte crossing NVLink exactly once), keeps it resident
// in 128B-swizzled shared memory as the UMMA B operand, and streams the local rows past it with TMA
// (L2-resident) through tcgen05.mma into double-buffered TMEM accumulators.  The epilogue warps turn
// each 128x128 score tile into per-row (max, sum-exp) partials in registers; a second tiny kernel merges
// the partials across column tiles, subtracts the diagonal and reduces deterministically.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace msclip {

namespace {

constexpr int kLossE = 512;            // embedding width (K of the similarity GEMM)
constexpr int kLossKB = kLossE / 64;   // 8 k-blocks
constexpr int kLossBN = 128;           // columns per CTA
constexpr int kLossStages = 4;
constexpr int kLossStageBytes = 128 * 64 * 2;  // one A k-block: 16 KB
constexpr int kLossBBytes = kLossKB * kLossBN * 64 * 2;  // 128 KB resident column tile
constexpr int kLossSmem = kLossBBytes + kLossStages * kLossStageBytes + 256 + 1024;
constexpr int kLossThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct LossParams {
  const emb16* const* col_shards[2];  // [dir] -> device table of `world` shard pointers (columns of that direction)
  const uint32_t* flags;             // optional: this rank's flag array [world], raised by the owners; null for world == 1
  uint32_t epoch;
  int world, rank, b_local, tiles_per_shard, n_col_tiles, n_row_blocks, b_pad;
  float scale_log2;                  // exp(logit_scale) * log2(e): scores are kept in the log2 domain
  float2* ws;                        // [2][n_col_tiles][b_pad] (max, sum) partials, log2 domain
  float* diag;                       // [2][b_pad] the diagonal scores s_ii taken from the same accumulators
  // MODE 1 (backward): instead of the partials the epilogue writes w_ij = exp(s_ij - lse_row_i) + exp(s_ij - lse_col_j)
  const float* row_lse;              // [2][b_pad] log2-domain lse of this rank's rows (image rows | text rows)
  const float* const* col_lse[2];    // [dir] -> device table of `world` pointers: the column owners' row lse of the OTHER direction
  emb16* wout[2];                    // [dir] -> [b_local][ldw] probabilities * wscale (fp16), column = owner * b_local + local column
  float wscale;                      // power of two that lifts probabilities of order 1 / G out of fp16's subnormal range
  long long ldw;
  int lse_pitch;                     // floats between the two directions of row_lse
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int MODE>
__global__ void __launch_bounds__(kLossThreads, 1)
contrastive_lse_kernel(const __grid_constant__ CUtensorMap tmap_rows_img, const __grid_constant__ CUtensorMap tmap_rows_txt,
                       const LossParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;
  uint8_t* sA = smem + kLossBBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sA + kLossStages * kLossStageBytes);
  uint64_t* empty_bar = full_bar + kLossStages;
  uint64_t* tfull_bar = empty_bar + kLossStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ct = blockIdx.x;   // column tile of the gathered matrix
  const int dir = blockIdx.y;  // 0: image rows x text columns, 1: text rows x image columns
  const CUtensorMap* tmap_rows = dir == 0 ? &tmap_rows_img : &tmap_rows_txt;
  const int owner = ct / p.tiles_per_shard;
  const int c0 = (ct % p.tiles_per_shard) * kLossBN;
  const int valid_cols = min(kLossBN, p.b_local - c0);

  if (warp == 0 && lane == 0) tma_prefetch_desc(tmap_rows);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kLossStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }

  // ---- in-kernel gather of this column tile from its owner (local or NVLink peer) -------------------
  if (p.flags != nullptr) {
    // the owner publishes `epoch` with a system-scope release once its embeddings are written
    if (threadIdx.x == 0) {
      const uint32_t* f = p.flags + owner;
      const uint64_t t0 = global_timer_ns();
      while (static_cast<int32_t>(ld_acquire_sys(f) - p.epoch) < 0) {
        if (global_timer_ns() - t0 > 20000000000ull) {  // bounded wait: a dead peer must not hang the GPU
          printf("msclip: peer %d never published epoch %u\n", owner, p.epoch);
          __trap();
        }
      }
    }
    __syncthreads();
  }
  {
    const emb16* src = p.col_shards[dir][owner] + static_cast<long long>(c0) * kLossE;
    // 128 rows x 64 chunks of 16 B; adjacent threads read adjacent chunks of one row (coalesced)
    for (int i = threadIdx.x; i < kLossBN * (kLossE / 8); i += kLossThreads) {
      const int r = i >> 6;
      const int cc = i & 63;
      const int kb = cc >> 3, c = cc & 7;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (r < valid_cols) v = __ldcg(reinterpret_cast<const uint4*>(src + static_cast<long long>(r) * kLossE + cc * 8));
      *reinterpret_cast<uint4*>(sB + kb * (kLossBN * 128) + r * 128 + ((c ^ (r & 7)) << 4)) = v;
    }
    fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int rb = 0; rb < p.n_row_blocks; ++rb) {
        for (int kb = 0; kb < kLossKB; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1, 11);
          mbar_arrive_expect_tx(&full_bar[s], kLossStageBytes);
          tma_load_2d(sA + s * kLossStageBytes, tmap_rows, &full_bar[s], kb * 64, rb * 128);
          if (++s == kLossStages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f32acc_fmt(128, kLossBN, kUmmaFormatF16);
      const uint32_t b_base = smem_u32(sB);
      int s = 0;
      uint32_t ph = 0;
      for (int rb = 0; rb < p.n_row_blocks; ++rb) {
        const int as = rb & 1;
        const uint32_t aph = (rb >> 1) & 1;
        mbar_wait(&tempty_bar[as], aph ^ 1, 12);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kLossBN;
        for (int kb = 0; kb < kLossKB; ++kb) {
          mbar_wait(&full_bar[s], ph, 13);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + s * kLossStageBytes);
          const uint32_t b_addr = b_base + kb * (kLossBN * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          if (++s == kLossStages) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit(&tfull_bar[as]);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    float2* ws = p.ws + (static_cast<long long>(dir) * p.n_col_tiles + ct) * p.b_pad;
    for (int rb = 0; rb < p.n_row_blocks; ++rb) {
      const int as = rb & 1;
      const uint32_t aph = (rb >> 1) & 1;
      mbar_wait(&tfull_bar[as], aph, 14);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kLossBN;
      if (MODE == 1) {
        // backward: softmax probabilities of both directions for this 128 x 128 tile of (local rows) x (gathered columns)
        const int row = rb * 128 + q * 32 + lane;
        const float lse_r = row < p.b_local ? p.row_lse[dir * p.lse_pitch + row] : 0.f;
        const float* lse_c = p.col_lse[dir][owner] + c0;
        emb16* wrow = p.wout[dir] + static_cast<long long>(row) * p.ldw + static_cast<long long>(owner) * p.b_local + c0;
        // the positive pair of a row: its "- 2 x_i" term is folded into the matrix in fp32 (w_ii - 2 is small when the
        // softmax is peaked; rounding w_ii ~ 2 to 16 bits first would swamp the gradient)
        const int diag_col = (owner == p.rank) ? row - c0 : -1;
#pragma unroll 1
        for (int c = 0; c < kLossBN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (row < p.b_local) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const int col = c * 32 + j;
              float w0 = 0.f, w1 = 0.f;
              if (col < valid_cols) {
                const float v = __uint_as_float(r[j]) * p.scale_log2;
                w0 = exp2f(v - lse_r) + exp2f(v - __ldg(lse_c + col));
                if (col == diag_col) w0 = (exp2f(v - lse_r) - 1.0f) + (exp2f(v - __ldg(lse_c + col)) - 1.0f);
              }
              if (col + 1 < valid_cols) {
                const float v = __uint_as_float(r[j + 1]) * p.scale_log2;
                w1 = exp2f(v - lse_r) + exp2f(v - __ldg(lse_c + col + 1));
                if (col + 1 == diag_col) w1 = (exp2f(v - lse_r) - 1.0f) + (exp2f(v - __ldg(lse_c + col + 1)) - 1.0f);
              }
              // element-wise stores: b_local (hence the column offset of a shard) need not be even
              if (col < valid_cols) wrow[col] = __float2half_rn(w0 * p.wscale);
              if (col + 1 < valid_cols) wrow[col + 1] = __float2half_rn(w1 * p.wscale);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[as]);
        continue;
      }
      float m = -INFINITY, l = 0.f;
      // the tile that holds the positives of these rows: row i of the block meets column i of the tile.
      // Taking s_ii from the very accumulator that also enters the log-sum-exp keeps lse_i - s_ii free of
      // cancellation error (it is >= 0 by construction), which matters at large logit scales.
      const bool diag_tile = (owner == p.rank) && (rb * 128 == c0);
      float diag = 0.f;
#pragma unroll 1
      for (int c = 0; c < kLossBN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
        float v[32];
        float cm = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = (c * 32 + j < valid_cols) ? __uint_as_float(r[j]) * p.scale_log2 : -INFINITY;
          cm = fmaxf(cm, v[j]);
        }
        if (diag_tile && c == q) {
          float d = v[0];
#pragma unroll
          for (int j = 1; j < 32; ++j) d = (lane == j) ? v[j] : d;
          diag = d;
        }
        const float mn = fmaxf(m, cm);
        if (mn != -INFINITY) {
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) acc += exp2f(v[j] - mn);
          l = l * exp2f(m - mn) + acc;
          m = mn;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      const int row = rb * 128 + q * 32 + lane;
      if (row < p.b_local) {
        ws[row] = make_float2(m, l);
        if (diag_tile) p.diag[dir * p.b_pad + row] = diag;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

// per local row and direction: merge the column-tile partials, subtract the diagonal score
__global__ void __launch_bounds__(256)
lse_combine_kernel(const float2* __restrict__ ws, const float* __restrict__ diag, int b_local, int b_pad,
                   int n_col_tiles, float* __restrict__ row_loss, float* __restrict__ lse_out, int lse_pitch) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= 2 * b_local) return;
  const int dir = idx / b_local, row = idx % b_local;
  const float2* w = ws + static_cast<long long>(dir) * n_col_tiles * b_pad + row;
  float m = -INFINITY;
  for (int t = 0; t < n_col_tiles; ++t) m = fmaxf(m, w[static_cast<long long>(t) * b_pad].x);
  float l = 0.f;
  for (int t = 0; t < n_col_tiles; ++t) {
    const float2 e = w[static_cast<long long>(t) * b_pad];
    l += e.y * exp2f(e.x - m);
  }
  // scores are in log2 units (scale * log2 e folded in): lse_i - s_ii = ln2 * ((m - d) + log2 l)
  row_loss[idx] = kLn2 * ((m - diag[dir * b_pad + row]) + log2f(l));
  if (lse_out != nullptr) lse_out[dir * lse_pitch + row] = m + log2f(l);  // log2 domain; the backward pass reads it
}

// out[k][owner * b_local + r] = shards[owner][r][k] (fp16, exactly as exchanged): the gathered embeddings, transposed, as the K-major B operand
// of the gradient GEMM (K = global batch); 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256)
gather_transpose_kernel(const emb16* const* __restrict__ shards, int world, int b_local, int E, emb16* __restrict__ out,
                        long long ld) {
  __shared__ emb16 tile[32][34];
  const int owner = blockIdx.z;
  const int r0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const emb16* src = shards[owner];
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i;
    tile[i][tx] = (r < b_local && k0 + tx < E) ? src[static_cast<long long>(r) * E + k0 + tx] : __float2half_rn(0.f);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int k = k0 + i, r = r0 + tx;
    if (k < E && r < b_local) out[static_cast<long long>(k) * ld + static_cast<long long>(owner) * b_local + r] = tile[tx][i];
  }
}

// grad = coef * raw: d loss / d (normalised embedding), see launch_contrastive_loss_backward
__global__ void __launch_bounds__(256)
loss_grad_finish_kernel(const float* __restrict__ raw, float coef, float* __restrict__ grad, long long total) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) grad[i] = coef * raw[i];
}

// deterministic reduction: out[dir] = sum_row row_loss[dir][row]
__global__ void __launch_bounds__(1024)
loss_reduce_kernel(const float* __restrict__ row_loss, int b_local, float* __restrict__ out) {
  __shared__ double sh[1024];
  const int dir = blockIdx.x;
  double acc = 0.0;
  for (int i = threadIdx.x; i < b_local; i += 1024) acc += static_cast<double>(row_loss[dir * b_local + i]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[dir] = static_cast<float>(sh[0]);
}

inline int pad128(int x) { return (x + 127) / 128 * 128; }

}  // namespace

size_t contrastive_loss_workspace_bytes(int world, int b_local) {
  const size_t tiles = static_cast<size_t>(world) * ((b_local + kLossBN - 1) / kLossBN);
  return 2 * tiles * pad128(b_local) * sizeof(float2) + 2 * static_cast<size_t>(pad128(b_local)) * sizeof(float) +
         2 * static_cast<size_t>(b_local) * sizeof(float);
}

static int configure_loss_kernels() {
  static bool configured = false;
  if (!configured) {
    MSCLIP_CHECK_CUDA(cudaFuncSetAttribute(contrastive_lse_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLossSmem));
    MSCLIP_CHECK_CUDA(cudaFuncSetAttribute(contrastive_lse_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLossSmem));
    configured = true;
  }
  return 0;
}

int launch_contrastive_loss_ex(const emb16* img_local, const emb16* txt_local, const emb16* const* img_shards,
                               const emb16* const* txt_shards, const uint32_t* flags, uint32_t epoch, int world,
                               int rank, int b_local, int E, float scale, void* workspace, float* loss_parts,
                               float* lse_out, int lse_pitch, cudaStream_t stream) {
  MSCLIP_REQUIRE(E == kLossE, "contrastive loss: embedding width must be 512");
  MSCLIP_REQUIRE(world >= 1 && b_local >= 1, "contrastive loss: empty problem");
  MSCLIP_TRY(configure_loss_kernels());
  LossParams p = {};
  p.col_shards[0] = txt_shards;  // image rows see text columns
  p.col_shards[1] = img_shards;  // text rows see image columns
  p.flags = world > 1 ? flags : nullptr;
  p.epoch = epoch;
  p.world = world;
  p.rank = rank;
  p.b_local = b_local;
  p.tiles_per_shard = (b_local + kLossBN - 1) / kLossBN;
  p.n_col_tiles = world * p.tiles_per_shard;
  p.n_row_blocks = (b_local + 127) / 128;
  p.b_pad = pad128(b_local);
  p.scale_log2 = scale * kLog2e;
  p.ws = reinterpret_cast<float2*>(workspace);
  p.diag = reinterpret_cast<float*>(p.ws + 2ll * p.n_col_tiles * p.b_pad);
  float* row_loss = p.diag + 2 * p.b_pad;
  CUtensorMap ti, tt;
  MSCLIP_TRY(make_tmap_op16_2d(&ti, img_local, b_local, kLossE, kLossE, 128));
  MSCLIP_TRY(make_tmap_op16_2d(&tt, txt_local, b_local, kLossE, kLossE, 128));
  contrastive_lse_kernel<0><<<dim3(p.n_col_tiles, 2), kLossThreads, kLossSmem, stream>>>(ti, tt, p);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  lse_combine_kernel<<<(2 * b_local + 255) / 256, 256, 0, stream>>>(p.ws, p.diag, b_local, p.b_pad, p.n_col_tiles,
                                                                    row_loss, lse_out, lse_pitch);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  loss_reduce_kernel<<<2, 1024, 0, stream>>>(row_loss, b_local, loss_parts);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---- backward of the contrastive loss with respect to this rank's (normalised) embeddings --------------------------------
// loss = (1 / 2G) sum_i [lse_j s_ij - s_ii] + (1 / 2G) sum_j [lse_i s_ij - s_jj], s = scale * I . T^T.  With
// w_ij = softmax_j(s_i.)_j + softmax_i(s_.j)_i - 2 [i == j]:   dL/dI_i = scale / 2G * sum_j w_ij T_j   (i local, j global)
//                                                              dL/dT_j = scale / 2G * sum_i w_ij I_i   (j local, i global)
// - the gradient reaches only the local shard, exactly as with gather_tensors (lib/utils/comm.py:151-152).  The column
// softmax needs the row lse of the columns' owners: a second peer read (col_lse tables).  Three steps: (1) the forward
// kernel re-run in MODE 1 writes w for (local rows) x (all columns) in both directions, (2) the gathered embeddings are
// transposed into K-major operands, (3) two tcgen05 GEMMs over K = G and a scaling pass.
// Workspace: contrastive_backward_workspace_bytes(world, b_local).
static long long padded_g(int world, int b_local) { return (static_cast<long long>(world) * b_local + 63) / 64 * 64; }

size_t contrastive_backward_workspace_bytes(int world, int b_local) {
  const long long ldg = padded_g(world, b_local);
  return static_cast<size_t>(2 * (static_cast<long long>(b_local) + kLossE) * ldg * 2 + 2ll * b_local * kLossE * 4 + 256);
}

int launch_contrastive_loss_backward(const emb16* img_local, const emb16* txt_local, const emb16* const* img_shards,
                                     const emb16* const* txt_shards, const float* row_lse, int lse_pitch,
                                     const float* const* img_lse_shards, const float* const* txt_lse_shards, int world,
                                     int rank, int b_local, float scale, void* workspace, float* d_img, float* d_txt,
                                     cudaStream_t stream) {
  MSCLIP_REQUIRE(world >= 1 && b_local >= 1 && d_img && d_txt, "contrastive loss backward: bad arguments");
  MSCLIP_REQUIRE(lse_pitch >= b_local, "contrastive loss backward: lse pitch smaller than the local batch");
  MSCLIP_TRY(configure_loss_kernels());
  const long long G = static_cast<long long>(world) * b_local, ldg = padded_g(world, b_local);
  uint8_t* w8 = static_cast<uint8_t*>(workspace);
  // 16-bit operands of the two gradient GEMMs are IEEE fp16 in BOTH builds: the embeddings are used exactly as exchanged, and
  // w (|w| <= 2, typical magnitude 1 / G) is stored times 2^k, k = floor(log2 G) - 1, so that it keeps fp16's 11 significant
  // bits (8 x less rounding error than bf16) instead of sinking into the subnormals; 2^-k is folded into the final coefficient.
  emb16* wprob[2] = {reinterpret_cast<emb16*>(w8), reinterpret_cast<emb16*>(w8) + static_cast<long long>(b_local) * ldg};
  emb16* xt[2] = {wprob[1] + static_cast<long long>(b_local) * ldg, wprob[1] + static_cast<long long>(b_local) * ldg + kLossE * ldg};
  int wexp = 0;
  while ((2ll << (wexp + 1)) <= G) ++wexp;   // 2 * 2^wexp <= G < 4 * 2^wexp  ->  |w| * 2^wexp <= G <= fp16 max for G <= 65504
  if (wexp > 14) wexp = 14;
  const float wscale = static_cast<float>(1 << wexp);
  float* raw = reinterpret_cast<float*>(xt[1] + kLossE * ldg);
  // zero padding columns (K of the GEMMs is ldg): clear the operands once per call
  MSCLIP_CHECK_CUDA(cudaMemsetAsync(w8, 0, static_cast<size_t>(2 * (static_cast<long long>(b_local) + kLossE) * ldg * 2), stream));
  LossParams p = {};
  p.col_shards[0] = txt_shards;
  p.col_shards[1] = img_shards;
  p.world = world;
  p.rank = rank;
  p.b_local = b_local;
  p.tiles_per_shard = (b_local + kLossBN - 1) / kLossBN;
  p.n_col_tiles = world * p.tiles_per_shard;
  p.n_row_blocks = (b_local + 127) / 128;
  p.b_pad = pad128(b_local);
  p.scale_log2 = scale * kLog2e;
  p.row_lse = row_lse;
  p.lse_pitch = lse_pitch;
  p.col_lse[0] = txt_lse_shards;  // image rows x text columns: the columns' lse is the text rows' lse of their owner
  p.col_lse[1] = img_lse_shards;
  p.wout[0] = wprob[0];
  p.wout[1] = wprob[1];
  p.ldw = ldg;
  p.wscale = wscale;
  CUtensorMap ti, tt;
  MSCLIP_TRY(make_tmap_op16_2d(&ti, img_local, b_local, kLossE, kLossE, 128));
  MSCLIP_TRY(make_tmap_op16_2d(&tt, txt_local, b_local, kLossE, kLossE, 128));
  contrastive_lse_kernel<1><<<dim3(p.n_col_tiles, 2), kLossThreads, kLossSmem, stream>>>(ti, tt, p);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  const dim3 tg((b_local + 31) / 32, kLossE / 32, world);
  gather_transpose_kernel<<<tg, 256, 0, stream>>>(txt_shards, world, b_local, kLossE, xt[0], ldg);
  gather_transpose_kernel<<<tg, 256, 0, stream>>>(img_shards, world, b_local, kLossE, xt[1], ldg);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  // raw_img = w[0] . T_all, raw_txt = w[1] . I_all   (K = padded global batch)
  MSCLIP_TRY(launch_gemm_f16(wprob[0], ldg, xt[0], ldg, b_local, kLossE, static_cast<int>(ldg), 1.0f, raw, kLossE, EPI_F32, stream));
  MSCLIP_TRY(launch_gemm_f16(wprob[1], ldg, xt[1], ldg, b_local, kLossE, static_cast<int>(ldg), 1.0f, raw + static_cast<long long>(b_local) * kLossE,
                             kLossE, EPI_F32, stream));
  const float coef = scale / (2.0f * static_cast<float>(G) * wscale);
  const long long total = static_cast<long long>(b_local) * kLossE;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 8));
  loss_grad_finish_kernel<<<grid, 256, 0, stream>>>(raw, coef, d_img, total);
  loss_grad_finish_kernel<<<grid, 256, 0, stream>>>(raw + total, coef, d_txt, total);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace msclip
