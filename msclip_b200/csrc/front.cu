// Fused front end of the image tower: everything that touches the 112 x 112 feature maps in one pass.
//
//   stem  = relu(bn1(conv3x3_s2(img)))                      EarlyconvRes.conv1/bn1            M.py:1993
//   p0    = relu(bn(conv3x3_s2(img)))                       parallel_branch_v.0               M.py:2260-2273
//   y1    = relu(bn1(conv1x1(p0)))                          parallel_branch_v.1 ConvResBlock  M.py:1842-1846
//   p0s   = p0[:, ::2, ::2]                                 input of its strided 1x1 shortcut M.py:1857
//   pool0 = bn(dwconv_{k x k, stride k}(p0))                Lateral_Adapter.0 top2bottom_dw   M.py:1756
//
// Unfused this is im2col + GEMM (N = 96) + GEMM (48 -> 48) + patch pooling: the 2.4 MB / image of 112 x 112
// activations are written and re-read three times.  Here a CTA owns a 16 x 16 tile of output pixels: the 33 x 36
// input window of the three colour planes is staged once in shared memory (cp.async, zero fill for the padding),
// each warp builds the im2col fragments of its two pixel rows straight from that window, runs both first
// convolutions (K = 27 -> 32, N = 48 + 48) and the 1x1 bottleneck entry (K = 48, N = 48, fed from the accumulator
// registers of p0) on the tensor cores with mma.sync m16n8k16 - the work per byte is far too small for a 128-row
// tcgen05 tile pipeline to pay off; measured, the kernel moves exactly its algorithmic bytes at ~3.4 TB/s and is
// bound by shared-memory wavefronts (weight fragments, output staging), profiles/r01_conv_front_ncu.md - and
// accumulates the depth-wise patch pooling of p0 in registers.  p0 itself never reaches HBM (only its even pixels do).  Outputs are staged per warp
// in shared memory and leave as full 16-byte vectors of contiguous NHWC rows.
#include "common.cuh"
#include "kernels.h"

namespace msclip {

namespace {

constexpr int kT = 16;               // output tile side
constexpr int kFrontThreads = 256;   // 8 warps; warp w owns tile rows 2w and 2w+1
constexpr int kInRows = 2 * kT + 1;  // staged input rows: iy = 2*oy0 - 1 .. 2*oy0 + 31
constexpr int kInPitch = 36;         // staged input columns: ix = 2*ox0 - 4 .. 2*ox0 + 31 (16-byte aligned start)
constexpr int kC = 48;               // channels of each first conv and of the bottleneck entry (width / 16)
constexpr int kW0Pitch = 40;         // op16 per staged row of w0 [96][32]  (80 B: conflict-free ldmatrix)
constexpr int kW1Pitch = 56;         // op16 per staged row of w1 [48][48]  (112 B)
constexpr int kStagePitch = 48;      // op16 per staged output pixel (dense; 16-byte chunk c of pixel p sits at c ^ ((p >> 2) & 1))
constexpr int kPoolPitch = 56;       // f32 per staged pooling tap [k*k][48] (224 B: conflict-free 8-byte loads)

constexpr int kOffIn = 0;
constexpr int kOffW0 = kOffIn + 3 * kInRows * kInPitch * 4;    // 14256
constexpr int kOffW1 = kOffW0 + 2 * kC * kW0Pitch * 2;         // + 7680
constexpr int kOffBias = kOffW1 + kC * kW1Pitch * 2;           // + 5376
constexpr int kOffStage = kOffBias + 4 * kC * 4;               // b0[96] | b1[48] | pool_b[48]
constexpr int kOffPart = kOffStage + 8 * kT * kStagePitch * 2;  // + 12288
constexpr int kOffPoolW = kOffPart + 8 * 2 * kC * 4;           // + 3072
static_assert(kOffW0 % 16 == 0 && kOffW1 % 16 == 0 && kOffBias % 16 == 0 && kOffStage % 16 == 0 && kOffPart % 16 == 0 &&
                  kOffPoolW % 16 == 0,
              "shared-memory regions keep 16-byte alignment");

struct FrontParams {
  const void* img;
  int img_dtype;  // 0 f32, 1 bf16, 2 f16
  int H, W;       // input height / width
  int Ho, Wo, tiles_x, tiles_per_img, total_tiles;
  int k;          // lateral kernel = stride of the patch pooling (8 or 16)
  const op16* w0;
  const float* b0;
  const op16* w1;
  const float* b1;
  const float* pool_w;
  const float* pool_b;
  op16* stem;
  op16* y1;
  op16* p0s;
  int p0s_pitch;  // elements between consecutive p0s pixels (48 = dense; 96 when it is the second half of a [y2 | p0s] operand)
  op16* pooled;
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32." MSCLIP_MMA_OPERANDS ".f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float px_to_float(float v, int) { return v; }
__device__ __forceinline__ float px_to_float(uint16_t v, int dtype) {
  return dtype == 1 ? __uint_as_float(static_cast<uint32_t>(v) << 16) : __half2float(__ushort_as_half(v));
}

// relu(acc) of one 16-pixel row (6 n-tiles = 48 channels; the bias is already in the accumulator) -> staged ->
// 1536 contiguous bytes at gdst
__device__ __forceinline__ void store_row(const float (&acc)[6][4], op16* stage, op16* gdst, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const int sw = (g >> 2) & 1;  // same for pixels g and g + 8
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    // dense 96-byte pixels would put pixels g and g + 4 on the same banks; swapping the chunk pairs of every
    // second group of four pixels keeps both the 4-byte stores and the 16-byte read-back conflict-free
    *reinterpret_cast<uint32_t*>(stage + g * kStagePitch + 8 * (j ^ sw) + 2 * t) =
        pack16(fmaxf(acc[j][0], 0.f), fmaxf(acc[j][1], 0.f));
    *reinterpret_cast<uint32_t*>(stage + (g + 8) * kStagePitch + 8 * (j ^ sw) + 2 * t) =
        pack16(fmaxf(acc[j][2], 0.f), fmaxf(acc[j][3], 0.f));
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int idx = lane + 32 * i;  // 16 pixels x 6 vectors of 8 channels, physical order
    const int px = idx / 6, pc = idx - px * 6;
    *reinterpret_cast<uint4*>(gdst + (px * 6 + (pc ^ ((px >> 2) & 1))) * 8) = *reinterpret_cast<const uint4*>(stage + idx * 8);
  }
  __syncwarp();
}

template <typename IMG>
__global__ void __launch_bounds__(kFrontThreads, 2) front_conv_kernel(const FrontParams p) {
  extern __shared__ __align__(16) uint8_t fsm[];
  IMG* in_s = reinterpret_cast<IMG*>(fsm + kOffIn);
  op16* w0_s = reinterpret_cast<op16*>(fsm + kOffW0);
  op16* w1_s = reinterpret_cast<op16*>(fsm + kOffW1);
  float* bias_s = reinterpret_cast<float*>(fsm + kOffBias);
  op16* stage_s = reinterpret_cast<op16*>(fsm + kOffStage);
  float* part_s = reinterpret_cast<float*>(fsm + kOffPart);
  float* poolw_s = reinterpret_cast<float*>(fsm + kOffPoolW);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int mi = lane >> 3, r8 = lane & 7;
  const int k = p.k;

  // stage the next tile's input window: 3 planes x 33 rows x 9 groups of 4 pixels
  auto issue_load = [&](int tile) {
    const int b = tile / p.tiles_per_img;
    const int rem = tile - b * p.tiles_per_img;
    const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
    const int iy0 = 2 * kT * ty - 1, ix0 = 2 * kT * tx - 4;
    for (int i = tid; i < 3 * kInRows * 9; i += kFrontThreads) {
      const int c = i / (kInRows * 9);
      const int r2 = i - c * (kInRows * 9);
      const int row = r2 / 9, grp = r2 - row * 9;
      const int iy = iy0 + row, ix = ix0 + 4 * grp;
      const bool ok = iy >= 0 && ix >= 0;
      const long long e = ((static_cast<long long>(b) * 3 + c) * p.H + iy) * p.W + ix;
      const IMG* src = reinterpret_cast<const IMG*>(p.img) + (ok ? e : 0);
      const uint32_t dst = smem_u32(in_s + (c * kInRows + row) * kInPitch + grp * 4);
      if constexpr (sizeof(IMG) == 4)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16u : 0u) : "memory");
      else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(ok ? 8u : 0u) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int tile = blockIdx.x;
  if (tile < p.total_tiles) issue_load(tile);

  // weights, biases and pooling taps: once per CTA
  for (int i = tid; i < 2 * kC * 4; i += kFrontThreads) {
    const int row = i >> 2, pc = i & 3;
    *reinterpret_cast<uint4*>(w0_s + row * kW0Pitch + pc * 8) = reinterpret_cast<const uint4*>(p.w0)[i];
  }
  for (int i = tid; i < kC * 6; i += kFrontThreads) {
    const int row = i / 6, pc = i - row * 6;
    *reinterpret_cast<uint4*>(w1_s + row * kW1Pitch + pc * 8) = reinterpret_cast<const uint4*>(p.w1)[i];
  }
  __syncthreads();  // w0 rows are in place before their bias columns are patched
  // the first convs' biases ride in the two spare K columns of w0 (A supplies 1.0 there): hi + lo split keeps them
  // exact to 2^-17; columns 29..31 stay zero
  for (int i = tid; i < 2 * kC; i += kFrontThreads) {
    const float bv = p.b0[i];
    const op16 hi = to_op16(bv);
    w0_s[i * kW0Pitch + 27] = hi;
    w0_s[i * kW0Pitch + 28] = to_op16(bv - op16_to_float(hi));
  }
  for (int i = tid; i < 4 * kC; i += kFrontThreads)
    bias_s[i] = i < 2 * kC ? p.b0[i] : (i < 3 * kC ? p.b1[i - 2 * kC] : p.pool_b[i - 3 * kC]);
  for (int i = tid; i < k * k * kC / 4; i += kFrontThreads) {
    const int tapi = i / (kC / 4), c4 = i - tapi * (kC / 4);
    *reinterpret_cast<float4*>(poolw_s + tapi * kPoolPitch + c4 * 4) = reinterpret_cast<const float4*>(p.pool_w)[i];
  }

  // window offsets of the 8 im2col columns this thread feeds: k = 16*kt + 8*hf + 2*t + e, k = c*9 + ky*3 + kx;
  // columns 27..31 meet zero weights, any finite value will do
  int koff[2][2][2];
#pragma unroll
  for (int kt = 0; kt < 2; ++kt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        int kk = 16 * kt + 8 * hf + 2 * t + e;
        if (kk > 26) kk = 26;
        const int c = kk / 9, r9 = kk - 9 * c;
        const int ky = r9 / 3, kx = r9 - 3 * ky;
        koff[kt][hf][e] = (c * kInRows + ky) * kInPitch + kx + 3;
      }

  op16* stage = stage_s + warp * (kT * kStagePitch);
  const int gy = p.Ho / k, gx = p.Wo / k;  // pooled grid

  for (; tile < p.total_tiles; tile += gridDim.x) {
    const int b = tile / p.tiles_per_img;
    const int rem = tile - b * p.tiles_per_img;
    const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
    const int oy0 = kT * ty, ox0 = kT * tx;

    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // window (and, the first time, the weights) visible to every warp

    // ---- im2col fragments of this warp's two pixel rows
    uint32_t a[2][2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const IMG* bp = in_s + (4 * warp + 2 * r) * kInPitch + 2 * (g + 8 * rh);
#pragma unroll
        for (int kt = 0; kt < 2; ++kt)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            float lo = px_to_float(bp[koff[kt][hf][0]], p.img_dtype), hi = px_to_float(bp[koff[kt][hf][1]], p.img_dtype);
            if (kt == 1 && hf == 1) {  // k = 24 + 2t + e: columns 27 (t = 1, e = 1) and 28 (t = 2, e = 0) carry the bias
              if (t == 1) hi = 1.0f;
              if (t == 2) lo = 1.0f;
            }
            a[r][kt][hf * 2 + rh] = pack16(lo, hi);
          }
      }
    __syncthreads();  // the window is consumed: the next tile's copy may overwrite it while this tile computes
    if (tile + static_cast<int>(gridDim.x) < p.total_tiles) issue_load(tile + gridDim.x);

    const long long pix0 = (static_cast<long long>(b) * p.Ho + oy0 + 2 * warp) * p.Wo + ox0;  // first pixel of row r = 0

    float acc[2][6][4];
    // ---- stem half of the first conv (output channels 0..47)
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int j = 0; j < 6; ++j) acc[r][j][0] = acc[r][j][1] = acc[r][j][2] = acc[r][j][3] = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      uint32_t bf[4];
      ldsm_x4(bf, smem_u32(w0_s + (8 * j + r8) * kW0Pitch + mi * 8));
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mma16816(acc[r][j], a[r][0], bf[0], bf[1]);
        mma16816(acc[r][j], a[r][1], bf[2], bf[3]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) store_row(acc[r], stage, p.stem + (pix0 + static_cast<long long>(r) * p.Wo) * kC, lane);

    // ---- branch half (output channels 48..95): p0 stays in registers
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int j = 0; j < 6; ++j) acc[r][j][0] = acc[r][j][1] = acc[r][j][2] = acc[r][j][3] = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      uint32_t bf[4];
      ldsm_x4(bf, smem_u32(w0_s + (kC + 8 * j + r8) * kW0Pitch + mi * 8));
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mma16816(acc[r][j], a[r][0], bf[0], bf[1]);
        mma16816(acc[r][j], a[r][1], bf[2], bf[3]);
      }
    }
    float pa[2][12];  // patch-pool partial sums: pixel columns g (0) / g + 8 (1), channels 8j + 2t + e
#pragma unroll
    for (int i = 0; i < 12; ++i) pa[0][i] = pa[1][i] = 0.f;
    uint32_t a2[2][3][4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        acc[r][j][0] = fmaxf(acc[r][j][0], 0.f);
        acc[r][j][1] = fmaxf(acc[r][j][1], 0.f);
        acc[r][j][2] = fmaxf(acc[r][j][2], 0.f);
        acc[r][j][3] = fmaxf(acc[r][j][3], 0.f);
      }
      const int ky = (2 * warp + r) & (k - 1);
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const float* wp = poolw_s + (ky * k + ((g + 8 * rh) & (k - 1))) * kPoolPitch + 2 * t;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const float2 wv = *reinterpret_cast<const float2*>(wp + 8 * j);
          pa[rh][2 * j] = fmaf(acc[r][j][2 * rh], wv.x, pa[rh][2 * j]);
          pa[rh][2 * j + 1] = fmaf(acc[r][j][2 * rh + 1], wv.y, pa[rh][2 * j + 1]);
        }
      }
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) {
        a2[r][kk][0] = pack16(acc[r][2 * kk][0], acc[r][2 * kk][1]);
        a2[r][kk][1] = pack16(acc[r][2 * kk][2], acc[r][2 * kk][3]);
        a2[r][kk][2] = pack16(acc[r][2 * kk + 1][0], acc[r][2 * kk + 1][1]);
        a2[r][kk][3] = pack16(acc[r][2 * kk + 1][2], acc[r][2 * kk + 1][3]);
      }
    }
    // even pixels of the even row (r = 0) feed the strided shortcut: 8 pixels x 48 channels, contiguous in p0s
    if ((g & 1) == 0) {
      // staged pixels g/2 (chunks as they are) and g/2 + 4 (chunk pairs swapped, see store_row)
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) {
        *reinterpret_cast<uint32_t*>(stage + (g >> 1) * kStagePitch + 16 * kk + 2 * t) = a2[0][kk][0];
        *reinterpret_cast<uint32_t*>(stage + ((g >> 1) + 4) * kStagePitch + 16 * kk + 8 + 2 * t) = a2[0][kk][1];
        *reinterpret_cast<uint32_t*>(stage + (g >> 1) * kStagePitch + 16 * kk + 8 + 2 * t) = a2[0][kk][2];
        *reinterpret_cast<uint32_t*>(stage + ((g >> 1) + 4) * kStagePitch + 16 * kk + 2 * t) = a2[0][kk][3];
      }
    }
    __syncwarp();
    {
      op16* gdst = p.p0s + ((static_cast<long long>(b) * (p.Ho / 2) + (oy0 / 2 + warp)) * (p.Wo / 2) + ox0 / 2) * p.p0s_pitch;
      for (int idx = lane; idx < 48; idx += 32) {
        const int px = idx / 6, pc = idx - px * 6;
        *reinterpret_cast<uint4*>(gdst + px * p.p0s_pitch + (pc ^ ((px >> 2) & 1)) * 8) =
            *reinterpret_cast<const uint4*>(stage + idx * 8);
      }
    }
    __syncwarp();

    // ---- y1 = relu(bn1(conv1x1(p0))): K = 48 from the registers above, accumulators start at the bias
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const float2 bb = *reinterpret_cast<const float2*>(bias_s + 2 * kC + 8 * j + 2 * t);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        acc[r][j][0] = acc[r][j][2] = bb.x;
        acc[r][j][1] = acc[r][j][3] = bb.y;
      }
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      uint32_t bf[4], bg[4];
      ldsm_x4(bf, smem_u32(w1_s + (8 * j + r8) * kW1Pitch + mi * 8));             // k 0..31
      ldsm_x4(bg, smem_u32(w1_s + (8 * j + r8) * kW1Pitch + 32 + (mi & 1) * 8));  // k 32..47 (matrices 2, 3 repeat 0, 1)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mma16816(acc[r][j], a2[r][0], bf[0], bf[1]);
        mma16816(acc[r][j], a2[r][1], bf[2], bf[3]);
        mma16816(acc[r][j], a2[r][2], bg[0], bg[1]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
      store_row(acc[r], stage, p.y1 + (pix0 + static_cast<long long>(r) * p.Wo) * kC, lane);

    // ---- patch pooling: reduce over the 8 pixel columns held by the lanes of equal t, then over warps (fixed order)
    if (k == 16) {
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        pa[0][i] += pa[1][i];
        pa[1][i] = 0.f;
      }
    }
#pragma unroll
    for (int rh = 0; rh < 2; ++rh)
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        float v = pa[rh][i];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        pa[rh][i] = v;
      }
    if (g == 0) {
#pragma unroll
      for (int rh = 0; rh < 2; ++rh)
#pragma unroll
        for (int j = 0; j < 6; ++j)
          *reinterpret_cast<float2*>(part_s + (warp * 2 + rh) * kC + 8 * j + 2 * t) = make_float2(pa[rh][2 * j], pa[rh][2 * j + 1]);
    }
    __syncthreads();
    if (k == 16) {
      if (tid < kC) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += part_s[(w * 2) * kC + tid];
        p.pooled[((static_cast<long long>(b) * gy + ty) * gx + tx) * kC + tid] = to_op16(s + bias_s[3 * kC + tid]);
      }
    } else {  // k == 8: the tile holds 2 x 2 cells; warps 0-3 cover the upper cells, 4-7 the lower ones
      if (tid < 4 * kC) {
        const int cell = tid / kC, c = tid - cell * kC;
        const int py = cell >> 1, px = cell & 1;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) s += part_s[((4 * py + w) * 2 + px) * kC + c];
        p.pooled[((static_cast<long long>(b) * gy + 2 * ty + py) * gx + 2 * tx + px) * kC + c] =
            to_op16(s + bias_s[3 * kC + c]);
      }
    }
    // part_s is rewritten only after the next iteration's two __syncthreads
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace

bool front_conv_supported(int H, int W, int c0, int k) {
  return H > 0 && W > 0 && H % (2 * kT) == 0 && W % (2 * kT) == 0 && c0 == kC && (k == 8 || k == 16);
}

int launch_front_conv(const void* img, int img_dtype, int batch, int H, int W, const op16* w0, const float* b0,
                      const op16* w1, const float* b1, const float* pool_w, const float* pool_b, int k, op16* stem,
                      op16* y1, op16* p0s, int p0s_pitch, op16* pooled, cudaStream_t stream) {
  if (batch <= 0) return 0;
  MSCLIP_REQUIRE(img_dtype >= 0 && img_dtype <= 2, "front_conv: image dtype must be 0 (f32), 1 (bf16) or 2 (f16)");
  MSCLIP_REQUIRE(front_conv_supported(H, W, kC, k), "front_conv: needs H, W multiples of 32 and a lateral kernel of 8 or 16");
  const uintptr_t al = reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(w0) | reinterpret_cast<uintptr_t>(w1) |
                       reinterpret_cast<uintptr_t>(pool_w) | reinterpret_cast<uintptr_t>(stem) |
                       reinterpret_cast<uintptr_t>(y1) | reinterpret_cast<uintptr_t>(p0s);
  MSCLIP_REQUIRE((al & 15) == 0, "front_conv: image, weights and outputs must be 16-byte aligned");
  MSCLIP_REQUIRE(p0s_pitch >= kC && p0s_pitch % 8 == 0, "front_conv: p0s pixel pitch must be a multiple of 8 and at least 48");
  FrontParams p;
  p.img = img;
  p.img_dtype = img_dtype;
  p.H = H;
  p.W = W;
  p.Ho = H / 2;
  p.Wo = W / 2;
  p.tiles_x = p.Wo / kT;
  p.tiles_per_img = (p.Ho / kT) * p.tiles_x;
  const long long total = static_cast<long long>(batch) * p.tiles_per_img;
  MSCLIP_REQUIRE(total < (1ll << 31), "front_conv: too many tiles for one launch");
  p.total_tiles = static_cast<int>(total);
  p.k = k;
  p.w0 = w0;
  p.b0 = b0;
  p.w1 = w1;
  p.b1 = b1;
  p.pool_w = pool_w;
  p.pool_b = pool_b;
  p.stem = stem;
  p.y1 = y1;
  p.p0s = p0s;
  p.p0s_pitch = p0s_pitch;
  p.pooled = pooled;
  const int smem = kOffPoolW + k * k * kPoolPitch * 4;
  const int grid = p.total_tiles < 2 * num_sms() ? p.total_tiles : 2 * num_sms();
  if (img_dtype == 0) {
    static int configured = 0;
    if (configured < smem) {
      MSCLIP_CHECK_CUDA(cudaFuncSetAttribute(front_conv_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      configured = smem;
    }
    front_conv_kernel<float><<<grid, kFrontThreads, smem, stream>>>(p);
  } else {
    static int configured = 0;
    if (configured < smem) {
      MSCLIP_CHECK_CUDA(cudaFuncSetAttribute(front_conv_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      configured = smem;
    }
    front_conv_kernel<uint16_t><<<grid, kFrontThreads, smem, stream>>>(p);
  }
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace msclip
