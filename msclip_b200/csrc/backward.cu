// HBM-bound kernels of the BACKWARD pass of the shared ResidualAttentionBlock stack and its heads (SURVEY.md section
// 8f-1; the reference ships no backward - the oracle is torch.autograd on the reference module, M.py:1027-1028, 204-224):
//   * LayerNorm backward (M.py:204-219) fused with the residual-gradient accumulation, the 16-bit copy the next dgrad /
//     wgrad GEMMs consume and the column sums that are the bias gradient of the linear layer in front of it
//   * QuickGELU backward (M.py:222-224) fused with the fc1 bias gradient
//   * column sums of 16-bit gradient matrices (bias gradients), deterministic two-stage reduction
//   * pooled-row LayerNorm backward (ln_final at the EOT row M.py:3057-3072, ln_post at the CLS row M.py:2685-2690),
//     L2-normalisation backward (M.py:2982-2983, 3076-3077), embedding backward (M.py:3047-3048, 2418-2426)
//   * fused multi-tensor AdamW (decoupled weight decay; optimiser of experiments/model/b32.yaml:32-53)
// One warp owns one 768-wide row, as in the forward row kernels (rowops.cuh).  Every cross-row reduction is done in a
// fixed order (registers -> shared memory by warp turns -> per-CTA partial rows -> reduce_partials), so gradients are
// bit-reproducible run to run - except the token-embedding scatter, which uses fp32 atomics.
#include "common.cuh"
#include "kernels.h"
#include "rowops.cuh"

namespace msclip {

namespace {

using namespace rowops;
constexpr int kWarps = 8;
constexpr int kCols = kVec * 4;  // columns per lane (24)

// Column c of a lane's value j: float4 index (lane + 32 * (j / 4)), component j % 4
__device__ __forceinline__ int lane_col(int lane, int j) { return 4 * (lane + 32 * (j >> 2)) + (j & 3); }

// acc[NS][24] of every warp -> part[NS * 768] of this CTA, warps taking turns (fixed summation order)
template <int NS>
__device__ __forceinline__ void cta_reduce_columns(const float (&acc)[NS][kCols], float* smem, float* __restrict__ part) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int w = 0; w < kWarps; ++w) {
    if (warp == w) {
#pragma unroll
      for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int j = 0; j < kCols; ++j) {
          const int c = s * kD + lane_col(lane, j);
          smem[c] = (w == 0) ? acc[s][j] : smem[c] + acc[s][j];
        }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < NS * kD; i += blockDim.x) part[i] = smem[i];
}

__device__ __forceinline__ void unpack(const float4 (&v)[kVec], float (&f)[kCols]) {
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    f[4 * i + 0] = v[i].x;
    f[4 * i + 1] = v[i].y;
    f[4 * i + 2] = v[i].z;
    f[4 * i + 3] = v[i].w;
  }
}

// x -> (xhat, rstd) in registers: the forward's two-pass statistics (rowops::layer_norm_row)
__device__ __forceinline__ float normalise_row(float (&x)[kCols]) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kCols; ++j) s += x[j];
  const float mean = warp_sum(s) * (1.0f / kD);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < kCols; ++j) {
    x[j] -= mean;
    q = fmaf(x[j], x[j], q);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / kD) + kLnEps);
#pragma unroll
  for (int j = 0; j < kCols; ++j) x[j] *= rstd;
  return rstd;
}

// dy (gradient of gamma * xhat + beta) -> gradient of the row; accumulates dgamma / dbeta
__device__ __forceinline__ void ln_backward_row(const float (&xhat)[kCols], float rstd, float (&dy)[kCols], const float (&g)[kCols],
                                                float (&dgamma)[kCols], float (&dbeta)[kCols]) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < kCols; ++j) {
    dgamma[j] = fmaf(dy[j], xhat[j], dgamma[j]);
    dbeta[j] += dy[j];
    dy[j] *= g[j];  // d xhat
    s1 += dy[j];
    s2 = fmaf(dy[j], xhat[j], s2);
  }
  s1 = warp_sum(s1) * (1.0f / kD);
  s2 = warp_sum(s2) * (1.0f / kD);
#pragma unroll
  for (int j = 0; j < kCols; ++j) dy[j] = rstd * (dy[j] - s1 - xhat[j] * s2);
}

__device__ __forceinline__ void load_gamma(const float* __restrict__ gamma, int lane, float (&g)[kCols]) {
  float4 v[kVec];
  load_row(gamma, lane, v);
  unpack(v, g);
}

// dx[r] (+)= LN_backward(dy[r]; x[r], gamma).  g16 (optional) = op16(updated dx); part = [gridDim.x][3 * 768]:
// dgamma | dbeta | column sums of the updated dx (the bias gradient of the linear layer whose output joins the stream here).
// accumulate = 0: dx is overwritten (no residual path).
__global__ void __launch_bounds__(32 * kWarps)
ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma, float* __restrict__ dx,
              op16* __restrict__ g16, float* __restrict__ part, long long rows, int accumulate) {
  __shared__ float red[3 * kD];
  const int lane = threadIdx.x & 31;
  float g[kCols], acc[3][kCols];
  load_gamma(gamma, lane, g);
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int j = 0; j < kCols; ++j) acc[s][j] = 0.f;
  for (long long r = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5); r < rows;
       r += static_cast<long long>(gridDim.x) * kWarps) {
    float4 v[kVec];
    float xh[kCols], d[kCols];
    load_row(x + r * kD, lane, v);
    unpack(v, xh);
    load_row(dy + r * kD, lane, v);
    unpack(v, d);
    float4 o[kVec];
    if (accumulate) load_row(dx + r * kD, lane, o);
    const float rstd = normalise_row(xh);
    ln_backward_row(xh, rstd, d, g, acc[0], acc[1]);
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      if (accumulate) {
        o[i].x += d[4 * i + 0];
        o[i].y += d[4 * i + 1];
        o[i].z += d[4 * i + 2];
        o[i].w += d[4 * i + 3];
      } else {
        o[i] = make_float4(d[4 * i + 0], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3]);
      }
      acc[2][4 * i + 0] += o[i].x;
      acc[2][4 * i + 1] += o[i].y;
      acc[2][4 * i + 2] += o[i].z;
      acc[2][4 * i + 3] += o[i].w;
    }
    float4* d4 = reinterpret_cast<float4*>(dx + r * kD);
#pragma unroll
    for (int i = 0; i < kVec; ++i) d4[lane + 32 * i] = o[i];
    if (g16 != nullptr) store_row_bf16(g16 + r * kD, lane, o);
  }
  cta_reduce_columns<3>(acc, red, part + static_cast<size_t>(blockIdx.x) * 3 * kD);
}

// g16 = op16(dx); part[gridDim.x][768] = column sums of dx
__global__ void __launch_bounds__(32 * kWarps)
cast_colsum_kernel(const float* __restrict__ dx, op16* __restrict__ g16, float* __restrict__ part, long long rows) {
  __shared__ float red[kD];
  const int lane = threadIdx.x & 31;
  float acc[1][kCols];
#pragma unroll
  for (int j = 0; j < kCols; ++j) acc[0][j] = 0.f;
  for (long long r = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5); r < rows;
       r += static_cast<long long>(gridDim.x) * kWarps) {
    float4 v[kVec];
    load_row(dx + r * kD, lane, v);
    store_row_bf16(g16 + r * kD, lane, v);
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      acc[0][4 * i + 0] += v[i].x;
      acc[0][4 * i + 1] += v[i].y;
      acc[0][4 * i + 2] += v[i].z;
      acc[0][4 * i + 3] += v[i].w;
    }
  }
  cta_reduce_columns<1>(acc, red, part + static_cast<size_t>(blockIdx.x) * kD);
}

// ---- 16-bit matrices of arbitrary width (multiple of 256): CTA = 256 columns x a slab of rows; thread = 8 columns of
// every 8th row of the slab; per-CTA partial = 256 column sums (fixed order: thread registers, then the 8 row phases)
constexpr int kSlabThreads = 256;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const op162* h = reinterpret_cast<const op162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = op162_to_float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__device__ __forceinline__ void slab_reduce_store(const float (&acc)[8], float* __restrict__ part, int width) {
  __shared__ float red[8][256];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ry][cx * 8 + j] = acc[j];
  __syncthreads();
  const int c = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int y = 0; y < 8; ++y) s += red[y][c];
  part[static_cast<size_t>(blockIdx.y) * width + blockIdx.x * 256 + c] = s;
}

// QuickGELU precise forms (the backward and the recomputed activation use exp, not the forward epilogue's tanh.approx:
// these kernels are HBM-bound)
__device__ __forceinline__ float sigmoid_1702(float u) { return 1.0f / (1.0f + __expf(-1.702f * u)); }

// a = quickgelu(u)   (recomputation of the fc1 activation for the fc2 weight gradient)
__global__ void __launch_bounds__(kSlabThreads)
qgelu_fwd_kernel(const op16* __restrict__ u, op16* __restrict__ a, long long n8) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 uu = reinterpret_cast<const uint4*>(u)[i];
    float f[8];
    unpack8(uu, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = f[j] * sigmoid_1702(f[j]);
    reinterpret_cast<uint4*>(a)[i] = make_uint4(pack16(f[0], f[1]), pack16(f[2], f[3]), pack16(f[4], f[5]), pack16(f[6], f[7]));
  }
}

// du = da * quickgelu'(u) in place over da; part[gridDim.y][width] = column sums of du (fc1 bias gradient)
// quickgelu'(u) = s (1 + 1.702 u (1 - s)), s = sigmoid(1.702 u)
__global__ void __launch_bounds__(kSlabThreads)
qgelu_bwd_kernel(op16* __restrict__ da, const op16* __restrict__ u, float* __restrict__ part, long long rows, int width,
                 int rows_per_slab) {
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const long long col = static_cast<long long>(blockIdx.x) * 256 + cx * 8;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_slab;
  const long long r1 = r0 + rows_per_slab < rows ? r0 + rows_per_slab : rows;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (long long r = r0 + ry; r < r1; r += 8) {
    uint4* pd = reinterpret_cast<uint4*>(da + r * width + col);
    const uint4 dd = *pd;
    const uint4 uu = *reinterpret_cast<const uint4*>(u + r * width + col);
    float d[8], f[8];
    unpack8(dd, d);
    unpack8(uu, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = sigmoid_1702(f[j]);
      d[j] *= s * fmaf(1.702f * f[j], 1.0f - s, 1.0f);
      acc[j] += d[j];
    }
    *pd = make_uint4(pack16(d[0], d[1]), pack16(d[2], d[3]), pack16(d[4], d[5]), pack16(d[6], d[7]));
  }
  slab_reduce_store(acc, part, width);
}

__global__ void __launch_bounds__(kSlabThreads)
colsum16_kernel(const op16* __restrict__ g, float* __restrict__ part, long long rows, int width, int rows_per_slab) {
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const long long col = static_cast<long long>(blockIdx.x) * 256 + cx * 8;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_slab;
  const long long r1 = r0 + rows_per_slab < rows ? r0 + rows_per_slab : rows;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (long long r = r0 + ry; r < r1; r += 8) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(g + r * width + col), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += f[j];
  }
  slab_reduce_store(acc, part, width);
}

// dst[j] (+)= scale_j * sum_p part[p * pitch + j];  scale_j = head_scale for j < head_n, else 1 (the q rows of the packed
// QKV projection carry the folded 1/8)
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ part, int nparts, long long pitch, float* __restrict__ dst, long long n,
                       int accumulate, long long head_n, float head_scale) {
  for (long long j = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; j < n;
       j += static_cast<long long>(gridDim.x) * blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += part[p * pitch + j];
    if (j < head_n) s *= head_scale;
    dst[j] = accumulate ? dst[j] + s : s;
  }
}

// up to three destinations of width n from one partial buffer whose rows hold them side by side (dgamma | dbeta | colsum)
__global__ void __launch_bounds__(256)
reduce_partials3_kernel(const float* __restrict__ part, int nparts, long long pitch, float* __restrict__ d0, float* __restrict__ d1,
                        float* __restrict__ d2, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= 3 * n) return;
  const int slot = j / n;
  float* dst = slot == 0 ? d0 : (slot == 1 ? d1 : d2);
  if (dst == nullptr) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[p * pitch + j];
  dst[j - slot * n] += s;
}

// ---- heads ---------------------------------------------------------------------------------------------------------
// dg = (df - f (f . df)) / ||g||, f = g / ||g||   (normalise = 0: dg = df); also the 16-bit copy for the projection GEMMs
__global__ void __launch_bounds__(32 * kWarps)
l2norm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ df, float* __restrict__ dg, op16* __restrict__ dg16,
                  int rows, int E, int normalise) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* gr = g + static_cast<long long>(r) * E;
  const float* dr = df + static_cast<long long>(r) * E;
  float ss = 0.f, dot = 0.f;
  for (int c = lane; c < E; c += 32) {
    ss = fmaf(gr[c], gr[c], ss);
    dot = fmaf(gr[c], dr[c], dot);
  }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float inv = normalise ? 1.0f / sqrtf(ss) : 1.0f;
  const float k = normalise ? dot * inv * inv * inv : 0.f;  // (f . df) / ||g|| * (1 / ||g||) applied to g
  for (int c = lane; c < E; c += 32) {
    const float v = dr[c] * inv - gr[c] * k;
    if (dg) dg[static_cast<long long>(r) * E + c] = v;
    dg16[static_cast<long long>(r) * E + c] = to_op16(v);
  }
}

// argmax over the token ids of sequence r, first occurrence on ties (eot_layernorm_kernel, M.py:3059)
__device__ __forceinline__ int eot_position(const int64_t* __restrict__ tok, int L, int r, int lane) {
  long long best = INT64_MIN;
  int best_i = 0x7fffffff;
  for (int j = lane; j < L; j += 32) {
    const long long t = tok[static_cast<long long>(r) * L + j];
    if (t > best) {
      best = t;
      best_i = j;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ob > best || (ob == best && oi < best_i)) {
      best = ob;
      best_i = oi;
    }
  }
  return best_i;
}

// Pooled rows: sequence b contributes row b * Lx + pos_b (pos_b = EOT position for text, 0 = CLS for the image tower).
// dx (zero-filled beforehand) receives LN_backward(dz[b]) at that row; part[gridDim.x][2 * 768] = dgamma | dbeta.
__global__ void __launch_bounds__(32 * kWarps)
pooled_ln_bwd_kernel(const float* __restrict__ x, int Lx, const int64_t* __restrict__ tok, int Ltok, const float* __restrict__ dz,
                     const float* __restrict__ gamma, float* __restrict__ dx, float* __restrict__ part, int batch) {
  __shared__ float red[2 * kD];
  const int lane = threadIdx.x & 31;
  float g[kCols], acc[2][kCols];
  load_gamma(gamma, lane, g);
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int j = 0; j < kCols; ++j) acc[s][j] = 0.f;
  for (int b = blockIdx.x * kWarps + (threadIdx.x >> 5); b < batch; b += gridDim.x * kWarps) {
    int pos = 0;
    if (tok != nullptr) {
      pos = eot_position(tok, Ltok, b, lane);
      if (pos >= Lx) pos = Lx - 1;
    }
    const long long row = static_cast<long long>(b) * Lx + pos;
    float4 v[kVec];
    float xh[kCols], d[kCols];
    load_row(x + row * kD, lane, v);
    unpack(v, xh);
    load_row(dz + static_cast<long long>(b) * kD, lane, v);
    unpack(v, d);
    const float rstd = normalise_row(xh);
    ln_backward_row(xh, rstd, d, g, acc[0], acc[1]);
    float4* d4 = reinterpret_cast<float4*>(dx + row * kD);
#pragma unroll
    for (int i = 0; i < kVec; ++i) d4[lane + 32 * i] = make_float4(d[4 * i + 0], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3]);
  }
  cta_reduce_columns<2>(acc, red, part + static_cast<size_t>(blockIdx.x) * 2 * kD);
}

// ---- embeddings ----------------------------------------------------------------------------------------------------
// text: x[b, l] = tok_emb[tok[b, l]] + pos[l]  ->  d pos[l] (+)= sum_b dx[b, l] (one CTA per position, fixed order over b),
// d tok_emb[tok[b, l]] += dx[b, l] (fp32 atomics: many sequences share a token id)
__global__ void __launch_bounds__(192)
text_embed_bwd_kernel(const float* __restrict__ dx, const int64_t* __restrict__ tok, int Ltok, int L, int batch, int vocab,
                      float* __restrict__ dpos, float* __restrict__ demb) {
  const int l = blockIdx.x;
  const int c = threadIdx.x;  // float4 column
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = 0; b < batch; ++b) {
    const float4 v = reinterpret_cast<const float4*>(dx + (static_cast<long long>(b) * L + l) * kD)[c];
    s.x += v.x;
    s.y += v.y;
    s.z += v.z;
    s.w += v.w;
    long long t = tok[static_cast<long long>(b) * Ltok + l];
    if (t < 0 || t >= vocab) t = 0;
    float* e = demb + t * kD + 4 * c;
    atomicAdd(e + 0, v.x);
    atomicAdd(e + 1, v.y);
    atomicAdd(e + 2, v.z);
    atomicAdd(e + 3, v.w);
  }
  float4* d = reinterpret_cast<float4*>(dpos + static_cast<long long>(l) * kD) + c;
  const float4 o = *d;
  *d = make_float4(o.x + s.x, o.y + s.y, o.z + s.z, o.w + s.w);
}

// image: x[b, l] = ln_pre(e[b, l]), e[b, 0] = cls + pos[0], e[b, l] = grid[b, l - 1] + pos[l]   (M.py:2418-2426).
// de = LN_backward(dx; e) is written over dx; part[gridDim.x][2 * 768] = dgamma | dbeta of ln_pre.
__global__ void __launch_bounds__(32 * kWarps)
image_embed_bwd_kernel(const float* __restrict__ grid, const float* __restrict__ cls, const float* __restrict__ pos,
                       const float* __restrict__ gamma, float* __restrict__ dx, float* __restrict__ part, long long rows, int L) {
  __shared__ float red[2 * kD];
  const int lane = threadIdx.x & 31;
  float g[kCols], acc[2][kCols];
  load_gamma(gamma, lane, g);
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int j = 0; j < kCols; ++j) acc[s][j] = 0.f;
  for (long long r = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5); r < rows;
       r += static_cast<long long>(gridDim.x) * kWarps) {
    const long long b = r / L;
    const int l = static_cast<int>(r - b * L);
    float4 v[kVec], p4[kVec];
    float e[kCols], pp[kCols], d[kCols];
    load_row(l == 0 ? cls : grid + (b * (L - 1) + l - 1) * kD, lane, v);
    load_row(pos + static_cast<long long>(l) * kD, lane, p4);
    unpack(v, e);
    unpack(p4, pp);
#pragma unroll
    for (int j = 0; j < kCols; ++j) e[j] += pp[j];
    load_row(dx + r * kD, lane, v);
    unpack(v, d);
    const float rstd = normalise_row(e);
    ln_backward_row(e, rstd, d, g, acc[0], acc[1]);
    float4* d4 = reinterpret_cast<float4*>(dx + r * kD);
#pragma unroll
    for (int i = 0; i < kVec; ++i) d4[lane + 32 * i] = make_float4(d[4 * i + 0], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3]);
  }
  cta_reduce_columns<2>(acc, red, part + static_cast<size_t>(blockIdx.x) * 2 * kD);
}

// d pos[l] (+)= sum_b de[b, l]; d cls (+)= sum_b de[b, 0]   (one CTA per position, fixed order over b)
__global__ void __launch_bounds__(192)
image_pos_bwd_kernel(const float* __restrict__ de, int L, int batch, float* __restrict__ dpos, float* __restrict__ dcls) {
  const int l = blockIdx.x;
  const int c = threadIdx.x;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = 0; b < batch; ++b) {
    const float4 v = reinterpret_cast<const float4*>(de + (static_cast<long long>(b) * L + l) * kD)[c];
    s.x += v.x;
    s.y += v.y;
    s.z += v.z;
    s.w += v.w;
  }
  float4* d = reinterpret_cast<float4*>(dpos + static_cast<long long>(l) * kD) + c;
  float4 o = *d;
  *d = make_float4(o.x + s.x, o.y + s.y, o.z + s.z, o.w + s.w);
  if (l == 0) {
    float4* dc = reinterpret_cast<float4*>(dcls) + c;
    o = *dc;
    *dc = make_float4(o.x + s.x, o.y + s.y, o.z + s.z, o.w + s.w);
  }
}

// ---- lateral adapter, bottom path (M.py:1760-1777): x_out = ln_adapt(s), s[b, 0] = 2 x[b, 0],
// s[b, 1 + p] = dw3x3(grid(x))[p] (BN folded: w9 [9][768], bias [768]) + t[b, p].
// Kernel 1: ds = LN_backward(dx_out; s) with s recomputed from x and t (nine neighbour rows), written over dx_out;
// part = dgamma | dbeta of ln_adapt.  Kernel 2: dx[b, 0] = 2 ds[b, 0]; dx[b, 1 + p] = sum_taps w9[tap] * ds[b, 1 + p - tap].
__global__ void __launch_bounds__(32 * kWarps)
adapter_ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ t, const float* __restrict__ w9,
                      const float* __restrict__ bias, const float* __restrict__ gamma, float* __restrict__ dxo,
                      float* __restrict__ part, long long rows, int gsz) {
  __shared__ float red[2 * kD];
  const int lane = threadIdx.x & 31;
  const int L = gsz * gsz + 1;
  float g[kCols], acc[2][kCols];
  load_gamma(gamma, lane, g);
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int j = 0; j < kCols; ++j) acc[s][j] = 0.f;
  for (long long r = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5); r < rows;
       r += static_cast<long long>(gridDim.x) * kWarps) {
    const long long b = r / L;
    const int l = static_cast<int>(r - b * L);
    float4 v[kVec];
    float sv[kCols], d[kCols];
    if (l == 0) {
      load_row(x + r * kD, lane, v);
      unpack(v, sv);
#pragma unroll
      for (int j = 0; j < kCols; ++j) sv[j] *= 2.0f;
    } else {
      const int py = (l - 1) / gsz, px = (l - 1) % gsz;
      load_row(t + (b * (L - 1) + l - 1) * kD, lane, v);
      unpack(v, sv);
      load_row(bias, lane, v);
      float tmp[kCols];
      unpack(v, tmp);
#pragma unroll
      for (int j = 0; j < kCols; ++j) sv[j] += tmp[j];
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = py + ky - 1;
        if (yy < 0 || yy >= gsz) continue;
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = px + kx - 1;
          if (xx < 0 || xx >= gsz) continue;
          float4 w4[kVec];
          load_row(x + (b * L + 1 + yy * gsz + xx) * kD, lane, v);
          load_row(w9 + (ky * 3 + kx) * kD, lane, w4);
          float xv[kCols], wv[kCols];
          unpack(v, xv);
          unpack(w4, wv);
#pragma unroll
          for (int j = 0; j < kCols; ++j) sv[j] = fmaf(xv[j], wv[j], sv[j]);
        }
      }
    }
    load_row(dxo + r * kD, lane, v);
    unpack(v, d);
    const float rstd = normalise_row(sv);
    ln_backward_row(sv, rstd, d, g, acc[0], acc[1]);
    float4* d4 = reinterpret_cast<float4*>(dxo + r * kD);
#pragma unroll
    for (int i = 0; i < kVec; ++i) d4[lane + 32 * i] = make_float4(d[4 * i + 0], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3]);
  }
  cta_reduce_columns<2>(acc, red, part + static_cast<size_t>(blockIdx.x) * 2 * kD);
}

__global__ void __launch_bounds__(32 * kWarps)
adapter_dw_bwd_kernel(const float* __restrict__ ds, const float* __restrict__ w9, float* __restrict__ dx, long long rows, int gsz) {
  const int lane = threadIdx.x & 31;
  const int L = gsz * gsz + 1;
  for (long long r = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5); r < rows;
       r += static_cast<long long>(gridDim.x) * kWarps) {
    const long long b = r / L;
    const int l = static_cast<int>(r - b * L);
    float4 v[kVec];
    float o[kCols];
    if (l == 0) {
      load_row(ds + r * kD, lane, v);
      unpack(v, o);
#pragma unroll
      for (int j = 0; j < kCols; ++j) o[j] *= 2.0f;
    } else {
#pragma unroll
      for (int j = 0; j < kCols; ++j) o[j] = 0.f;
      const int py = (l - 1) / gsz, px = (l - 1) % gsz;
      // output pixel (qy, qx) reads input (qy + ky - 1, qx + kx - 1): input (py, px) feeds outputs (py - ky + 1, px - kx + 1)
      for (int ky = 0; ky < 3; ++ky) {
        const int qy = py - ky + 1;
        if (qy < 0 || qy >= gsz) continue;
        for (int kx = 0; kx < 3; ++kx) {
          const int qx = px - kx + 1;
          if (qx < 0 || qx >= gsz) continue;
          float4 w4[kVec];
          load_row(ds + (b * L + 1 + qy * gsz + qx) * kD, lane, v);
          load_row(w9 + (ky * 3 + kx) * kD, lane, w4);
          float dv[kCols], wv[kCols];
          unpack(v, dv);
          unpack(w4, wv);
#pragma unroll
          for (int j = 0; j < kCols; ++j) o[j] = fmaf(dv[j], wv[j], o[j]);
        }
      }
    }
    float4* d4 = reinterpret_cast<float4*>(dx + r * kD);
#pragma unroll
    for (int i = 0; i < kVec; ++i) d4[lane + 32 * i] = make_float4(o[4 * i + 0], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
  }
}

// ---- fused multi-tensor AdamW ----------------------------------------------------------------------------------------
// p -= lr * (m_hat / (sqrt(v_hat) + eps) + wd * p), torch.optim.AdamW semantics (decoupled decay applied first)
__global__ void __launch_bounds__(256)
adamw_kernel(const AdamwTensor* __restrict__ tab, const int* __restrict__ chunk_tensor, const long long* __restrict__ chunk_off,
             int nchunks, int chunk, float beta1, float beta2, float eps, float bc1, float bc2) {
  for (int ci = blockIdx.x; ci < nchunks; ci += gridDim.x) {
    const AdamwTensor t = tab[chunk_tensor[ci]];
    const long long o0 = chunk_off[ci];
    const long long n = t.numel - o0 < chunk ? t.numel - o0 : chunk;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const long long k = o0 + i;
      const float g = t.grad[k];
      float p = t.param[k];
      const float m = beta1 * t.m[k] + (1.0f - beta1) * g;
      const float v = beta2 * t.v[k] + (1.0f - beta2) * g * g;
      p *= 1.0f - t.lr * t.wd;
      p -= t.lr * (m / bc1) / (sqrtf(v / bc2) + eps);
      t.m[k] = m;
      t.v[k] = v;
      t.param[k] = p;
    }
  }
}

int grid_for_rows(long long rows) {
  const long long want = (rows + kWarps - 1) / kWarps;
  const long long cap = static_cast<long long>(num_sms()) * 4;
  return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace

int bwd_row_parts(long long rows) { return grid_for_rows(rows); }

int launch_reduce_partials(const float* part, int nparts, long long pitch, float* dst, long long n, int accumulate,
                           long long head_n, float head_scale, cudaStream_t stream) {
  if (n <= 0) return 0;
  const int grid = static_cast<int>((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  reduce_partials_kernel<<<grid, 256, 0, stream>>>(part, nparts, pitch, dst, n, accumulate, head_n, head_scale);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// d0 / d1 / d2 [n] += column sums of part[:, 0:n] / [:, n:2n] / [:, 2n:3n] (null destinations are skipped)
int launch_reduce_partials3(const float* part, int nparts, long long pitch, float* d0, float* d1, float* d2, int n, cudaStream_t stream) {
  reduce_partials3_kernel<<<(3 * n + 255) / 256, 256, 0, stream>>>(part, nparts, pitch, d0, d1, d2, n);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_ln_bwd(const float* x, const float* dy, const float* gamma, float* dx, op16* g16, float* part, long long rows,
                  int accumulate, cudaStream_t stream) {
  ln_bwd_kernel<<<grid_for_rows(rows), 32 * kWarps, 0, stream>>>(x, dy, gamma, dx, g16, part, rows, accumulate);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_cast_colsum(const float* dx, op16* g16, float* part, long long rows, cudaStream_t stream) {
  cast_colsum_kernel<<<grid_for_rows(rows), 32 * kWarps, 0, stream>>>(dx, g16, part, rows);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int bwd_slab_parts(long long rows) {
  // slabs of rows: enough CTAs to fill the machine a few times over, at least 64 rows each
  long long slabs = (rows + 63) / 64;
  const long long cap = 128;
  if (slabs > cap) slabs = cap;
  return static_cast<int>(slabs < 1 ? 1 : slabs);
}

int launch_qgelu_fwd(const op16* u, op16* a, long long n, cudaStream_t stream) {
  MSCLIP_REQUIRE(n % 8 == 0, "qgelu_fwd: element count must be a multiple of 8");
  const long long n8 = n / 8;
  const long long want = (n8 + kSlabThreads - 1) / kSlabThreads;
  const int grid = static_cast<int>(want < num_sms() * 8 ? (want > 0 ? want : 1) : num_sms() * 8);
  qgelu_fwd_kernel<<<grid, kSlabThreads, 0, stream>>>(u, a, n8);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_qgelu_bwd(op16* da, const op16* u, float* part, long long rows, int width, cudaStream_t stream) {
  MSCLIP_REQUIRE(width % 256 == 0, "qgelu_bwd: width must be a multiple of 256");
  const int slabs = bwd_slab_parts(rows);
  const int rps = static_cast<int>((rows + slabs - 1) / slabs);
  qgelu_bwd_kernel<<<dim3(width / 256, slabs), kSlabThreads, 0, stream>>>(da, u, part, rows, width, rps);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_colsum16(const op16* g, float* part, long long rows, int width, cudaStream_t stream) {
  MSCLIP_REQUIRE(width % 256 == 0, "colsum16: width must be a multiple of 256");
  const int slabs = bwd_slab_parts(rows);
  const int rps = static_cast<int>((rows + slabs - 1) / slabs);
  colsum16_kernel<<<dim3(width / 256, slabs), kSlabThreads, 0, stream>>>(g, part, rows, width, rps);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_l2norm_bwd(const float* g, const float* df, float* dg, op16* dg16, int rows, int E, int normalise, cudaStream_t stream) {
  l2norm_bwd_kernel<<<(rows + kWarps - 1) / kWarps, 32 * kWarps, 0, stream>>>(g, df, dg, dg16, rows, E, normalise);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_pooled_ln_bwd(const float* x, int Lx, const int64_t* tok, int Ltok, const float* dz, const float* gamma, float* dx,
                         float* part, int batch, cudaStream_t stream) {
  pooled_ln_bwd_kernel<<<grid_for_rows(batch), 32 * kWarps, 0, stream>>>(x, Lx, tok, Ltok, dz, gamma, dx, part, batch);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_text_embed_bwd(const float* dx, const int64_t* tok, int Ltok, int L, int batch, int vocab, float* dpos, float* demb,
                          cudaStream_t stream) {
  text_embed_bwd_kernel<<<L, 192, 0, stream>>>(dx, tok, Ltok, L, batch, vocab, dpos, demb);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_image_embed_bwd(const float* grid, const float* cls, const float* pos, const float* gamma, float* dx, float* part,
                           int batch, int L, float* dpos, float* dcls, cudaStream_t stream) {
  const long long rows = static_cast<long long>(batch) * L;
  image_embed_bwd_kernel<<<grid_for_rows(rows), 32 * kWarps, 0, stream>>>(grid, cls, pos, gamma, dx, part, rows, L);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  image_pos_bwd_kernel<<<L, 192, 0, stream>>>(dx, L, batch, dpos, dcls);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_adapter_bwd(const float* x, const float* t, const float* w9, const float* bias, const float* gamma, float* dxo,
                       float* dx, float* part, int batch, int gsz, cudaStream_t stream) {
  const long long rows = static_cast<long long>(batch) * (gsz * gsz + 1);
  adapter_ln_bwd_kernel<<<grid_for_rows(rows), 32 * kWarps, 0, stream>>>(x, t, w9, bias, gamma, dxo, part, rows, gsz);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  adapter_dw_bwd_kernel<<<grid_for_rows(rows), 32 * kWarps, 0, stream>>>(dxo, w9, dx, rows, gsz);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_adamw(const AdamwTensor* tab_dev, const int* chunk_tensor_dev, const long long* chunk_off_dev, int nchunks, int chunk,
                 float beta1, float beta2, float eps, int step, cudaStream_t stream) {
  if (nchunks <= 0) return 0;
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  const int grid = nchunks < num_sms() * 8 ? nchunks : num_sms() * 8;
  adamw_kernel<<<grid, 256, 0, stream>>>(tab_dev, chunk_tensor_dev, chunk_off_dev, nchunks, chunk, beta1, beta2, eps, bc1, bc2);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace msclip
