// HBM-bound row kernels of the MS-CLIP-S path: LayerNorm (M.py:204-219), token / image embedding
// (M.py:3047-3048, 2418-2426), the lateral-adapter tail (M.py:1760-1777), EOT pooling
// (M.py:3057-3060) and the final L2 normalisation (M.py:2982-2983, 3076-3077).
// One warp owns one 768-wide row: 6 float4 per lane, statistics by warp shuffle in fp32, two-pass
// (mean, then centred variance) exactly like the reference's (x-u).pow(2).mean().
#include "common.cuh"
#include "kernels.h"
#include "rowops.cuh"

namespace msclip {

namespace {

using namespace rowops;
constexpr int kRowsPerBlock = 8;  // 8 warps per CTA

__device__ __forceinline__ void store_row_f32(float* __restrict__ dst, int lane, const float4 (&v)[kVec]) {
  float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int i = 0; i < kVec; ++i) d4[lane + 32 * i] = v[i];
}

// LN fold (gemm_common.cuh): centred 16-bit copy of a freshly produced residual row and its record -
// shift = exact row mean, slice 0 = (sum, sum of squares) of the centred row, slices 1..5 empty
__device__ __forceinline__ void emit_centred_row(const float4 (&v)[kVec], int lane, op16* __restrict__ xc,
                                                 float* __restrict__ rec) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / kD);
  float4 c[kVec];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    c[i] = make_float4(v[i].x - mean, v[i].y - mean, v[i].z - mean, v[i].w - mean);
    s1 += (c[i].x + c[i].y) + (c[i].z + c[i].w);
    s2 += (c[i].x * c[i].x + c[i].y * c[i].y) + (c[i].z * c[i].z + c[i].w * c[i].w);
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  store_row_bf16(xc, lane, c);
  if (lane < kLnRecordFloats) rec[lane] = lane == 0 ? mean : (lane == 4 ? s1 : (lane == 5 ? s2 : 0.f));
}

__global__ void __launch_bounds__(32 * kRowsPerBlock)
layernorm_kernel(const float* __restrict__ x, int row_stride, const float* __restrict__ w,
                      const float* __restrict__ b, op16* __restrict__ y, int rows) {
  const int lane = threadIdx.x & 31;
  for (long long r = static_cast<long long>(blockIdx.x) * kRowsPerBlock + (threadIdx.x >> 5); r < rows;
       r += static_cast<long long>(gridDim.x) * kRowsPerBlock) {
    const long long src = r * row_stride;
    float4 v[kVec];
    load_row(x + src * kD, lane, v);
    layer_norm_row(v, w, b, lane);
    store_row_bf16(y + r * kD, lane, v);
  }
}

__global__ void __launch_bounds__(32 * kRowsPerBlock)
eot_layernorm_kernel(const float* __restrict__ x, int Lx, const int64_t* __restrict__ tok, int L,
                          const float* __restrict__ w, const float* __restrict__ b, op16* __restrict__ y, int batch) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (r >= batch) return;
  // argmax over the token ids, first occurrence on ties (torch.argmax, M.py:3059)
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
  float4 v[kVec];
  if (best_i >= Lx) best_i = Lx - 1;  // cannot happen when Lx covers the longest sequence; keeps the read in bounds
  load_row(x + (static_cast<long long>(r) * Lx + best_i) * kD, lane, v);
  layer_norm_row(v, w, b, lane);
  store_row_bf16(y + static_cast<long long>(r) * kD, lane, v);
}

// longest live prefix of a batch: max over sequences of argmax_j tok[b, j] + 1 (everything after the EOT token cannot
// influence the pooled output of a causal tower, M.py:2965-2971 + 3059)
__global__ void __launch_bounds__(32 * kRowsPerBlock)
text_max_len_kernel(const int64_t* __restrict__ tok, int L, int batch, int* __restrict__ out_max) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (r >= batch) return;
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
  if (lane == 0) atomicMax(out_max, best_i + 1);
}

// The producers of the residual stream can also emit the LayerNorm the next GEMM consumes (ln_1 of the block that
// follows): the row is in registers anyway, so the separate LayerNorm launch (one more read of the fp32 row) disappears.
// Same row arithmetic as layernorm_kernel -> bit-identical h.
struct NextLn {
  const float* w;
  const float* b;
  op16* h;  // null: not requested
};
__device__ __forceinline__ void emit_next_ln(float4 (&v)[kVec], const NextLn& nl, long long r, int lane) {
  if (nl.h == nullptr) return;
  layer_norm_row(v, nl.w, nl.b, lane);
  store_row_bf16(nl.h + r * kD, lane, v);
}

__global__ void __launch_bounds__(32 * kRowsPerBlock)
text_embed_kernel(const int64_t* __restrict__ tok, const float* __restrict__ emb, const float* __restrict__ pos,
                  float* __restrict__ x, long long rows, int L, int Ltok, int vocab, int* __restrict__ err,
                  op16* __restrict__ xc, float* __restrict__ rec, NextLn nl) {
  const int lane = threadIdx.x & 31;
  for (long long r = static_cast<long long>(blockIdx.x) * kRowsPerBlock + (threadIdx.x >> 5); r < rows;
       r += static_cast<long long>(gridDim.x) * kRowsPerBlock) {
    long long t = tok[(r / L) * Ltok + (r % L)];  // rows: L live positions per sequence, tokens: pitch Ltok
    if (t < 0 || t >= vocab) {  // nn.Embedding raises; we flag and clamp so the kernel stays in bounds
      if (lane == 0) atomicExch(err, 1);
      t = 0;
    }
    const int l = static_cast<int>(r % L);
    float4 v[kVec], p[kVec];
    load_row(emb + t * kD, lane, v);
    load_row(pos + static_cast<long long>(l) * kD, lane, p);
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      v[i].x += p[i].x;
      v[i].y += p[i].y;
      v[i].z += p[i].z;
      v[i].w += p[i].w;
    }
    store_row_f32(x + r * kD, lane, v);
    if (xc != nullptr) emit_centred_row(v, lane, xc + r * kD, rec + r * kLnRecordFloats);
    emit_next_ln(v, nl, r, lane);
  }
}

__global__ void __launch_bounds__(32 * kRowsPerBlock)
image_embed_ln_pre_kernel(const float* __restrict__ grid, const float* __restrict__ cls, const float* __restrict__ pos,
                          const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ x,
                          long long rows, int L, op16* __restrict__ xc, float* __restrict__ rec, NextLn nl) {
  const int lane = threadIdx.x & 31;
  for (long long r = static_cast<long long>(blockIdx.x) * kRowsPerBlock + (threadIdx.x >> 5); r < rows;
       r += static_cast<long long>(gridDim.x) * kRowsPerBlock) {
    const long long bi = r / L;
    const int l = static_cast<int>(r % L);
    float4 v[kVec], p[kVec];
    load_row(l == 0 ? cls : grid + (bi * (L - 1) + (l - 1)) * kD, lane, v);
    load_row(pos + static_cast<long long>(l) * kD, lane, p);
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      v[i].x += p[i].x;
      v[i].y += p[i].y;
      v[i].z += p[i].z;
      v[i].w += p[i].w;
    }
    layer_norm_row(v, w, b, lane);
    store_row_f32(x + r * kD, lane, v);
    if (xc != nullptr) emit_centred_row(v, lane, xc + r * kD, rec + r * kLnRecordFloats);
    emit_next_ln(v, nl, r, lane);
  }
}

// dw_w9: [9][768] (BN scale folded), dw_bias: [768] (BN shift).  Tap (dy,dx) index = (dy+1)*3 + (dx+1).
__global__ void __launch_bounds__(32 * kRowsPerBlock)
adapter_fuse_ln_kernel(const float* __restrict__ x, const float* __restrict__ t, const float* __restrict__ dw_w9,
                       const float* __restrict__ dw_bias, const float* __restrict__ w, const float* __restrict__ b,
                       float* __restrict__ x_out, long long rows, int g, op16* __restrict__ xc, float* __restrict__ rec,
                       NextLn nl) {
  const int lane = threadIdx.x & 31;
  const int L = g * g + 1;
  for (long long r = static_cast<long long>(blockIdx.x) * kRowsPerBlock + (threadIdx.x >> 5); r < rows;
       r += static_cast<long long>(gridDim.x) * kRowsPerBlock) {
    const long long bi = r / L;
    const int l = static_cast<int>(r % L);
    float4 v[kVec];
    if (l == 0) {
      load_row(x + r * kD, lane, v);  // cls + cls (PRALLEL_T2B_USECLS, M.py:1770-1771)
#pragma unroll
      for (int i = 0; i < kVec; ++i) {
        v[i].x += v[i].x;
        v[i].y += v[i].y;
        v[i].z += v[i].z;
        v[i].w += v[i].w;
      }
    } else {
      // depth-wise 3x3 over the token grid + t, one third of the channels at a time with all nine neighbour rows
      // in flight (the serial tap loop paid one L2 round trip per tap); out-of-grid taps load zeros
      const int gy = (l - 1) / g, gx = (l - 1) % g;
      const float4* xg = reinterpret_cast<const float4*>(x + (bi * L + 1) * kD);
      const float4* t4 = reinterpret_cast<const float4*>(t + (bi * (L - 1) + (l - 1)) * kD);
      const float4* w4 = reinterpret_cast<const float4*>(dw_w9);
      const float4* b4 = reinterpret_cast<const float4*>(dw_bias);
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      static_assert(kVec % 2 == 0, "channel thirds are pairs of float4");
#pragma unroll
      for (int part = 0; part < kVec / 2; ++part) {
        const int c0 = lane + 64 * part, c1 = c0 + 32;
        float4 a[9][2];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int yy = gy + tap / 3 - 1, xx = gx + tap % 3 - 1;
          const bool ok = yy >= 0 && yy < g && xx >= 0 && xx < g;
          const float4* src = xg + static_cast<long long>(ok ? yy * g + xx : 0) * (kD / 4);
          a[tap][0] = ok ? src[c0] : zero;
          a[tap][1] = ok ? src[c1] : zero;
        }
        const float4 tt0 = t4[c0], tt1 = t4[c1];
        float4 acc0 = b4[c0], acc1 = b4[c1];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const float4 k0 = w4[tap * (kD / 4) + c0], k1 = w4[tap * (kD / 4) + c1];
          acc0.x = fmaf(a[tap][0].x, k0.x, acc0.x); acc0.y = fmaf(a[tap][0].y, k0.y, acc0.y);
          acc0.z = fmaf(a[tap][0].z, k0.z, acc0.z); acc0.w = fmaf(a[tap][0].w, k0.w, acc0.w);
          acc1.x = fmaf(a[tap][1].x, k1.x, acc1.x); acc1.y = fmaf(a[tap][1].y, k1.y, acc1.y);
          acc1.z = fmaf(a[tap][1].z, k1.z, acc1.z); acc1.w = fmaf(a[tap][1].w, k1.w, acc1.w);
        }
        v[2 * part] = make_float4(acc0.x + tt0.x, acc0.y + tt0.y, acc0.z + tt0.z, acc0.w + tt0.w);
        v[2 * part + 1] = make_float4(acc1.x + tt1.x, acc1.y + tt1.y, acc1.z + tt1.z, acc1.w + tt1.w);
      }
    }
    layer_norm_row(v, w, b, lane);
    store_row_f32(x_out + r * kD, lane, v);
    if (xc != nullptr) emit_centred_row(v, lane, xc + r * kD, rec + r * kLnRecordFloats);
    emit_next_ln(v, nl, r, lane);
  }
}

// one warp per row of width E (multiple of 4, <= 1024)
__global__ void __launch_bounds__(32 * kRowsPerBlock)
l2norm_kernel(const float* __restrict__ x, float* __restrict__ out_f32, emb16* __restrict__ out_f16, int rows, int E,
              int normalise) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float4* x4 = reinterpret_cast<const float4*>(x + static_cast<long long>(r) * E);
  const int n4 = E / 4;
  float4 v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < n4 ? x4[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float inv = normalise ? 1.0f / sqrtf(warp_sum(s)) : 1.0f;  // no eps, like x / x.norm()
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    if (c >= n4) continue;
    const float4 o = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
    if (out_f32) reinterpret_cast<float4*>(out_f32 + static_cast<long long>(r) * E)[c] = o;
    if (out_f16) {
      const __half2 a = __floats2half2_rn(o.x, o.y), b = __floats2half2_rn(o.z, o.w);
      reinterpret_cast<uint2*>(out_f16 + static_cast<long long>(r) * E)[c] =
          make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    }
  }
}

inline int row_grid(long long rows) {
  long long blocks = (rows + kRowsPerBlock - 1) / kRowsPerBlock;
  const long long cap = static_cast<long long>(num_sms()) * 16;  // grid-stride beyond 16 CTAs / SM
  return static_cast<int>(blocks < cap ? blocks : cap);
}

}  // namespace

int launch_layernorm_op16(const float* x, int row_stride, const float* w, const float* b, op16* y, int rows,
                          cudaStream_t stream) {
  if (rows <= 0) return 0;
  layernorm_kernel<<<row_grid(rows), 32 * kRowsPerBlock, 0, stream>>>(x, row_stride, w, b, y, rows);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_eot_layernorm_op16(const float* x, int x_len, const int64_t* tok, int L, const float* w, const float* b, op16* y,
                              int batch, cudaStream_t stream) {
  if (batch <= 0) return 0;
  eot_layernorm_kernel<<<(batch + kRowsPerBlock - 1) / kRowsPerBlock, 32 * kRowsPerBlock, 0, stream>>>(
      x, x_len, tok, L, w, b, y, batch);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_text_max_len(const int64_t* tok, int L, int batch, int* out_max, cudaStream_t stream) {
  if (batch <= 0) return 0;
  text_max_len_kernel<<<(batch + kRowsPerBlock - 1) / kRowsPerBlock, 32 * kRowsPerBlock, 0, stream>>>(tok, L, batch, out_max);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int* token_error_flag() {
  static int* flag = nullptr;
  if (!flag) {
    if (cudaMalloc(&flag, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(flag, 0, sizeof(int));
  }
  return flag;
}

int launch_text_embed(const int64_t* tok, int tok_pitch, const float* tok_emb, const float* pos, float* x, int batch, int L,
                      int vocab, op16* xc, float* rec, const float* next_ln_w, const float* next_ln_b, op16* next_h,
                      cudaStream_t stream) {
  if (batch <= 0) return 0;
  int* flag = token_error_flag();
  MSCLIP_REQUIRE(flag != nullptr, "cudaMalloc of the token error flag failed");
  const long long rows = static_cast<long long>(batch) * L;
  text_embed_kernel<<<row_grid(rows), 32 * kRowsPerBlock, 0, stream>>>(tok, tok_emb, pos, x, rows, L, tok_pitch, vocab, flag,
                                                                       xc, xc ? rec : nullptr, NextLn{next_ln_w, next_ln_b, next_h});
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int check_token_error(cudaStream_t stream) {
  int* flag = token_error_flag();
  if (!flag) return 0;
  int h = 0;
  MSCLIP_CHECK_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
  MSCLIP_CHECK_CUDA(cudaStreamSynchronize(stream));
  if (h) {
    cudaMemsetAsync(flag, 0, sizeof(int), stream);
    set_last_error("encode_text: token id out of range [0, vocab)");
    return 3;
  }
  return 0;
}

int launch_image_embed_ln_pre(const float* grid, const float* cls, const float* pos, const float* w, const float* b,
                              float* x, int batch, int L, op16* xc, float* rec, const float* next_ln_w,
                              const float* next_ln_b, op16* next_h, cudaStream_t stream) {
  if (batch <= 0) return 0;
  const long long rows = static_cast<long long>(batch) * L;
  image_embed_ln_pre_kernel<<<row_grid(rows), 32 * kRowsPerBlock, 0, stream>>>(grid, cls, pos, w, b, x, rows, L, xc, rec,
                                                                               NextLn{next_ln_w, next_ln_b, next_h});
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_adapter_fuse_ln(const float* x, const float* t, const float* dw_w9, const float* dw_bias, const float* w,
                           const float* b, float* x_out, int batch, int g, op16* xc, float* rec, const float* next_ln_w,
                           const float* next_ln_b, op16* next_h, cudaStream_t stream) {
  if (batch <= 0) return 0;
  MSCLIP_REQUIRE(x != x_out, "adapter tail cannot run in place (3x3 neighbourhood reads)");
  const long long rows = static_cast<long long>(batch) * (g * g + 1);
  adapter_fuse_ln_kernel<<<row_grid(rows), 32 * kRowsPerBlock, 0, stream>>>(x, t, dw_w9, dw_bias, w, b, x_out, rows, g, xc,
                                                                            rec, NextLn{next_ln_w, next_ln_b, next_h});
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---- zero-shot classifier head (tools/zero_shot.py:125-131, 266, 150-163) -------------------------------------------
// weights[c] = normalise(mean_t feat[row_of[c * n_templates + t]]): one warp per class, E = 512 (4 float4 per lane)
__global__ void __launch_bounds__(256)
class_mean_renorm_kernel(const float* __restrict__ feat, const int* __restrict__ row_of, int n_classes, int n_templates, int E,
                         float* __restrict__ weights) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= n_classes) return;
  const int n4 = E / 4;
  float4 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < n_templates; ++t) {
    const long long idx = static_cast<long long>(c) * n_templates + t;
    const long long r = row_of ? row_of[idx] : idx;
    const float4* f4 = reinterpret_cast<const float4*>(feat + r * E);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int col = lane + 32 * i;
      if (col < n4) {
        const float4 v = f4[col];
        acc[i].x += v.x;
        acc[i].y += v.y;
        acc[i].z += v.z;
        acc[i].w += v.w;
      }
    }
  }
  const float inv_t = 1.0f / static_cast<float>(n_templates);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i].x *= inv_t;
    acc[i].y *= inv_t;
    acc[i].z *= inv_t;
    acc[i].w *= inv_t;
    s += (acc[i].x * acc[i].x + acc[i].y * acc[i].y) + (acc[i].z * acc[i].z + acc[i].w * acc[i].w);
  }
  const float inv = 1.0f / sqrtf(warp_sum(s));  // class_embedding /= class_embedding.norm(), no eps
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int col = lane + 32 * i;
    if (col < n4)
      reinterpret_cast<float4*>(weights + static_cast<long long>(c) * E)[col] =
          make_float4(acc[i].x * inv, acc[i].y * inv, acc[i].z * inv, acc[i].w * inv);
  }
}

// top-k (k <= 8) class indices per row of logits [rows, n]: one warp per row; every lane keeps the k best of its strided
// share (sorted insertion), then k rounds of warp arg-max merge them (ties -> lower index, like torch.topk on CUDA is
// NOT guaranteed to; callers compare scores, not tie order)
__global__ void __launch_bounds__(256)
topk_rows_kernel(const float* __restrict__ logits, int rows, int n, int k, int* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  float bv[8];
  int bi[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    bv[i] = -INFINITY;
    bi[i] = 0x7fffffff;
  }
  const float* row = logits + static_cast<long long>(r) * n;
  for (int j = lane; j < n; j += 32) {
    float v = row[j];
    int idx = j;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < k && (v > bv[i] || (v == bv[i] && idx < bi[i]))) {
        const float tv = bv[i];
        const int ti = bi[i];
        bv[i] = v;
        bi[i] = idx;
        v = tv;
        idx = ti;
      }
    }
  }
  for (int round = 0; round < k; ++round) {
    float v = bv[0];
    int idx = bi[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) {
        v = ov;
        idx = oi;
      }
    }
    if (lane == 0) out[static_cast<long long>(r) * k + round] = idx;
    if (bi[0] == idx) {  // the winning lane pops its head
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        bv[i] = bv[i + 1];
        bi[i] = bi[i + 1];
      }
      bv[7] = -INFINITY;
      bi[7] = 0x7fffffff;
    }
  }
}

int launch_class_mean_renorm(const float* feat, const int* row_of, int n_classes, int n_templates, int E, float* weights,
                             cudaStream_t stream) {
  if (n_classes <= 0) return 0;
  MSCLIP_REQUIRE(E % 4 == 0 && E <= 1024 && n_templates >= 1, "class_mean_renorm: width must be a multiple of 4 and <= 1024");
  class_mean_renorm_kernel<<<(n_classes + 7) / 8, 256, 0, stream>>>(feat, row_of, n_classes, n_templates, E, weights);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_topk_rows(const float* logits, int rows, int n, int k, int* out, cudaStream_t stream) {
  if (rows <= 0) return 0;
  MSCLIP_REQUIRE(k >= 1 && k <= 8 && k <= n, "topk_rows: k must be in [1, min(8, n)]");
  topk_rows_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(logits, rows, n, k, out);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_l2norm(const float* x, float* out_f32, emb16* out_f16, int rows, int E, int normalise,
                  cudaStream_t stream) {
  if (rows <= 0) return 0;
  MSCLIP_REQUIRE(E % 4 == 0 && E <= 1024, "l2norm: width must be a multiple of 4 and <= 1024");
  l2norm_kernel<<<(rows + kRowsPerBlock - 1) / kRowsPerBlock, 32 * kRowsPerBlock, 0, stream>>>(x, out_f32, out_f16,
                                                                                              rows, E, normalise);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace msclip
