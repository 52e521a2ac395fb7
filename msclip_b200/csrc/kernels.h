// Internal launcher interface between the engine (engine.cu), the C ABI (api.cu) and the kernels.
// Every launcher is stream-ordered, returns 0 on success and records a message with
// set_last_error() otherwise.  Activations are batch-major: token row r = b*L + l.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace msclip {

// 16-bit operand type of every MMA (activations, packed weights, embeddings): bf16 by default, IEEE fp16 when
// the library is compiled with -DMSCLIP_FP16 (same tensor-core rate, 3 more mantissa bits, 5-bit less exponent;
// conversions saturate at +-65504).  Accumulation is fp32 in both builds.
#ifdef MSCLIP_FP16
typedef __half op16;
typedef __half2 op162;
#define MSCLIP_OPERAND_NAME "fp16"
#define MSCLIP_MMA_OPERANDS "f16.f16"
#define MSCLIP_TMA_DTYPE CU_TENSOR_MAP_DATA_TYPE_FLOAT16
constexpr uint32_t kUmmaOperandFormat = 0;  // kind::f16 A/B format field: 0 = F16, 1 = BF16
#else
typedef __nv_bfloat16 op16;
typedef __nv_bfloat162 op162;
#define MSCLIP_OPERAND_NAME "bf16"
#define MSCLIP_MMA_OPERANDS "bf16.bf16"
#define MSCLIP_TMA_DTYPE CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
constexpr uint32_t kUmmaOperandFormat = 1;
#endif

// Normalised embeddings handed to the contrastive loss (and exchanged between ranks) are IEEE fp16 in BOTH builds:
// |x| <= 1, so fp16's 11-bit significand applies (8x less rounding error than bf16 at the same size and tensor-core
// rate) and its narrow exponent costs nothing (elements below 6e-5 lose < 3e-8 absolute).
typedef __half emb16;
constexpr uint32_t kUmmaFormatF16 = 0;

__host__ __device__ inline op16 to_op16(float x) {
#ifdef MSCLIP_FP16
  return __float2half_rn(fminf(fmaxf(x, -65504.0f), 65504.0f));
#else
  return __float2bfloat16_rn(x);
#endif
}
__host__ __device__ inline float op16_to_float(op16 x) {
#ifdef MSCLIP_FP16
  return __half2float(x);
#else
  return __bfloat162float(x);
#endif
}

int num_sms();

// ---- tcgen05 GEMM: out[M,N] = epi(A[M,K] . W[N,K]^T + bias)  (gemm.cu) ----------------------
enum GemmEpilogue {
  EPI_BF16 = 0,        // op16 out = acc + bias
  EPI_QGELU_BF16 = 1,  // op16 out = quickgelu(acc + bias)          (M.py:222-224)
  EPI_RELU_BF16 = 2,   // op16 out = relu(acc + bias)               (conv + folded BN + ReLU)
  EPI_RESID_F32 = 3,   // f32  out = resid + acc + bias             (residual stream, may alias out)
  EPI_F32 = 4,         // f32  out = acc + bias
  // backward of the MLP (M.py:794-798): acc = d(fc2 input); aux = u, the fc1 pre-activation.  op16 out = acc * quickgelu'(u)
  // (= d u) and op16 out2 = quickgelu(u) (the fc2 input the weight gradient needs): both QuickGELU passes of the backward
  // ride on the dgrad GEMM's epilogue instead of two more trips over the [M, 3072] tensors
  EPI_DGELU_BF16 = 5,
  // training forward of fc1: op16 out = quickgelu(acc + bias) as EPI_QGELU_BF16, and op16 out2 = acc + bias (the
  // pre-activation the backward needs) - saves the backward a whole fc1 GEMM
  EPI_QGELU_DUAL_BF16 = 6,
};
int launch_gemm(const op16* A, int64_t lda, const op16* W, int64_t ldw, int M, int N, int K, const float* bias,
                void* out, int64_t ldo, const float* resid, int64_t ldr, int epi, cudaStream_t stream);

int launch_gemm_split(const op16* A, int64_t lda, const op16* W, int64_t ldw, int M, int N, int K, const float* bias,
                      void* out, int split_cols, int epi, cudaStream_t stream);
// fp16 operands in either build (only 16-bit-out / f32-out epilogues without bias or residual): the loss backward
int launch_gemm_f16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, float alpha, void* out, int64_t ldo,
                    int epi, cudaStream_t stream);
// a = quickgelu(A . W^T + bias), u = A . W^T + bias   (EPI_QGELU_DUAL_BF16; N % 256 == 0, all pitches = N)
int launch_gemm_qgelu_dual(const op16* A, int64_t lda, const op16* W, int64_t ldw, int M, int N, int K, const float* bias, op16* a,
                           op16* u, cudaStream_t stream);
// du = (A . W^T) * quickgelu'(u), a = quickgelu(u)   (EPI_DGELU_BF16; N % 256 == 0, all pitches = N)
int launch_gemm_dgelu(const op16* A, int64_t lda, const op16* W, int64_t ldw, int M, int N, int K, const op16* u, op16* du, op16* a,
                      cudaStream_t stream);
void gemm_set_pair_mode(int mode);  // 0: 1 CTA per tile, 1: CTA pairs, 2|4: multicast clusters of 2|4 pairs
int launch_gemm_scaled(const op16* A, int64_t lda, const op16* W, int64_t ldw, int M, int N, int K, float alpha,
                       const float* bias, void* out, int64_t ldo, const float* resid, int64_t ldr, int epi,
                       cudaStream_t stream);

// Residual GEMM + LayerNorm of the updated rows in one kernel (gemm.cu, LN = 3): x += A . W^T + bias (fp32, in place);
// h = LayerNorm(x) * gamma + beta (op16, bit-identical to launch_layernorm_op16 on the same x).  N = 768, M >= 256.
// counters: gemm_resid_ln_counters(M) zero-initialised 32-bit words of device memory (every launch leaves them zero again)
int launch_gemm_resid_ln(const op16* A, int64_t lda, const op16* W, int64_t ldw, int M, int N, int K, const float* bias, float* x,
                         int64_t ldx, const float* gamma, const float* beta, op16* h, int64_t ldh, uint32_t* counters,
                         cudaStream_t stream);
size_t gemm_resid_ln_counters(int M);

// LayerNorm folded into the GEMMs around it (gemm_common.cuh).  ln_mode 1 (QKV, fc1; epi EPI_BF16 / EPI_QGELU_BF16):
// A = op16(x - shift), W = W * diag(gamma), bias = b + W . beta, colsum[n] = sum_k W'[n][k]; the epilogue applies
// rstd * (acc - mean_c * colsum) from the row records ln_in.  ln_mode 2 (out-proj, fc2; EPI_RESID_F32, N = 768): also
// writes out16 = op16(x_new - shift_new) and the records ln_out of x_new (shift_new = row mean of resid, from ln_in).
// Row record: 16 floats, [0] shift, [4 + 2 s], [5 + 2 s] = sum / sum of squares of the centred row over 128-column slice s.
constexpr int kLnRecordFloats = 16;
int launch_gemm_ln(const op16* A, int64_t lda, const op16* W, int64_t ldw, int M, int N, int K, const float* bias, void* out,
                   int64_t ldo, const float* resid, int64_t ldr, int epi, int ln_mode, const float* ln_in, float* ln_out,
                   op16* out16, int64_t ldo16, const float* colsum, cudaStream_t stream);
// dst[n][k] = op16(src[n][k] * row_scale[n] * gamma[k]); colsum[n] = sum_k dst[n][k]; bias_out[n] = row_scale[n] *
// (bias[n] + sum_k src[n][k] * beta[k])   (row_scale may be null)
int launch_pack_ln_fold(const float* src, const float* row_scale, const float* gamma, const float* beta, const float* bias,
                        op16* dst, float* colsum, float* bias_out, int N, int K, cudaStream_t stream);

// ---- LayerNorm family, D = 768 (elementwise.cu) ----------------------------------------------
// y[r] = LN(x[r * row_stride]) ; out op16 (GEMM operand).  row_stride = L picks the CLS rows.
int launch_layernorm_op16(const float* x, int row_stride, const float* w, const float* b, op16* y, int rows,
                          cudaStream_t stream);
// text pooling: row b <- LN(x[b*L + argmax_j tok[b,j]])  (M.py:3057-3060, 3072)
// x holds x_len (<= L) live positions per sequence, tok has pitch L
int launch_eot_layernorm_op16(const float* x, int x_len, const int64_t* tok, int L, const float* w, const float* b, op16* y,
                              int batch, cudaStream_t stream);
// *out_max = max(*out_max, max_b argmax_j tok[b, j] + 1): the longest live prefix of the batch
int launch_text_max_len(const int64_t* tok, int L, int batch, int* out_max, cudaStream_t stream);
// x[b*L+l] = tok_emb[tok[b*tok_pitch + l]] + pos[l], l < L <= tok_pitch   (M.py:3047-3048)
// xc / rec (optional, all three producers of the residual stream): op16(x - mean) and the row record for the LN fold.
// next_h (optional, all three): also emit LayerNorm(x) * next_ln_w + next_ln_b as op16 - ln_1 of the block that follows -
// so that block needs no separate LayerNorm launch (bit-identical to launch_layernorm_op16 on the same x).
int launch_text_embed(const int64_t* tok, int tok_pitch, const float* tok_emb, const float* pos, float* x, int batch, int L,
                      int vocab, op16* xc, float* rec, const float* next_ln_w, const float* next_ln_b, op16* next_h,
                      cudaStream_t stream);
// x[b*L+l] = ln_pre((l == 0 ? cls : grid[b*(L-1)+l-1]) + pos[l])   (M.py:2418-2426)
int launch_image_embed_ln_pre(const float* grid, const float* cls, const float* pos, const float* w, const float* b,
                              float* x, int batch, int L, op16* xc, float* rec, const float* next_ln_w,
                              const float* next_ln_b, op16* next_h, cudaStream_t stream);
// Lateral adapter tail (M.py:1760-1777): x_out = ln_adapt(concat(2*cls, BN(dw3x3(grid(x))) + t))
int launch_adapter_fuse_ln(const float* x, const float* t, const float* dw_w9, const float* dw_bias, const float* w,
                           const float* b, float* x_out, int batch, int g, op16* xc, float* rec, const float* next_ln_w,
                           const float* next_ln_b, op16* next_h, cudaStream_t stream);
// out[r] = x[r] / ||x[r]|| (optional) as f32 and op16 copies; width E (<= 1024, multiple of 4)
int check_token_error(cudaStream_t stream);  // synchronises; non-zero if an out-of-range token id was seen
int launch_l2norm(const float* x, float* out_f32, emb16* out_f16, int rows, int E, int normalise, cudaStream_t stream);

// zero-shot head (tools/zero_shot.py:125-131, 150-163): weights[c] = normalise(mean over templates of feat[row_of[c*T + t]])
// (row_of null = identity); top-k (k <= 8) column indices of every row of logits [rows, n], best first
int launch_class_mean_renorm(const float* feat, const int* row_of, int n_classes, int n_templates, int E, float* weights,
                             cudaStream_t stream);
int launch_topk_rows(const float* logits, int rows, int n, int k, int* out, cudaStream_t stream);

// ---- conv helpers (conv.cu) -------------------------------------------------------------------
// NCHW fp32 image -> im2col rows [B*Ho*Wo, 32] op16 for the 3x3 stride-2 pad-1 first convs; k = c*9+ky*3+kx
int launch_im2col_first(const void* img, int img_dtype, op16* out, int batch, int H, int W, cudaStream_t stream);
// NHWC op16 (pixel pitch cpix, channel offset c_off, C channels, C % 8 == 0) -> rows [B*Ho*Wo] of
// ksize*ksize*C columns (k = (ky*ksize+kx)*C + c) written at out[row*out_ld + out_off + k]
int launch_im2col_nhwc(const op16* in, int batch, int H, int W, int cpix, int c_off, int C, int ksize, int stride,
                       int pad, op16* out, int64_t out_ld, int out_off, cudaStream_t stream);
// depth-wise k x k / stride k patch pooling + folded BN (Lateral_Adapter top2bottom_dw_conv, M.py:1756)
// in NHWC op16 [B,H,W,cpix] (+c_off), w [k*k][C] f32, bias [C] f32 -> out [B*(H/k)*(W/k), C] op16
int launch_patch_pool(const op16* in, int batch, int H, int W, int cpix, int c_off, int C, int k, const float* w,
                      const float* bias, op16* out, cudaStream_t stream);

// ---- fused front end (front.cu): both 3x3 / stride-2 first convs (stem | parallel branch, w0 [96][32] with
// k = c*9 + ky*3 + kx, BN folded), the branch's first bottleneck 1x1 (w1 [48][48]), the even-pixel copy of the
// branch activation for its strided shortcut and the depth-wise k x k patch pooling of adapter 0, in one pass over
// the image.  NCHW image (f32 / bf16 / f16) -> stem, y1 NHWC op16 [B, H/2, W/2, 48]; p0s [B, H/4, W/4, 48] with pixel
// pitch p0s_pitch (96 lets it be the right half of the [y2 | p0s] operand of the ConvResBlock tail GEMM);
// pooled [B * (H/2/k) * (W/2/k), 48].
bool front_conv_supported(int H, int W, int c0, int k);
int launch_front_conv(const void* img, int img_dtype, int batch, int H, int W, const op16* w0, const float* b0,
                      const op16* w1, const float* b1, const float* pool_w, const float* pool_b, int k, op16* stem,
                      op16* y1, op16* p0s, int p0s_pitch, op16* pooled, cudaStream_t stream);

// ---- implicit-GEMM convolution (conv_gemm.cu): out[B*Ho*Wo, N] = epi(patches . W^T + bias) ---------------
// K is the concatenation of the sources' (ky, kx, c) patch vectors; every source must map onto the same
// Ho x Wo output grid.  NHWC op16 inputs with pixel pitch cpix, channel window [c_off, c_off + C).
struct ConvSource {
  const void* in;
  int H, W, cpix, c_off, C, ksize, stride, pad;
};
int launch_conv_gemm(const ConvSource* src, int nsrc, int batch, int Ho, int Wo, const op16* W, int64_t ldw, int N,
                     const float* bias, void* out, int64_t ldo, int epi, cudaStream_t stream);

// The same convolution with the A operand fetched by im2col-mode TMA (gemm.cu, gemm_tcgen05_kernel<..., CONV = 1>): no
// gather warps, one TMA instruction per filter tap x 64-channel block.  Wp is the weight in the padded K layout
// [N][source][ky][kx][64-channel block][64] (conv_tma_kpad columns; launch_pack_conv_kpad builds it from the dense
// (ky, kx, c) layout); N a multiple of 48; epi EPI_RELU_BF16 or EPI_F32.
int conv_tma_kpad(const ConvSource* src, int nsrc);
int launch_conv_tma(const ConvSource* src, int nsrc, int batch, int Ho, int Wo, const op16* Wp, int64_t ldw, int N,
                    const float* bias, void* out, int64_t ldo, int epi, cudaStream_t stream);
// dense [N][sum_src ksize^2 * C] op16 -> padded layout [N][conv_tma_kpad] (zeros in the padding columns)
int launch_pack_conv_kpad(const op16* dense, int64_t ldd, const ConvSource* src, int nsrc, int N, op16* padded,
                          cudaStream_t stream);

// ---- attention (attention.cu): qkv op16 [B*L, 3*768] (q pre-scaled) -> out op16 [B*L, 768] ------
int launch_attention(const op16* qkv, op16* out, int batch, int L, int heads, int causal, cudaStream_t stream);

// ---- contrastive loss (loss.cu) ---------------------------------------------------------------
// Fused similarity + log-sum-exp for the local rows against all G = world*b_local columns, both
// directions; the column shards are read in-kernel from their owners (local or NVLink peer pointers).
// loss_parts[0] = sum_i (lse_i - s_ii) over local image rows, [1] = same over local text rows;
// loss = (sum over ranks of parts[0] + parts[1]) / (2 G).  img_shards / txt_shards: device arrays of
// `world` rank-ordered shard pointers; flags: this rank's device array of `world` publish flags or null.
int launch_contrastive_loss_ex(const emb16* img_local, const emb16* txt_local, const emb16* const* img_shards,
                               const emb16* const* txt_shards, const uint32_t* flags, uint32_t epoch, int world,
                               int rank, int b_local, int E, float scale, void* workspace, float* loss_parts,
                               float* lse_out, int lse_pitch, cudaStream_t stream);
size_t contrastive_loss_workspace_bytes(int world, int b_local);
// Backward of that loss with respect to this rank's normalised embeddings (only the local shard receives a gradient,
// like gather_tensors, comm.py:151-152).  row_lse: the [2][lse_pitch] log2-domain lse the forward wrote (lse_out);
// img_lse_shards / txt_lse_shards: device tables of `world` pointers to every rank's image-row / text-row lse (peer
// pointers allowed).  d_img / d_txt: [b_local, 512] f32.
int launch_contrastive_loss_backward(const emb16* img_local, const emb16* txt_local, const emb16* const* img_shards,
                                     const emb16* const* txt_shards, const float* row_lse, int lse_pitch,
                                     const float* const* img_lse_shards, const float* const* txt_lse_shards, int world,
                                     int rank, int b_local, float scale, void* workspace, float* d_img, float* d_txt,
                                     cudaStream_t stream);
size_t contrastive_backward_workspace_bytes(int world, int b_local);
// ---- backward pass: wgrad.cu, attention_bwd.cu, backward.cu ----------------------------
// dW [N, K] f32 (+)= dY[tokens, N]^T . X[tokens, K]  (tcgen05 with MN-major operands, token range split S ways,
// deterministic reduction of the S partial matrices).  N % 128 == 0, K % 256 == 0.
int launch_wgrad(const op16* dY, int64_t ldy, const op16* X, int64_t ldx, int tokens, int N, int K, float* dW, int accumulate,
                 void* workspace, cudaStream_t stream);
size_t wgrad_workspace_bytes(int tokens, int N, int K);
int wgrad_pick_splits(int tokens, int N, int K);
void wgrad_set_desc(uint32_t lbo, uint32_t sbo);  // bring-up knob (0 = default)
// qkv [B*L, 3*64*heads] (q pre-scaled), dctx [B*L, 64*heads] -> dqkv = gradient of the unscaled (q | k | v); L <= 80
int launch_attention_bwd(const op16* qkv, const op16* dctx, op16* dqkv, int batch, int L, int heads, int causal,
                         cudaStream_t stream);
// Row kernels write per-CTA partial column sums ([parts][slots * 768] / [parts][width]); launch_reduce_partials folds them:
// dst[j] (+)= (j < head_n ? head_scale : 1) * sum_p part[p * pitch + j]
int bwd_row_parts(long long rows);   // CTAs (= partial rows) of the 768-wide row kernels below
int bwd_slab_parts(long long rows);  // partial rows of the 16-bit slab kernels (qgelu_bwd, colsum16)
int launch_reduce_partials(const float* part, int nparts, long long pitch, float* dst, long long n, int accumulate,
                           long long head_n, float head_scale, cudaStream_t stream);
int launch_reduce_partials3(const float* part, int nparts, long long pitch, float* d0, float* d1, float* d2, int n, cudaStream_t stream);
// dx (+)= LN_backward(dy; x, gamma); g16 (optional) = op16(dx); part [bwd_row_parts][3 * 768] = dgamma | dbeta | colsum(dx)
int launch_ln_bwd(const float* x, const float* dy, const float* gamma, float* dx, op16* g16, float* part, long long rows,
                  int accumulate, cudaStream_t stream);
// g16 = op16(dx); part [bwd_row_parts][768] = colsum(dx)
int launch_cast_colsum(const float* dx, op16* g16, float* part, long long rows, cudaStream_t stream);
int launch_qgelu_fwd(const op16* u, op16* a, long long n, cudaStream_t stream);
// da <- da * quickgelu'(u); part [bwd_slab_parts][width] = colsum
int launch_qgelu_bwd(op16* da, const op16* u, float* part, long long rows, int width, cudaStream_t stream);
int launch_colsum16(const op16* g, float* part, long long rows, int width, cudaStream_t stream);
int launch_l2norm_bwd(const float* g, const float* df, float* dg, op16* dg16, int rows, int E, int normalise, cudaStream_t stream);
// pooled rows (tok != null: EOT position of tok [batch, Ltok]; null: row 0 = CLS): dx[b * Lx + pos_b] = LN_backward(dz[b]);
// part [bwd_row_parts(batch)][2 * 768] = dgamma | dbeta
int launch_pooled_ln_bwd(const float* x, int Lx, const int64_t* tok, int Ltok, const float* dz, const float* gamma, float* dx,
                         float* part, int batch, cudaStream_t stream);
int launch_text_embed_bwd(const float* dx, const int64_t* tok, int Ltok, int L, int batch, int vocab, float* dpos, float* demb,
                          cudaStream_t stream);
// dx <- LN_backward(dx; e) with e recomputed from (grid, cls, pos); part [bwd_row_parts][2 * 768]; dpos / dcls accumulate
int launch_image_embed_bwd(const float* grid, const float* cls, const float* pos, const float* gamma, float* dx, float* part,
                           int batch, int L, float* dpos, float* dcls, cudaStream_t stream);
// lateral adapter, bottom path: dxo (gradient of the adapter output, overwritten with ds = gradient of the pre-LayerNorm sum,
// whose grid rows are also the gradient of the top path's t) -> dx = gradient of the adapter input; part [..][2 * 768]
int launch_adapter_bwd(const float* x, const float* t, const float* w9, const float* bias, const float* gamma, float* dxo,
                       float* dx, float* part, int batch, int gsz, cudaStream_t stream);
// fused multi-tensor AdamW (torch.optim.AdamW semantics)
struct AdamwTensor {
  float* param;
  const float* grad;
  float* m;
  float* v;
  long long numel;
  float lr, wd;
};
int launch_adamw(const AdamwTensor* tab_dev, const int* chunk_tensor_dev, const long long* chunk_off_dev, int nchunks, int chunk,
                 float beta1, float beta2, float eps, int step, cudaStream_t stream);

// ---- weight packing (pack.cu) -------------------------------------------------------------------
// dst[n, k] (op16, pitch ldd) = src[n*sn + k*sk] * (row_scale ? row_scale[n] : 1)
int launch_pack_op16(const float* src, int64_t sn, int64_t sk, const float* row_scale, op16* dst, int64_t ldd, int N,
                     int K, cudaStream_t stream);

}  // namespace msclip
