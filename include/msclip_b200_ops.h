/* msclip_b200 — single-kernel entry points of the C ABI (device pointers only, stream-ordered).
 *
 * These expose the individual sm_100a kernels behind the model-level calls of msclip_b200.h so that each
 * one can be parity-tested against the operation of the reference it replaces
 * (lib/models/clip_openai_pe_res_v1.py = "M.py"):
 *
 *   msclip_op_gemm            F.linear + fused epilogue             M.py:612, 747, 794-798, 2690, 3074
 *   msclip_op_layernorm       LayerNorm.forward                     M.py:204-219
 *   msclip_op_gemm_ln         LayerNorm folded into QKV / fc1 and    M.py:204-219 + 612, 747, 794-798
 *                             emitted by out-proj / fc2
 *   msclip_op_attention       Attention_CUST.forward core           M.py:707-738
 *   msclip_op_im2col_first    gather for the 3x3/s2 first convs     M.py:1952, 2154 (EarlyconvRes / branch)
 *   msclip_op_im2col_nhwc     gather for the later convs            M.py:1920-1936, 1842-1861
 *   msclip_op_conv_gemm       conv3x3 / strided conv1x1 (+BN+ReLU)  M.py:1920-1936, 1842-1861 (implicit GEMM)
 *   msclip_op_patch_pool      Lateral_Adapter.top2bottom_dw_conv    M.py:1756
 *   msclip_op_front_conv      first convs of stem + branch, branch   M.py:1993, 2260-2273, 1842-1846, 1857, 1756
 *                             bottleneck entry, adapter-0 pooling (one fused pass over the image)
 *   msclip_op_adapter_fuse_ln Lateral_Adapter tail                  M.py:1760-1777
 *   msclip_op_contrastive_lse similarity + symmetric CE partials    M.py:3141 + north-star loss
 *   msclip_op_wgrad / _attention_bwd / _layernorm_bwd / _qgelu_bwd   backward of the ops above (section 8f-1)
 *   msclip_op_adamw           fused multi-tensor AdamW              experiments/model/b32.yaml:32-53
 *   msclip_num_keys / msclip_key_info   the state-dict contract     SURVEY.md section 8c
 */
#ifndef MSCLIP_B200_OPS_H_
#define MSCLIP_B200_OPS_H_

#include "msclip_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* epilogue codes of msclip_op_gemm */
enum {
  MSCLIP_EPI_BF16 = 0,       /* bf16 out = alpha*acc + bias */
  MSCLIP_EPI_QGELU_BF16 = 1, /* bf16 out = quickgelu(alpha*acc + bias) */
  MSCLIP_EPI_RELU_BF16 = 2,  /* bf16 out = relu(alpha*acc + bias) */
  MSCLIP_EPI_RESID_F32 = 3,  /* f32 out = resid + alpha*acc + bias (resid may alias out) */
  MSCLIP_EPI_F32 = 4         /* f32 out = alpha*acc + bias */
};

/* out[m,n] = epi(alpha * a[m,k] . w[n,k]^T + bias[n]); a, w bf16 with row pitches lda, ldw (elements). */
int msclip_op_gemm(const void* a, int64_t lda, const void* w, int64_t ldw, int m, int n, int k, float alpha,
                   const float* bias, void* out, int64_t ldo, const float* resid, int64_t ldr, int epilogue,
                   void* stream);
/* Tiling of the large GEMMs (N % 256 == 0): 0 = one CTA per 128x256 tile, 1 = CTA pairs (cta_group::2, 256x256),
 * 2 / 4 = clusters of 2 / 4 pairs that share the weight tile by TMA multicast; any other value = library default. */
void msclip_op_set_gemm_pair_mode(int mode);
/* LayerNorm (M.py:204-219) folded into the GEMMs around it: LN(x) . W^T + b = rstd * (xc . W'^T - mean_c * colsum) + b'
 * with xc = x - shift, W' = W diag(gamma), b' = b + W . beta.
 * Row record = 16 floats: [0] shift, [4 + 2s] / [5 + 2s] = sum / sum of squares of the centred row over 128-column slice s.
 *   ln_mode 1 (epilogue BF16 or QGELU_BF16): a = bf16(x - shift) [m,k], w = W', bias = b', colsum[n]; records ln_in.
 *   ln_mode 2 (epilogue RESID_F32, n = 768): out = resid + a . w^T + bias as msclip_op_gemm, and additionally
 *     out16 = bf16(out - shift') with shift' = row mean of resid (from ln_in), records of out -> ln_out (!= ln_in). */
int msclip_op_gemm_ln(const void* a, int64_t lda, const void* w, int64_t ldw, int m, int n, int k, const float* bias, void* out,
                      int64_t ldo, const float* resid, int64_t ldr, int epilogue, int ln_mode, const float* ln_in,
                      float* ln_out, void* out16, int64_t ldo16, const float* colsum, void* stream);
/* Residual GEMM + LayerNorm of the updated rows in ONE kernel (out-proj + ln_2, fc2 + the next block's ln_1,
 * M.py:1027-1028): x[m, 768] (f32, in place) += a . w^T + bias, then h_out = LayerNorm(x) * gamma + beta (16-bit), written by
 * four extra warps per CTA that re-read each finished 128-row block from L2.  n must be 768, m >= 256.
 * h_out is bit-identical to msclip_op_layernorm applied to the updated x.  `counters`: msclip_op_gemm_resid_ln_counters(m)
 * zero-initialised uint32 words of device memory (left zero by every call). */
int msclip_op_gemm_resid_ln(const void* a, int64_t lda, const void* w, int64_t ldw, int m, int n, int k, const float* bias, float* x,
                            int64_t ldx, const float* gamma, const float* beta, void* h_out, int64_t ldh, void* counters,
                            void* stream);
size_t msclip_op_gemm_resid_ln_counters(int m);

/* w_out[n,k] = bf16(w[n,k] * row_scale[n] * gamma[k]); colsum[n] = sum_k w_out[n,k]; bias_out[n] = row_scale[n] *
 * (bias[n] + sum_k w[n,k] * beta[k]); row_scale / bias may be NULL. */
int msclip_op_pack_ln_fold(const float* w, const float* row_scale, const float* gamma, const float* beta, const float* bias,
                           void* w_out_bf16, float* colsum, float* bias_out, int n, int k, void* stream);
/* y[r,:] (bf16) = LN(x[r*row_stride,:]) over 768 columns, eps 1e-12 inside the sqrt. */
int msclip_op_layernorm(const float* x, int row_stride, const float* w, const float* b, void* y_bf16, int rows,
                        void* stream);
/* qkv bf16 [batch*seq_len, 3*64*heads] (q already scaled) -> out bf16 [batch*seq_len, 64*heads]. */
int msclip_op_attention(const void* qkv_bf16, void* out_bf16, int batch, int seq_len, int heads, int causal,
                        void* stream);
int msclip_op_im2col_first(const void* img, int dtype, void* out_bf16, int batch, int height, int width, void* stream);
int msclip_op_im2col_nhwc(const void* in_bf16, int batch, int height, int width, int cpix, int c_off, int channels,
                          int ksize, int stride, int pad, void* out_bf16, int64_t out_ld, int out_off, void* stream);
/* msclip_op_conv_gemm with the A operand fetched by im2col-mode TMA (no gather warps): same arguments and the same
 * dense (ky, kx, c) weight; `w_padded_scratch` (n * msclip_op_conv_tma_kpad(c0, k0, c1, k1) 16-bit elements) receives the
 * padded-K weight copy the kernel consumes.  n a multiple of 48; epilogue EPI_RELU_BF16 or EPI_F32. */
int msclip_op_conv_tma(const void* in0, int h0, int w0, int cpix0, int coff0, int c0, int k0, int s0, int p0, const void* in1,
                       int h1, int w1, int cpix1, int coff1, int c1, int k1, int s1, int p1, int batch, int ho, int wo,
                       const void* w_dense, int64_t ldw, int n, const float* bias, void* out, int64_t ldo, int epilogue,
                       void* w_padded_scratch, void* stream);
int msclip_op_conv_tma_kpad(int c0, int k0, int c1, int k1);

/* Implicit-GEMM convolution: out[batch*ho*wo, n] = epi(patches . w^T + bias); K = (ky,kx,c) patch of source 0
 * followed by that of the optional source 1 (in1 = NULL: single source).  NHWC bf16 inputs. */
int msclip_op_conv_gemm(const void* in0, int h0, int w0, int cpix0, int coff0, int c0, int k0, int s0, int p0, const void* in1,
                        int h1, int w1, int cpix1, int coff1, int c1, int k1, int s1, int p1, int batch, int ho, int wo,
                        const void* w_bf16, int64_t ldw, int n, const float* bias, void* out, int64_t ldo, int epilogue,
                        void* stream);
int msclip_op_patch_pool(const void* in_bf16, int batch, int height, int width, int cpix, int c_off, int channels,
                         int k, const float* w, const float* bias, void* out_bf16, void* stream);
/* Fused 112x112 stage: NCHW image (dtype code as msclip_encode_image) -> stem = relu(conv3x3_s2 . w0[0:48]),
 * p0 = relu(conv3x3_s2 . w0[48:96]) (kept on chip), y1 = relu(w1 . p0 + b1), p0s = p0[:, ::2, ::2],
 * pooled = depth-wise k x k / stride k pooling of p0 (+ pool_b).  w0 bf16 [96][32] (k = c*9 + ky*3 + kx, zero padded),
 * w1 bf16 [48][48], pool_w f32 [k*k][48]; outputs NHWC bf16 (p0s with a pixel pitch of p0s_pitch >= 48 elements);
 * height, width multiples of 32; k = 8 or 16. */
int msclip_op_front_conv(const void* img, int dtype, int batch, int height, int width, const void* w0_bf16, const float* b0,
                         const void* w1_bf16, const float* b1, const float* pool_w, const float* pool_b, int k,
                         void* stem_bf16, void* y1_bf16, void* p0s_bf16, int p0s_pitch, void* pooled_bf16, void* stream);
int msclip_op_adapter_fuse_ln(const float* x, const float* t, const float* dw_w9, const float* dw_bias, const float* w,
                              const float* b, float* x_out, int batch, int grid, void* stream);
/* parts2[0] = sum_i (lse_j s_ij - s_ii), parts2[1] = sum_j (lse_i s_ij - s_jj), s = scale * img . txt^T */
int msclip_op_contrastive_lse(const void* img_f16, const void* txt_f16, int b, float scale, void* workspace,
                              float* parts2, void* stream);
size_t msclip_op_contrastive_lse_workspace(int b);

/* ---- backward kernels (oracle = torch.autograd of the forward op) -------------------------------
 * dw[n, k] (f32, dense) (+)= sum_t dy[t, n] * x[t, k]: weight gradient of F.linear (M.py:612, 747, 794-798) by tcgen05 with
 * MN-major operands (no transposes); n % 128 == 0, k % 256 == 0; workspace = msclip_op_wgrad_workspace(tokens, n, k) bytes. */
int msclip_op_wgrad(const void* dy_bf16, int64_t ldy, const void* x_bf16, int64_t ldx, int tokens, int n, int k, float* dw,
                    int accumulate, void* workspace, void* stream);
size_t msclip_op_wgrad_workspace(int tokens, int n, int k);
void msclip_op_set_wgrad_desc(unsigned lbo_bytes, unsigned sbo_bytes); /* bring-up knob; 0 = default */
/* backward of msclip_op_attention: dqkv = gradient of the UNSCALED (q | k | v) given qkv (q pre-scaled) and the gradient of
 * the attention output; seq_len <= 80 */
int msclip_op_attention_bwd(const void* qkv_bf16, const void* dctx_bf16, void* dqkv_bf16, int batch, int seq_len, int heads,
                            int causal, void* stream);
/* dx (+)= LayerNorm_backward(dy; x, gamma) over 768 columns; dx16 (optional) = bf16(dx after the update); dgamma / dbeta /
 * dcolsum [768] (+)= sum_rows dy*xhat / dy / updated dx (each may be NULL); workspace = msclip_op_bwd_workspace(rows) bytes */
int msclip_op_layernorm_bwd(const float* x, const float* dy, const float* gamma, float* dx, void* dx16, float* dgamma, float* dbeta,
                            float* dcolsum, int rows, int accumulate, void* workspace, void* stream);
/* da <- da * quickgelu'(u) (bf16, in place); dbias [width] (+)= column sums of the result; width % 256 == 0 */
int msclip_op_qgelu_bwd(void* da_bf16, const void* u_bf16, float* dbias, int rows, int width, void* workspace, void* stream);
size_t msclip_op_bwd_workspace(int rows);
/* Fused multi-tensor AdamW (torch.optim.AdamW semantics; optimiser of experiments/model/b32.yaml:32-53): host arrays of n
 * device pointers / element counts / per-tensor lr and weight decay; `step` >= 1 (bias correction). */
int msclip_op_adamw(int n, float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                    const int64_t* numel, const float* lr, const float* weight_decay, float beta1, float beta2, float eps,
                    int step, void* stream);

/* state-dict contract of a handle: number of keys, and key / rank / shape (up to 4 dims) of entry i. */
int msclip_num_keys(msclip_handle h);
int msclip_key_info(msclip_handle h, int index, const char** key, int* ndim, int64_t* shape4);

#ifdef __cplusplus
}
#endif
#endif /* MSCLIP_B200_OPS_H_ */
