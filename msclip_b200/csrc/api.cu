// extern "C" surface declared in include/msclip_b200.h and include/msclip_b200_ops.h.
#include "engine.h"

#include <new>
#include <cstring>
#include <vector>

#include "../../include/msclip_b200_ops.h"
#include "common.cuh"

using namespace msclip;

static cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

const char* msclip_version(void) { return "msclip_b200 0.1.0 (sm_100a, " MSCLIP_OPERAND_NAME " operands)"; }
const char* msclip_last_error(void) { return get_last_error(); }

int msclip_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ++ok;
  }
  return ok;
}

int msclip_create(const msclip_config* cfg, msclip_handle* out) {
  MSCLIP_REQUIRE(cfg != nullptr && out != nullptr, "msclip_create: null argument");
  MSCLIP_REQUIRE(cfg->patch_size == 16 || cfg->patch_size == 32, "MS-CLIP-S envelope: patch_size must be 16 or 32");
  MSCLIP_REQUIRE(cfg->width == 768, "MS-CLIP-S envelope: width must be 768 (12 heads of 64)");
  MSCLIP_REQUIRE(cfg->embed_dim == 512, "MS-CLIP-S envelope: embed_dim must be 512");
  MSCLIP_REQUIRE(cfg->layers >= 1 && cfg->layers <= 64, "layers out of range");
  MSCLIP_REQUIRE(cfg->image_resolution == 224, "MS-CLIP-S envelope: image_resolution must be 224");
  MSCLIP_REQUIRE(cfg->context_length >= 1 && cfg->context_length <= 208, "context_length must be in [1, 208]");
  MSCLIP_REQUIRE(cfg->vocab_size >= 2, "vocab_size out of range");
  MSCLIP_REQUIRE(cfg->parallel_strides[0] == 2, "the first parallel-branch conv must have stride 2 (it shares the stem's im2col)");
  int g = cfg->image_resolution / 2;
  for (int i = 0; i < 4; ++i) {
    MSCLIP_REQUIRE(cfg->early_strides[i] == 1 || cfg->early_strides[i] == 2, "early-conv strides must be 1 or 2");
    g /= cfg->early_strides[i];
  }
  MSCLIP_REQUIRE(g == cfg->image_resolution / cfg->patch_size, "early-conv strides do not produce the patch grid");
  int r = cfg->image_resolution;
  for (int j = 0; j < 5; ++j) {
    MSCLIP_REQUIRE(cfg->parallel_strides[j] == 1 || cfg->parallel_strides[j] == 2, "parallel strides must be 1 or 2");
    r /= cfg->parallel_strides[j];
    MSCLIP_REQUIRE(cfg->t2b_kernels[j] >= 1 && r % cfg->t2b_kernels[j] == 0 && r / cfg->t2b_kernels[j] == g,
                   "lateral adapter kernel does not land on the token grid");
  }
  msclip_ctx* h = new (std::nothrow) msclip_ctx();
  MSCLIP_REQUIRE(h != nullptr, "out of host memory");
  h->cfg = *cfg;
  h->grid = g;
  h->l_img = g * g + 1;
  h->heads = cfg->width / 64;
  build_spec(h);
  *out = h;
  return 0;
}

int msclip_destroy(msclip_handle h) {
  delete h;
  return 0;
}

// ---- backward kernels ----------------------------------------------------------------------------------------------
int msclip_op_wgrad(const void* dy, int64_t ldy, const void* x, int64_t ldx, int tokens, int n, int k, float* dw, int accumulate,
                    void* workspace, void* stream) {
  return launch_wgrad(static_cast<const op16*>(dy), ldy, static_cast<const op16*>(x), ldx, tokens, n, k, dw, accumulate, workspace,
                      as_stream(stream));
}
size_t msclip_op_wgrad_workspace(int tokens, int n, int k) { return wgrad_workspace_bytes(tokens, n, k); }
void msclip_op_set_wgrad_desc(unsigned lbo, unsigned sbo) { wgrad_set_desc(lbo, sbo); }
int msclip_op_attention_bwd(const void* qkv, const void* dctx, void* dqkv, int batch, int seq_len, int heads, int causal,
                            void* stream) {
  return launch_attention_bwd(static_cast<const op16*>(qkv), static_cast<const op16*>(dctx), static_cast<op16*>(dqkv), batch, seq_len,
                              heads, causal, as_stream(stream));
}
size_t msclip_op_bwd_workspace(int rows) {
  const size_t a = static_cast<size_t>(bwd_row_parts(rows)) * 3 * 768, b = static_cast<size_t>(bwd_slab_parts(rows)) * 4 * 768;
  return (a > b ? a : b) * sizeof(float);
}
int msclip_op_layernorm_bwd(const float* x, const float* dy, const float* gamma, float* dx, void* dx16, float* dgamma, float* dbeta,
                            float* dcolsum, int rows, int accumulate, void* workspace, void* stream) {
  MSCLIP_REQUIRE(workspace != nullptr && rows > 0, "msclip_op_layernorm_bwd: bad arguments");
  float* part = static_cast<float*>(workspace);
  cudaStream_t s = as_stream(stream);
  MSCLIP_TRY(launch_ln_bwd(x, dy, gamma, dx, static_cast<op16*>(dx16), part, rows, accumulate, s));
  const int rp = bwd_row_parts(rows);
  float* dst[3] = {dgamma, dbeta, dcolsum};
  for (int i = 0; i < 3; ++i)
    if (dst[i]) MSCLIP_TRY(launch_reduce_partials(part + i * 768, rp, 3 * 768, dst[i], 768, 1, 0, 1.0f, s));
  return 0;
}
int msclip_op_qgelu_bwd(void* da, const void* u, float* dbias, int rows, int width, void* workspace, void* stream) {
  MSCLIP_REQUIRE(workspace != nullptr && rows > 0 && width <= 4 * 768, "msclip_op_qgelu_bwd: bad arguments");
  float* part = static_cast<float*>(workspace);
  cudaStream_t s = as_stream(stream);
  MSCLIP_TRY(launch_qgelu_bwd(static_cast<op16*>(da), static_cast<const op16*>(u), part, rows, width, s));
  if (dbias) MSCLIP_TRY(launch_reduce_partials(part, bwd_slab_parts(rows), width, dbias, width, 1, 0, 1.0f, s));
  return 0;
}
// Tensor table + chunk map of the last calls, kept on the device: two slots (pinned staging + device copy each).  A call
// whose table equals the one already resident (same tensors, same lr / weight decay - every step of a constant-lr run)
// uploads nothing; otherwise the other slot is refilled once the kernel that last read it has finished.  No stream
// synchronisation, no allocation in the steady state: the optimiser step never stalls the host.
namespace {
struct AdamwSlot {
  uint8_t* pinned = nullptr;
  uint8_t* dev = nullptr;
  size_t cap = 0;
  std::vector<uint8_t> content;
  cudaEvent_t used = nullptr;
};
AdamwSlot g_adamw_slot[2];
int g_adamw_cur = 0;
}  // namespace

int msclip_op_adamw(int n, float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                    const int64_t* numel, const float* lr, const float* weight_decay, float beta1, float beta2, float eps, int step,
                    void* stream) {
  MSCLIP_REQUIRE(n >= 0 && step >= 1, "msclip_op_adamw: bad arguments");
  if (n == 0) return 0;
  constexpr int kChunk = 65536;  // elements per work item
  std::vector<AdamwTensor> tab(n);
  std::vector<int> ct;
  std::vector<long long> co;
  for (int i = 0; i < n; ++i) {
    tab[i] = {params[i], grads[i], exp_avg[i], exp_avg_sq[i], static_cast<long long>(numel[i]), lr[i], weight_decay[i]};
    for (long long o = 0; o < numel[i]; o += kChunk) {
      ct.push_back(i);
      co.push_back(o);
    }
  }
  const size_t b0 = tab.size() * sizeof(AdamwTensor), b1 = ct.size() * sizeof(int), b2 = co.size() * sizeof(long long);
  const size_t o1 = (b0 + 15) & ~size_t(15), o2 = (o1 + b1 + 15) & ~size_t(15);
  std::vector<uint8_t> host(o2 + b2, 0);
  memcpy(host.data(), tab.data(), b0);
  memcpy(host.data() + o1, ct.data(), b1);
  memcpy(host.data() + o2, co.data(), b2);
  cudaStream_t s = as_stream(stream);
  AdamwSlot* slot = &g_adamw_slot[g_adamw_cur];
  if (slot->content != host) {
    g_adamw_cur ^= 1;
    slot = &g_adamw_slot[g_adamw_cur];
    if (slot->used) MSCLIP_CHECK_CUDA(cudaEventSynchronize(slot->used));  // last reader of this slot (two steps ago)
    if (slot->cap < host.size()) {
      if (slot->pinned) cudaFreeHost(slot->pinned);
      if (slot->dev) cudaFree(slot->dev);
      slot->cap = host.size() * 2;
      MSCLIP_CHECK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&slot->pinned), slot->cap, cudaHostAllocDefault));
      MSCLIP_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&slot->dev), slot->cap));
    }
    if (!slot->used) MSCLIP_CHECK_CUDA(cudaEventCreateWithFlags(&slot->used, cudaEventDisableTiming));
    memcpy(slot->pinned, host.data(), host.size());
    MSCLIP_CHECK_CUDA(cudaMemcpyAsync(slot->dev, slot->pinned, host.size(), cudaMemcpyHostToDevice, s));
    slot->content = host;
  }
  const uint8_t* dev = slot->dev;
  MSCLIP_TRY(launch_adamw(reinterpret_cast<const AdamwTensor*>(dev), reinterpret_cast<const int*>(dev + o1),
                          reinterpret_cast<const long long*>(dev + o2), static_cast<int>(ct.size()), kChunk, beta1, beta2, eps, step, s));
  MSCLIP_CHECK_CUDA(cudaEventRecord(slot->used, s));
  return 0;
}

int msclip_num_keys(msclip_handle h) { return h ? static_cast<int>(h->spec.size()) : -1; }

int msclip_key_info(msclip_handle h, int index, const char** key, int* ndim, int64_t* shape4) {
  MSCLIP_REQUIRE(h != nullptr && index >= 0 && index < static_cast<int>(h->spec.size()), "msclip_key_info: bad index");
  auto it = h->spec.begin();
  std::advance(it, index);
  *key = it->first.c_str();
  *ndim = static_cast<int>(it->second.size());
  for (size_t i = 0; i < it->second.size() && i < 4; ++i) shape4[i] = it->second[i];
  return 0;
}

int msclip_set_weight(msclip_handle h, const char* key, const void* data, int dtype, int ndim, const int64_t* shape) {
  MSCLIP_REQUIRE(h != nullptr && key != nullptr, "msclip_set_weight: null argument");
  auto it = h->spec.find(key);
  MSCLIP_REQUIRE(it != h->spec.end(), std::string("unexpected state-dict key: ") + key);
  MSCLIP_REQUIRE(ndim == static_cast<int>(it->second.size()), std::string("rank mismatch for ") + key);
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    MSCLIP_REQUIRE(shape[i] == it->second[i], std::string("size mismatch for ") + key);
    n *= static_cast<size_t>(shape[i]);
  }
  const std::string k(key);
  const bool counter = k.size() > 19 && k.compare(k.size() - 19, 19, "num_batches_tracked") == 0;
  RawTensor& t = h->raw[k];
  if (t.dev) {
    cudaFree(t.dev);
    t.dev = nullptr;
  }
  t.shape = it->second;
  t.numel = counter ? 0 : n;
  h->finalized = false;
  if (counter) return 0;  // BatchNorm's step counter does not enter the eval-mode forward
  MSCLIP_REQUIRE(dtype == MSCLIP_F32, std::string("weights must be float32: ") + key);
  MSCLIP_REQUIRE(data != nullptr, std::string("null data for ") + key);
  MSCLIP_CHECK_CUDA(cudaMalloc(&t.dev, (n > 0 ? n : 1) * 4));
  MSCLIP_CHECK_CUDA(cudaMemcpy(t.dev, data, n * 4, cudaMemcpyDefault));
  return 0;
}

int msclip_finalize_weights(msclip_handle h, void* stream) {
  MSCLIP_REQUIRE(h != nullptr, "null handle");
  return engine_finalize(h, as_stream(stream));
}

int msclip_logit_scale_exp(msclip_handle h, float* out) {
  MSCLIP_REQUIRE(h != nullptr && out != nullptr && h->finalized, "msclip_logit_scale_exp: weights not finalized");
  *out = expf(current_logit_scale(h));
  return 0;
}

int msclip_encode_image(msclip_handle h, const void* image, int image_dtype, int batch, float* out, int normalize,
                        void* stream) {
  MSCLIP_REQUIRE(h != nullptr, "null handle");
  return engine_encode_image(h, image, image_dtype, batch, out, normalize, as_stream(stream));
}

int msclip_stage_images(msclip_handle h, const void* image_host, int image_dtype, int batch, void* stream) {
  MSCLIP_REQUIRE(h != nullptr, "null handle");
  return engine_stage_images(h, image_host, image_dtype, batch, as_stream(stream));
}

int msclip_encode_text(msclip_handle h, const int64_t* tokens, int batch, float* out, int normalize, void* stream) {
  MSCLIP_REQUIRE(h != nullptr, "null handle");
  return engine_encode_text(h, tokens, batch, out, normalize, as_stream(stream));
}

int msclip_set_text_trim(msclip_handle h, int enable) { return engine_set_text_trim(h, enable); }

int msclip_similarity_logits(msclip_handle h, const float* img_feat, int n_img, const float* txt_feat, int n_txt,
                             float scale, float* logits, void* stream) {
  return engine_similarity_logits(h, img_feat, n_img, txt_feat, n_txt, scale, logits, as_stream(stream));
}

int msclip_zeroshot_classifier(msclip_handle h, const int64_t* tokens, int n_classes, int n_templates, float* weights_out,
                               void* stream) {
  MSCLIP_REQUIRE(h != nullptr, "null handle");
  return engine_zeroshot_classifier(h, tokens, n_classes, n_templates, weights_out, as_stream(stream));
}
int msclip_zeroshot_predict(msclip_handle h, const float* img_feat, int n_img, const float* weights, int n_classes, float scale,
                            int topk, int32_t* topk_out, float* logits_out, void* stream) {
  MSCLIP_REQUIRE(h != nullptr, "null handle");
  return engine_zeroshot_predict(h, img_feat, n_img, weights, n_classes, scale, topk, topk_out, logits_out, as_stream(stream));
}

int msclip_forward(msclip_handle h, const void* image, int image_dtype, const int64_t* tokens, int batch,
                   float* logits, void* stream) {
  MSCLIP_REQUIRE(h != nullptr, "null handle");
  return engine_forward(h, image, image_dtype, tokens, batch, logits, as_stream(stream));
}

int msclip_comm_init(msclip_handle h, int rank, int world, int max_b_local) {
  MSCLIP_REQUIRE(h != nullptr, "null handle");
  return comm_init(h, rank, world, max_b_local);
}
int msclip_comm_export(msclip_handle h, void* handle_out_64) {
  MSCLIP_REQUIRE(h != nullptr && handle_out_64 != nullptr, "null argument");
  return comm_export(h, handle_out_64);
}
int msclip_comm_import(msclip_handle h, const void* handles) {
  MSCLIP_REQUIRE(h != nullptr && handles != nullptr, "null argument");
  return comm_import(h, handles);
}

int msclip_comm_buffer(msclip_handle h, void** base_out) {
  MSCLIP_REQUIRE(h != nullptr && base_out != nullptr && h->xchg != nullptr, "msclip_comm_buffer: call msclip_comm_init first");
  *base_out = h->xchg;
  return 0;
}
int msclip_comm_import_pointers(msclip_handle h, void* const* bases_world) {
  MSCLIP_REQUIRE(h != nullptr && bases_world != nullptr, "null argument");
  return comm_import_pointers(h, bases_world);
}

int msclip_contrastive_loss(msclip_handle h, int b_local, float scale, float* partial_out, float* loss_out,
                            void* stream) {
  MSCLIP_REQUIRE(h != nullptr, "null handle");
  return engine_contrastive_loss(h, b_local, scale, partial_out, loss_out, as_stream(stream));
}

int msclip_contrastive_loss_features(msclip_handle h, const float* img_feat, const float* txt_feat, int b_local, float scale,
                                     float* partial_out, float* loss_out, void* stream) {
  MSCLIP_REQUIRE(h != nullptr, "null handle");
  return engine_contrastive_loss_features(h, img_feat, txt_feat, b_local, scale, partial_out, loss_out, as_stream(stream));
}

int msclip_contrastive_loss_backward(msclip_handle h, float* d_img_feat, float* d_txt_feat, void* stream) {
  MSCLIP_REQUIRE(h != nullptr, "null handle");
  return engine_contrastive_loss_backward(h, d_img_feat, d_txt_feat, as_stream(stream));
}

int msclip_encode_pairs(msclip_handle h, const void* image, int image_dtype, const int64_t* tokens, int b_micro,
                        int row_offset, void* stream) {
  MSCLIP_REQUIRE(h != nullptr, "null handle");
  return engine_encode_pairs(h, image, image_dtype, tokens, b_micro, row_offset, as_stream(stream));
}

int msclip_forward_loss(msclip_handle h, const void* image, int image_dtype, const int64_t* tokens, int b_local,
                        float* partial_out, float* loss_out, void* stream) {
  MSCLIP_REQUIRE(h != nullptr, "null handle");
  return engine_forward_loss(h, image, image_dtype, tokens, b_local, partial_out, loss_out, as_stream(stream));
}

int msclip_preprocess_images(msclip_handle h, const uint8_t* pixels, const int64_t* offsets, const int* heights, const int* widths, int n,
                             int out_size, const float* mean3, const float* std3, void* out, int out_dtype, uint8_t* out_u8, void* stream) {
  return engine_preprocess(h, pixels, offsets, heights, widths, n, out_size, mean3, std3, out, out_dtype, out_u8, as_stream(stream));
}
int msclip_train_enable(msclip_handle h, int enable) { return train_enable(h, enable); }
int msclip_backward(msclip_handle h, const float* d_img_feat, const float* d_txt_feat, void* stream) {
  return engine_backward(h, d_img_feat, d_txt_feat, as_stream(stream));
}
int msclip_zero_grad(msclip_handle h, void* stream) { return engine_zero_grad(h, as_stream(stream)); }
int msclip_taped_features(msclip_handle h, int modality, float* out, int b, void* stream) {
  MSCLIP_REQUIRE(h != nullptr && out != nullptr && (modality == 0 || modality == 1), "msclip_taped_features: bad arguments");
  const msclip_ctx::TapeInfo& t = modality == 0 ? h->tape_img : h->tape_txt;
  MSCLIP_REQUIRE(t.valid && b == t.batch, "msclip_taped_features: no taped call with that batch size");
  const float* raw = static_cast<const float*>(tape_get(h, modality == 0 ? "v_feat" : "t_feat"));
  MSCLIP_REQUIRE(raw != nullptr, "msclip_taped_features: incomplete tape");
  return launch_l2norm(raw, out, nullptr, b, h->cfg.embed_dim, 1, as_stream(stream));
}
int msclip_num_grads(msclip_handle h) { return h ? static_cast<int>(h->grad_list.size()) : 0; }
int msclip_grad_info(msclip_handle h, int index, const char** key, float** dev_ptr, int64_t* numel) {
  MSCLIP_REQUIRE(h != nullptr && index >= 0 && index < static_cast<int>(h->grad_list.size()), "msclip_grad_info: index out of range");
  if (key) *key = h->grad_list[index].first.c_str();
  if (dev_ptr) *dev_ptr = h->grad_list[index].second;
  if (numel) *numel = h->grad_numel[index];
  return 0;
}
int msclip_update_weight(msclip_handle h, const char* key, const float* dev_ptr, void* stream) {
  return engine_update_weight(h, key, dev_ptr, as_stream(stream));
}

int64_t msclip_launch_count(msclip_handle) { return launch_count(); }
int64_t msclip_device_bytes(msclip_handle h) {
  return h ? static_cast<int64_t>(h->weight_bytes + h->ws_bytes + h->xchg_bytes) : 0;
}

// ---- single-kernel entry points (device pointers only) --------------------------------------------------
int msclip_op_gemm(const void* a, int64_t lda, const void* w, int64_t ldw, int m, int n, int k, float alpha,
                   const float* bias, void* out, int64_t ldo, const float* resid, int64_t ldr, int epilogue,
                   void* stream) {
  return launch_gemm_scaled(static_cast<const op16*>(a), lda, static_cast<const op16*>(w), ldw, m, n, k, alpha, bias,
                            out, ldo, resid, ldr, epilogue, as_stream(stream));
}
void msclip_op_set_gemm_pair_mode(int enable) { gemm_set_pair_mode(enable); }
int msclip_op_gemm_ln(const void* a, int64_t lda, const void* w, int64_t ldw, int m, int n, int k, const float* bias, void* out,
                      int64_t ldo, const float* resid, int64_t ldr, int epilogue, int ln_mode, const float* ln_in,
                      float* ln_out, void* out16, int64_t ldo16, const float* colsum, void* stream) {
  return launch_gemm_ln(static_cast<const op16*>(a), lda, static_cast<const op16*>(w), ldw, m, n, k, bias, out, ldo, resid, ldr,
                        epilogue, ln_mode, ln_in, ln_out, static_cast<op16*>(out16), ldo16, colsum, as_stream(stream));
}
int msclip_op_gemm_resid_ln(const void* a, int64_t lda, const void* w, int64_t ldw, int m, int n, int k, const float* bias, float* x,
                            int64_t ldx, const float* gamma, const float* beta, void* h_out, int64_t ldh, void* counters,
                            void* stream) {
  return launch_gemm_resid_ln(static_cast<const op16*>(a), lda, static_cast<const op16*>(w), ldw, m, n, k, bias, x, ldx, gamma, beta,
                              static_cast<op16*>(h_out), ldh, static_cast<uint32_t*>(counters), as_stream(stream));
}
size_t msclip_op_gemm_resid_ln_counters(int m) { return gemm_resid_ln_counters(m); }
int msclip_op_pack_ln_fold(const float* w, const float* row_scale, const float* gamma, const float* beta, const float* bias,
                           void* w_out_bf16, float* colsum, float* bias_out, int n, int k, void* stream) {
  return launch_pack_ln_fold(w, row_scale, gamma, beta, bias, static_cast<op16*>(w_out_bf16), colsum, bias_out, n, k,
                             as_stream(stream));
}
int msclip_op_layernorm(const float* x, int row_stride, const float* w, const float* b, void* y_bf16, int rows,
                        void* stream) {
  return launch_layernorm_op16(x, row_stride, w, b, static_cast<op16*>(y_bf16), rows, as_stream(stream));
}
int msclip_op_attention(const void* qkv_bf16, void* out_bf16, int batch, int seq_len, int heads, int causal,
                        void* stream) {
  return launch_attention(static_cast<const op16*>(qkv_bf16), static_cast<op16*>(out_bf16), batch, seq_len, heads,
                          causal, as_stream(stream));
}
int msclip_op_im2col_first(const void* img, int dtype, void* out_bf16, int batch, int height, int width, void* stream) {
  return launch_im2col_first(img, dtype, static_cast<op16*>(out_bf16), batch, height, width, as_stream(stream));
}
int msclip_op_im2col_nhwc(const void* in_bf16, int batch, int height, int width, int cpix, int c_off, int channels,
                          int ksize, int stride, int pad, void* out_bf16, int64_t out_ld, int out_off, void* stream) {
  return launch_im2col_nhwc(static_cast<const op16*>(in_bf16), batch, height, width, cpix, c_off, channels, ksize, stride,
                            pad, static_cast<op16*>(out_bf16), out_ld, out_off, as_stream(stream));
}
int msclip_op_conv_gemm(const void* in0, int h0, int w0, int cpix0, int coff0, int c0, int k0, int s0, int p0, const void* in1,
                        int h1, int w1, int cpix1, int coff1, int c1, int k1, int s1, int p1, int batch, int ho, int wo,
                        const void* w_bf16, int64_t ldw, int n, const float* bias, void* out, int64_t ldo, int epilogue,
                        void* stream) {
  ConvSource src[2] = {{in0, h0, w0, cpix0, coff0, c0, k0, s0, p0}, {in1, h1, w1, cpix1, coff1, c1, k1, s1, p1}};
  return launch_conv_gemm(src, in1 ? 2 : 1, batch, ho, wo, static_cast<const op16*>(w_bf16), ldw, n, bias, out, ldo, epilogue,
                          as_stream(stream));
}
int msclip_op_conv_tma(const void* in0, int h0, int w0, int cpix0, int coff0, int c0, int k0, int s0, int p0, const void* in1,
                       int h1, int w1, int cpix1, int coff1, int c1, int k1, int s1, int p1, int batch, int ho, int wo,
                       const void* w_dense, int64_t ldw, int n, const float* bias, void* out, int64_t ldo, int epilogue,
                       void* w_padded_scratch, void* stream) {
  // same arguments as msclip_op_conv_gemm (dense (ky, kx, c) weight); the padded-K copy the TMA kernel consumes is
  // built in the caller's scratch buffer of n * msclip_op_conv_tma_kpad(...) elements
  ConvSource src[2] = {{in0, h0, w0, cpix0, coff0, c0, k0, s0, p0}, {in1, h1, w1, cpix1, coff1, c1, k1, s1, p1}};
  const int nsrc = in1 ? 2 : 1;
  const int kp = conv_tma_kpad(src, nsrc);
  MSCLIP_TRY(launch_pack_conv_kpad(static_cast<const op16*>(w_dense), ldw, src, nsrc, n, static_cast<op16*>(w_padded_scratch),
                                   as_stream(stream)));
  return launch_conv_tma(src, nsrc, batch, ho, wo, static_cast<const op16*>(w_padded_scratch), kp, n, bias, out, ldo, epilogue,
                         as_stream(stream));
}
int msclip_op_conv_tma_kpad(int c0, int k0, int c1, int k1) {
  ConvSource src[2] = {{nullptr, 0, 0, 0, 0, c0, k0, 1, 0}, {nullptr, 0, 0, 0, 0, c1, k1, 1, 0}};
  return conv_tma_kpad(src, c1 > 0 ? 2 : 1);
}
int msclip_op_patch_pool(const void* in_bf16, int batch, int height, int width, int cpix, int c_off, int channels,
                         int k, const float* w, const float* bias, void* out_bf16, void* stream) {
  return launch_patch_pool(static_cast<const op16*>(in_bf16), batch, height, width, cpix, c_off, channels, k, w, bias,
                           static_cast<op16*>(out_bf16), as_stream(stream));
}
int msclip_op_front_conv(const void* img, int dtype, int batch, int height, int width, const void* w0_bf16, const float* b0,
                         const void* w1_bf16, const float* b1, const float* pool_w, const float* pool_b, int k,
                         void* stem_bf16, void* y1_bf16, void* p0s_bf16, int p0s_pitch, void* pooled_bf16, void* stream) {
  return launch_front_conv(img, dtype, batch, height, width, static_cast<const op16*>(w0_bf16), b0,
                           static_cast<const op16*>(w1_bf16), b1, pool_w, pool_b, k, static_cast<op16*>(stem_bf16),
                           static_cast<op16*>(y1_bf16), static_cast<op16*>(p0s_bf16), p0s_pitch,
                           static_cast<op16*>(pooled_bf16), as_stream(stream));
}
int msclip_op_adapter_fuse_ln(const float* x, const float* t, const float* dw_w9, const float* dw_bias, const float* w,
                              const float* b, float* x_out, int batch, int grid, void* stream) {
  return launch_adapter_fuse_ln(x, t, dw_w9, dw_bias, w, b, x_out, batch, grid, nullptr, nullptr, nullptr, nullptr, nullptr,
                                as_stream(stream));
}
int msclip_op_contrastive_lse(const void* img_f16, const void* txt_f16, int b, float scale, void* workspace,
                              float* parts2, void* stream) {
  // single-process form: shard tables with one entry each, built in the caller-provided workspace tail
  const size_t need = contrastive_loss_workspace_bytes(1, b);
  const void** tab = reinterpret_cast<const void**>(static_cast<uint8_t*>(workspace) + ((need + 15) & ~size_t(15)));
  const void* host_tab[2] = {img_f16, txt_f16};
  MSCLIP_CHECK_CUDA(cudaMemcpyAsync(tab, host_tab, sizeof(host_tab), cudaMemcpyHostToDevice, as_stream(stream)));
  return launch_contrastive_loss_ex(static_cast<const emb16*>(img_f16), static_cast<const emb16*>(txt_f16),
                                    reinterpret_cast<const emb16* const*>(tab), reinterpret_cast<const emb16* const*>(tab + 1),
                                    nullptr, 0, 1, 0, b, 512, scale, workspace, parts2, nullptr, 0, as_stream(stream));
}
size_t msclip_op_contrastive_lse_workspace(int b) { return ((contrastive_loss_workspace_bytes(1, b) + 15) & ~size_t(15)) + 64; }

}  // extern "C"
