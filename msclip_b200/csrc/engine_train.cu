// Training side of the engine: backward of the encode-and-contrast path from the gradient of
// the (normalised) embeddings down to the embeddings of both towers, and in-place refresh of packed weights after an
// optimiser step.  The reference ships no backward (M.py:3155 returns logits); the oracle is torch.autograd on the
// reference module in eval mode (tests/golden/grad_*.npz, oracle/make_golden_grads.py).
//
// Scope: both heads (L2 norm, projection, ln_final / ln_post; M.py:3057-3077, 2685-2690, 2982-2983), every
// ResidualAttentionBlock of both towers (M.py:1027-1028; shared weights accumulate the gradients of both modalities,
// M.py:2786-2830), the bottom path of the lateral adapters (M.py:1760-1777, so the gradient reaches every vision block),
// token / positional / class embeddings and ln_pre (M.py:3047-3048, 2418-2426).  The convolutional front (stem, parallel
// branch, the adapters' convolutions and BatchNorms) is treated as frozen: it would need train-mode BatchNorm, which the
// reference's eval forward - the parity target of this repo - does not exercise.
//
// Memory plan: the training forward keeps only the fp32 input of every block (3 KB per token and block, "tape");
// block_backward recomputes the block's intermediates with the forward kernels, then runs
//   cast+colsum -> dgrad fc2 -> wgrad fc2 -> QuickGELU' (+ fc1 bias grad) -> wgrad fc1 -> dgrad fc1 -> LN2 backward
//   -> dgrad out-proj -> wgrad out-proj -> attention backward -> bias grad -> wgrad QKV -> dgrad QKV -> LN1 backward.
// dgrad = the forward tcgen05 GEMM on transposed weight copies; wgrad = wgrad.cu (MN-major operands).
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "engine.h"

namespace msclip {

namespace {

constexpr int kW = 768;

// MSCLIP_BWD_FUSED_GELU=0: separate QuickGELU forward / backward kernels around a plain dgrad GEMM (A/B timing and the
// comparator of the fused epilogue's parity test)
const bool g_fused_dgelu = [] {
  const char* e = getenv("MSCLIP_BWD_FUSED_GELU");
  return e == nullptr || e[0] != '0';
}();

float* grad_find(msclip_ctx* h, const std::string& key) {
  for (auto& kv : h->grad_list)
    if (kv.first == key) return kv.second;
  return nullptr;
}

int grad_new(msclip_ctx* h, const std::string& key, int64_t numel, float** out) {
  float* p = nullptr;
  const size_t bytes = static_cast<size_t>(numel) * sizeof(float);
  MSCLIP_CHECK_CUDA(cudaMalloc(&p, bytes));
  MSCLIP_CHECK_CUDA(cudaMemset(p, 0, bytes));
  h->grad_allocs.emplace_back(p, bytes);
  h->grad_list.emplace_back(key, p);
  h->grad_numel.push_back(numel);
  *out = p;
  return 0;
}

void grad_alias(msclip_ctx* h, const std::string& key, float* p, int64_t numel) {
  h->grad_list.emplace_back(key, p);
  h->grad_numel.push_back(numel);
}

int make_block_grads(msclip_ctx* h, const std::string& p, BlockGrads& g, const BlockGrads* share) {
  const int64_t w = kW;
  struct Lin {
    const char* suffix;
    float* BlockGrads::*field;
    int64_t numel;
  };
  const Lin lin[8] = {{".attn.in_proj_weight", &BlockGrads::w_qkv, 3 * w * w}, {".attn.in_proj_bias", &BlockGrads::b_qkv, 3 * w},
                      {".attn.out_proj.weight", &BlockGrads::w_o, w * w},      {".attn.out_proj.bias", &BlockGrads::b_o, w},
                      {".mlp.c_fc.weight", &BlockGrads::w_fc1, 4 * w * w},     {".mlp.c_fc.bias", &BlockGrads::b_fc1, 4 * w},
                      {".mlp.c_proj.weight", &BlockGrads::w_fc2, 4 * w * w},   {".mlp.c_proj.bias", &BlockGrads::b_fc2, w}};
  for (const Lin& l : lin) {
    if (share) {
      g.*(l.field) = share->*(l.field);
      grad_alias(h, p + l.suffix, g.*(l.field), l.numel);
    } else {
      MSCLIP_TRY(grad_new(h, p + l.suffix, l.numel, &(g.*(l.field))));
    }
  }
  MSCLIP_TRY(grad_new(h, p + ".ln_1.weight", w, &g.ln1_w));
  MSCLIP_TRY(grad_new(h, p + ".ln_1.bias", w, &g.ln1_b));
  MSCLIP_TRY(grad_new(h, p + ".ln_2.weight", w, &g.ln2_w));
  MSCLIP_TRY(grad_new(h, p + ".ln_2.bias", w, &g.ln2_b));
  return 0;
}

int active_adapters(const msclip_ctx* h) {
  int n = 0;
  for (int j = 0; j < 5; ++j)
    if (kLateralLayers[j] < h->cfg.layers) n = j + 1;
  return n;
}

// dst[j] (+)= sum over the per-CTA partial rows
int fold(const float* part, int nparts, long long pitch, float* dst, long long n, cudaStream_t s) {
  return launch_reduce_partials(part, nparts, pitch, dst, n, 1, 0, 1.0f, s);
}

struct BwdScratch {
  op16 *h1, *qkv, *ctx, *u, *a, *da, *g16, *dctx, *dqkv, *h2;
  float *x1, *tmp32, *part;
  void* wg;
};

int get_scratch(msclip_ctx* h, int M, BwdScratch& b) {
  const size_t m = static_cast<size_t>(M);
  WS(h1, op16, "h", m * kW);
  WS(qkv, op16, "qkv", m * 3 * kW);
  WS(ctx, op16, "attn", m * kW);
  WS(u, op16, "fc1", m * 4 * kW);
  WS(a, op16, "bw_a", m * 4 * kW);
  WS(da, op16, "bw_da", m * 4 * kW);
  WS(g16, op16, "bw_g16", m * kW);
  WS(dctx, op16, "bw_dctx", m * kW);
  WS(dqkv, op16, "bw_dqkv", m * 3 * kW);
  WS(h2, op16, "bw_h2", m * kW);
  WS(x1, float, "bw_x1", m * kW);
  WS(tmp32, float, "bw_tmp32", m * kW);
  const size_t part_floats = std::max(static_cast<size_t>(bwd_row_parts(M)) * 3 * kW, static_cast<size_t>(bwd_slab_parts(M)) * 4 * kW);
  WS(part, float, "bw_part", part_floats);
  size_t wgb = 0;
  const int shapes[5][2] = {{3 * kW, kW}, {kW, kW}, {4 * kW, kW}, {kW, 4 * kW}, {kW, 512}};
  for (const auto& sh : shapes) wgb = std::max(wgb, wgrad_workspace_bytes(M, sh[0], sh[1]));
  WS(wg, uint8_t, "bw_wgrad", wgb);
  b = {h1, qkv, ctx, u, a, da, g16, dctx, dqkv, h2, x1, tmp32, part, wg};
  return 0;
}

// dx (gradient of the block output, fp32 [M, 768]) -> gradient of the block input (in place); parameter gradients accumulate.
// g16_ready: b.g16 already holds op16(dx) and the fc2 bias gradient of this block has been taken (both were emitted by the
// LayerNorm backward that ended the block above).  below: the block that consumes this block's dx next with nothing in
// between (its fc2 bias gradient = column sums of our final dx), or null.
// kept: QKV / attention output / mid-block stream of this block from the tape (all three or none): their recompute is skipped.
struct KeptActs {
  const op16* qkv = nullptr;
  const op16* ctx = nullptr;
  const float* x1 = nullptr;
  const op16* u = nullptr;  // fc1 pre-activation (optional on top of the other three)
};
int block_backward(msclip_ctx* h, const BlockWeights& bw, const BlockWeightsT& bt, const BlockGrads& bg, const float* x_in,
                   float* dx, int batch, int L, int causal, const BwdScratch& b_in, bool g16_ready, const BlockGrads* below,
                   const KeptActs& kept, cudaStream_t s) {
  const int w = kW, M = batch * L;
  const int rp = bwd_row_parts(M), sp = bwd_slab_parts(M);
  BwdScratch b = b_in;
  // ---- recompute the block's intermediates (the forward's kernels; fc1 keeps the pre-activation u)
  MSCLIP_TRY(launch_layernorm_op16(x_in, 1, bw.ln1_w, bw.ln1_b, b.h1, M, s));
  if (kept.qkv != nullptr) {
    b.qkv = const_cast<op16*>(kept.qkv);
    b.ctx = const_cast<op16*>(kept.ctx);
    b.x1 = const_cast<float*>(kept.x1);
    count_launch(-3);
  } else {
    MSCLIP_TRY(launch_gemm(b.h1, w, bw.w_qkv, w, M, 3 * w, w, bw.b_qkv, b.qkv, 3 * w, nullptr, 0, EPI_BF16, s));
    MSCLIP_TRY(launch_attention(b.qkv, b.ctx, batch, L, h->heads, causal, s));
    MSCLIP_TRY(launch_gemm(b.ctx, w, bw.w_o, w, M, w, w, bw.b_o, b.x1, w, x_in, w, EPI_RESID_F32, s));
  }
  MSCLIP_TRY(launch_layernorm_op16(b.x1, 1, bw.ln2_w, bw.ln2_b, b.h2, M, s));
  if (kept.u != nullptr) {
    b.u = const_cast<op16*>(kept.u);
    count_launch(-1);
  } else {
    MSCLIP_TRY(launch_gemm(b.h2, w, bw.w_fc1, w, M, 4 * w, w, bw.b_fc1, b.u, 4 * w, nullptr, 0, EPI_BF16, s));
  }
  // ---- MLP (M.py:794-798, 1028)
  if (!g16_ready) {
    MSCLIP_TRY(launch_cast_colsum(dx, b.g16, b.part, M, s));
    MSCLIP_TRY(fold(b.part, rp, w, bg.b_fc2, w, s));
  }
  if (g_fused_dgelu) {
    // d u = (d x2 . W2) * quickgelu'(u) and a = quickgelu(u) from ONE kernel: the dgrad GEMM's epilogue reads u
    MSCLIP_TRY(launch_gemm_dgelu(b.g16, w, bt.w_fc2_t, w, M, 4 * w, w, b.u, b.da, b.a, s));
    MSCLIP_TRY(launch_colsum16(b.da, b.part, M, 4 * w, s));
  } else {
    MSCLIP_TRY(launch_qgelu_fwd(b.u, b.a, static_cast<long long>(M) * 4 * w, s));
    MSCLIP_TRY(launch_gemm(b.g16, w, bt.w_fc2_t, w, M, 4 * w, w, nullptr, b.da, 4 * w, nullptr, 0, EPI_BF16, s));
    MSCLIP_TRY(launch_qgelu_bwd(b.da, b.u, b.part, M, 4 * w, s));
  }
  MSCLIP_TRY(fold(b.part, sp, 4 * w, bg.b_fc1, 4 * w, s));
  MSCLIP_TRY(launch_wgrad(b.g16, w, b.a, 4 * w, M, w, 4 * w, bg.w_fc2, 1, b.wg, s));
  MSCLIP_TRY(launch_wgrad(b.da, 4 * w, b.h2, w, M, 4 * w, w, bg.w_fc1, 1, b.wg, s));
  MSCLIP_TRY(launch_gemm(b.da, 4 * w, bt.w_fc1_t, 4 * w, M, w, 4 * w, nullptr, b.tmp32, w, nullptr, 0, EPI_F32, s));
  // dx1 = dx2 + LN2'(dh2); its 16-bit copy feeds out-proj's dgrad / wgrad, its column sums are out-proj's bias gradient
  MSCLIP_TRY(launch_ln_bwd(b.x1, b.tmp32, bw.ln2_w, dx, b.g16, b.part, M, 1, s));
  MSCLIP_TRY(launch_reduce_partials3(b.part, rp, 3 * w, bg.ln2_w, bg.ln2_b, bg.b_o, w, s));
  // ---- attention (M.py:612, 707-747, 1027)
  MSCLIP_TRY(launch_gemm(b.g16, w, bt.w_o_t, w, M, w, w, nullptr, b.dctx, w, nullptr, 0, EPI_BF16, s));
  MSCLIP_TRY(launch_wgrad(b.g16, w, b.ctx, w, M, w, w, bg.w_o, 1, b.wg, s));
  MSCLIP_TRY(launch_attention_bwd(b.qkv, b.dctx, b.dqkv, batch, L, h->heads, causal, s));
  MSCLIP_TRY(launch_colsum16(b.dqkv, b.part, M, 3 * w, s));
  MSCLIP_TRY(fold(b.part, sp, 3 * w, bg.b_qkv, 3 * w, s));
  MSCLIP_TRY(launch_wgrad(b.dqkv, 3 * w, b.h1, w, M, 3 * w, w, bg.w_qkv, 1, b.wg, s));
  MSCLIP_TRY(launch_gemm(b.dqkv, 3 * w, bt.w_qkv_t, 3 * w, M, w, 3 * w, nullptr, b.tmp32, w, nullptr, 0, EPI_F32, s));
  // the block below starts from this dx: hand it the 16-bit copy and its fc2 bias gradient now
  MSCLIP_TRY(launch_ln_bwd(x_in, b.tmp32, bw.ln1_w, dx, below ? b.g16 : nullptr, b.part, M, 1, s));
  MSCLIP_TRY(launch_reduce_partials3(b.part, rp, 3 * w, bg.ln1_w, bg.ln1_b, below ? below->b_fc2 : nullptr, w, s));
  count_launch(6 + (g16_ready ? 0 : 2) + (g_fused_dgelu ? 2 : 3) + 1 + 4 + 1 + 2 + 2 + 2 + 1 + 2 + 2 + 1 + 2);
  return 0;
}

// head: d(normalised features) -> dx rows of the pooled positions (dx zero elsewhere); projection / final-LN gradients
int head_backward(msclip_ctx* h, const float* d_feat, const float* feat_raw, const float* x_final, int batch, int L, int normalize,
                  const int64_t* tok, const float* ln_w, const float* ln_b, const op16* proj_n, float* g_proj, float* g_ln_w,
                  float* g_ln_b, float* dx, const BwdScratch& b, cudaStream_t s) {
  const msclip_config& c = h->cfg;
  const int E = c.embed_dim;
  WS(dg16, op16, "bw_dg16", static_cast<size_t>(batch) * E);
  WS(z16, op16, "bw_z16", static_cast<size_t>(batch) * kW);
  WS(dz, float, "bw_dz", static_cast<size_t>(batch) * kW);
  MSCLIP_TRY(launch_l2norm_bwd(feat_raw, d_feat, nullptr, dg16, batch, E, normalize, s));
  if (tok != nullptr)
    MSCLIP_TRY(launch_eot_layernorm_op16(x_final, L, tok, c.context_length, ln_w, ln_b, z16, batch, s));
  else
    MSCLIP_TRY(launch_layernorm_op16(x_final, L, ln_w, ln_b, z16, batch, s));
  // features = z . P, P [768, 512]:  dP = z^T dg,  dz = dg . P^T
  MSCLIP_TRY(launch_wgrad(z16, kW, dg16, E, batch, kW, E, g_proj, 1, b.wg, s));
  MSCLIP_TRY(launch_gemm(dg16, E, proj_n, E, batch, kW, E, nullptr, dz, kW, nullptr, 0, EPI_F32, s));
  MSCLIP_CHECK_CUDA(cudaMemsetAsync(dx, 0, static_cast<size_t>(batch) * L * kW * sizeof(float), s));
  MSCLIP_TRY(launch_pooled_ln_bwd(x_final, L, tok, c.context_length, dz, ln_w, dx, b.part, batch, s));
  const int rp = bwd_row_parts(batch);
  MSCLIP_TRY(fold(b.part, rp, 2 * kW, g_ln_w, kW, s));
  MSCLIP_TRY(fold(b.part + kW, rp, 2 * kW, g_ln_b, kW, s));
  count_launch(9);
  return 0;
}

int text_backward(msclip_ctx* h, const float* d_txt, cudaStream_t s) {
  const msclip_config& c = h->cfg;
  MSCLIP_REQUIRE(h->tape_txt.valid, "backward: no taped encode_text (enable training, then encode at most 4096 sequences per call)");
  const int B = h->tape_txt.batch, L = h->tape_txt.L, M = B * L;
  BwdScratch b;
  MSCLIP_TRY(get_scratch(h, M, b));
  WS(dx, float, "bw_dx", static_cast<size_t>(M) * kW);
  const int64_t* tok = static_cast<const int64_t*>(tape_get(h, "t_tok"));
  const float* feat = static_cast<const float*>(tape_get(h, "t_feat"));
  const float* xfin = static_cast<const float*>(tape_get(h, "t_x" + std::to_string(c.layers)));
  MSCLIP_REQUIRE(tok && feat && xfin, "backward: incomplete text tape");
  MSCLIP_TRY(head_backward(h, d_txt, feat, xfin, B, L, h->tape_txt.normalize, tok, h->ln_final_w, h->ln_final_b, h->tproj_n,
                           grad_find(h, "text_projection"), grad_find(h, "ln_final.weight"), grad_find(h, "ln_final.bias"), dx, b, s));
  for (int idx = c.layers - 1; idx >= 0; --idx) {
    const float* xin = static_cast<const float*>(tape_get(h, "t_x" + std::to_string(idx)));
    MSCLIP_REQUIRE(xin != nullptr, "backward: incomplete text tape");
    KeptActs kept;
    if (h->tape_txt.keep) {
      const std::string t = std::to_string(idx);
      kept.qkv = static_cast<const op16*>(tape_get(h, "t_qkv" + t));
      kept.ctx = static_cast<const op16*>(tape_get(h, "t_ctx" + t));
      kept.x1 = static_cast<const float*>(tape_get(h, "t_mid" + t));
      MSCLIP_REQUIRE(kept.qkv && kept.ctx && kept.x1, "backward: incomplete text tape (kept activations)");
      if (h->tape_txt.keep_u) {
        kept.u = static_cast<const op16*>(tape_get(h, "t_u" + t));
        MSCLIP_REQUIRE(kept.u != nullptr, "backward: incomplete text tape (kept pre-activations)");
      }
    }
    MSCLIP_TRY(block_backward(h, h->tblocks[idx], h->tblocks_t[idx], h->tgrads[idx], xin, dx, B, L, 1, b, idx != c.layers - 1,
                              idx > 0 ? &h->tgrads[idx - 1] : nullptr, kept, s));
  }
  MSCLIP_TRY(launch_text_embed_bwd(dx, tok, c.context_length, L, B, c.vocab_size, grad_find(h, "positional_embedding"),
                                   grad_find(h, "token_embedding.weight"), s));
  count_launch(1);
  return 0;
}

int image_backward(msclip_ctx* h, const float* d_img, cudaStream_t s) {
  const msclip_config& c = h->cfg;
  MSCLIP_REQUIRE(h->tape_img.valid, "backward: no taped encode_image (enable training, then encode at most 4096 images per call)");
  MSCLIP_REQUIRE(h->l_img <= 208, "backward: image sequences longer than 208 tokens have no attention backward");
  const int B = h->tape_img.batch, L = h->l_img, M = B * L, g = h->grid;
  BwdScratch b;
  MSCLIP_TRY(get_scratch(h, M, b));
  WS(dx, float, "bw_dx", static_cast<size_t>(M) * kW);
  WS(dx2, float, "bw_dx2", static_cast<size_t>(M) * kW);
  const float* feat = static_cast<const float*>(tape_get(h, "v_feat"));
  const float* xfin = static_cast<const float*>(tape_get(h, "v_x" + std::to_string(c.layers)));
  const float* grid = static_cast<const float*>(tape_get(h, "v_grid"));
  MSCLIP_REQUIRE(feat && xfin && grid, "backward: incomplete image tape");
  MSCLIP_TRY(head_backward(h, d_img, feat, xfin, B, L, h->tape_img.normalize, nullptr, h->ln_post_w, h->ln_post_b, h->vproj_n,
                           grad_find(h, "visual.proj"), grad_find(h, "visual.ln_post.weight"), grad_find(h, "visual.ln_post.bias"), dx,
                           b, s));
  const int n_active = active_adapters(h);
  bool g16_ready = false;
  for (int idx = c.layers - 1; idx >= 1; --idx) {
    const float* xin = static_cast<const float*>(tape_get(h, "v_x" + std::to_string(idx)));
    MSCLIP_REQUIRE(xin != nullptr, "backward: incomplete image tape");
    bool adapter_here = false;
    for (int j = 0; j < n_active; ++j) adapter_here |= kLateralLayers[j] == idx;
    KeptActs kept;
    if (h->tape_img.keep) {
      const std::string t = std::to_string(idx);
      kept.qkv = static_cast<const op16*>(tape_get(h, "v_qkv" + t));
      kept.ctx = static_cast<const op16*>(tape_get(h, "v_ctx" + t));
      kept.x1 = static_cast<const float*>(tape_get(h, "v_mid" + t));
      MSCLIP_REQUIRE(kept.qkv && kept.ctx && kept.x1, "backward: incomplete image tape (kept activations)");
      if (h->tape_img.keep_u) {
        kept.u = static_cast<const op16*>(tape_get(h, "v_u" + t));
        MSCLIP_REQUIRE(kept.u != nullptr, "backward: incomplete image tape (kept pre-activations)");
      }
    }
    MSCLIP_TRY(block_backward(h, h->vblocks[idx], h->vblocks_t[idx], h->vgrads[idx], xin, dx, B, L, 0, b, g16_ready,
                              (idx > 1 && !adapter_here) ? &h->vgrads[idx - 1] : nullptr, kept, s));
    g16_ready = idx > 1 && !adapter_here;
    for (int j = 0; j < n_active; ++j) {
      if (kLateralLayers[j] != idx) continue;
      // lateral adapter in front of this block: x_in = ln_adapt(2 cls | dw3x3(x) + t)  (M.py:1760-1777)
      const AdapterWeights& a = h->adapters[j];
      const float* ax = static_cast<const float*>(tape_get(h, "v_ax" + std::to_string(j)));
      const float* at = static_cast<const float*>(tape_get(h, "v_at" + std::to_string(j)));
      MSCLIP_REQUIRE(ax && at, "backward: incomplete image tape (adapter)");
      MSCLIP_TRY(launch_adapter_bwd(ax, at, a.bdw_w9, a.bdw_b, a.ln_w, dx, dx2, b.part, B, g, s));
      const std::string p = "visual.transformer.parallel_lateral_adapter." + std::to_string(j) + ".ln_adapt.";
      const int rp = bwd_row_parts(M);
      MSCLIP_TRY(fold(b.part, rp, 2 * kW, grad_find(h, p + "weight"), kW, s));
      MSCLIP_TRY(fold(b.part + kW, rp, 2 * kW, grad_find(h, p + "bias"), kW, s));
      std::swap(dx, dx2);
      count_launch(4);
    }
  }
  MSCLIP_TRY(launch_image_embed_bwd(grid, h->cls, h->vpos, h->ln_pre_w, dx, b.part, B, L, grad_find(h, "visual.positional_embedding"),
                                    grad_find(h, "visual.class_embedding"), s));
  const int rp = bwd_row_parts(M);
  MSCLIP_TRY(fold(b.part, rp, 2 * kW, grad_find(h, "visual.ln_pre.weight"), kW, s));
  MSCLIP_TRY(fold(b.part + kW, rp, 2 * kW, grad_find(h, "visual.ln_pre.bias"), kW, s));
  count_launch(4);
  return 0;
}

}  // namespace

int train_enable(msclip_ctx* h, int enable) {
  MSCLIP_REQUIRE(h != nullptr, "null handle");
#ifdef MSCLIP_FP16
  MSCLIP_REQUIRE(!enable, "training needs the bf16 build (fp16 gradients would need loss scaling)");
#endif
  if (!enable) {
    h->train = false;
    return 0;
  }
  MSCLIP_REQUIRE(h->cfg.width == kW, "training: width must be 768");
  if (!h->train) {
    h->train = true;
    // packed weights lack the transposed copies: the caller re-sends the state dict (the raw tensors are gone)
    if (h->finalized) h->finalized = false;
  }
  if (!h->grad_list.empty()) return 0;
  const msclip_config& c = h->cfg;
  h->vgrads.assign(c.layers, BlockGrads());
  h->tgrads.assign(c.layers, BlockGrads());
  for (int i = 1; i < c.layers; ++i)
    MSCLIP_TRY(make_block_grads(h, "visual.transformer.resblocks." + std::to_string(i), h->vgrads[i], nullptr));
  for (int i = 0; i < c.layers; ++i)
    MSCLIP_TRY(make_block_grads(h, "transformer.resblocks." + std::to_string(i), h->tgrads[i], i >= 1 ? &h->vgrads[i] : nullptr));
  float* p = nullptr;
  const int64_t w = kW;
  MSCLIP_TRY(grad_new(h, "positional_embedding", static_cast<int64_t>(c.context_length) * w, &p));
  MSCLIP_TRY(grad_new(h, "text_projection", w * c.embed_dim, &p));
  MSCLIP_TRY(grad_new(h, "token_embedding.weight", static_cast<int64_t>(c.vocab_size) * w, &p));
  MSCLIP_TRY(grad_new(h, "ln_final.weight", w, &p));
  MSCLIP_TRY(grad_new(h, "ln_final.bias", w, &p));
  MSCLIP_TRY(grad_new(h, "visual.class_embedding", w, &p));
  MSCLIP_TRY(grad_new(h, "visual.positional_embedding", static_cast<int64_t>(h->l_img) * w, &p));
  MSCLIP_TRY(grad_new(h, "visual.proj", w * c.embed_dim, &p));
  MSCLIP_TRY(grad_new(h, "visual.ln_pre.weight", w, &p));
  MSCLIP_TRY(grad_new(h, "visual.ln_pre.bias", w, &p));
  MSCLIP_TRY(grad_new(h, "visual.ln_post.weight", w, &p));
  MSCLIP_TRY(grad_new(h, "visual.ln_post.bias", w, &p));
  for (int j = 0; j < active_adapters(h); ++j) {
    const std::string a = "visual.transformer.parallel_lateral_adapter." + std::to_string(j) + ".ln_adapt.";
    MSCLIP_TRY(grad_new(h, a + "weight", w, &p));
    MSCLIP_TRY(grad_new(h, a + "bias", w, &p));
  }
  return 0;
}

void train_free(msclip_ctx* h) {
  if (h->ls_pinned) cudaFreeHost(h->ls_pinned);
  if (h->ls_event) cudaEventDestroy(h->ls_event);
  h->ls_pinned = nullptr;
  h->ls_event = nullptr;
  h->ls_pending = false;
  for (auto& a : h->grad_allocs) cudaFree(a.first);
  h->grad_allocs.clear();
  h->grad_list.clear();
  h->grad_numel.clear();
}

int engine_zero_grad(msclip_ctx* h, cudaStream_t s) {
  MSCLIP_REQUIRE(h != nullptr && h->train, "zero_grad: training is not enabled");
  for (auto& a : h->grad_allocs) MSCLIP_CHECK_CUDA(cudaMemsetAsync(a.first, 0, a.second, s));
  return 0;
}

// Gradients of the LAST taped encode_image / encode_text of this handle with respect to every trainable parameter,
// accumulated into the handle's gradient buffers; d_img / d_txt [batch, 512] f32 = gradient of the (normalised) features,
// e.g. from msclip_contrastive_loss_backward; either may be null.
int engine_backward(msclip_ctx* h, const float* d_img, const float* d_txt, cudaStream_t s) {
  MSCLIP_REQUIRE(h != nullptr && h->train, "backward: training is not enabled (msclip_train_enable before the weights are finalized)");
  MSCLIP_REQUIRE(h->finalized && !h->vblocks_t.empty(), "backward: weights must be (re)finalized after msclip_train_enable");
  if (d_txt != nullptr) MSCLIP_TRY(text_backward(h, d_txt, s));
  if (d_img != nullptr) MSCLIP_TRY(image_backward(h, d_img, s));
  return 0;
}

// Refresh the packed copies of ONE trainable parameter from its fp32 master (after an optimiser step): device-side
// re-pack, no allocation.  Keys of the frozen convolutional front are rejected.
int engine_update_weight(msclip_ctx* h, const char* key_c, const float* src, cudaStream_t s) {
  MSCLIP_REQUIRE(h != nullptr && key_c != nullptr && src != nullptr, "update_weight: null argument");
  MSCLIP_REQUIRE(h->finalized && h->train && !h->vblocks_t.empty(), "update_weight: needs finalized weights of a training handle");
  const std::string key(key_c);
  const msclip_config& c = h->cfg;
  const int w = kW, E = c.embed_dim;
  auto copy = [&](float* dst, size_t n) -> int {
    MSCLIP_CHECK_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return 0;
  };
  // ---- transformer blocks
  for (int tower = 0; tower < 2; ++tower) {
    const std::string pre = tower == 0 ? "visual.transformer.resblocks." : "transformer.resblocks.";
    if (key.compare(0, pre.size(), pre) != 0) continue;
    const size_t dot = key.find('.', pre.size());
    MSCLIP_REQUIRE(dot != std::string::npos, "update_weight: malformed block key " + key);
    const int idx = std::atoi(key.substr(pre.size(), dot - pre.size()).c_str());
    const std::string suf = key.substr(dot);
    MSCLIP_REQUIRE(idx >= (tower == 0 ? 1 : 0) && idx < c.layers, "update_weight: " + key + " is not a trainable block parameter");
    BlockWeights& bw = tower == 0 ? h->vblocks[idx] : h->tblocks[idx];
    BlockWeightsT& bt = tower == 0 ? h->vblocks_t[idx] : h->tblocks_t[idx];
    if (suf == ".attn.in_proj_weight") {
      MSCLIP_TRY(launch_pack_op16(src, w, 1, h->qscale_dev, bw.w_qkv, w, 3 * w, w, s));
      return launch_pack_op16(src, 1, w, nullptr, bt.w_qkv_t, 3 * w, w, 3 * w, s);
    }
    if (suf == ".attn.in_proj_bias") return launch_reduce_partials(src, 1, 0, bw.b_qkv, 3 * w, 0, w, 0.125f, s);
    if (suf == ".attn.out_proj.weight") {
      MSCLIP_TRY(launch_pack_op16(src, w, 1, nullptr, bw.w_o, w, w, w, s));
      return launch_pack_op16(src, 1, w, nullptr, bt.w_o_t, w, w, w, s);
    }
    if (suf == ".attn.out_proj.bias") return copy(bw.b_o, w);
    if (suf == ".mlp.c_fc.weight") {
      MSCLIP_TRY(launch_pack_op16(src, w, 1, nullptr, bw.w_fc1, w, 4 * w, w, s));
      return launch_pack_op16(src, 1, w, nullptr, bt.w_fc1_t, 4 * w, w, 4 * w, s);
    }
    if (suf == ".mlp.c_fc.bias") return copy(bw.b_fc1, 4 * w);
    if (suf == ".mlp.c_proj.weight") {
      MSCLIP_TRY(launch_pack_op16(src, 4 * w, 1, nullptr, bw.w_fc2, 4 * w, w, 4 * w, s));
      return launch_pack_op16(src, 1, 4 * w, nullptr, bt.w_fc2_t, w, 4 * w, w, s);
    }
    if (suf == ".mlp.c_proj.bias") return copy(bw.b_fc2, w);
    if (suf == ".ln_1.weight") return copy(bw.ln1_w, w);
    if (suf == ".ln_1.bias") return copy(bw.ln1_b, w);
    if (suf == ".ln_2.weight") return copy(bw.ln2_w, w);
    if (suf == ".ln_2.bias") return copy(bw.ln2_b, w);
    MSCLIP_REQUIRE(false, "update_weight: " + key + " belongs to the frozen convolutional front");
  }
  if (key == "positional_embedding") return copy(h->tpos, static_cast<size_t>(c.context_length) * w);
  if (key == "token_embedding.weight") return copy(h->tok_emb, static_cast<size_t>(c.vocab_size) * w);
  if (key == "ln_final.weight") return copy(h->ln_final_w, w);
  if (key == "ln_final.bias") return copy(h->ln_final_b, w);
  if (key == "visual.class_embedding") return copy(h->cls, w);
  if (key == "visual.positional_embedding") return copy(h->vpos, static_cast<size_t>(h->l_img) * w);
  if (key == "visual.ln_pre.weight") return copy(h->ln_pre_w, w);
  if (key == "visual.ln_pre.bias") return copy(h->ln_pre_b, w);
  if (key == "visual.ln_post.weight") return copy(h->ln_post_w, w);
  if (key == "visual.ln_post.bias") return copy(h->ln_post_b, w);
  if (key == "text_projection" || key == "visual.proj") {
    op16* t = key == "text_projection" ? h->tproj : h->vproj;      // [E, w] (forward operand)
    op16* n = key == "text_projection" ? h->tproj_n : h->vproj_n;  // [w, E] (dgrad operand)
    MSCLIP_TRY(launch_pack_op16(src, 1, E, nullptr, t, w, E, w, s));
    return launch_pack_op16(src, E, 1, nullptr, n, E, w, E, s);
  }
  if (key == "logit_scale") {
    // asynchronous read-back into pinned memory; whoever needs the value next waits for it (current_logit_scale)
    if (!h->ls_pinned) {
      MSCLIP_CHECK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h->ls_pinned), sizeof(float), cudaHostAllocDefault));
      MSCLIP_CHECK_CUDA(cudaEventCreateWithFlags(&h->ls_event, cudaEventDisableTiming));
    }
    if (h->ls_pending) MSCLIP_CHECK_CUDA(cudaEventSynchronize(h->ls_event));
    MSCLIP_CHECK_CUDA(cudaMemcpyAsync(h->ls_pinned, src, sizeof(float), cudaMemcpyDeviceToHost, s));
    MSCLIP_CHECK_CUDA(cudaEventRecord(h->ls_event, s));
    h->ls_pending = true;
    return 0;
  }
  for (int j = 0; j < active_adapters(h); ++j) {
    const std::string a = "visual.transformer.parallel_lateral_adapter." + std::to_string(j) + ".ln_adapt.";
    if (key == a + "weight") return copy(h->adapters[j].ln_w, w);
    if (key == a + "bias") return copy(h->adapters[j].ln_b, w);
  }
  MSCLIP_REQUIRE(false, "update_weight: " + key + " is not trainable in this build (frozen convolutional front)");
  return 2;
}

}  // namespace msclip
