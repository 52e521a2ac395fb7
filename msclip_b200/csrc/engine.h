// Engine: owns the packed weights and the workspace of one model instance and sequences the kernels of
// the MS-CLIP-S forward (vision tower M.py:2621-2697 / 2357-2471, text tower M.py:3043-3079, contrast
// M.py:3126-3155).  Host-only C++; all device work goes through the launchers in kernels.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/msclip_b200.h"
#include "kernels.h"

namespace msclip {

struct RawTensor {
  std::vector<int64_t> shape;
  float* dev = nullptr;
  size_t numel = 0;
};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

struct BlockWeights {
  op16 *w_qkv = nullptr, *w_o = nullptr, *w_fc1 = nullptr, *w_fc2 = nullptr;
  float *b_qkv = nullptr, *b_o = nullptr, *b_fc1 = nullptr, *b_fc2 = nullptr;
  float *ln1_w = nullptr, *ln1_b = nullptr, *ln2_w = nullptr, *ln2_b = nullptr;
  // LN fold (gemm_common.cuh): weights with ln_1 / ln_2 gamma folded in (per tower: the LayerNorms are not shared),
  // their column sums and the biases with W . beta added
  op16 *w_qkv_ln = nullptr, *w_fc1_ln = nullptr;
  float *cs_qkv = nullptr, *cs_fc1 = nullptr, *b_qkv_ln = nullptr, *b_fc1_ln = nullptr;
};

// transposed copies of the block's linear weights: the B operand of the input-gradient GEMMs (dX = dY . W needs W^T
// K-major); w_qkv_t is packed from the UNSCALED in_proj weight (attention_bwd emits the gradient of the unscaled q)
struct BlockWeightsT {
  op16 *w_qkv_t = nullptr, *w_o_t = nullptr, *w_fc1_t = nullptr, *w_fc2_t = nullptr;  // [768,2304] [768,768] [768,3072] [3072,768]
};
// fp32 gradient buffers of one block (text blocks 1.. share the linear-layer buffers of the vision block, like the weights)
struct BlockGrads {
  float *w_qkv = nullptr, *b_qkv = nullptr, *w_o = nullptr, *b_o = nullptr, *w_fc1 = nullptr, *b_fc1 = nullptr, *w_fc2 = nullptr,
        *b_fc2 = nullptr, *ln1_w = nullptr, *ln1_b = nullptr, *ln2_w = nullptr, *ln2_b = nullptr;
};

struct ConvWeights {
  op16* w = nullptr;   // [N, K] op16, BN scale folded
  float* b = nullptr;  // [N] BN shift (null = no bias)
  int N = 0, K = 0;
  op16* w_tma = nullptr;  // the same weight in the padded K layout of the im2col-TMA kernel (conv_tma_kpad columns)
  int K_tma = 0;
};

struct AdapterWeights {
  float *dw_w = nullptr, *dw_b = nullptr;    // top2bottom depth-wise k x k (+BN): [k*k][C], [C]
  op16* pw = nullptr;                        // top2bottom point-wise: [768, C]
  float *bdw_w9 = nullptr, *bdw_b = nullptr; // bottom depth-wise 3x3 (+BN): [9][768], [768]
  float *ln_w = nullptr, *ln_b = nullptr;
  int C = 0, k = 0;
};

}  // namespace msclip

struct msclip_ctx {
  msclip_config cfg;
  int grid = 0, l_img = 0, heads = 0;
  std::map<std::string, std::vector<int64_t>> spec;  // expected state-dict keys -> shapes
  std::map<std::string, msclip::RawTensor> raw;      // copies made by msclip_set_weight (freed by finalize)
  bool finalized = false;
  bool text_trim = true;  // encode_text runs the causal tower only over the longest live prefix (<= EOT) of the batch
  float logit_scale = 0.f;
  // logit_scale refreshed from the device after an optimiser step: the read-back lands in pinned memory and is waited for
  // only when the value is next needed (no host synchronisation inside the training step)
  float* ls_pinned = nullptr;
  cudaEvent_t ls_event = nullptr;
  bool ls_pending = false;
  float* qscale_dev = nullptr;  // [3 * width]: 1/8 for the q rows of in_proj (M.py:707), 1 elsewhere

  // packed weights
  std::vector<msclip::BlockWeights> vblocks, tblocks;  // index = block id (vblocks[0] unused: it is the stem)
  msclip::ConvWeights first, stem[4], last_conv, br1[5], br2[5], br3[5];
  msclip::AdapterWeights adapters[5];
  float *cls = nullptr, *vpos = nullptr, *ln_pre_w = nullptr, *ln_pre_b = nullptr, *ln_post_w = nullptr,
        *ln_post_b = nullptr;
  float *tok_emb = nullptr, *tpos = nullptr, *ln_final_w = nullptr, *ln_final_b = nullptr;
  msclip::op16 *vproj = nullptr, *tproj = nullptr;  // [embed, width] (transposed projections)
  std::vector<void*> weight_allocs;
  size_t weight_bytes = 0;

  // workspace (grown on demand)
  std::map<std::string, msclip::DevBuf> ws;
  size_t ws_bytes = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copy = nullptr, ev_main = nullptr;
  // images handed over early with msclip_stage_images (input prefetch, double buffered)
  struct Staged {
    const void* host = nullptr;
    void* dev = nullptr;
    size_t bytes = 0, capacity = 0;
    cudaEvent_t ready = nullptr, consumed = nullptr;
    bool pending = false;
    uint64_t seq = 0;
  } staged[2];
  int stage_next = 0;
  uint64_t stage_seq = 0;

  // embedding exchange (data-parallel contrastive loss)
  int rank = 0, world = 1, max_b_local = 0;
  void* xchg = nullptr;  // [2 parity][2 modality][max_b_local, E] fp16, flags[64], flags2[64], lse [2 parity][2 dir][pitch] f32
  size_t xchg_bytes = 0;
  std::vector<void*> peer_base;    // imported peer bases (own base at [rank])
  bool peers_borrowed = false;     // peer bases are same-process device pointers (not IPC mappings to close)
  void* shard_tables = nullptr;    // device: [2 parity][2 modality][world] pointers
  uint32_t** peer_flag_tables = nullptr;  // device: [world] pointers to each rank's flag array
  uint32_t epoch = 0;
  // rows of this rank's shard (exchange slot of the NEXT loss) filled so far by encode_image / encode_text: one call
  // with the whole local batch, or several micro-batches appended back to back (msclip_encode_pairs)
  int img_rows = 0, txt_rows = 0;
  int loss_b = 0;          // local batch of the last contrastive loss (0 = none): what msclip_contrastive_loss_backward uses
  float loss_scale = 0.f;

  // ---- training (engine_train.cu): backward of heads, transformer blocks, adapter bottom paths and
  // embeddings.  `train` makes finalize keep transposed weight copies and makes the towers keep their per-block inputs.
  bool train = false;
  std::vector<msclip::BlockWeightsT> vblocks_t, tblocks_t;
  msclip::op16 *vproj_n = nullptr, *tproj_n = nullptr;  // projections as [width, embed] K-major (dgrad operand)
  std::vector<msclip::BlockGrads> vgrads, tgrads;
  std::vector<std::pair<std::string, float*>> grad_list;  // state-dict key -> gradient buffer (aliased keys share one buffer)
  std::vector<int64_t> grad_numel;
  std::vector<std::pair<float*, size_t>> grad_allocs;     // unique allocations (pointer, bytes)
  struct TapeInfo {
    int batch = 0, L = 0, normalize = 1;
    bool valid = false;
    bool keep = false;  // the tape also holds every block's QKV, attention output and mid-block stream (no recompute of those)
    bool keep_u = false;  // ... and fc1's pre-activation (written by the dual-output epilogue of the taped forward)
  } tape_txt, tape_img;

  ~msclip_ctx();
};

namespace msclip {

void build_spec(msclip_ctx* h);
int engine_finalize(msclip_ctx* h, cudaStream_t stream);
int engine_encode_image(msclip_ctx* h, const void* image, int dtype, int batch, float* out, int normalize,
                        cudaStream_t stream);
int engine_set_text_trim(msclip_ctx* h, int enable);
int engine_encode_text(msclip_ctx* h, const int64_t* tokens, int batch, float* out, int normalize,
                       cudaStream_t stream);
int engine_similarity_logits(msclip_ctx* h, const float* img, int n_img, const float* txt, int n_txt, float scale,
                             float* logits, cudaStream_t stream);
int engine_zeroshot_classifier(msclip_ctx* h, const int64_t* tokens, int n_classes, int n_templates, float* weights_out,
                               cudaStream_t stream);
int engine_zeroshot_predict(msclip_ctx* h, const float* img_feat, int n_img, const float* weights, int n_classes, float scale,
                            int topk, int32_t* topk_out, float* logits_out, cudaStream_t stream);
int engine_forward(msclip_ctx* h, const void* image, int dtype, const int64_t* tokens, int batch, float* logits,
                   cudaStream_t stream);
int engine_contrastive_loss(msclip_ctx* h, int b_local, float scale, float* partial_out, float* loss_out,
                            cudaStream_t stream);
int engine_contrastive_loss_features(msclip_ctx* h, const float* img_feat, const float* txt_feat, int b_local, float scale,
                                     float* partial_out, float* loss_out, cudaStream_t stream);
int engine_contrastive_loss_backward(msclip_ctx* h, float* d_img, float* d_txt, cudaStream_t stream);
int engine_encode_pairs(msclip_ctx* h, const void* image, int dtype, const int64_t* tokens, int b_micro, int row_offset,
                        cudaStream_t stream);
int engine_forward_loss(msclip_ctx* h, const void* image, int dtype, const int64_t* tokens, int b_local,
                        float* partial_out, float* loss_out, cudaStream_t stream);
int engine_stage_images(msclip_ctx* h, const void* image_host, int dtype, int batch, cudaStream_t stream);
int comm_init(msclip_ctx* h, int rank, int world, int max_b_local);
int comm_export(msclip_ctx* h, void* handle_out);
int comm_import(msclip_ctx* h, const void* handles);
int comm_import_pointers(msclip_ctx* h, void* const* bases);

// workspace buffers are named and grow on demand (engine.cu)
int ws_get(msclip_ctx* h, const char* name, size_t bytes, void** out);
#define WS(var, type, name, count) \
  type* var = nullptr;             \
  MSCLIP_TRY(::msclip::ws_get(h, name, static_cast<size_t>(count) * sizeof(type), reinterpret_cast<void**>(&var)))
// tape of the training forward: named copies of activations the backward pass needs ("tape:" + name)
int tape_save(msclip_ctx* h, const std::string& name, const void* src, size_t bytes, cudaStream_t s);
void* tape_get(msclip_ctx* h, const std::string& name);
constexpr int kLateralLayers[5] = {2, 4, 6, 8, 10};  // PARALLEL_LATERAL_LAYER, b32-yfcc-msclips.yaml:18
constexpr int kTrainMaxBatch = 4096;                 // one transformer chunk per tower call

int train_enable(msclip_ctx* h, int enable);
int train_pack_transposed(msclip_ctx* h, cudaStream_t stream);   // called by engine_finalize while the raw tensors exist
int engine_backward(msclip_ctx* h, const float* d_img, const float* d_txt, cudaStream_t stream);
int engine_zero_grad(msclip_ctx* h, cudaStream_t stream);
int engine_update_weight(msclip_ctx* h, const char* key, const float* src, cudaStream_t stream);
void train_free(msclip_ctx* h);

int engine_preprocess(msclip_ctx* h, const uint8_t* pixels, const int64_t* offsets, const int* heights, const int* widths, int n, int S,
                      const float* mean, const float* stdv, void* out, int out_dtype, uint8_t* out_u8, cudaStream_t s);

float current_logit_scale(msclip_ctx* h);  // h->logit_scale, after any pending read-back has landed

int64_t launch_count();
void count_launch(int n);
bool is_device_pointer(const void* p);
const void* pinned_device_alias(const void* p);

}  // namespace msclip
