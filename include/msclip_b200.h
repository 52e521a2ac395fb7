/* msclip_b200 — C ABI of the B200-native MS-CLIP-S encode-and-contrast path.
 *
 * This is the drop-in boundary.  The reference (Hxyou/MSCLIP) is pure Python and
 * has no FFI of its own; the entry points below are what a binding for its model API
 * (lib/models/clip_openai_pe_res_v1.py, "M.py") binds, one per reference method:
 *
 *   msclip_create               <- get_clip_model(config) -> CLIP(...)            M.py:3182-3227, 2700-2852
 *   msclip_set_weight           <- model.load_state_dict(sd), one call per key    tools/zero_shot.py:223-224
 *   msclip_finalize_weights     <- model.to(device) / .eval()                     tools/zero_shot.py:225-228
 *   msclip_encode_image         <- CLIP.encode_image(image, norm)                 M.py:2979-2985
 *   msclip_stage_images         <- images.cuda(non_blocking=True) (input prefetch)    tools/zero_shot.py:262
 *   msclip_encode_text          <- CLIP.encode_text(text, norm)                   M.py:3043-3079
 *   msclip_similarity_logits    <- logit_scale.exp() * I @ T.t()                  M.py:3136, 3141/3146; zero_shot.py:266
 *   msclip_forward              <- CLIP.forward(image, text) -> logits            M.py:3126-3155
 *   msclip_contrastive_loss     <- gather_tensors x2 + logits + symmetric CE      M.py:3139-3141, comm.py:140-154
 *   msclip_forward_loss         <- forward + loss in one call (the benchmarked step)
 *   msclip_comm_*               <- lib/utils/comm.py gather_tensors replacement: peer shard registration
 *
 * Conventions: plain pointers and sizes only (no torch / CUDA types); `stream` is a cudaStream_t passed
 * as void* (NULL = default stream); every call is stream-ordered and returns 0 on success, non-zero on
 * failure with a message available from msclip_last_error() (thread-local).  Data pointers may be device
 * pointers or host pointers (pageable or pinned) — the library inspects them with
 * cudaPointerGetAttributes and stages host data through pinned buffers itself; a call with a host
 * output pointer returns after the result has landed.  The library borrows caller memory only for the
 * duration of a call and owns its workspace and its re-packed weights.
 * There is no CPU fallback: every compute entry point fails if no sm_100 device is present.
 */
#ifndef MSCLIP_B200_H_
#define MSCLIP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct msclip_ctx* msclip_handle;

/* dtype codes for msclip_set_weight / image inputs */
enum { MSCLIP_F32 = 0, MSCLIP_BF16 = 1, MSCLIP_F16 = 2, MSCLIP_I64 = 3 };

/* The MS-CLIP-S envelope (experiments/model/b32-yfcc-msclips.yaml, b16-yfcc-msclips.yaml). */
typedef struct msclip_config {
  int32_t patch_size;          /* 32 or 16 */
  int32_t layers;              /* vision layer 0 is the conv stem; text has `layers` blocks */
  int32_t width;               /* 768 */
  int32_t embed_dim;           /* 512 */
  int32_t image_resolution;    /* 224 */
  int32_t context_length;      /* 77 */
  int32_t vocab_size;          /* 49408 */
  int32_t early_strides[4];    /* EARLY_CONV_RES_STRIDES: {2,2,2,2} (B/32) or {2,2,2,1} (B/16) */
  int32_t parallel_strides[5]; /* PARALLEL_STRIDES: {2,2,2,2,2} or {2,2,2,2,1} */
  int32_t t2b_kernels[5];      /* PRALLEL_T2B_KERNELS (= strides): {16,8,4,2,1} or {8,4,2,1,1} */
} msclip_config;

const char* msclip_version(void);
const char* msclip_last_error(void);
/* number of visible CUDA devices with compute capability 10.x (0 on a CPU-only box; never fails) */
int msclip_device_count(void);

int msclip_create(const msclip_config* cfg, msclip_handle* out);
int msclip_destroy(msclip_handle h);

/* Hand over one entry of the reference state_dict (same key names, SURVEY.md section 8c "state-dict
 * contract"); data is copied.  dtype must be MSCLIP_F32 except for num_batches_tracked (MSCLIP_I64,
 * ignored).  Unknown keys and shape mismatches are errors (load_state_dict(strict=True) behaviour). */
int msclip_set_weight(msclip_handle h, const char* key, const void* data, int dtype, int ndim, const int64_t* shape);
/* Validate that every key arrived, fold BatchNorm, fold the 1/8 query scaling, re-pack to bf16 K-major. */
int msclip_finalize_weights(msclip_handle h, void* stream);
/* exp(logit_scale) of the loaded weights (M.py:3136) */
int msclip_logit_scale_exp(msclip_handle h, float* out);

/* image: [batch, 3, R, R] NCHW of `image_dtype`; out: [batch, embed_dim] f32. */
int msclip_encode_image(msclip_handle h, const void* image, int image_dtype, int batch, float* out, int normalize,
                        void* stream);
/* Input prefetch (the DataLoader's `.cuda(non_blocking=True)`, tools/zero_shot.py:262): start the host->device copy
 * of a batch of images now, on a private copy stream.  A later msclip_encode_image / msclip_forward_loss call that
 * is given the SAME host pointer and batch consumes the staged copy instead of transferring again, so the copy
 * of step i+1 overlaps the compute of step i.  Two slots (double buffering); the host buffer must stay unchanged
 * until it has been consumed. */
int msclip_stage_images(msclip_handle h, const void* image_host, int image_dtype, int batch, void* stream);
/* tokens: [batch, context_length] int64; out: [batch, embed_dim] f32.  Out-of-range ids are an error.
 * The text tower is causal and pools the row at argmax(tokens) (M.py:2965-2971, 3057-3060), so positions after a
 * sequence's EOT token cannot influence its output: by default the call runs the tower only over the longest live
 * prefix of the batch (bit-identical results; costs one 4-byte device->host read per call).
 * msclip_set_text_trim(h, 0) switches this off; msclip_forward / msclip_forward_loss never trim (no host sync). */
int msclip_encode_text(msclip_handle h, const int64_t* tokens, int batch, float* out, int normalize, void* stream);
int msclip_set_text_trim(msclip_handle h, int enable);
/* logits[n_img, n_txt] = scale * img_feat . txt_feat^T (f32 in, f32 out; split-bf16 tensor-core product). */
int msclip_similarity_logits(msclip_handle h, const float* img_feat, int n_img, const float* txt_feat, int n_txt,
                             float scale, float* logits, void* stream);
/* ---- zero-shot evaluation fast path (what tools/zero_shot.py:121-132, 265-266, 150-163 compute) ----
 * msclip_zeroshot_classifier: tokens [n_classes * n_templates, context_length] int64, class-major (host or device) ->
 * weights [n_classes, embed_dim] f32 (host or device): per class the mean over its templates of the normalised prompt
 * embeddings, re-normalised (the tool's zeroshot_weights, transposed).  One call instead of n_classes encode_text calls:
 * prompts are sorted by live length and encoded in chunks, each only over its longest live prefix (bit-identical
 * embeddings).  Synchronises the stream.
 * msclip_zeroshot_predict: logits = scale * img_feat . weights^T (scale = 100 in the tool) and the top-k (k <= 8) class
 * indices per image, best first, all on the device (no per-batch host round trip); logits_out may be NULL. */
int msclip_zeroshot_classifier(msclip_handle h, const int64_t* tokens, int n_classes, int n_templates, float* weights_out,
                               void* stream);
int msclip_zeroshot_predict(msclip_handle h, const float* img_feat, int n_img, const float* weights, int n_classes, float scale,
                            int topk, int32_t* topk_out, float* logits_out, void* stream);
/* CLIP.forward for one process: logits[batch, batch] = exp(logit_scale) * I . T^T. */
int msclip_forward(msclip_handle h, const void* image, int image_dtype, const int64_t* tokens, int batch,
                   float* logits, void* stream);

/* ---- sharded contrastive loss (data parallel, one process per GPU) -----------------------------------
 * Each rank owns `b_local` pairs; global row index = rank * b_local + i (rank order of gather_tensors,
 * comm.py:150-153).  The embedding exchange is NOT a collective: every rank publishes its normalised
 * embeddings (IEEE fp16 in both library builds: |x| <= 1, 11-bit significand) in a library-owned, IPC-exported
 * buffer and the loss kernel of every other rank loads them over NVLink; the stream waits for the peers' publish
 * flags with stream memory operations (no kernel spins, no time-out: a late rank is waited for like in a collective).
 * At most 64 ranks of one NVLink domain (CUDA IPC); beyond that use gather_tensors + msclip_similarity_logits.
 * Setup (once):
 *   1. msclip_comm_init(h, rank, world, max_b_local)              allocates the exchange buffer
 *   2. msclip_comm_export(h, handle_bytes[64])                    -> exchange the 64-byte handles out of band
 *   3. msclip_comm_import(h, all_handles[world*64])               (torch.distributed / MPI / files ...)
 * world == 1 needs none of this. */
int msclip_comm_init(msclip_handle h, int rank, int world, int max_b_local);
int msclip_comm_export(msclip_handle h, void* handle_out_64);
int msclip_comm_import(msclip_handle h, const void* handles_world_x_64);
/* Same-process peers (one process driving several GPUs with peer access enabled, or several handles on one GPU):
 * instead of steps 2-3 hand every handle the others' exchange buffers as plain device pointers. */
int msclip_comm_buffer(msclip_handle h, void** base_out);
int msclip_comm_import_pointers(msclip_handle h, void* const* bases_world);

/* Loss of the features produced by the LAST msclip_encode_image / msclip_encode_text calls of this handle
 * (kept on the device in bf16), batch b_local each.  partial_out[2] (host or device) receives this rank's
 * partial sums: loss = sum over ranks (partial[0] + partial[1]) / (2 * world * b_local).
 * With world == 1 loss_out (optional, may be NULL) receives the final loss. */
int msclip_contrastive_loss(msclip_handle h, int b_local, float scale, float* partial_out, float* loss_out,
                            void* stream);
/* The same loss for embeddings computed elsewhere: img_feat / txt_feat [b_local, embed_dim] f32, already normalised
 * (what gather_tensors would be given, M.py:3139-3140); device or host pointers. */
int msclip_contrastive_loss_features(msclip_handle h, const float* img_feat, const float* txt_feat, int b_local,
                                     float scale, float* partial_out, float* loss_out, void* stream);
/* Backward of the LAST msclip_contrastive_loss / msclip_forward_loss call of this handle with respect to this rank's
 * normalised embeddings (first piece): d_img_feat / d_txt_feat [b_local, embed_dim] f32 (host or
 * device) receive d loss / d image_features and d loss / d text_features of the global symmetric cross-entropy.  Only the
 * local shard receives a gradient, as with gather_tensors (lib/utils/comm.py:151-152); every rank must call it (the
 * column-wise softmax needs the peers' row log-sum-exps: a second in-kernel peer read, no collective). */
int msclip_contrastive_loss_backward(msclip_handle h, float* d_img_feat, float* d_txt_feat, void* stream);
/* ---- training (the reference ships no backward - what a training loop around CLIP.forward would get
 * from torch.autograd: loss.backward() through lib/utils/comm.py:151-152, M.py:3126-3155, 3043-3079, 2621-2697) ------------
 * msclip_train_enable(h, 1), called BEFORE msclip_finalize_weights (a finalized handle must be re-sent its weights),
 * makes the handle (i) keep transposed copies of the block weights, (ii) allocate one fp32 gradient buffer per trainable
 * state-dict key (aliased keys share a buffer, like the parameters, M.py:2786-2830) and (iii) keep the fp32 input of
 * every block during encode_image / encode_text (calls of at most 4096 rows).  bf16 build only.
 * Trainable: both heads, every ResidualAttentionBlock, ln_pre / ln_post / ln_final / ln_adapt, token / positional / class
 * embeddings.  The convolutional front (stem, parallel branch, adapter convolutions, BatchNorms) is frozen: the gradient
 * flows THROUGH the adapters' bottom path but their convolution parameters receive none. */
int msclip_train_enable(msclip_handle h, int enable);
/* Accumulate d loss / d parameter for the LAST taped msclip_encode_image / msclip_encode_text (or msclip_forward_loss /
 * msclip_encode_pairs) call into the gradient buffers, given d loss / d features [batch, embed_dim] f32 on the device
 * (of the normalised features when the call normalised) - e.g. the output of msclip_contrastive_loss_backward.
 * Either gradient may be NULL (that tower is skipped). */
int msclip_backward(msclip_handle h, const float* d_img_feat, const float* d_txt_feat, void* stream);
int msclip_zero_grad(msclip_handle h, void* stream);
/* Normalised features [b, embed_dim] f32 of the last taped tower call (modality 0 = image, 1 = text), re-read from the
 * tape - e.g. for d loss / d logit_scale = sum_i I_i . dI_i after msclip_forward_loss, which returns no features. */
int msclip_taped_features(msclip_handle h, int modality, float* out, int b, void* stream);
/* The gradient buffers: count, and key / device pointer / element count of entry i (layout = the parameter's). */
int msclip_num_grads(msclip_handle h);
int msclip_grad_info(msclip_handle h, int index, const char** key, float** dev_ptr, int64_t* numel);
/* Re-pack ONE trainable parameter from its updated fp32 master copy `dev_ptr` (after an optimiser step): no allocation,
 * no host round trip; keys of the frozen front are rejected. */
int msclip_update_weight(msclip_handle h, const char* key, const float* dev_ptr, void* stream);

/* ---- input pipeline on the GPU ---------------------------------------------------------------
 * The tool's transform, tools/zero_shot.py:202-207: Resize(out_size, BICUBIC) -> CenterCrop(out_size) -> ToTensor ->
 * Normalize(mean, std), for n decoded RGB images (uint8, HWC) of arbitrary sizes, BIT-EXACT with torchvision on PIL images
 * (Pillow's two-pass 22-bit fixed-point resampling, restated in csrc/preprocess.cu).  `pixels`: the images back to back in
 * host or device memory, image i at byte offset offsets[i] with heights[i] x widths[i] x 3 bytes (offsets / heights /
 * widths are host arrays).  `out` [n, 3, out_size, out_size] in out_dtype (MSCLIP_F32 / _BF16 / _F16) - ready for
 * msclip_encode_image - and / or `out_u8` [n, out_size, out_size, 3] (the resized, cropped bytes); device memory; either
 * may be NULL. */
int msclip_preprocess_images(msclip_handle h, const uint8_t* pixels, const int64_t* offsets, const int* heights, const int* widths,
                             int n, int out_size, const float* mean3, const float* std3, void* out, int out_dtype, uint8_t* out_u8,
                             void* stream);

/* ---- native BPE tokenizer (host code, no GPU) -----------------------------------------------------
 * lib/dataset/languages/simple_tokenizer.py:64-166 restated: whitespace_clean -> lower -> CLIP pattern -> byte alphabet ->
 * greedy pair merges -> ids.  `merges_text` = the decompressed text of the reference's own bpe_simple_vocab_16e6.txt.gz
 * (not shipped here).  basic_clean (ftfy + html.unescape, :53-56) is left to the caller. */
int msclip_tokenizer_create(const char* merges_text, int64_t length, void** tokenizer_out);
int msclip_tokenizer_destroy(void* tokenizer);
int msclip_tokenizer_info(void* tokenizer, int* vocab_size, int* sot_token, int* eot_token);
/* SimpleTokenizer.encode of one cleaned UTF-8 text: returns the number of ids (or -1), writes at most `capacity` of them. */
int64_t msclip_tokenizer_encode(void* tokenizer, const char* text, int64_t length, int32_t* ids_out, int64_t capacity);
/* SimpleTokenizer.tokenize for n texts (text i = texts[offsets[i] .. offsets[i + 1])): out [n, context_length] int64 =
 * [SOT] + ids + [EOT] truncated to context_length and zero padded; `threads` <= 0: all host cores. */
int msclip_tokenizer_tokenize(void* tokenizer, const char* texts, const int64_t* offsets, int n, int context_length, int64_t* out,
                              int threads);

/* Micro-batching (BASELINE.json: global batch 32 768 on 1 / 2 / 4 GPUs): run both towers
 * for b_micro pairs and keep their normalised embeddings as rows [row_offset, row_offset + b_micro) of this rank's
 * shard of the NEXT msclip_contrastive_loss.  row_offset must be 0 (new shard) or the number of rows encoded so far;
 * the shard must fit the exchange buffer (msclip_comm_init(h, rank, world, b_local) first, also with world == 1).
 * n_micro calls followed by msclip_contrastive_loss(h, n_micro * b_micro, ...) give the loss over the whole global
 * batch - the same value as one msclip_forward_loss over all rows at once. */
int msclip_encode_pairs(msclip_handle h, const void* image, int image_dtype, const int64_t* tokens, int b_micro,
                        int row_offset, void* stream);
/* msclip_encode_pairs(row_offset = 0) + msclip_contrastive_loss with scale = exp(logit_scale). */
int msclip_forward_loss(msclip_handle h, const void* image, int image_dtype, const int64_t* tokens, int b_local,
                        float* partial_out, float* loss_out, void* stream);

/* Number of kernels this handle has launched since creation (for the benchmark's gpu_launches claim). */
int64_t msclip_launch_count(msclip_handle h);
/* Bytes of device memory currently owned by the handle (weights + workspace). */
int64_t msclip_device_bytes(msclip_handle h);

#ifdef __cplusplus
}
#endif
#endif /* MSCLIP_B200_H_ */
