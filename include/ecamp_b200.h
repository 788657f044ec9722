/*
 * ecamp_b200 — C ABI of the B200-native ECAMP pre-training hot path.
 *
 * The reference (ToniChopp/ECAMP) has no FFI or operator registry for this path: the boundary is
 * the nn.Module returned by `module.model_ecamp.ecamp(**kwargs)`
 * (ECAMP/Pre-training/module/model_ecamp.py:328-333, called at
 * ECAMP/Pre-training/main_pretrain.py:141,233).  The Python mirror of that module
 * (ecamp_b200/model_ecamp.py) binds the entry points below with ctypes; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.  Each entry point cites the reference code whose
 * library kernels it replaces (paths relative to ECAMP/Pre-training/).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless documented otherwise; `stream` is a cudaStream_t
 *     passed as void* (0 = legacy default stream);
 *   - every function returns 0 on success and a negative code on failure, never throws, never
 *     exits; `ecamp_last_error()` returns a thread-local description of the last failure;
 *   - work is enqueued asynchronously on `stream`; asynchronous CUDA errors surface at the
 *     caller's next synchronisation exactly as for torch ops;
 *   - the caller (PyTorch) owns every buffer.  A context only remembers the base pointers it was
 *     bound to; rebind after `.to()`, `load_state_dict` into new storage, etc.
 */
#ifndef ECAMP_B200_H_
#define ECAMP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECAMP_ABI_VERSION 5  /* 5: + ecamp_ce_rows_bias, ecamp_set_side_stream and three measurement switches (additive) */
#if defined(__GNUC__)
#define ECAMP_API __attribute__((visibility("default")))
#else
#define ECAMP_API
#endif

ECAMP_API int ecamp_abi_version(void);
ECAMP_API const char* ecamp_last_error(void);
/* number of CUDA kernels this library has launched in this process (bench.py reports the delta per timed region) */
ECAMP_API int64_t ecamp_launch_count(void);

/* ==========================================================================================
 * 1. Operators (each usable on its own; the parity tests call these)
 * ========================================================================================== */

/* GEMM (tcgen05 / TMEM / TMA): replaces every nn.Linear on the path — timm Block qkv/proj/fc1/fc2
 * (module/model_ecamp.py:66-68,80-82), decoder_embed / decoder_pred / bert_mlp (:73,85,100),
 * HF BertSelfAttention/BertSelfOutput/BertIntermediate/BertOutput (module/context_fusion.py:12-19),
 * the LM head (module/bert_modeling.py:209) — forward, dgrad and wgrad. */
enum {
  ECAMP_GEMM_GELU = 1,    /* v = gelu(bf16(v)), rounded pre-activation stored to aux_out          */
  ECAMP_GEMM_DGELU = 2,   /* v *= gelu'(aux_in[m, n])                                              */
  ECAMP_GEMM_DROPOUT = 4, /* inverted dropout with a Philox mask keyed by (seed, site, m * N + n)  */
  ECAMP_GEMM_AUX_GRAD = 8 /* with GELU: aux_out = bf16(gelu'(pre-activation)); with DGELU: v *= aux_in (that
                           * stored derivative) - what the step runtime uses for fc1 / fc2                  */
};
typedef struct ecamp_epilogue {
  const float* bias;     /* [N] or NULL                                   */
  const void* aux_in;    /* bf16 [M, ld_aux] or NULL                      */
  void* aux_out;         /* bf16 [M, ld_aux] or NULL                      */
  int32_t ld_aux;
  const float* residual; /* fp32 [M, ld_res] or NULL; may alias out_f32   */
  int32_t ld_res;
  float* out_f32;        /* fp32 [M, ld_f32] or NULL                      */
  int32_t ld_f32;
  void* out_bf16;        /* bf16 [M, ld_bf16] or NULL                     */
  int32_t ld_bf16;
  int32_t flags;
  float drop_p;
  uint64_t seed;
  uint64_t site;
  float* colsum_out;     /* fp32 [N] or NULL: += column sums of the emitted bf16 values (atomic adds): the bias
                          * gradient of the Linear that consumes out_bf16.  Needs out_bf16; excludes split-K.   */
} ecamp_epilogue;
/* D[M,N] = epilogue(A . B^T).  a_mn / b_mn = 0: operand stored [rows, contraction] (contraction
 * contiguous); = 1: stored [contraction, rows].  tile_n = 0 lets the library choose. */
/* 0 = automatic, 1 = single-CTA tcgen05 kernel only, 2 = CTA-pair (cta_group::2) kernel always (testing knob) */
ECAMP_API void ecamp_gemm_set_cta_pair(int32_t mode);
/* 1 = fp32-output epilogues of the CTA-pair kernel go through TMA (residual tiles loaded and results stored with
 * cp.async.bulk.tensor); 0 (default) = LSU epilogue through the per-warp transpose tile.  Same results either way. */
ECAMP_API void ecamp_gemm_set_tma_epilogue(int32_t on);
/* bf16-output epilogues written straight from the TMEM row-per-thread layout with 256-bit stores instead of through the
 * per-warp shared-memory transpose tile: 0 = never, 1 (default) = the plain bias -> bf16 epilogue, 2 = also GELU + stored
 * GELU' and x stored GELU' with column sums (measured slower).  Same results. */
ECAMP_API void ecamp_gemm_set_direct_epilogue(int32_t on);
ECAMP_API int ecamp_gemm_bf16(const void* A, int32_t lda, int32_t a_mn, const void* B, int32_t ldb, int32_t b_mn,
                              int32_t M, int32_t N, int32_t K, const ecamp_epilogue* ep, int32_t tile_n, void* stream);

/* The same product from fp32 operands in the fp32-accurate parity mode: each operand is split into three bf16 planes,
 * the six significant plane pairs are accumulated by the same tcgen05 kernel, the epilogue runs in fp32 (exact erf GELU).
 * In `ep`, aux_in / aux_out / out_bf16 point to FP32 arrays here.  `ws`: ecamp_gemm_fp32_ws_bytes(M, N, K) bytes, 256-byte
 * aligned. */
ECAMP_API int64_t ecamp_gemm_fp32_ws_bytes(int32_t M, int32_t N, int32_t K);
ECAMP_API int ecamp_gemm_fp32(const float* A, int32_t lda, int32_t a_mn, const float* B, int32_t ldb, int32_t b_mn,
                              int32_t M, int32_t N, int32_t K, const ecamp_epilogue* ep, void* ws, int64_t ws_bytes,
                              void* stream);

/* random_masking on caller-supplied noise (module/model_ecamp.py:168-193): stable ascending argsort.
 * ids_restore / ids_keep int64 as in the reference, mask fp32 {0,1} (1 = removed). */
ECAMP_API int ecamp_random_masking(const float* noise, int32_t B, int32_t L, int32_t len_keep, int64_t* ids_restore,
                                   int64_t* ids_keep, float* mask, void* scratch_i32 /* (B*L + B*len_keep) int32 */,
                                   void* stream);

/* ---- image half of the loader (module/pretrain_datasets.py:47-52): RandomResizedCrop(448, scale 0.2-1, BICUBIC) +
 * RandomHorizontalFlip of a decoded 8-bit grayscale frame.  The crop box and the flip are drawn on the host with torch's
 * generator in torchvision's order (ecamp_b200/image_pipeline.py); the crop boxes of a batch are packed back to back in
 * `crops`, image b being [h, w] row-major at desc[b].src_off.  The resampling is Pillow's (two passes with an 8-bit
 * intermediate, antialiased bicubic a = -0.5, 22-bit fixed-point weights), so the bytes equal what the reference transform
 * produces on the PIL image; out is [B, out, out] uint8 - ecamp_image_u8_normalize (or the uint8 input of the step) does
 * Grayscale(3) + ToTensor + Normalize.  tmp_off: offsets into the intermediate image buffer ([h, out] bytes per image,
 * `tmp_bytes` in total); kmax = max over the batch of ecamp_image_resample_kmax(max(h, w), out). */
typedef struct ecamp_crop_desc {
  int64_t src_off, tmp_off;
  int32_t h, w, flip, pad;
} ecamp_crop_desc;
ECAMP_API int32_t ecamp_image_resample_kmax(int32_t in_size, int32_t out);
ECAMP_API int64_t ecamp_image_resized_crop_ws_bytes(int32_t B, int32_t out, int32_t kmax, int64_t tmp_bytes);
ECAMP_API int ecamp_image_resized_crop(const uint8_t* crops, const ecamp_crop_desc* desc_dev, int32_t B, int32_t hmax,
                                       int32_t out, int32_t kmax, void* ws, int64_t ws_bytes, int64_t tmp_bytes,
                                       uint8_t* dst, void* stream);
/* the same arithmetic on the host for ONE crop box: test infrastructure (checked against Pillow without a GPU), never
 * called by the product */
ECAMP_API int ecamp_image_resized_crop_host(const uint8_t* crop, int32_t h, int32_t w, int32_t flip, int32_t out,
                                            uint8_t* dst);

/* bicubic 448 -> 224 of torchvision Resize (module/model_ecamp.py:318), result in patch layout
 * tgt[b, l, (p*16+q)*3+c] == patchify(resized) with p = 16. */
/* Tail of the image transform of module/pretrain_datasets.py:47-52 on the GPU: Grayscale(3) + ToTensor + Normalize of the
 * loader's 8-bit grayscale crop, out[n, c, :] = (gray[n, :] / 255 - mean) / std for c = 0..2, bit-exact with the CPU
 * transform.  A step then copies 1 byte per pixel to the device instead of 12.  pixels_per_image % 16 == 0. */
ECAMP_API int ecamp_image_u8_normalize(const uint8_t* gray, int64_t n_images, int64_t pixels_per_image, float mean,
                                       float std_, float* out /* [n_images, 3, pixels_per_image] */, void* stream);
ECAMP_API int ecamp_resize_patchify(const float* big, int32_t B, int32_t side_in, float* tgt, void* stream);

/* LayerNorm forward / backward (fp32 in, bf16 and/or fp32 out). */
ECAMP_API int ecamp_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int32_t M,
                                  int32_t D, void* out_bf16, float* out_f32, float* mean, float* rstd, void* stream);
/* dgamma / dbeta / colsum_out (each may be NULL) are ADDED to with atomics; accumulate == 0 zeroes them first.
 * colsum_out [D] = column sums of dx_bf16 (the bias gradient of the Linear that produced the LayerNorm input's
 * branch).  ws is unused since ABI 2 (kept so that callers need not change their allocation code). */
ECAMP_API int ecamp_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd,
                                  const float* gamma, int32_t M, int32_t D, const float* addend, float* dx_f32,
                                  void* dx_bf16, float* dgamma, float* dbeta, float* colsum_out, int32_t accumulate,
                                  float* ws /* unused */, void* stream);
ECAMP_API size_t ecamp_layernorm_ws_floats(void);
/* Measurement switch: 1 selects the slab form of the backward kernel (740 CTAs, per-warp shared-memory column sums), 0 (default)
 * the staged form (one CTA per SM, rows staged by cp.async.bulk, column sums in registers).  Same results up to fp32 summation order. */
ECAMP_API void ecamp_layernorm_set_bwd_slab(int32_t on);

/* fused attention (timm Attention; HF BertSelfAttention eager path incl. key-padding mask and
 * probability dropout; cross-attention of module/context_fusion.py:45-53). */
typedef struct ecamp_attn {
  const void *q, *k, *v;      /* bf16; head h at columns [h*D, (h+1)*D) */
  int32_t ldq, ldk, ldv;
  void* o;                    /* bf16 */
  int32_t ldo;
  float* lse;                 /* [B, H, Sq] */
  const int64_t* key_mask;    /* [B, Sk] nonzero = attend, or NULL */
  int32_t B, H, Sq, Sk, D;
  float scale;
  float drop_p;
  uint64_t seed, site;
  const void* d_o;            /* backward: bf16 */
  int32_t ld_do;
  float* delta;               /* backward scratch [B, H, Sq] */
  void *dq, *dk, *dv;         /* backward outputs, bf16 */
  int32_t lddq, lddk, lddv;
  float *cs_q, *cs_k, *cs_v;  /* backward, each [H*D] fp32 or NULL: += column sums of dq / dk / dv over all rows (atomic
                               * adds): the bias gradients of the q / k / v projections                                */
} ecamp_attn;
/* 1 (default): head_dim 64 / 128 problems that fit use the tcgen05 / TMEM kernels; 0: always the mma.sync kernels */
ECAMP_API void ecamp_attention_set_tcgen05(int32_t on);
ECAMP_API int ecamp_attention_fwd(const ecamp_attn* a, void* stream);
ECAMP_API int ecamp_attention_bwd(const ecamp_attn* a, void* stream);
/* probs [B, H, Sq, Sk] fp32 = exp(scale * q . k - lse), 0 at masked keys; needs q, k, lse (of a preceding
 * ecamp_attention_fwd), the shape and scale.  Replaces `output_attentions=True` of the cross-attention in
 * Visualization/module/context_fusion.py:45-57 (inference tool; the training kernels never materialise P). */
ECAMP_API int ecamp_attention_probs(const ecamp_attn* a, float* probs, void* stream);

/* losses.  pred is [B, 197, 768] fp32 (row 0 of each sample = cls, ignored), tgt [B, 196, 768]. */
ECAMP_API int ecamp_mim_loss(const float* pred, const float* tgt, const float* mask, int32_t B, float* loss,
                             float* ws /* B*196 floats */, void* stream);
ECAMP_API int ecamp_sr_loss_fwd(const float* pred, const float* big, const int64_t* column, const int64_t* row,
                                const float* w1, const float* b1, const float* w2, const float* b2, int32_t B,
                                float* loss, float* ws /* ecamp_sr_ws_floats(B) */, void* stream);
ECAMP_API int ecamp_sr_loss_bwd(const float* pred, const float* big, const int64_t* column, const int64_t* row,
                                const float* w1, const float* b1, const float* w2, const float* b2, int32_t B,
                                const float* g_res, float* d_u /* [B,3,448,448] */, float* d_conv /* 168 */,
                                int32_t accumulate, float* ws, void* stream);
ECAMP_API size_t ecamp_sr_ws_floats(int32_t B);
ECAMP_API int ecamp_pred_grad(const float* pred, const float* tgt, const float* mask, const float* d_u,
                              const float* g_mim, int32_t B, void* d_pred_bf16 /* [B,197,768] */, void* stream);
/* weighted cross-entropy on materialised bf16 logits [rows, V] (module/bert_modeling.py:211-217);
 * with write_grad the logits are overwritten by (softmax - onehot) * w * g / total_rows. */
ECAMP_API int ecamp_ce_rows(void* logits_bf16, int32_t ld, int32_t rows, int32_t V, const int64_t* labels,
                            const float* weights, float* row_loss, const float* g, float inv_total_rows,
                            int32_t write_grad, void* stream);
/* The same with the gradient pass always on and the gradient of cls.predictions.bias (module/bert_modeling.py:208-210: the
 * decoder Linear's bias) folded in: bias_grad [V] (may be NULL) is ADDED to with the column sums of the written gradient. */
ECAMP_API int ecamp_ce_rows_bias(void* logits_bf16, int32_t ld, int32_t rows, int32_t V, const int64_t* labels,
                                 const float* weights, float* row_loss, const float* g, float inv_total_rows,
                                 float* bias_grad, void* stream);
/* Measurement switch: 1 (default) = persistent kernel, rows staged by cp.async.bulk, column sums in registers; 0 = one CTA per
 * row followed by a column-sum pass over the written gradient. */
ECAMP_API void ecamp_ce_set_fused(int32_t on);
/* Measurement switch: 1 (default) = the SR backward kernel limits every stage of a tile to what the loss window can reach
 * (tiles that merely border the window become cheap); 0 = every stage on the whole 36 x 36 neighbourhood. */
ECAMP_API void ecamp_sr_set_window_skip(int32_t on);
/* 1 (default; ECAMP_SIDE_WGRAD=0 in the environment also turns it off): ecamp_ctx_backward queues the weight-gradient GEMMs of the
 * transformer blocks and of the vocabulary head on an internal side stream, forked from / joined to the caller's stream with
 * events inside the call (every stage ends joined: when the call returns, all work is ordered on the caller's stream as
 * before).  0: everything on the caller's stream.  2 (tests): as 1, with every side-stream GEMM held back by ~0.2 ms so that a
 * missing ordering edge shows up as wrong gradients.  Takes effect at the next backward call. */
ECAMP_API void ecamp_set_side_stream(int32_t on);

/* ==========================================================================================
 * 2. The step runtime (what the nn.Module calls): parameter table, context, forward / backward /
 *    optimizer.  Replaces ECAMP.forward (module/model_ecamp.py:303-325), autograd backward
 *    (util/misc.py:258) and torch.optim.AdamW.step (main_pretrain.py:254; util/misc.py:267).
 * ========================================================================================== */
ECAMP_API int32_t ecamp_param_count(void);
ECAMP_API const char* ecamp_param_name(int32_t i);     /* reference state_dict key            */
ECAMP_API int64_t ecamp_param_numel(int32_t i);
ECAMP_API int32_t ecamp_param_decay(int32_t i);        /* timm add_weight_decay group         */
ECAMP_API int64_t ecamp_param_grad_offset(int32_t i);  /* floats into the flat grad buffer    */
ECAMP_API int64_t ecamp_grad_floats(void);
ECAMP_API int64_t ecamp_shadow_bytes(void);            /* GEMM copies (sized for either precision) + fused fp32 biases */
ECAMP_API int64_t ecamp_adam_table_bytes(void);
ECAMP_API int64_t ecamp_adam_chunk_bytes(void);

typedef struct ecamp_ctx ecamp_ctx;
ECAMP_API int ecamp_ctx_create(ecamp_ctx** out);
ECAMP_API void ecamp_ctx_destroy(ecamp_ctx* ctx);
/* params_host: HOST array of ecamp_param_count() device pointers in table order.  grads / adam_m /
 * adam_v: flat fp32 buffers of ecamp_grad_floats() (adam_* may be NULL for inference). */
ECAMP_API int ecamp_ctx_bind(ecamp_ctx* ctx, float* const* params_host, int32_t n, float* grads, float* adam_m,
                             float* adam_v, void* shadows, const float* pos_embed, const float* decoder_pos_embed,
                             void* adam_table, void* adam_chunks);
typedef struct ecamp_shape {
  int32_t B, T, len_keep, has_big, ce_rows;
} ecamp_shape;
ECAMP_API int64_t ecamp_workspace_bytes(const ecamp_shape* s);   /* production precision */
/* Precision of the step: 0 = production (bf16 GEMM / attention operands with fp32 accumulation, the reference's autocast
 * in bf16), 1 = fp32-accurate parity mode: the SAME schedule and kernels instantiated with fp32 activations, every GEMM on
 * the same tcgen05 kernel with error-compensated bf16 x 3 split operands (gemm.cu: gemm_hp), attention in fp32.  It exists
 * to pin the algebra of the step against the fp32 oracle at 1e-5 (north_star's fp32 tolerance); it is ~10 x slower.
 * Changing the precision resets the context: ecamp_ctx_bind() and ecamp_ctx_set_workspace() must follow; the workspace
 * size of the current precision comes from ecamp_ctx_workspace_bytes(). */
ECAMP_API int ecamp_ctx_set_precision(ecamp_ctx* ctx, int32_t fp32_accurate);
ECAMP_API int64_t ecamp_ctx_workspace_bytes(ecamp_ctx* ctx, const ecamp_shape* s);
ECAMP_API int ecamp_ctx_set_workspace(ecamp_ctx* ctx, void* ws, int64_t bytes, const ecamp_shape* s);
/* re-derive the bf16 / fused copies from the fp32 parameters (after load_state_dict, an external
 * optimizer step, ...). */
ECAMP_API int ecamp_refresh_shadows(ecamp_ctx* ctx, void* stream);

typedef struct ecamp_batch {
  const float* image;             /* [B,3,448,448] (has_big) or [B,3,224,224]                         */
  const int64_t* ids;             /* [B,T]  batch["ids"]                                              */
  const int64_t* labels;          /* [B,T]  batch["labels"]                                           */
  const int64_t* attention_mask;  /* [B,T]                                                            */
  const int64_t* type_ids;        /* [B,T]                                                            */
  const float* weights;           /* [B,T]                                                            */
  const int64_t* column;          /* [B] (has_big)                                                    */
  const int64_t* row;             /* [B] (has_big)                                                    */
  const float* noise;             /* [B,196] uniform noise of random_masking                          */
} ecamp_batch;
enum { ECAMP_FWD_TRAIN = 1, ECAMP_FWD_DEFER_MLM = 2 };
/* losses: 3 floats (mim, res, mlm).  mask [B,196] fp32, ids_restore [B,196] / ids_keep [B,len_keep]
 * int64 are optional outputs. */
ECAMP_API int ecamp_forward(ecamp_ctx* ctx, const ecamp_batch* batch, int32_t flags, float drop_p, uint64_t seed,
                            float* losses, float* mask, int64_t* ids_restore, int64_t* ids_keep, void* stream);
ECAMP_API int32_t ecamp_backward_stage_count(void);
/* [begin, end) floats of the flat gradient buffer that are final once `stage` has run */
ECAMP_API int ecamp_backward_stage_range(int32_t stage, int64_t* begin, int64_t* end);
/* g3: the three upstream gradients d(total)/d(mim, res, mlm).  stage = -1 runs all stages in order. */
ECAMP_API int ecamp_backward(ecamp_ctx* ctx, const float* g3, int32_t accumulate, int32_t stage, void* stream);
/* Stages [first_stage, end_stage) in one call: every slice those stages announce is final when the call returns.  A caller that
 * only needs finality at a few points (the bucket boundaries of a gradient all-reduce) gives the library room to overlap the
 * tail of one transformer block with the head of the next (see ecamp_set_side_stream).  Stages must be run in order, each
 * exactly once per backward pass; first_stage == 0 starts the pass (zeroing, as ecamp_backward with stage 0). */
ECAMP_API int ecamp_backward_stages(ecamp_ctx* ctx, const float* g3, int32_t accumulate, int32_t first_stage, int32_t end_stage,
                                    void* stream);
ECAMP_API int ecamp_adamw_step(ecamp_ctx* ctx, float lr, float beta1, float beta2, float eps, float weight_decay,
                               int32_t step, float grad_scale, void* stream);
/* the same with one learning rate per timm add_weight_decay group (PT/main_pretrain.py:253: [no-decay, decay]) */
ECAMP_API int ecamp_adamw_step_groups(ecamp_ctx* ctx, float lr_decay, float lr_no_decay, float beta1, float beta2, float eps,
                                      float weight_decay, int32_t step, float grad_scale, void* stream);
/* The same update for the tensors whose gradients occupy flat[grad_begin, grad_end) only (both on tensor boundaries: a
 * backward stage range or a union of consecutive ones).  A small-footprint launch (128 threads, no shared memory) meant to run
 * on another stream WHILE later backward stages still compute: once a slice is final (and all-reduced), nothing in this backward
 * pass reads those parameters or their bf16 copies again.  Every parameter must be covered exactly once per optimizer step. */
ECAMP_API int ecamp_adamw_step_range(ecamp_ctx* ctx, float lr_decay, float lr_no_decay, float beta1, float beta2, float eps,
                                     float weight_decay, int32_t step, float grad_scale, int64_t grad_begin, int64_t grad_end,
                                     void* stream);
/* after ecamp_forward: probs [B, 6, T, keep] fp32 of the fusion layer's text -> image cross-attention, columns in
 * ids_keep order — what Visualization/module/model_ecamp.py:308-319 returns at mask_ratio = 0 */
ECAMP_API int ecamp_cross_attention_probs(ecamp_ctx* ctx, float* probs, void* stream);
/* named intermediate buffers for the parity tests ("latent", "pred", "tgt", ...), NULL if unknown */
ECAMP_API const void* ecamp_debug_buffer(ecamp_ctx* ctx, const char* name);

/* ---- fine-tune classification (ECAMP/Fine-tuning/Classification/models_vit.py:60-98 with global_pool, train.py:438-465) ----
 * The same context (ecamp_ctx_create / ecamp_ctx_bind): the encoder parameters are the table entries patch_embed.*,
 * cls_token and blocks.* (the remaining entries of the pre-training table may point at any scratch buffer of at
 * least their size); their gradients land in the flat gradient buffer at the same offsets.  pos_embed (learnable
 * here), fc_norm and the head come in through ecamp_cls_io; the head is padded to 16 outputs. */
typedef struct ecamp_cls_io {
  const float* image;        /* [B, 3, 224, 224]                                                   */
  const float* pos_embed;    /* [197, 768]                                                         */
  const float* fc_norm_w;    /* [768]                                                              */
  const float* fc_norm_b;    /* [768]                                                              */
  const void* head_w16;      /* bf16 [16, 768], rows >= num_classes zero                           */
  const float* head_b;       /* [16]                                                               */
  const float* dp_scale;     /* DropPath mask / keep_prob, [12][2][B], or NULL (eval / rate 0)     */
  float* logits;             /* out [B, 16]                                                        */
  const float* d_logits;     /* backward: [B, 16], columns >= num_classes zero                     */
  float* g_pos_embed;        /* backward outputs: [197*768], [768], [768], [16*768], [16]          */
  float* g_fc_norm_w;
  float* g_fc_norm_b;
  float* g_head_w;
  float* g_head_b;
} ecamp_cls_io;
ECAMP_API int64_t ecamp_cls_workspace_bytes(int32_t B);
ECAMP_API int ecamp_cls_set_workspace(ecamp_ctx* ctx, void* ws, int64_t bytes, int32_t B);
ECAMP_API int ecamp_cls_forward(ecamp_ctx* ctx, const ecamp_cls_io* io, void* stream);
ECAMP_API int ecamp_cls_backward(ecamp_ctx* ctx, const ecamp_cls_io* io, int32_t accumulate, void* stream);

/* ---- host-side report masking and loss re-weighting (HOST pointers, no CUDA) ----
 * Native form of `ContextBertDataset._context_mask` and of the template re-weighting in `__getitem__`
 * (ECAMP/Pre-training/module/pretrain_datasets.py:60-110, 141-184): same decisions in the same order.  The random numbers are
 * passed in pre-drawn: draw exactly ecamp_text_mask_draw_count(...) values with `random.random()` and Python's generator
 * ends where the reference leaves it.  is_sub[v] / is_entity[v]: word piece v starts with "##" / is one of the 44 entity
 * words (:17-22).  mask_pos must hold T entries.  Bit-exact against tests/golden/text_masking.json. */
ECAMP_API int32_t ecamp_text_mask_draw_count(const int64_t* ids, int32_t T, const uint8_t* is_sub, const uint8_t* is_entity,
                                             int32_t vocab);
ECAMP_API int ecamp_text_context_mask(const int64_t* ids, int32_t T, const uint8_t* is_sub, const uint8_t* is_entity,
                                      int32_t vocab, const double* draws, int32_t n_draws, int64_t* masked,
                                      int32_t* mask_pos, int32_t* n_mask_pos);
ECAMP_API int ecamp_text_template_weights(const int64_t* ids, int32_t n_ids, const int32_t* mask_pos, int32_t n_mask_pos,
                                          int32_t max_len, float* weights);
/* both steps of one padded report (T = max_caption_length) in one call: masked[T], weights[T], mask_pos[T] */
ECAMP_API int ecamp_text_mask_and_weights(const int64_t* ids, int32_t T, const uint8_t* is_sub, const uint8_t* is_entity,
                                          int32_t vocab, const double* draws, int32_t n_draws, int64_t* masked,
                                          float* weights, int32_t* mask_pos, int32_t* n_mask_pos);

/* ---- fused SGD-momentum + global gradient-norm clip (fine-tune trainer) ----
 * Replaces `torch.nn.utils.clip_grad_norm_(model.parameters(), args.max_grad_norm)` followed by
 * `torch.optim.SGD(model.parameters(), lr, momentum=0.9, weight_decay=wd).step()`
 * (ECAMP/Fine-tuning/Classification/train.py:377-380,459-463).  The caller owns two device buffers of
 * ecamp_sgd_table_bytes(n) / ecamp_sgd_chunk_bytes(numel, n) bytes; ecamp_sgd_build_tables fills them
 * (synchronous, HOST array in).  Per step: ecamp_grad_sumsq writes sum g^2 over all tensors to *sumsq (device),
 * ecamp_sgd_momentum_step applies  coef = min(1, max_norm / (sqrt(*sumsq) + 1e-6))  (max_norm <= 0: no clip),
 * d = coef * g + wd * p, buf = first_step ? d : momentum * buf + d, p -= lr * buf;  the clipped gradients are
 * written back only if write_clipped_grads != 0.  No host synchronisation. */
typedef struct ecamp_sgd_tensor {
  float* p;      /* parameter            */
  float* g;      /* gradient             */
  float* buf;    /* momentum buffer      */
  int64_t numel;
} ecamp_sgd_tensor;
ECAMP_API int64_t ecamp_sgd_table_bytes(int32_t n);
ECAMP_API int64_t ecamp_sgd_chunk_bytes(const int64_t* numel /* host */, int32_t n);
ECAMP_API int ecamp_sgd_build_tables(const ecamp_sgd_tensor* host, int32_t n, void* dev_table, void* dev_chunks,
                                     int64_t* n_chunks /* host out */);
ECAMP_API int ecamp_grad_sumsq(const void* dev_table, const void* dev_chunks, int64_t n_chunks, float* sumsq, void* stream);
ECAMP_API int ecamp_sgd_momentum_step(const void* dev_table, const void* dev_chunks, int64_t n_chunks, float lr,
                                      float momentum, float weight_decay, int32_t first_step, float max_grad_norm,
                                      const float* sumsq, int32_t write_clipped_grads, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ECAMP_B200_H_ */
