// Internal C++ launch interface of the non-GEMM kernels of the ECAMP hot path (all sm_100a, all
// asynchronous on the given stream, all returning 0 or a negative error code).
#pragma once
#include "common.cuh"

namespace ecamp {

struct DropoutCfg {
  float p = 0.f;
  unsigned long long seed = 0, site = 0;
};

// ---- layernorm.cu -----------------------------------------------------------------------------
// y = LN(x) * gamma + beta, fp32 statistics; D in {512, 768}.  Any of out_bf16 / out_f32 may be null.
template <typename AT>
int layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int M, int D, AT* out_bf16,
                  float* out_f32, float* mean, float* rstd, cudaStream_t st);
// dx = addend + LNbwd(dy); dx_bf16 = dropout_bwd(dx) (optional, mask of the forward's dense-output dropout);
// dgamma/dbeta partials are reduced and ADDED to dgamma/dbeta when accumulate != 0, else stored.
template <typename AT>
int layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, int M,
                  int D, const float* addend, float* dx_f32, AT* dx_bf16, DropoutCfg drop, float* dgamma,
                  float* dbeta, float* colsum_out, int accumulate, cudaStream_t st, const float* out_row_scale = nullptr,
                  int rows_per_scale = 1);
size_t layernorm_bwd_ws_floats(int D);
void layernorm_set_bwd_slab(int on);  // 1: the round-1 slab kernel instead of the staged one (A/B measurements)

// ---- elementwise.cu ---------------------------------------------------------------------------
// model_ecamp.py:168-193 on given noise: stable ascending rank == ids_restore; ids_keep; mask (1 = removed).
int random_masking(const float* noise, int B, int L, int len_keep, int32_t* ids_restore, int32_t* ids_keep,
                   float* mask, int64_t* ids_restore64, int64_t* ids_keep64, cudaStream_t st);
// bicubic 448 -> 224 (antialias off, A = -0.75, align_corners = False) written in patch layout
// tgt[b, l, (p*16+q)*3 + c]  (model_ecamp.py:318 + the layout of patchify with p = 16).
int resize_bicubic_patchify(const float* big, int B, int Hin, float* tgt, cudaStream_t st);
int patchify224(const float* imgs, int B, float* tgt, cudaStream_t st);
// A_pe[b*keep + j, :] = bf16(tgt[b, ids_keep[b, j], :])
template <typename AT>
int gather_patches(const float* tgt, const int32_t* ids_keep, int B, int L, int keep, int PD, AT* out,
                   cudaStream_t st);
// x0[b, 0] = cls + pos[0]; x0[b, 1 + j] = pe[b*keep + j] + pos[1 + ids_keep[b, j]]      (model_ecamp.py:222-230)
int assemble_encoder_input(const float* pe, const float* cls, const float* pos, const int32_t* ids_keep, int B,
                           int keep, int D, float* x0, cudaStream_t st);
// backward of the above: d_pe (bf16, GEMM operand) and d_cls (sum over batch of row 0)
template <typename AT>
int assemble_encoder_input_bwd(const float* dx0, int B, int keep, int D, AT* d_pe, float* d_cls, int accumulate,
                               cudaStream_t st);
// xd[b, 0] = e[b, 0] + dpos[0]; xd[b, 1 + l] = (r = ids_restore[b, l]) < keep ? e[b, 1 + r] : mask_token, + dpos[1 + l]
template <typename AT>
int assemble_decoder_input(const AT* e, const float* mask_token, const float* dpos, const int32_t* ids_restore,
                           int B, int L, int keep, int D, float* xd, cudaStream_t st);
template <typename AT>
int assemble_decoder_input_bwd(const float* dxd, const int32_t* ids_restore, int B, int L, int keep, int D, AT* d_e,
                               float* d_mask_token, int accumulate, float* ws, cudaStream_t st);
// img_tok[b*keep + j] = lat2[b, 1 + j]; gap[b] = mean_j lat2[b, 1 + j]                  (model_ecamp.py:269-271)
template <typename AT>
int split_latent_gap(const AT* lat2, int B, int keep, int D, AT* img_tok, AT* gap, cudaStream_t st);
// d_lat2[b, 0] = 0; d_lat2[b, 1 + j] = d_img_tok[b*keep + j] + d_gap[b] / keep
template <typename AT>
int split_latent_gap_bwd(const AT* d_img_tok, const AT* d_gap, int B, int keep, int D, AT* d_lat2,
                         cudaStream_t st);
// y[b*T + t, :] += vec[b, :]   (context_fusion.py:54-55)
template <typename AT>
int add_batch_rowvec(AT* y, const AT* vec, int B, int T, int D, cudaStream_t st);
// out[b, :] = sum_t x[b*T + t, :]  (bf16 in, bf16 out, fp32 accumulate)
template <typename AT>
int batch_colsum(const AT* x, int B, int T, int D, AT* out, cudaStream_t st);
// BertEmbeddings: e = word[id] + type[tt] + pos[t] -> pre (fp32, saved); LN(1e-12) -> dropout -> bf16 + fp32
template <typename AT>
int bert_embeddings_fwd(const int64_t* ids, const int64_t* type_ids, const float* word, const float* type,
                        const float* pos, const float* gamma, const float* beta, float eps, int B, int T, int D,
                        DropoutCfg drop, float* pre, float* mean, float* rstd, AT* out_bf16, float* out_f32,
                        cudaStream_t st);
// in-place inverted-dropout backward on an fp32 gradient (n % 4 == 0)
int dropout_bwd_f32(float* g, size_t n, DropoutCfg drop, cudaStream_t st);
// scatter-add of d_pre into the three tables (word row padding_idx = 0 receives nothing)
int bert_embeddings_bwd(const float* d_pre, const int64_t* ids, const int64_t* type_ids, int B, int T, int D,
                        float* d_word, float* d_type, float* d_pos, int accumulate, float* ws, cudaStream_t st);
// bias gradient: out[n] (+)= sum_m x[m, n]
template <typename AT>
int colsum_bf16(const AT* x, int ld, int M, int N, float* out, int accumulate, float* ws, cudaStream_t st);
size_t colsum_ws_floats(int N);
// y = dropout(x) elementwise on bf16 (used for the dropout behind LayerNorm-less sites), in place allowed
int scale_f32(float* x, const float* scale_dev, size_t n, cudaStream_t st);
int cast_f32_to_bf16(const float* x, bf16* y, size_t n, cudaStream_t st);
// fine-tune classification helpers (FT/Classification/models_vit.py:78-98)
int iota_mod_i32(int32_t* out, size_t n, int mod, cudaStream_t st);                      // out[i] = i % mod
int mean_pool_tokens(const float* x, int B, int S, int D, float* out, cudaStream_t st);  // mean over tokens 1..S-1
// dx[b, 0] = 0, dx[b, t >= 1] = d_pooled[b] / (S - 1); gx = bf16(dx * scale[b]) (scale may be null)
int mean_pool_tokens_bwd(const float* d_pooled, int B, int S, int D, float* dx, bf16* gx, const float* scale, cudaStream_t st);
int strided_rowsum(const float* x, int B, size_t stride, int D, float* out, int accumulate, cudaStream_t st);
// patch-embed weight: canonical [768, (c, p, q)] <-> GEMM K-order [768, (p, q, c)]
int permute_pe_weight_grad(const float* dw_pqc, float* grad_cpq, int accumulate, cudaStream_t st);
// out = bf16(d * gelu'(pre))   (LM-head transform: dense -> GELU -> LayerNorm, bert_modeling.py:208)
template <typename AT>
int gelu_bwd_bf16(const float* d, const AT* pre, AT* out, size_t n, cudaStream_t st);
// y = x + vec[b] broadcast over the T rows of each batch element (out of place)
template <typename AT>
int add_batch_rowvec_oop(const AT* x, const AT* vec, int B, int T, int D, AT* y, cudaStream_t st);

// ---- attention.cu -----------------------------------------------------------------------------
template <typename AT>
struct AttnArgsT {
  const AT *q = nullptr, *k = nullptr, *v = nullptr;  // head h lives at columns [h*D, (h+1)*D) of each row
  int ldq = 0, ldk = 0, ldv = 0;
  AT* o = nullptr;
  int ldo = 0;
  float* lse = nullptr;              // [B, H, Sq]
  const int64_t* key_mask = nullptr;  // [B, Sk], nonzero = attend (HF attention_mask); null = all valid
  int B = 0, H = 0, Sq = 0, Sk = 0, D = 0;
  float scale = 1.f;
  DropoutCfg drop;
  // backward only
  const AT* d_o = nullptr;
  int ld_do = 0;
  float* delta = nullptr;  // [B, H, Sq] scratch
  AT *dq = nullptr, *dk = nullptr, *dv = nullptr;
  int lddq = 0, lddk = 0, lddv = 0;
  // optional [H * D] fp32 each: += column sums over all (batch, position) rows of dq / dk / dv (atomic adds) = the bias
  // gradients of the query / key / value projections, folded into the kernel that produces their operand
  float *cs_q = nullptr, *cs_k = nullptr, *cs_v = nullptr;
  // fp32-accurate mode only: scratch of 2 * B * H * Sq * Sk floats (dS and the dropped probabilities between the two
  // backward kernels)
  float* hp_ws = nullptr;
};
typedef AttnArgsT<bf16> AttnArgs;
// fp32-accurate mode (attention_hp.cu): the same attention on fp32 q / k / v with fp32 CUDA-core arithmetic
int attention_fwd(const AttnArgsT<float>& a, cudaStream_t st);
int attention_bwd(const AttnArgsT<float>& a, cudaStream_t st);
int attention_fwd(const AttnArgs& a, cudaStream_t st);
int attention_bwd(const AttnArgs& a, cudaStream_t st);
// probs[B, H, Sq, Sk] fp32 = exp(scale * q . k - lse), masked keys 0 (needs the lse of a preceding attention_fwd)
int attention_probs(const AttnArgs& a, float* probs, cudaStream_t st);

// Grayscale(3) + ToTensor + Normalize of the loader's 8-bit crop on the GPU: out[n, c, :] = (gray[n, :] / 255 - mean) / std
int image_u8_normalize(const uint8_t* gray, long long n_images, long long pixels_per_image, float mean, float stdv, float* out,
                       cudaStream_t st);

// ---- image_pipeline.cu: RandomResizedCrop (Pillow-exact antialiased bicubic) + horizontal flip of 8-bit frames ------------
size_t image_resized_crop_ws_bytes(int B, int out, int kmax, long long tmp_bytes);
int image_resample_kmax(int in_size, int out);
int image_resized_crop(const uint8_t* crops, const void* desc_dev, int B, int hmax, int out, int kmax, void* ws, size_t ws_bytes,
                       long long tmp_bytes, uint8_t* dst, cudaStream_t st);
int image_resized_crop_host(const uint8_t* crop, int h, int w, int flip, int out, uint8_t* dst);  // host restatement (tests)

// ---- losses.cu --------------------------------------------------------------------------------
// mim = sum_{masked patches} (pred - tgt)^2 / (B*3*224*224)       (model_ecamp.py:288-297)
int mim_loss_fwd(const float* pred, int ld_pred_rows, const float* tgt, const float* mask, int B, int L, int PD,
                 float* loss_out, float* ws, cudaStream_t st);
// SR branch forward loss (model_ecamp.py:37-46,196-215,286,291-299) and backward (d_u + conv grads)
int sr_loss_fwd(const float* pred, const float* big, const int64_t* column, const int64_t* row, const float* w1,
                const float* b1, const float* w2, const float* b2, int B, float* loss_out, float* ws,
                cudaStream_t st);
int sr_loss_bwd(const float* pred, const float* big, const int64_t* column, const int64_t* row, const float* w1,
                const float* b1, const float* w2, const float* b2, int B, const float* g_res, float* d_u,
                float* d_conv /*168: w1,b1,w2,b2*/, int accumulate, float* ws, cudaStream_t st,
                float* loss_tiles = nullptr /*B*196 scratch*/, float* loss_out = nullptr /*also emit the loss*/);
size_t sr_ws_floats(int B);
// d_pred[b, 0] = 0; d_pred[b, 1 + l, e] = g_mim * 2 * mask * (pred - tgt) / Nmim + bilinear^T(d_u) (if d_u)
template <typename AT>
int pred_grad(const float* pred, const float* tgt, const float* mask, const float* d_u, const float* g_mim, int B,
              AT* d_pred, cudaStream_t st);
// weighted cross-entropy over a chunk of rows, logits bf16 [rows, V] (ld = ldl); optionally overwrites the
// logits with d_logits = (softmax - onehot) * w * g / total_rows           (bert_modeling.py:211-217)
void sr_set_window_skip(int on);  // measurement switch: 0 = SR backward computes every stage on whole tiles
void ce_set_fused(int on);  // measurement switch: 0 = one CTA per row + separate column-sum pass
// bias_grad (optional, [V], ADDED to): column sums of the gradient = gradient of the vocabulary bias, folded into the same pass
int ce_chunk(bf16* logits, int ldl, int rows, int V, const int64_t* labels, const float* weights, float* row_loss,
             const float* g_mlm, float inv_total, int write_grad, cudaStream_t st, float* bias_grad = nullptr);
int ce_chunk(float* logits, int ldl, int rows, int V, const int64_t* labels, const float* weights, float* row_loss,
             const float* g_mlm, float inv_total, int write_grad, cudaStream_t st);  // fp32-accurate mode
int sum_to_scalar(const float* x, size_t n, float scale, float* out, cudaStream_t st);

// ---- adamw.cu ---------------------------------------------------------------------------------
struct AdamTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  bf16* shadow;        // bf16 GEMM copy, may be null
  float* shadow_f;     // fp32-accurate mode: the GEMM copy in fp32 (same layouts), may be null
  float* shadow32;     // fp32 copy (fused q|k|v bias vectors), may be null
  long long numel;
  int decay;           // 1 = apply weight decay
  int shadow_kind;     // 0 plain, 1 patch-embed (c,p,q)->(p,q,c)
};
int adamw_build_tables(const AdamTensor* host, int n, void* dev_table, void* dev_chunks, long long* n_chunks);
// lr: learning rate of the weight-decay group, lr_nodecay: of the no-decay group (timm add_weight_decay's two groups)
int adamw_step(const void* dev_table, const void* dev_chunks, long long n_chunks, float lr, float lr_nodecay, float beta1,
               float beta2, float eps, float wd, int step, float grad_scale, cudaStream_t st);
int adamw_step_range(const void* dev_table, const void* dev_chunks, long long chunk_begin, long long chunk_end, float lr,
                     float lr_nodecay, float beta1, float beta2, float eps, float wd, int step, float grad_scale,
                     cudaStream_t st);
long long adamw_chunks_of(long long numel);
int refresh_shadows(const void* dev_table, const void* dev_chunks, long long n_chunks, cudaStream_t st);
size_t adamw_table_bytes(int n);
size_t adamw_chunk_bytes(const AdamTensor* host, int n);

// ---- text_host.cu: host-side report masking / re-weighting (pretrain_datasets.py:60-110,141-184), no CUDA ----------
int text_mask_draw_count(const long long* ids, int T, const uint8_t* is_sub, const uint8_t* is_entity, int vocab);
int text_context_mask(const long long* ids, int T, const uint8_t* is_sub, const uint8_t* is_entity, int vocab,
                      const double* draws, int n_draws, long long* masked, int* mask_pos, int* n_mask_pos);
int text_template_weights(const long long* ids, int n_ids, const int* mask_pos, int n_mask_pos, int max_len, float* w);

// ---- sgd.cu: fused SGD-momentum + global grad-norm clip (fine-tune trainer) ---------------------
struct SgdTensor {  // mirrors ecamp_sgd_tensor
  float* p;
  float* g;
  float* buf;
  long long numel;
};
size_t sgd_table_bytes(int n);
size_t sgd_chunk_bytes(const long long* numel, int n);
int sgd_build_tables(const SgdTensor* host, int n, void* dev_table, void* dev_chunks, long long* n_chunks);
int grad_sumsq(const void* dev_table, const void* dev_chunks, long long n_chunks, float* sumsq, cudaStream_t st);
int sgd_momentum_step(const void* dev_table, const void* dev_chunks, long long n_chunks, float lr, float momentum,
                      float wd, int first, float max_norm, const float* sumsq, int write_grads, cudaStream_t st);

}  // namespace ecamp
