// C ABI (extern "C") of libecamp_b200.so — see include/ecamp_b200.h for the contract.
#include "../../include/ecamp_b200.h"

#include "common.cuh"
#include "gemm.cuh"
#include "kernels.cuh"
#include "model.cuh"

namespace ecamp {
const char* last_error_cstr();
long long launch_count();
void set_cta_pair_mode(int mode);
void set_tma_epilogue(int on);
void set_direct_epilogue(int on);
void set_attention_tc(int on);
}
using namespace ecamp;

struct ecamp_ctx {
  Ctx* impl;
};

namespace {
cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }
AttnArgs to_args(const ecamp_attn* a) {
  AttnArgs r;
  r.q = static_cast<const bf16*>(a->q); r.k = static_cast<const bf16*>(a->k); r.v = static_cast<const bf16*>(a->v);
  r.ldq = a->ldq; r.ldk = a->ldk; r.ldv = a->ldv;
  r.o = static_cast<bf16*>(a->o); r.ldo = a->ldo; r.lse = a->lse; r.key_mask = a->key_mask;
  r.B = a->B; r.H = a->H; r.Sq = a->Sq; r.Sk = a->Sk; r.D = a->D; r.scale = a->scale;
  r.drop.p = a->drop_p; r.drop.seed = a->seed; r.drop.site = a->site;
  r.d_o = static_cast<const bf16*>(a->d_o); r.ld_do = a->ld_do; r.delta = a->delta;
  r.dq = static_cast<bf16*>(a->dq); r.dk = static_cast<bf16*>(a->dk); r.dv = static_cast<bf16*>(a->dv);
  r.lddq = a->lddq; r.lddk = a->lddk; r.lddv = a->lddv;
  r.cs_q = a->cs_q; r.cs_k = a->cs_k; r.cs_v = a->cs_v;
  return r;
}
Shape to_shape(const ecamp_shape* s) {
  Shape r;
  r.B = s->B; r.T = s->T; r.keep = s->len_keep; r.has_big = s->has_big; r.ce_rows = s->ce_rows > 0 ? s->ce_rows : 2048;
  return r;
}
}  // namespace

extern "C" {

int ecamp_abi_version(void) { return ECAMP_ABI_VERSION; }
const char* ecamp_last_error(void) { return ecamp::last_error_cstr(); }
int64_t ecamp_launch_count(void) { return ecamp::launch_count(); }
void ecamp_gemm_set_cta_pair(int32_t mode) { ecamp::set_cta_pair_mode(mode); }
void ecamp_gemm_set_tma_epilogue(int32_t on) { ecamp::set_tma_epilogue(on); }
void ecamp_gemm_set_direct_epilogue(int32_t on) { ecamp::set_direct_epilogue(on); }
void ecamp_attention_set_tcgen05(int32_t on) { ecamp::set_attention_tc(on); }

int ecamp_gemm_bf16(const void* A, int32_t lda, int32_t a_mn, const void* B, int32_t ldb, int32_t b_mn, int32_t M,
                    int32_t N, int32_t K, const ecamp_epilogue* ep, int32_t tile_n, void* stream) {
  ECAMP_REQUIRE(A && B && ep, "ecamp_gemm_bf16: null argument");
  GemmEpilogue e;
  e.bias = ep->bias;
  e.aux_in = static_cast<const bf16*>(ep->aux_in);
  e.aux_out = static_cast<bf16*>(ep->aux_out);
  e.ld_aux = ep->ld_aux;
  e.residual = ep->residual;
  e.ld_res = ep->ld_res;
  e.out_f32 = ep->out_f32;
  e.ld_f32 = ep->ld_f32;
  e.out_bf16 = static_cast<bf16*>(ep->out_bf16);
  e.ld_bf16 = ep->ld_bf16;
  e.flags = ep->flags;
  e.drop_p = ep->drop_p;
  e.seed = ep->seed;
  e.stream = ep->site;
  e.colsum_out = ep->colsum_out;
  return gemm_bf16(static_cast<const bf16*>(A), lda, a_mn, static_cast<const bf16*>(B), ldb, b_mn, M, N, K, e, tile_n,
                   S(stream));
}

int64_t ecamp_gemm_fp32_ws_bytes(int32_t M, int32_t N, int32_t K) { return (int64_t)gemm_hp_ws_bytes(M, N, K); }
int ecamp_gemm_fp32(const float* A, int32_t lda, int32_t a_mn, const float* B, int32_t ldb, int32_t b_mn, int32_t M,
                    int32_t N, int32_t K, const ecamp_epilogue* ep, void* ws, int64_t ws_bytes, void* stream) {
  ECAMP_REQUIRE(A && B && ep && ws, "ecamp_gemm_fp32: null argument");
  GemmEpilogueT<float> e;
  e.bias = ep->bias;
  e.aux_in = static_cast<const float*>(ep->aux_in);
  e.aux_out = static_cast<float*>(ep->aux_out);
  e.ld_aux = ep->ld_aux;
  e.residual = ep->residual;
  e.ld_res = ep->ld_res;
  e.out_f32 = ep->out_f32;
  e.ld_f32 = ep->ld_f32;
  e.out_bf16 = static_cast<float*>(ep->out_bf16);
  e.ld_bf16 = ep->ld_bf16;
  e.flags = ep->flags;
  e.drop_p = ep->drop_p;
  e.seed = ep->seed;
  e.stream = ep->site;
  e.colsum_out = ep->colsum_out;
  return gemm_hp(A, lda, a_mn, B, ldb, b_mn, M, N, K, e, ws, (size_t)ws_bytes, S(stream));
}

int ecamp_random_masking(const float* noise, int32_t B, int32_t L, int32_t len_keep, int64_t* ids_restore,
                         int64_t* ids_keep, float* mask, void* scratch_i32, void* stream) {
  ECAMP_REQUIRE(noise && ids_restore && ids_keep && mask && scratch_i32, "ecamp_random_masking: null argument");
  int32_t* r32 = static_cast<int32_t*>(scratch_i32);
  return random_masking(noise, B, L, len_keep, r32, r32 + (size_t)B * L, mask, ids_restore, ids_keep, S(stream));
}

int ecamp_image_u8_normalize(const uint8_t* gray, int64_t n_images, int64_t pixels_per_image, float mean, float std_,
                             float* out, void* stream) {
  return image_u8_normalize(gray, n_images, pixels_per_image, mean, std_, out, S(stream));
}
int32_t ecamp_image_resample_kmax(int32_t in_size, int32_t out) { return image_resample_kmax(in_size, out); }
int64_t ecamp_image_resized_crop_ws_bytes(int32_t B, int32_t out, int32_t kmax, int64_t tmp_bytes) {
  return (int64_t)image_resized_crop_ws_bytes(B, out, kmax, tmp_bytes);
}
int ecamp_image_resized_crop(const uint8_t* crops, const ecamp_crop_desc* desc_dev, int32_t B, int32_t hmax, int32_t out,
                             int32_t kmax, void* ws, int64_t ws_bytes, int64_t tmp_bytes, uint8_t* dst, void* stream) {
  static_assert(sizeof(ecamp_crop_desc) == 32, "ecamp_crop_desc layout");
  return image_resized_crop(crops, desc_dev, B, hmax, out, kmax, ws, (size_t)ws_bytes, tmp_bytes, dst, S(stream));
}
int ecamp_image_resized_crop_host(const uint8_t* crop, int32_t h, int32_t w, int32_t flip, int32_t out, uint8_t* dst) {
  return image_resized_crop_host(crop, h, w, flip, out, dst);
}
int ecamp_resize_patchify(const float* big, int32_t B, int32_t side_in, float* tgt, void* stream) {
  ECAMP_REQUIRE(big && tgt, "ecamp_resize_patchify: null argument");
  if (side_in == 224) return patchify224(big, B, tgt, S(stream));
  return resize_bicubic_patchify(big, B, side_in, tgt, S(stream));
}

int ecamp_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int32_t M, int32_t D,
                        void* out_bf16, float* out_f32, float* mean, float* rstd, void* stream) {
  return layernorm_fwd(x, gamma, beta, eps, M, D, static_cast<bf16*>(out_bf16), out_f32, mean, rstd, S(stream));
}
int ecamp_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                        int32_t M, int32_t D, const float* addend, float* dx_f32, void* dx_bf16, float* dgamma,
                        float* dbeta, float* colsum_out, int32_t accumulate, float* /*ws*/, void* stream) {
  return layernorm_bwd(dy, x, mean, rstd, gamma, M, D, addend, dx_f32, static_cast<bf16*>(dx_bf16), DropoutCfg(),
                       dgamma, dbeta, colsum_out, accumulate, S(stream));
}
size_t ecamp_layernorm_ws_floats(void) { return layernorm_bwd_ws_floats(768); }
void ecamp_layernorm_set_bwd_slab(int32_t on) { layernorm_set_bwd_slab(on); }

int ecamp_attention_fwd(const ecamp_attn* a, void* stream) {
  ECAMP_REQUIRE(a && a->q && a->k && a->v && a->o, "ecamp_attention_fwd: null argument");
  return attention_fwd(to_args(a), S(stream));
}
int ecamp_attention_bwd(const ecamp_attn* a, void* stream) {
  ECAMP_REQUIRE(a && a->q && a->k && a->v && a->o, "ecamp_attention_bwd: null argument");
  return attention_bwd(to_args(a), S(stream));
}

int ecamp_attention_probs(const ecamp_attn* a, float* probs, void* stream) {
  ECAMP_REQUIRE(a && a->q && a->k && a->lse && probs, "ecamp_attention_probs: null argument");
  return attention_probs(to_args(a), probs, S(stream));
}

int ecamp_mim_loss(const float* pred, const float* tgt, const float* mask, int32_t B, float* loss, float* ws,
                   void* stream) {
  return mim_loss_fwd(pred, 197, tgt, mask, B, 196, 768, loss, ws, S(stream));
}
int ecamp_sr_loss_fwd(const float* pred, const float* big, const int64_t* column, const int64_t* row, const float* w1,
                      const float* b1, const float* w2, const float* b2, int32_t B, float* loss, float* ws,
                      void* stream) {
  return sr_loss_fwd(pred, big, column, row, w1, b1, w2, b2, B, loss, ws, S(stream));
}
int ecamp_sr_loss_bwd(const float* pred, const float* big, const int64_t* column, const int64_t* row, const float* w1,
                      const float* b1, const float* w2, const float* b2, int32_t B, const float* g_res, float* d_u,
                      float* d_conv, int32_t accumulate, float* ws, void* stream) {
  return sr_loss_bwd(pred, big, column, row, w1, b1, w2, b2, B, g_res, d_u, d_conv, accumulate, ws, S(stream));
}
size_t ecamp_sr_ws_floats(int32_t B) { return sr_ws_floats(B); }
int ecamp_pred_grad(const float* pred, const float* tgt, const float* mask, const float* d_u, const float* g_mim,
                    int32_t B, void* d_pred_bf16, void* stream) {
  return pred_grad(pred, tgt, mask, d_u, g_mim, B, static_cast<bf16*>(d_pred_bf16), S(stream));
}
int ecamp_ce_rows(void* logits_bf16, int32_t ld, int32_t rows, int32_t V, const int64_t* labels, const float* weights,
                  float* row_loss, const float* g, float inv_total_rows, int32_t write_grad, void* stream) {
  return ce_chunk(static_cast<bf16*>(logits_bf16), ld, rows, V, labels, weights, row_loss, g, inv_total_rows,
                  write_grad, S(stream));
}

int ecamp_ce_rows_bias(void* logits_bf16, int32_t ld, int32_t rows, int32_t V, const int64_t* labels, const float* weights,
                       float* row_loss, const float* g, float inv_total_rows, float* bias_grad, void* stream) {
  return ce_chunk(static_cast<bf16*>(logits_bf16), ld, rows, V, labels, weights, row_loss, g, inv_total_rows, 1, S(stream),
                  bias_grad);
}
void ecamp_ce_set_fused(int32_t on) { ce_set_fused(on); }
void ecamp_sr_set_window_skip(int32_t on) { sr_set_window_skip(on); }
void ecamp_set_side_stream(int32_t on) { set_side_stream(on); }

// ---- runtime -----------------------------------------------------------------------------------
int32_t ecamp_param_count(void) { return (int32_t)param_specs().size(); }
const char* ecamp_param_name(int32_t i) {
  return (i >= 0 && i < ecamp_param_count()) ? param_specs()[i].name.c_str() : nullptr;
}
int64_t ecamp_param_numel(int32_t i) { return (i >= 0 && i < ecamp_param_count()) ? param_specs()[i].numel : -1; }
int32_t ecamp_param_decay(int32_t i) { return (i >= 0 && i < ecamp_param_count()) ? param_specs()[i].decay : -1; }
int64_t ecamp_param_grad_offset(int32_t i) { return (i >= 0 && i < ecamp_param_count()) ? param_specs()[i].g_off : -1; }
int64_t ecamp_grad_floats(void) { return grad_total_floats(); }
int64_t ecamp_shadow_bytes(void) { return ((shadow_bf16_elems() + 7) & ~7LL) * 4 + shadow_f32_elems() * 4 + 64; }
int64_t ecamp_adam_table_bytes(void) { return (int64_t)ctx_adam_table_bytes(); }
int64_t ecamp_adam_chunk_bytes(void) { return (int64_t)ctx_adam_chunk_bytes(); }

int ecamp_ctx_create(ecamp_ctx** out) {
  ECAMP_REQUIRE(out != nullptr, "ecamp_ctx_create: null output");
  *out = new ecamp_ctx{ctx_new()};
  return 0;
}
void ecamp_ctx_destroy(ecamp_ctx* ctx) {
  if (!ctx) return;
  ctx_free(ctx->impl);
  delete ctx;
}
int ecamp_ctx_bind(ecamp_ctx* ctx, float* const* params_host, int32_t n, float* grads, float* adam_m, float* adam_v,
                   void* shadows, const float* pos_embed, const float* decoder_pos_embed, void* adam_table,
                   void* adam_chunks) {
  ECAMP_REQUIRE(ctx && params_host, "ecamp_ctx_bind: null argument");
  return ctx_bind(ctx->impl, params_host, n, grads, adam_m, adam_v, shadows, pos_embed, decoder_pos_embed, adam_table,
                  adam_chunks);
}
int64_t ecamp_workspace_bytes(const ecamp_shape* s) { return s ? (int64_t)workspace_bytes(to_shape(s), 0) : -1; }
int ecamp_ctx_set_precision(ecamp_ctx* ctx, int32_t fp32_accurate) {
  ECAMP_REQUIRE(ctx, "ecamp_ctx_set_precision: null context");
  return ctx_set_precision(ctx->impl, fp32_accurate);
}
int64_t ecamp_ctx_workspace_bytes(ecamp_ctx* ctx, const ecamp_shape* s) {
  return (ctx && s) ? (int64_t)workspace_bytes(to_shape(s), ctx_precision(ctx->impl)) : -1;
}
int ecamp_ctx_set_workspace(ecamp_ctx* ctx, void* ws, int64_t bytes, const ecamp_shape* s) {
  ECAMP_REQUIRE(ctx && ws && s, "ecamp_ctx_set_workspace: null argument");
  return ctx_set_workspace(ctx->impl, ws, (size_t)bytes, to_shape(s));
}
int ecamp_refresh_shadows(ecamp_ctx* ctx, void* stream) {
  ECAMP_REQUIRE(ctx, "ecamp_refresh_shadows: null context");
  return ctx_refresh_shadows(ctx->impl, S(stream));
}
int ecamp_forward(ecamp_ctx* ctx, const ecamp_batch* b, int32_t flags, float drop_p, uint64_t seed, float* losses,
                  float* mask, int64_t* ids_restore, int64_t* ids_keep, void* stream) {
  ECAMP_REQUIRE(ctx && b, "ecamp_forward: null argument");
  Batch bb;
  bb.image = b->image; bb.ids = b->ids; bb.labels = b->labels; bb.attention_mask = b->attention_mask;
  bb.type_ids = b->type_ids; bb.weights = b->weights; bb.column = b->column; bb.row = b->row; bb.noise = b->noise;
  return ctx_forward(ctx->impl, bb, flags, drop_p, seed, losses, mask, ids_restore, ids_keep, S(stream));
}
int32_t ecamp_backward_stage_count(void) { return backward_stage_count(); }
int ecamp_backward_stage_range(int32_t stage, int64_t* begin, int64_t* end) {
  long long b = 0, e = 0;
  const int rc = backward_stage_range(stage, &b, &e);
  if (rc) return rc;
  *begin = b; *end = e;
  return 0;
}
int ecamp_backward(ecamp_ctx* ctx, const float* g3, int32_t accumulate, int32_t stage, void* stream) {
  ECAMP_REQUIRE(ctx, "ecamp_backward: null context");
  return ctx_backward(ctx->impl, g3, accumulate, stage, 0, S(stream));
}
int ecamp_backward_stages(ecamp_ctx* ctx, const float* g3, int32_t accumulate, int32_t first_stage, int32_t end_stage,
                          void* stream) {
  ECAMP_REQUIRE(ctx, "ecamp_backward_stages: null context");
  ECAMP_REQUIRE(first_stage >= 0 && end_stage > first_stage, "ecamp_backward_stages: empty or negative stage range [%d, %d)",
                first_stage, end_stage);
  return ctx_backward(ctx->impl, g3, accumulate, first_stage, end_stage, S(stream));
}
int ecamp_adamw_step(ecamp_ctx* ctx, float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                     float grad_scale, void* stream) {
  ECAMP_REQUIRE(ctx, "ecamp_adamw_step: null context");
  return ctx_adamw(ctx->impl, lr, lr, beta1, beta2, eps, weight_decay, step, grad_scale, S(stream));
}
int ecamp_adamw_step_groups(ecamp_ctx* ctx, float lr_decay, float lr_no_decay, float beta1, float beta2, float eps,
                            float weight_decay, int32_t step, float grad_scale, void* stream) {
  ECAMP_REQUIRE(ctx, "ecamp_adamw_step_groups: null context");
  return ctx_adamw(ctx->impl, lr_decay, lr_no_decay, beta1, beta2, eps, weight_decay, step, grad_scale, S(stream));
}
int ecamp_adamw_step_range(ecamp_ctx* ctx, float lr_decay, float lr_no_decay, float beta1, float beta2, float eps,
                           float weight_decay, int32_t step, float grad_scale, int64_t grad_begin, int64_t grad_end, void* stream) {
  ECAMP_REQUIRE(ctx, "ecamp_adamw_step_range: null context");
  return ctx_adamw_range(ctx->impl, lr_decay, lr_no_decay, beta1, beta2, eps, weight_decay, step, grad_scale, grad_begin, grad_end,
                         S(stream));
}
int ecamp_cross_attention_probs(ecamp_ctx* ctx, float* probs, void* stream) {
  ECAMP_REQUIRE(ctx && probs, "ecamp_cross_attention_probs: null argument");
  return ctx_cross_attention_probs(ctx->impl, probs, S(stream));
}
const void* ecamp_debug_buffer(ecamp_ctx* ctx, const char* name) {
  return (ctx && name) ? ctx_debug_ptr(ctx->impl, name) : nullptr;
}

// ---- fine-tune classification ---------------------------------------------------------------------------------
static ClsIO to_cls(const ecamp_cls_io* io) {
  ClsIO r;
  r.image = io->image; r.pos_embed = io->pos_embed; r.fc_norm_w = io->fc_norm_w; r.fc_norm_b = io->fc_norm_b;
  r.head_w16 = static_cast<const bf16*>(io->head_w16); r.head_b = io->head_b; r.dp_scale = io->dp_scale; r.logits = io->logits;
  r.d_logits = io->d_logits; r.g_pos_embed = io->g_pos_embed; r.g_fc_norm_w = io->g_fc_norm_w; r.g_fc_norm_b = io->g_fc_norm_b;
  r.g_head_w = io->g_head_w; r.g_head_b = io->g_head_b;
  return r;
}
int64_t ecamp_cls_workspace_bytes(int32_t B) { return B > 0 ? (int64_t)cls_workspace_bytes(B) : 0; }
int ecamp_cls_set_workspace(ecamp_ctx* ctx, void* ws, int64_t bytes, int32_t B) {
  ECAMP_REQUIRE(ctx && ws, "ecamp_cls_set_workspace: null argument");
  return ctx_set_cls_workspace(ctx->impl, ws, (size_t)bytes, B);
}
int ecamp_cls_forward(ecamp_ctx* ctx, const ecamp_cls_io* io, void* stream) {
  ECAMP_REQUIRE(ctx && io, "ecamp_cls_forward: null argument");
  return ctx_cls_forward(ctx->impl, to_cls(io), S(stream));
}
int ecamp_cls_backward(ecamp_ctx* ctx, const ecamp_cls_io* io, int32_t accumulate, void* stream) {
  ECAMP_REQUIRE(ctx && io, "ecamp_cls_backward: null argument");
  return ctx_cls_backward(ctx->impl, to_cls(io), accumulate, S(stream));
}

// ---- host-side report masking (HOST pointers) ---------------------------------------------------------------------------
int32_t ecamp_text_mask_draw_count(const int64_t* ids, int32_t T, const uint8_t* is_sub, const uint8_t* is_entity, int32_t vocab) {
  if (!ids || !is_sub || !is_entity || T < 2) return -1;
  return text_mask_draw_count(reinterpret_cast<const long long*>(ids), T, is_sub, is_entity, vocab);
}
int ecamp_text_context_mask(const int64_t* ids, int32_t T, const uint8_t* is_sub, const uint8_t* is_entity, int32_t vocab,
                            const double* draws, int32_t n_draws, int64_t* masked, int32_t* mask_pos, int32_t* n_mask_pos) {
  return text_context_mask(reinterpret_cast<const long long*>(ids), T, is_sub, is_entity, vocab, draws, n_draws,
                           reinterpret_cast<long long*>(masked), mask_pos, n_mask_pos);
}
int ecamp_text_template_weights(const int64_t* ids, int32_t n_ids, const int32_t* mask_pos, int32_t n_mask_pos, int32_t max_len,
                                float* weights) {
  return text_template_weights(reinterpret_cast<const long long*>(ids), n_ids, mask_pos, n_mask_pos, max_len, weights);
}

int ecamp_text_mask_and_weights(const int64_t* ids, int32_t T, const uint8_t* is_sub, const uint8_t* is_entity, int32_t vocab,
                                const double* draws, int32_t n_draws, int64_t* masked, float* weights, int32_t* mask_pos,
                                int32_t* n_mask_pos) {
  ECAMP_REQUIRE(weights, "ecamp_text_mask_and_weights: null weights");
  int rc = ecamp_text_context_mask(ids, T, is_sub, is_entity, vocab, draws, n_draws, masked, mask_pos, n_mask_pos);
  if (rc) return rc;
  return ecamp_text_template_weights(ids, T, mask_pos, *n_mask_pos, T, weights);
}

// ---- fused SGD-momentum + grad-norm clip ----------------------------------------------------------------------
static_assert(sizeof(ecamp_sgd_tensor) == sizeof(SgdTensor), "ecamp_sgd_tensor layout");
int64_t ecamp_sgd_table_bytes(int32_t n) { return (int64_t)sgd_table_bytes(n); }
int64_t ecamp_sgd_chunk_bytes(const int64_t* numel, int32_t n) {
  if (!numel || n <= 0) return 0;
  return (int64_t)sgd_chunk_bytes(reinterpret_cast<const long long*>(numel), n);
}
int ecamp_sgd_build_tables(const ecamp_sgd_tensor* host, int32_t n, void* dev_table, void* dev_chunks, int64_t* n_chunks) {
  long long c = 0;
  const int rc = sgd_build_tables(reinterpret_cast<const SgdTensor*>(host), n, dev_table, dev_chunks, &c);
  if (rc) return rc;
  *n_chunks = c;
  return 0;
}
int ecamp_grad_sumsq(const void* dev_table, const void* dev_chunks, int64_t n_chunks, float* sumsq, void* stream) {
  return grad_sumsq(dev_table, dev_chunks, n_chunks, sumsq, S(stream));
}
int ecamp_sgd_momentum_step(const void* dev_table, const void* dev_chunks, int64_t n_chunks, float lr, float momentum,
                            float weight_decay, int32_t first_step, float max_grad_norm, const float* sumsq,
                            int32_t write_clipped_grads, void* stream) {
  return sgd_momentum_step(dev_table, dev_chunks, n_chunks, lr, momentum, weight_decay, first_step, max_grad_norm, sumsq,
                           write_clipped_grads, S(stream));
}

}  // extern "C"
