// Runtime of the ECAMP pre-training step: parameter table, workspace plan, forward / backward
// schedules.  One context per process / device; PyTorch owns every byte (parameters, the flat
// gradient / Adam-state / shadow buffers and the workspace are torch tensors whose base pointers
// are bound here).
#pragma once
#include <string>
#include <vector>

#include "gemm.cuh"
#include "kernels.cuh"

namespace ecamp {

struct ParamSpec {
  std::string name;   // reference state_dict key (model_ecamp.py / HF naming)
  long long numel;
  int decay;          // timm add_weight_decay: ndim > 1 and not *.bias
  int shadow;         // 0 none, 1 bf16 copy, 2 bf16 patch-embed permuted copy, 3 fp32 copy (fused bias)
  long long g_off;    // offset (floats) in the flat gradient / Adam-state buffers
  long long sh_off;   // offset (bf16 elements) in the bf16 shadow region, or fp32 elements in the fp32 region
};

const std::vector<ParamSpec>& param_specs();
long long grad_total_floats();
long long shadow_bf16_elems();
long long shadow_f32_elems();
int param_index(const std::string& name);

struct Shape {
  int B = 0, T = 0, keep = 49, has_big = 1;
  int ce_rows = 2048;  // rows of the vocabulary projection materialised at a time
};

struct Batch {
  const float* image = nullptr;  // [B,3,448,448] if has_big else [B,3,224,224]
  const int64_t* ids = nullptr;
  const int64_t* labels = nullptr;
  const int64_t* attention_mask = nullptr;
  const int64_t* type_ids = nullptr;
  const float* weights = nullptr;
  const int64_t* column = nullptr;
  const int64_t* row = nullptr;
  const float* noise = nullptr;  // [B,196]
};

struct Ctx;
void set_side_stream(int on);  // 0: the whole backward on the caller's stream (default 1: block / vocabulary weight gradients on a side stream)
Ctx* ctx_new();
void ctx_free(Ctx*);
int ctx_bind(Ctx*, float* const* params, int n, float* G, float* M1, float* M2, void* shadows,
             const float* pos_embed, const float* dec_pos_embed, void* adam_table, void* adam_chunks);
size_t ctx_adam_table_bytes();
size_t ctx_adam_chunk_bytes();
size_t workspace_bytes(const Shape&, int hp);
// 0 = production (bf16 GEMM / attention operands), 1 = fp32-accurate parity mode (fp32 activations, bf16 x 3 split-operand
// GEMMs on the same tcgen05 kernel).  Changing it resets the context: bind() and set_workspace() must follow.
int ctx_set_precision(Ctx*, int hp);
int ctx_precision(Ctx*);
int ctx_set_workspace(Ctx*, void* ws, size_t bytes, const Shape&);
int ctx_refresh_shadows(Ctx*, cudaStream_t);
// flags: 1 = training (saves activations), 2 = defer the MLM loss to backward (fused head)
int ctx_forward(Ctx*, const Batch&, int flags, float drop_p, unsigned long long seed, float* losses3, float* mask_out,
                int64_t* ids_restore_out, int64_t* ids_keep_out, cudaStream_t);
// after ctx_forward: probs[B, 6, T, keep] fp32 of the fusion layer's text -> image cross-attention (columns in ids_keep order)
int ctx_cross_attention_probs(Ctx*, float* probs, cudaStream_t);
int backward_stage_count();
int backward_stage_range(int stage, long long* g_begin, long long* g_end);
// g3: device pointer to the three upstream gradients (mim, res, mlm).  stage = -1 runs every stage.
int ctx_backward(Ctx*, const float* g3, int accumulate, int stage, int stage_end, cudaStream_t);  // stage_end <= stage: one stage
int ctx_adamw(Ctx*, float lr, float lr_nodecay, float b1, float b2, float eps, float wd, int step, float grad_scale, cudaStream_t);
int ctx_adamw_range(Ctx*, float lr, float lr_nodecay, float b1, float b2, float eps, float wd, int step, float grad_scale,
                    long long g_lo, long long g_hi, cudaStream_t);
const void* ctx_debug_ptr(Ctx*, const char* name);

// ---- fine-tune classification (FT/Classification/models_vit.py, global_pool = True) on the same context ----------
struct ClsIO {
  const float* image = nullptr;      // [B, 3, 224, 224]
  const float* pos_embed = nullptr;  // [197, 768] (learnable in the fine-tune model)
  const float *fc_norm_w = nullptr, *fc_norm_b = nullptr;
  const bf16* head_w16 = nullptr;    // [16, 768]: the head weight padded to 16 rows (rows >= num_classes are zero)
  const float* head_b = nullptr;     // [16]
  const float* dp_scale = nullptr;   // DropPath: [12][2][B] mask / keep_prob per (block, branch, sample), or null
  float* logits = nullptr;           // out [B, 16] fp32
  // backward
  const float* d_logits = nullptr;   // [B, 16] fp32 (columns >= num_classes zero)
  float *g_pos_embed = nullptr, *g_fc_norm_w = nullptr, *g_fc_norm_b = nullptr, *g_head_w = nullptr /* [16, 768] */,
        *g_head_b = nullptr /* [16] */;
};
size_t cls_workspace_bytes(int B);
int ctx_set_cls_workspace(Ctx*, void* ws, size_t bytes, int B);
int ctx_cls_forward(Ctx*, const ClsIO&, cudaStream_t);
// encoder-parameter gradients go to the flat gradient buffer (same offsets as in pre-training), the rest to ClsIO
int ctx_cls_backward(Ctx*, const ClsIO&, int accumulate, cudaStream_t);

}  // namespace ecamp
