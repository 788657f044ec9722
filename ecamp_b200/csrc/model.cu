// Forward / backward schedule of the ECAMP pre-training step on the sm_100a kernels.
// Follows ECAMP.forward (module/model_ecamp.py:303-325) and its callees; see SURVEY.md §3.3 for the
// call stack this file restates as a fixed kernel schedule.  Mixed precision mirrors the reference's
// autocast: GEMM / attention operands in 16-bit (bf16 here), LayerNorm, residual streams, losses and
// all parameter gradients in fp32.
#include "model.cuh"

#include <algorithm>
#include <cstring>
#include <map>

namespace ecamp {

// =============================================================================================
// parameter table (canonical order = forward execution order, so that backward finishes contiguous
// suffixes of the flat gradient buffer and data-parallel buckets are plain ranges)
// =============================================================================================
namespace {

constexpr int E = 768, EH = 12, EHID = 3072, EL = 12;      // encoder
constexpr int DD = 512, DH = 16, DHID = 2048, DL = 4;      // decoder
constexpr int BH = 6, BHID = 1536, BL = 6, VOC = 30000;    // BERT (hidden 768)
constexpr int L196 = 196, PDIM = 768, MAXPOS = 256;
constexpr int VIT_BLOCK_PARAMS = 12, BERT_LAYER_PARAMS = 16, FUSION_PARAMS = 28;
const char* BERT = "bert_encoder.model.bert.";

struct Table {
  std::vector<ParamSpec> specs;
  std::map<std::string, int> index;
  long long g_total = 0, sh16 = 0, sh32 = 0;
  void add(const std::string& name, long long numel, int decay, int shadow) {
    ParamSpec s{name, numel, decay, shadow, g_total, -1};
    if (shadow == 1 || shadow == 2) { s.sh_off = sh16; sh16 += numel; }
    if (shadow == 3) { s.sh_off = sh32; sh32 += numel; }
    g_total += numel;
    index[name] = (int)specs.size();
    specs.push_back(s);
  }
  void linear(const std::string& p, int out, int in) {
    add(p + ".weight", (long long)out * in, 1, 1);
    add(p + ".bias", out, 0, 0);
  }
  void ln(const std::string& p, int d) {
    add(p + ".weight", d, 0, 0);
    add(p + ".bias", d, 0, 0);
  }
  void vit_block(const std::string& p, int d, int hid) {
    ln(p + ".norm1", d);
    linear(p + ".attn.qkv", 3 * d, d);
    linear(p + ".attn.proj", d, d);
    ln(p + ".norm2", d);
    linear(p + ".mlp.fc1", hid, d);
    linear(p + ".mlp.fc2", d, hid);
  }
  void qkv(const std::string& p) {  // three Linear(768,768) whose shadows / grads are laid out as one [2304,768]
    add(p + ".query.weight", 768 * 768, 1, 1);
    add(p + ".key.weight", 768 * 768, 1, 1);
    add(p + ".value.weight", 768 * 768, 1, 1);
    add(p + ".query.bias", 768, 0, 3);
    add(p + ".key.bias", 768, 0, 3);
    add(p + ".value.bias", 768, 0, 3);
  }
  void bert_layer(const std::string& p) {
    qkv(p + ".attention.self");
    linear(p + ".attention.output.dense", 768, 768);
    ln(p + ".attention.output.LayerNorm", 768);
    linear(p + ".intermediate.dense", BHID, 768);
    linear(p + ".output.dense", 768, BHID);
    ln(p + ".output.LayerNorm", 768);
  }
  Table() {
    add("patch_embed.proj.weight", 768 * 768, 1, 2);
    add("patch_embed.proj.bias", 768, 0, 0);
    add("cls_token", 768, 1, 0);
    for (int i = 0; i < EL; ++i) vit_block("blocks." + std::to_string(i), E, EHID);
    ln("norm", E);
    linear("decoder_embed", DD, E);
    add("mask_token", DD, 1, 0);
    for (int i = 0; i < DL; ++i) vit_block("decoder_blocks." + std::to_string(i), DD, DHID);
    ln("decoder_norm", DD);
    linear("decoder_pred", PDIM, DD);
    add("super_res.conv1.weight", 81, 1, 0);
    add("super_res.conv1.bias", 3, 0, 0);
    add("super_res.conv2.weight", 81, 1, 0);
    add("super_res.conv2.bias", 3, 0, 0);
    linear("bert_mlp", 768, 768);
    const std::string b = BERT;
    add(b + "embeddings.word_embeddings.weight", (long long)VOC * 768, 1, 0);
    add(b + "embeddings.position_embeddings.weight", MAXPOS * 768, 1, 0);
    add(b + "embeddings.token_type_embeddings.weight", 2 * 768, 1, 0);
    ln(b + "embeddings.LayerNorm", 768);
    const std::string f = b + "context_fusion_layer";
    qkv(f + ".attention.self");
    linear(f + ".attention.output.dense", 768, 768);
    ln(f + ".attention.output.LayerNorm", 768);
    add(f + ".cross_self_attention.query.weight", 768 * 768, 1, 1);
    add(f + ".cross_self_attention.query.bias", 768, 0, 0);
    add(f + ".cross_self_attention.key.weight", 768 * 768, 1, 1);
    add(f + ".cross_self_attention.value.weight", 768 * 768, 1, 1);
    add(f + ".cross_self_attention.key.bias", 768, 0, 3);
    add(f + ".cross_self_attention.value.bias", 768, 0, 3);
    linear(f + ".gap_mlp", 768, 768);
    linear(f + ".out_layer.dense", 768, 768);
    ln(f + ".out_layer.LayerNorm", 768);
    linear(f + ".intermediate.dense", BHID, 768);
    linear(f + ".output.dense", 768, BHID);
    ln(f + ".output.LayerNorm", 768);
    for (int i = 0; i < BL; ++i) bert_layer(b + "encoder.layer." + std::to_string(i));
    const std::string c = "bert_encoder.model.cls.predictions.";
    linear(c + "transform.dense", 768, 768);
    ln(c + "transform.LayerNorm", 768);
    add(c + "decoder.weight", (long long)VOC * 768, 1, 1);
    add(c + "bias", VOC, 0, 0);
  }
};
const Table& table() {
  static Table t;
  return t;
}

}  // namespace

const std::vector<ParamSpec>& param_specs() { return table().specs; }
long long grad_total_floats() { return table().g_total; }
long long shadow_bf16_elems() { return table().sh16; }
long long shadow_f32_elems() { return table().sh32; }
int param_index(const std::string& name) {
  auto it = table().index.find(name);
  return it == table().index.end() ? -1 : it->second;
}

// =============================================================================================
// context
// =============================================================================================
template <typename AT>
struct VitActT {
  float *x_in, *x_mid, *mean1, *rstd1, *mean2, *rstd2, *lse;
  AT *ln1, *qkv, *ao, *ln2, *pre, *act;
};
template <typename AT>
struct VitStackT {
  int L, D, H, S, hid, M, pbase;
  std::vector<VitActT<AT>> a;
  float* x_out;  // stream after the last block
};
template <typename AT>
struct BertActT {
  const AT* h_in;
  AT *qkv, *ao, *a, *pre, *act, *h_out;
  float *lse, *s1, *mean1, *rstd1, *s2, *mean2, *rstd2;
};

// AT = activation element type: AT (production) or float (fp32-accurate parity mode, the same schedule on fp32
// activations; GEMMs through gemm_hp, attention through attention_hp.cu)
// ---- side stream of the backward pass (see CtxT::side) ---------------------------------------------------------
struct SideStream {
  cudaStream_t st = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};
struct SideGuard {
  cudaEvent_t ev = nullptr;
  bool pending = false;
};
int g_side_wgrad = -1;  // -1: ECAMP_SIDE_WGRAD or 1; 2 = on, with every side-stream GEMM held back (tests, see side_delay_kernel)
// Test aid: makes the side stream lag far behind the main stream, so that a missing guard (a buffer rewritten on the main
// stream while a weight gradient still reads it) shows up as wrong gradients instead of depending on timing.
__global__ void side_delay_kernel(long long cycles) {
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) __nanosleep(200);
}
SideStream* side_stream_for_device() {  // one per device and process, created on first use, never destroyed
  static SideStream table[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& s = table[dev];
  if (!s.st) {
    if (cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.ev_join, cudaEventDisableTiming) != cudaSuccess) {
      s.st = nullptr;
      return nullptr;
    }
  }
  return &s;
}

template <typename AT>
struct CtxT {
  std::vector<float*> p;
  float *G = nullptr, *M1 = nullptr, *M2 = nullptr;
  AT* SH16 = nullptr;   // shadow copies of the GEMM weights in the activation type
  float* SH32 = nullptr;
  const float *pos = nullptr, *dpos = nullptr;
  void *adam_table = nullptr, *adam_chunks = nullptr;
  long long adam_nchunks = 0;
  bool bound = false;

  Shape sh;
  bool planned = false;
  std::map<std::string, void*> named;
  // ---- activations ----
  float *tgt, *maskf, *pe, *mean_n, *rstd_n, *mean_dn, *rstd_dn, *pred, *d_u;
  int32_t *ids_restore, *ids_keep;
  AT *a_pe, *latent, *dec_e, *dn, *lat2, *img_tok, *gap, *gp;
  VitStackT<AT> enc, dec;
  // bert
  float *emb_pre, *emb_mean, *emb_rstd, *hf[2];
  AT* emb_out;
  BertActT<AT> fus;            // fusion layer re-uses the BertAct fields for its self-attention / FFN halves
  AT *f_qc, *f_kv, *f_oc, *f_oc2, *f_a2;
  float *f_lse_c, *f_s_ol, *f_mean_ol, *f_rstd_ol;
  std::vector<BertActT<AT>> layers;
  float *t_act, *t_mean, *t_rstd, *row_loss, *loss_ws;
  AT *t_pre, *tl, *logits;
  // ---- backward scratch ----
  float *dX, *dH, *dLat, *dTL, *colsum_ws, *ln_ws, *misc_ws, *dw_pe, *delta;
  AT *gX, *dA, *dAO, *dQKV, *d_small;
  // ---- per-step state ----
  Batch batch;
  int flags = 0;
  float drop_p = 0.f;
  unsigned long long seed = 0;
  float* losses = nullptr;
  int acc = 0;
  const float* g3 = nullptr;
  cudaStream_t st = 0;
  // Weight-gradient GEMMs of the transformer blocks and of the vocabulary head run on a SIDE stream: nothing downstream in
  // backward reads dW, so they only have to be ordered after the producer of dY and before (a) the next writer of a buffer
  // they read and (b) the end of their stage, where the gradient slice is declared final.  Every GEMM is a persistent kernel
  // that occupies all SMs, so the two streams do not run side by side - the side stream's CTAs fill the tail wave (and the
  // pipeline ramp) of the main stream's kernels and vice versa, which a single stream cannot do.
  SideStream* side = nullptr;  // null = everything on one stream (fp32-accurate mode, classification path, switch off)
  bool side_busy = false;      // work has been queued on the side stream since the last join
  SideGuard guard_gx, guard_da, guard_dqkv, guard_logits;  // side-stream readers of c->gX / dA / dQKV / logits: the next writer waits
  // fp32-accurate mode only: scratch of gemm_hp (operand planes + accumulator) and of the attention backward
  void* hp_gemm_ws = nullptr;
  size_t hp_gemm_bytes = 0;
  float* hp_attn_ws = nullptr;
  // ---- fine-tune classification path (FT/Classification/models_vit.py): full 197-token encoder + pooled head ----
  const float* dp = nullptr;  // DropPath scales [12 blocks][2 branches][B] (mask / keep_prob), null = none
  int cls_B = 0;
  bool cls_planned = false;
  int32_t* cls_ids = nullptr;
  float *cls_pooled = nullptr, *cls_mean = nullptr, *cls_rstd = nullptr, *cls_dpooled = nullptr, *cls_dfeat = nullptr;
  AT *cls_feat = nullptr, *cls_dlogits = nullptr;
  // per-sample DropPath scale of (block l, branch br) of the ENCODER stack, or null
  const float* dps(const VitStackT<AT>* s, int l, int br) const {
    return (dp && s == &enc && l >= 0) ? dp + ((size_t)l * 2 + br) * sh.B : nullptr;
  }

  float* P(int i) const { return p[i]; }
  float* Gp(int i) const { return G + param_specs()[i].g_off; }
  AT* W(int i) const { return SH16 + param_specs()[i].sh_off; }
  float* B32(int i) const { return SH32 + param_specs()[i].sh_off; }
  DropoutCfg drop(unsigned long long site) const {
    DropoutCfg d;
    d.p = (flags & 1) ? drop_p : 0.f;
    d.seed = seed;
    d.site = site;
    return d;
  }
};


size_t ctx_adam_table_bytes() { return adamw_table_bytes((int)param_specs().size()); }
size_t ctx_adam_chunk_bytes() {
  std::vector<AdamTensor> t(param_specs().size());
  for (size_t i = 0; i < t.size(); ++i) t[i].numel = param_specs()[i].numel;
  return adamw_chunk_bytes(t.data(), (int)t.size());
}

inline void set_shadow(AdamTensor& t, bf16* p) { t.shadow = p; t.shadow_f = nullptr; }
inline void set_shadow(AdamTensor& t, float* p) { t.shadow = nullptr; t.shadow_f = p; }

template <typename AT>
int t_bind(CtxT<AT>* c, float* const* params, int n, float* G, float* M1, float* M2, void* shadows,
             const float* pos_embed, const float* dec_pos_embed, void* adam_table, void* adam_chunks) {
  const auto& specs = param_specs();
  ECAMP_REQUIRE(n == (int)specs.size(), "bind: expected %d parameter pointers, got %d", (int)specs.size(), n);
  ECAMP_REQUIRE(G && shadows && pos_embed && dec_pos_embed && adam_table && adam_chunks, "bind: null buffer");
  c->p.assign(params, params + n);
  for (int i = 0; i < n; ++i) ECAMP_REQUIRE(c->p[i] != nullptr, "bind: parameter %s is null", specs[i].name.c_str());
  c->G = G; c->M1 = M1; c->M2 = M2;
  c->SH16 = static_cast<AT*>(shadows);
  c->SH32 = reinterpret_cast<float*>(c->SH16 + ((shadow_bf16_elems() + 7) & ~7LL));
  c->pos = pos_embed; c->dpos = dec_pos_embed;
  c->adam_table = adam_table; c->adam_chunks = adam_chunks;
  std::vector<AdamTensor> t(n);
  for (int i = 0; i < n; ++i) {
    t[i].p = c->p[i];
    t[i].g = G + specs[i].g_off;
    t[i].m = M1 ? M1 + specs[i].g_off : nullptr;
    t[i].v = M2 ? M2 + specs[i].g_off : nullptr;
    set_shadow(t[i], (specs[i].shadow == 1 || specs[i].shadow == 2) ? c->SH16 + specs[i].sh_off : nullptr);
    t[i].shadow32 = specs[i].shadow == 3 ? c->SH32 + specs[i].sh_off : nullptr;
    t[i].numel = specs[i].numel;
    t[i].decay = specs[i].decay;
    t[i].shadow_kind = specs[i].shadow == 2 ? 1 : 0;
  }
  if (int rc = adamw_build_tables(t.data(), n, adam_table, adam_chunks, &c->adam_nchunks)) return rc;
  c->bound = true;
  return 0;
}

template <typename AT>
int t_refresh_shadows(CtxT<AT>* c, cudaStream_t st) {
  ECAMP_REQUIRE(c->bound, "refresh_shadows: context not bound");
  return refresh_shadows(c->adam_table, c->adam_chunks, c->adam_nchunks, st);
}
template <typename AT>
int t_adamw(CtxT<AT>* c, float lr, float lr_nodecay, float b1, float b2, float eps, float wd, int step, float grad_scale,
              cudaStream_t st) {
  ECAMP_REQUIRE(c->bound && c->M1 && c->M2, "adamw: context not bound with optimizer state");
  return adamw_step(c->adam_table, c->adam_chunks, c->adam_nchunks, lr, lr_nodecay, b1, b2, eps, wd, step, grad_scale, st);
}

// AdamW for the tensors whose gradients occupy flat[g_lo, g_hi) only (both ends on tensor boundaries, e.g. a backward stage
// range or a union of consecutive ones): what DataParallelStep runs on a side stream as soon as a bucket is final.
template <typename AT>
int t_adamw_range(CtxT<AT>* c, float lr, float lr_nodecay, float b1, float b2, float eps, float wd, int step,
                  float grad_scale, long long g_lo, long long g_hi, cudaStream_t st) {
  ECAMP_REQUIRE(c->bound && c->M1 && c->M2, "adamw: context not bound with optimizer state");
  const auto& specs = param_specs();
  long long chunk = 0, c0 = -1, c1 = -1;
  bool lo_ok = false, hi_ok = g_hi == g_lo;
  for (size_t i = 0; i < specs.size(); ++i) {  // the chunk table lists the tensors in this order (t_bind)
    if (specs[i].g_off == g_lo) { lo_ok = true; c0 = chunk; }
    chunk += adamw_chunks_of(specs[i].numel);
    if (specs[i].g_off + specs[i].numel == g_hi) { hi_ok = true; c1 = chunk; }
  }
  ECAMP_REQUIRE(lo_ok && hi_ok && c1 >= c0, "adamw_range: [%lld, %lld) does not start and end on tensor boundaries", g_lo, g_hi);
  return adamw_step_range(c->adam_table, c->adam_chunks, c0, c1, lr, lr_nodecay, b1, b2, eps, wd, step, grad_scale, st);
}

// ---------------------------------------------------------------------------------------------
// workspace plan: one bump allocation, run once with base = null to size it
// ---------------------------------------------------------------------------------------------
namespace {
struct Bump {
  uint8_t* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return r;
  }
};
size_t max3(size_t a, size_t b, size_t c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }

template <typename AT>
void plan_vit(Bump& bp, VitStackT<AT>& s, int L, int D, int H, int S, int hid, int M, int pbase, int B) {
  s.L = L; s.D = D; s.H = H; s.S = S; s.hid = hid; s.M = M; s.pbase = pbase;
  s.a.resize(L);
  float* x = bp.take<float>((size_t)M * D);
  for (int l = 0; l < L; ++l) {
    VitActT<AT>& a = s.a[l];
    a.x_in = x;
    a.mean1 = bp.take<float>(M); a.rstd1 = bp.take<float>(M);
    a.ln1 = bp.take<AT>((size_t)M * D);
    a.qkv = bp.take<AT>((size_t)M * 3 * D);
    a.ao = bp.take<AT>((size_t)M * D);
    a.lse = bp.take<float>((size_t)B * H * S);
    a.x_mid = bp.take<float>((size_t)M * D);
    a.mean2 = bp.take<float>(M); a.rstd2 = bp.take<float>(M);
    a.ln2 = bp.take<AT>((size_t)M * D);
    a.pre = bp.take<AT>((size_t)M * hid);
    a.act = bp.take<AT>((size_t)M * hid);
    x = bp.take<float>((size_t)M * D);
  }
  s.x_out = x;
}
template <typename AT>
void plan_bert_act(Bump& bp, BertActT<AT>& a, int Mt, int B, int T) {
  a.qkv = bp.take<AT>((size_t)Mt * 2304);
  a.ao = bp.take<AT>((size_t)Mt * 768);
  a.lse = bp.take<float>((size_t)B * BH * T);
  a.s1 = bp.take<float>((size_t)Mt * 768);
  a.mean1 = bp.take<float>(Mt); a.rstd1 = bp.take<float>(Mt);
  a.a = bp.take<AT>((size_t)Mt * 768);
  a.pre = bp.take<AT>((size_t)Mt * BHID);
  a.act = bp.take<AT>((size_t)Mt * BHID);
  a.s2 = bp.take<float>((size_t)Mt * 768);
  a.mean2 = bp.take<float>(Mt); a.rstd2 = bp.take<float>(Mt);
  a.h_out = bp.take<AT>((size_t)Mt * 768);
}

template <typename AT>
size_t plan(CtxT<AT>* c, uint8_t* base, const Shape& sh) {
  Bump bp{base};
  const int B = sh.B, T = sh.T, keep = sh.keep;
  const int Me = B * (keep + 1), Mi = B * keep, Md = B * 197, Mt = B * T;
  c->tgt = bp.take<float>((size_t)B * L196 * PDIM);
  c->maskf = bp.take<float>((size_t)B * L196);
  c->ids_restore = bp.take<int32_t>((size_t)B * L196);
  c->ids_keep = bp.take<int32_t>((size_t)B * keep + 1);
  c->a_pe = bp.take<AT>((size_t)Mi * PDIM + 8);
  c->pe = bp.take<float>((size_t)Mi * E + 4);
  const int enc_base = param_index("blocks.0.norm1.weight");
  plan_vit(bp, c->enc, EL, E, EH, keep + 1, EHID, Me, enc_base, B);
  c->mean_n = bp.take<float>(Me); c->rstd_n = bp.take<float>(Me);
  c->latent = bp.take<AT>((size_t)Me * E);
  c->dec_e = bp.take<AT>((size_t)Me * DD);
  const int dec_base = param_index("decoder_blocks.0.norm1.weight");
  plan_vit(bp, c->dec, DL, DD, DH, 197, DHID, Md, dec_base, B);
  c->mean_dn = bp.take<float>(Md); c->rstd_dn = bp.take<float>(Md);
  c->dn = bp.take<AT>((size_t)Md * DD);
  c->pred = bp.take<float>((size_t)Md * PDIM);
  c->d_u = sh.has_big ? bp.take<float>((size_t)B * 3 * 448 * 448) : nullptr;
  c->lat2 = bp.take<AT>((size_t)Me * 768);
  c->img_tok = bp.take<AT>((size_t)Mi * 768 + 8);
  c->gap = bp.take<AT>((size_t)B * 768);
  c->gp = bp.take<AT>((size_t)B * 768);
  c->emb_pre = bp.take<float>((size_t)Mt * 768);
  c->emb_mean = bp.take<float>(Mt); c->emb_rstd = bp.take<float>(Mt);
  c->emb_out = bp.take<AT>((size_t)Mt * 768);
  c->hf[0] = bp.take<float>((size_t)Mt * 768);
  c->hf[1] = bp.take<float>((size_t)Mt * 768);
  plan_bert_act(bp, c->fus, Mt, B, T);
  c->f_qc = bp.take<AT>((size_t)Mt * 768);
  c->f_kv = bp.take<AT>((size_t)Mi * 1536 + 8);
  c->f_oc = bp.take<AT>((size_t)Mt * 768);
  c->f_oc2 = bp.take<AT>((size_t)Mt * 768);
  c->f_a2 = bp.take<AT>((size_t)Mt * 768);
  c->f_lse_c = bp.take<float>((size_t)B * BH * T);
  c->f_s_ol = bp.take<float>((size_t)Mt * 768);
  c->f_mean_ol = bp.take<float>(Mt); c->f_rstd_ol = bp.take<float>(Mt);
  c->layers.resize(BL);
  for (int l = 0; l < BL; ++l) plan_bert_act(bp, c->layers[l], Mt, B, T);
  c->t_pre = bp.take<AT>((size_t)Mt * 768);
  c->t_act = bp.take<float>((size_t)Mt * 768);
  c->t_mean = bp.take<float>(Mt); c->t_rstd = bp.take<float>(Mt);
  c->tl = bp.take<AT>((size_t)Mt * 768);
  c->row_loss = bp.take<float>(Mt);
  const int ce_rows = sh.ce_rows < Mt ? sh.ce_rows : Mt;
  c->logits = bp.take<AT>((size_t)ce_rows * VOC);
  c->loss_ws = bp.take<float>(sr_ws_floats(B) + (size_t)B * L196);
  // backward scratch
  const size_t stream_elems = max3((size_t)Me * E, (size_t)Md * DD, (size_t)Mt * 768);
  const size_t hid_elems = max3((size_t)Me * EHID, (size_t)Md * DHID, (size_t)Mt * BHID);
  const size_t qkv_elems = max3((size_t)Me * 3 * E, (size_t)Md * 3 * DD, (size_t)Mt * 2304);
  c->dX = bp.take<float>(stream_elems);
  c->dH = bp.take<float>(max3(stream_elems, (size_t)Md * PDIM, 0));
  c->dLat = bp.take<float>((size_t)Me * E);
  c->dTL = bp.take<float>((size_t)Mt * 768);
  c->gX = bp.take<AT>(max3(stream_elems, (size_t)Md * PDIM, 0));
  c->dA = bp.take<AT>(hid_elems);
  c->dAO = bp.take<AT>(stream_elems);
  c->dQKV = bp.take<AT>(qkv_elems);
  c->d_small = bp.take<AT>((size_t)Mi * 1536 + (size_t)Me * 768 + (size_t)4 * B * 768 + 64);
  c->colsum_ws = bp.take<float>(colsum_ws_floats(VOC));
  c->ln_ws = bp.take<float>(layernorm_bwd_ws_floats(768));
  c->misc_ws = bp.take<float>((size_t)MAXPOS * 2 * 768 + (size_t)B * 768 + 1024);
  c->dw_pe = bp.take<float>((size_t)768 * 768);
  c->delta = bp.take<float>((size_t)B * max3((size_t)DH * 197, (size_t)EH * (keep + 1), (size_t)BH * T));
  if (is_hp<AT>::value) {
    // gemm_hp scratch: the largest (operand planes + accumulator) over the Linear shapes of the step
    const int rows_max = (int)max3((size_t)Me, (size_t)Md, (size_t)Mt);
    size_t need = gemm_hp_ws_bytes(rows_max, EHID, E);                       // widest activation GEMM (any of its three forms)
    need = std::max(need, gemm_hp_ws_bytes(rows_max, E, EHID));
    need = std::max(need, gemm_hp_ws_bytes(EHID, E, rows_max));
    need = std::max(need, gemm_hp_ws_bytes(3 * E, E, rows_max));
    need = std::max(need, gemm_hp_ws_bytes(ce_rows, VOC, 768));              // vocabulary head: forward, dgrad, wgrad
    need = std::max(need, gemm_hp_ws_bytes(ce_rows, 768, VOC));
    need = std::max(need, gemm_hp_ws_bytes(VOC, 768, ce_rows));
    c->hp_gemm_bytes = need;
    c->hp_gemm_ws = bp.take<uint8_t>(need);
    const size_t attn = max3((size_t)DH * 197 * 197, (size_t)EH * (keep + 1) * (keep + 1), (size_t)BH * T * (T > keep ? T : keep));
    c->hp_attn_ws = bp.take<float>(2 * (size_t)B * attn);
  }
  return bp.off + 256;
}
}  // namespace

template <typename AT>
size_t t_workspace_bytes(const Shape& sh) {
  CtxT<AT> tmp;
  return plan(&tmp, nullptr, sh);
}

template <typename AT>
int t_set_workspace(CtxT<AT>* c, void* ws, size_t bytes, const Shape& sh) {
  ECAMP_REQUIRE(sh.B > 0 && sh.T > 0 && sh.T <= MAXPOS && sh.keep > 0 && sh.keep <= L196 && sh.ce_rows > 0,
                "workspace: bad shape B=%d T=%d keep=%d", sh.B, sh.T, sh.keep);
  ECAMP_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "workspace: base must be 256-byte aligned");
  const size_t need = plan(c, nullptr, sh);
  ECAMP_REQUIRE(bytes >= need, "workspace: need %zu bytes, got %zu", need, bytes);
  plan(c, static_cast<uint8_t*>(ws), sh);
  c->sh = sh;
  c->planned = true;
  c->named.clear();
  c->named["tgt"] = c->tgt; c->named["latent"] = c->latent; c->named["pred"] = c->pred;
  c->named["x0"] = c->enc.a[0].x_in; c->named["enc_out"] = c->enc.x_out; c->named["dec_out"] = c->dec.x_out;
  c->named["emb_out"] = c->emb_out; c->named["fusion_out"] = c->fus.h_out; c->named["bert_out"] = c->layers[BL - 1].h_out;
  c->named["tl"] = c->tl; c->named["row_loss"] = c->row_loss; c->named["dLat"] = c->dLat; c->named["d_u"] = c->d_u;
  c->named["ids_restore32"] = c->ids_restore; c->named["lat2"] = c->lat2;
  return 0;
}
template <typename AT>
const void* t_debug_ptr(CtxT<AT>* c, const char* name) {
  auto it = c->named.find(name);
  return it == c->named.end() ? nullptr : it->second;
}

// =============================================================================================
// building blocks
// =============================================================================================
namespace {
// fc1 stores GELU'(pre-activation) for the backward pass instead of the pre-activation itself (the dGELU epilogue of the
// fc2 dgrad is instruction-bound: 39 -> 23 instructions per element); -DECAMP_NO_AUX_GRAD builds the old behaviour
#ifdef ECAMP_NO_AUX_GRAD
constexpr int kAuxGrad = 0;
#else
constexpr int kAuxGrad = GEMM_AUX_GRAD;  // same-box A/B: ~0.4 ms of the 42 ms step
#endif
#define RC(expr)          \
  do {                    \
    int _rc = (expr);     \
    if (_rc) return _rc;  \
  } while (0)

inline int gemm_any(CtxT<bf16>* c, const bf16* A, int lda, int a_mn, const bf16* B, int ldb, int b_mn, int M, int N, int K,
                    const GemmEpilogueT<bf16>& ep) {
  return gemm_bf16(A, lda, a_mn, B, ldb, b_mn, M, N, K, ep, 0, c->st);
}
inline int gemm_any(CtxT<float>* c, const float* A, int lda, int a_mn, const float* B, int ldb, int b_mn, int M, int N, int K,
                    const GemmEpilogueT<float>& ep) {
  return gemm_hp(A, lda, a_mn, B, ldb, b_mn, M, N, K, ep, c->hp_gemm_ws, c->hp_gemm_bytes, c->st);
}
inline void attn_scratch(CtxT<bf16>*, AttnArgsT<bf16>&) {}
inline void attn_scratch(CtxT<float>* c, AttnArgsT<float>& a) { a.hp_ws = c->hp_attn_ws; }

// y[M, N] = x[M, K] W^T (+ epilogue)
template <typename AT>
int lin_fwd(CtxT<AT>* c, const AT* x, int ldx, int M, const AT* W, int N, int K, GemmEpilogueT<AT> ep) {
  return gemm_any(c, x, ldx, 0, W, K, 0, M, N, K, ep);
}
// dx[M, K] = dy[M, N] W   (W stored [N, K]: the contraction index N is the slow one -> MN-major B)
template <typename AT>
int lin_dgrad(CtxT<AT>* c, const AT* dy, int ld_dy, int M, const AT* W, int N, int K, GemmEpilogueT<AT> ep) {
  return gemm_any(c, dy, ld_dy, 0, W, K, 1, M, K, N, ep);
}
// main stream: wait for the side-stream readers of a buffer before it is overwritten
template <typename AT>
int side_wait_readers(CtxT<AT>* c, SideGuard& g) {
  if (g.pending) {
    ECAMP_CUDA_OK(cudaStreamWaitEvent(c->st, g.ev, 0));
    g.pending = false;
  }
  return 0;
}
// main stream: wait for everything queued on the side stream (end of a stage: the gradient slice is final)
template <typename AT>
int side_join(CtxT<AT>* c) {
  if (c->side && c->side_busy) {
    ECAMP_CUDA_OK(cudaEventRecord(c->side->ev_join, c->side->st));
    ECAMP_CUDA_OK(cudaStreamWaitEvent(c->st, c->side->ev_join, 0));
  }
  c->side_busy = false;
  c->guard_gx.pending = c->guard_da.pending = c->guard_dqkv.pending = c->guard_logits.pending = false;
  return 0;
}
// dW[N, K] (+)= dy[M, N]^T x[M, K];  db[N] += colsum(dy)  (bias / LayerNorm gradients are zeroed by zero_small_grads at
// the start of a non-accumulating backward and only ever added to).
// side_ok: the call site has been checked for the side stream - x is a saved activation (never rewritten during backward) and
// dy is a scratch buffer whose next writer calls side_wait_readers(guard) first; every side_ok site passes the guard of its dy
// (gX / dA / dQKV / logits), because consecutive blocks of one stack are not joined in between (t_backward).
template <typename AT>
int lin_wgrad(CtxT<AT>* c, const AT* dy, int ld_dy, const AT* x, int ldx, int M, int N, int K, float* dW, float* db,
              int acc, int side_ok = 0, SideGuard* guard = nullptr) {
  GemmEpilogueT<AT> ep;
  ep.out_f32 = dW;
  ep.ld_f32 = K;
  if (acc) { ep.residual = dW; ep.ld_res = K; }
  cudaStream_t main_st = c->st;
  const bool on_side = c->side && side_ok;
  if (on_side) {  // fork: everything queued on the main stream so far (the producer of dy included) comes first
    ECAMP_CUDA_OK(cudaEventRecord(c->side->ev_fork, main_st));
    ECAMP_CUDA_OK(cudaStreamWaitEvent(c->side->st, c->side->ev_fork, 0));
    c->st = c->side->st;
    c->side_busy = true;
    if (g_side_wgrad == 2) side_delay_kernel<<<1, 1, 0, c->side->st>>>(400000);  // ~0.2 ms
  }
  int rc = gemm_any(c, dy, ld_dy, 1, x, ldx, 1, N, K, M, ep);
  // db == nullptr: the kernel that produced dy already folded the bias gradient in (LayerNorm backward, dGELU epilogue)
  if (!rc && db) rc = colsum_bf16(dy, ld_dy, M, N, db, 1, c->colsum_ws, c->st);
  c->st = main_st;
  if (!rc && on_side && guard) {
    if (!guard->ev) ECAMP_CUDA_OK(cudaEventCreateWithFlags(&guard->ev, cudaEventDisableTiming));
    ECAMP_CUDA_OK(cudaEventRecord(guard->ev, c->side->st));
    guard->pending = true;
  }
  return rc;
}
template <typename AT>
GemmEpilogueT<AT> ep_bias_bf16(const float* bias, AT* out, int ld) {
  GemmEpilogueT<AT> ep;
  ep.bias = bias; ep.out_bf16 = out; ep.ld_bf16 = ld;
  return ep;
}

template <typename AT>
AttnArgsT<AT> self_attn_args(const AT* qkv, int D3, AT* o, float* lse, int B, int H, int S, int hd) {
  AttnArgsT<AT> a;
  a.q = qkv; a.k = qkv + D3 / 3; a.v = qkv + 2 * (D3 / 3);
  a.ldq = a.ldk = a.ldv = D3;
  a.o = o; a.ldo = D3 / 3; a.lse = lse;
  a.B = B; a.H = H; a.Sq = S; a.Sk = S; a.D = hd;
  a.scale = 1.0f / sqrtf((float)hd);
  return a;
}

// ---- timm Block (pre-LN) ------------------------------------------------------------------------
template <typename AT>
int vit_block_fwd(CtxT<AT>* c, VitStackT<AT>& s, int l, float* x_next) {
  VitActT<AT>& a = s.a[l];
  const int pb = s.pbase + l * VIT_BLOCK_PARAMS, M = s.M, D = s.D, B = c->sh.B;
  RC(layernorm_fwd(a.x_in, c->P(pb + 0), c->P(pb + 1), 1e-6f, M, D, a.ln1, nullptr, a.mean1, a.rstd1, c->st));
  RC(lin_fwd(c, a.ln1, D, M, c->W(pb + 2), 3 * D, D, ep_bias_bf16(c->P(pb + 3), a.qkv, 3 * D)));
  RC(attention_fwd(self_attn_args(a.qkv, 3 * D, a.ao, a.lse, B, s.H, s.S, D / s.H), c->st));
  GemmEpilogueT<AT> ep;
  ep.bias = c->P(pb + 5); ep.residual = a.x_in; ep.ld_res = D; ep.out_f32 = a.x_mid; ep.ld_f32 = D;
  ep.row_scale = c->dps(&s, l, 0); ep.rows_per_scale = s.S;
  RC(lin_fwd(c, a.ao, D, M, c->W(pb + 4), D, D, ep));
  RC(layernorm_fwd(a.x_mid, c->P(pb + 6), c->P(pb + 7), 1e-6f, M, D, a.ln2, nullptr, a.mean2, a.rstd2, c->st));
  GemmEpilogueT<AT> e1;
  e1.bias = c->P(pb + 9); e1.flags = GEMM_GELU | kAuxGrad; e1.aux_out = a.pre; e1.ld_aux = s.hid; e1.out_bf16 = a.act; e1.ld_bf16 = s.hid;
  RC(lin_fwd(c, a.ln2, D, M, c->W(pb + 8), s.hid, D, e1));
  GemmEpilogueT<AT> e2;
  e2.bias = c->P(pb + 11); e2.residual = a.x_mid; e2.ld_res = D; e2.out_f32 = x_next; e2.ld_f32 = D;
  e2.row_scale = c->dps(&s, l, 1); e2.rows_per_scale = s.S;
  RC(lin_fwd(c, a.act, s.hid, M, c->W(pb + 10), D, s.hid, e2));
  return 0;
}
// in: c->dX = d(x_next) fp32, c->gX = its AT copy.  out: c->dX = d(x_in), c->gX = AT copy.
template <typename AT>
int vit_block_bwd(CtxT<AT>* c, VitStackT<AT>& s, int l) {
  VitActT<AT>& a = s.a[l];
  const int pb = s.pbase + l * VIT_BLOCK_PARAMS, M = s.M, D = s.D, B = c->sh.B, acc = c->acc;
  // fc2
  RC(lin_wgrad(c, c->gX, D, a.act, s.hid, M, D, s.hid, c->Gp(pb + 10), nullptr, acc, 1, &c->guard_gx));  // bias: by the producer of gX
  GemmEpilogueT<AT> e2;
  e2.flags = GEMM_DGELU | kAuxGrad; e2.aux_in = a.pre; e2.ld_aux = s.hid; e2.out_bf16 = c->dA; e2.ld_bf16 = s.hid;
  e2.colsum_out = c->Gp(pb + 9);  // fc1 bias gradient = column sums of dA
  RC(side_wait_readers(c, c->guard_da));  // the previous block's fc1 weight gradient reads dA
  RC(lin_dgrad(c, c->gX, D, M, c->W(pb + 10), D, s.hid, e2));
  // fc1
  RC(lin_wgrad(c, c->dA, s.hid, a.ln2, D, M, s.hid, D, c->Gp(pb + 8), nullptr, acc, 1, &c->guard_da));
  GemmEpilogueT<AT> e1;
  e1.out_f32 = c->dH; e1.ld_f32 = D;
  RC(lin_dgrad(c, c->dA, s.hid, M, c->W(pb + 8), s.hid, D, e1));
  RC(side_wait_readers(c, c->guard_gx));  // the fc2 weight gradient reads the gX this LayerNorm backward overwrites
  RC(layernorm_bwd(c->dH, a.x_mid, a.mean2, a.rstd2, c->P(pb + 6), M, D, c->dX, c->dX, c->gX, DropoutCfg(),
                   c->Gp(pb + 6), c->Gp(pb + 7), c->Gp(pb + 5) /* proj bias */, 1, c->st, c->dps(&s, l, 0), s.S));
  // proj
  RC(lin_wgrad(c, c->gX, D, a.ao, D, M, D, D, c->Gp(pb + 4), nullptr, acc, 1, &c->guard_gx));
  GemmEpilogueT<AT> ep;
  ep.out_bf16 = c->dAO; ep.ld_bf16 = D;
  RC(lin_dgrad(c, c->gX, D, M, c->W(pb + 4), D, D, ep));
  AttnArgsT<AT> at = self_attn_args(a.qkv, 3 * D, a.ao, a.lse, B, s.H, s.S, D / s.H);
  at.d_o = c->dAO; at.ld_do = D; at.delta = c->delta;
  at.dq = c->dQKV; at.dk = c->dQKV + D; at.dv = c->dQKV + 2 * D;
  at.lddq = at.lddk = at.lddv = 3 * D;
  // (the qkv bias gradient stays with colsum_kernel here: folded into the mma.sync kernels it costs more in atomics -
  //  +0.26 ms for the encoder, +0.16 ms for the decoder - than the 16 column-sum launches it saves, ~0.2 ms)
  attn_scratch(c, at);
  RC(side_wait_readers(c, c->guard_dqkv));  // the previous block's qkv weight gradient reads dQKV
  RC(attention_bwd(at, c->st));
  // qkv
  RC(lin_wgrad(c, c->dQKV, 3 * D, a.ln1, D, M, 3 * D, D, c->Gp(pb + 2), c->Gp(pb + 3), acc, 1, &c->guard_dqkv));
  GemmEpilogueT<AT> eq;
  eq.out_f32 = c->dH; eq.ld_f32 = D;
  RC(lin_dgrad(c, c->dQKV, 3 * D, M, c->W(pb + 2), 3 * D, D, eq));
  RC(side_wait_readers(c, c->guard_gx));  // the proj weight gradient reads gX
  // the AT gradient emitted here is the dY of the PREVIOUS block's fc2: its bias gradient is folded in
  RC(layernorm_bwd(c->dH, a.x_in, a.mean1, a.rstd1, c->P(pb + 0), M, D, c->dX, c->dX, c->gX, DropoutCfg(),
                   c->Gp(pb + 0), c->Gp(pb + 1), l > 0 ? c->Gp(pb - VIT_BLOCK_PARAMS + 11) : nullptr, 1, c->st,
                   c->dps(&s, l - 1, 1), s.S));
  return 0;
}

// ---- HF BertLayer (post-LN): self-attention half and FFN half, shared with the fusion layer -------
// pb indexes: q.w k.w v.w q.b k.b v.b | ao.dense w b | ao.LN w b
template <typename AT>
int bert_attn_half_fwd(CtxT<AT>* c, BertActT<AT>& a, int pb, const AT* h_in, const float* h_in_f32, unsigned long long site) {
  const int Mt = c->sh.B * c->sh.T, B = c->sh.B, T = c->sh.T;
  a.h_in = h_in;
  RC(lin_fwd(c, h_in, 768, Mt, c->W(pb + 0), 2304, 768, ep_bias_bf16(c->B32(pb + 3), a.qkv, 2304)));
  AttnArgsT<AT> at = self_attn_args(a.qkv, 2304, a.ao, a.lse, B, BH, T, 128);
  at.key_mask = c->batch.attention_mask;
  at.drop = c->drop(site);
  RC(attention_fwd(at, c->st));
  GemmEpilogueT<AT> ep;
  ep.bias = c->P(pb + 7); ep.residual = h_in_f32; ep.ld_res = 768; ep.out_f32 = a.s1; ep.ld_f32 = 768;
  const DropoutCfg d = c->drop(site + 1);
  if (d.p > 0.f) { ep.flags |= GEMM_DROPOUT; ep.drop_p = d.p; ep.seed = d.seed; ep.stream = d.site; }
  RC(lin_fwd(c, a.ao, 768, Mt, c->W(pb + 6), 768, 768, ep));
  return 0;
}
// in: c->dX = d(LN output of this half) fp32.  out: c->dX = d(h_in) fp32 (residual + qkv paths).
template <typename AT>
int bert_attn_half_bwd(CtxT<AT>* c, BertActT<AT>& a, int pb, unsigned long long site, int side_ok = 0) {
  const int Mt = c->sh.B * c->sh.T, B = c->sh.B, T = c->sh.T, acc = c->acc;
  RC(side_wait_readers(c, c->guard_gx));  // a weight gradient of the FFN half may still be reading gX
  RC(layernorm_bwd(c->dX, a.s1, a.mean1, a.rstd1, c->P(pb + 8), Mt, 768, nullptr, c->dX, c->gX, c->drop(site + 1),
                   c->Gp(pb + 8), c->Gp(pb + 9), c->Gp(pb + 7) /* attention.output.dense bias */, 1, c->st));
  RC(lin_wgrad(c, c->gX, 768, a.ao, 768, Mt, 768, 768, c->Gp(pb + 6), nullptr, acc, side_ok, &c->guard_gx));
  GemmEpilogueT<AT> ep;
  ep.out_bf16 = c->dAO; ep.ld_bf16 = 768;
  RC(lin_dgrad(c, c->gX, 768, Mt, c->W(pb + 6), 768, 768, ep));
  AttnArgsT<AT> at = self_attn_args(a.qkv, 2304, a.ao, a.lse, B, BH, T, 128);
  at.key_mask = c->batch.attention_mask;
  at.drop = c->drop(site);
  at.d_o = c->dAO; at.ld_do = 768; at.delta = c->delta;
  at.dq = c->dQKV; at.dk = c->dQKV + 768; at.dv = c->dQKV + 1536;
  at.lddq = at.lddk = at.lddv = 2304;
  at.cs_q = c->Gp(pb + 3); at.cs_k = c->Gp(pb + 3) + 768; at.cs_v = c->Gp(pb + 3) + 1536;  // q | k | v bias gradients
  attn_scratch(c, at);
  RC(side_wait_readers(c, c->guard_dqkv));
  RC(attention_bwd(at, c->st));
  RC(lin_wgrad(c, c->dQKV, 2304, a.h_in, 768, Mt, 2304, 768, c->Gp(pb + 0), nullptr, acc, side_ok, &c->guard_dqkv));
  GemmEpilogueT<AT> eq;
  eq.residual = c->dX; eq.ld_res = 768; eq.out_f32 = c->dX; eq.ld_f32 = 768;
  RC(lin_dgrad(c, c->dQKV, 2304, Mt, c->W(pb + 0), 2304, 768, eq));
  return 0;
}
// FFN half: pi = intermediate.dense.w (b = pi+1), output.dense w b = pi+2, pi+3, output.LN = pi+4, pi+5
template <typename AT>
int bert_ffn_half_fwd(CtxT<AT>* c, BertActT<AT>& a, int pi, const AT* x, const float* x_f32, unsigned long long site) {
  const int Mt = c->sh.B * c->sh.T;
  GemmEpilogueT<AT> e1;
  e1.bias = c->P(pi + 1); e1.flags = GEMM_GELU | kAuxGrad; e1.aux_out = a.pre; e1.ld_aux = BHID; e1.out_bf16 = a.act; e1.ld_bf16 = BHID;
  RC(lin_fwd(c, x, 768, Mt, c->W(pi), BHID, 768, e1));
  GemmEpilogueT<AT> e2;
  e2.bias = c->P(pi + 3); e2.residual = x_f32; e2.ld_res = 768; e2.out_f32 = a.s2; e2.ld_f32 = 768;
  const DropoutCfg d = c->drop(site);
  if (d.p > 0.f) { e2.flags |= GEMM_DROPOUT; e2.drop_p = d.p; e2.seed = d.seed; e2.stream = d.site; }
  RC(lin_fwd(c, a.act, BHID, Mt, c->W(pi + 2), 768, BHID, e2));
  return 0;
}
// in: c->dX = d(h_out).  out: c->dX = d(x) (the FFN input = LN output of the previous half)
template <typename AT>
int bert_ffn_half_bwd(CtxT<AT>* c, BertActT<AT>& a, int pi, const AT* x, unsigned long long site, int side_ok = 0) {
  const int Mt = c->sh.B * c->sh.T, acc = c->acc;
  RC(side_wait_readers(c, c->guard_gx));
  RC(layernorm_bwd(c->dX, a.s2, a.mean2, a.rstd2, c->P(pi + 4), Mt, 768, nullptr, c->dX, c->gX, c->drop(site),
                   c->Gp(pi + 4), c->Gp(pi + 5), c->Gp(pi + 3) /* output.dense bias */, 1, c->st));
  RC(lin_wgrad(c, c->gX, 768, a.act, BHID, Mt, 768, BHID, c->Gp(pi + 2), nullptr, acc, side_ok, &c->guard_gx));
  GemmEpilogueT<AT> e2;
  e2.flags = GEMM_DGELU | kAuxGrad; e2.aux_in = a.pre; e2.ld_aux = BHID; e2.out_bf16 = c->dA; e2.ld_bf16 = BHID;
  e2.colsum_out = c->Gp(pi + 1);  // intermediate.dense bias gradient
  RC(side_wait_readers(c, c->guard_da));
  RC(lin_dgrad(c, c->gX, 768, Mt, c->W(pi + 2), 768, BHID, e2));
  RC(lin_wgrad(c, c->dA, BHID, x, 768, Mt, BHID, 768, c->Gp(pi), nullptr, acc, side_ok, &c->guard_da));
  GemmEpilogueT<AT> e1;
  e1.residual = c->dX; e1.ld_res = 768; e1.out_f32 = c->dX; e1.ld_f32 = 768;
  RC(lin_dgrad(c, c->dA, BHID, Mt, c->W(pi), BHID, 768, e1));
  return 0;
}

template <typename AT>
int bert_layer_fwd(CtxT<AT>* c, int l, const AT* h_in, int hf_in) {
  BertActT<AT>& a = c->layers[l];
  const int pb = param_index(std::string(BERT) + "encoder.layer." + std::to_string(l) + ".attention.self.query.weight");
  const int Mt = c->sh.B * c->sh.T;
  const unsigned long long site = 100 + 10 * l;
  RC(bert_attn_half_fwd(c, a, pb, h_in, c->hf[hf_in], site));
  RC(layernorm_fwd(a.s1, c->P(pb + 8), c->P(pb + 9), 1e-12f, Mt, 768, a.a, c->hf[hf_in ^ 1], a.mean1, a.rstd1, c->st));
  RC(bert_ffn_half_fwd(c, a, pb + 10, a.a, c->hf[hf_in ^ 1], site + 2));
  RC(layernorm_fwd(a.s2, c->P(pb + 14), c->P(pb + 15), 1e-12f, Mt, 768, a.h_out, c->hf[hf_in], a.mean2, a.rstd2, c->st));
  return 0;
}
template <typename AT>
int bert_layer_bwd(CtxT<AT>* c, int l) {
  BertActT<AT>& a = c->layers[l];
  const int pb = param_index(std::string(BERT) + "encoder.layer." + std::to_string(l) + ".attention.self.query.weight");
  const unsigned long long site = 100 + 10 * l;
  // the four weight gradients of a layer go to the side stream (the fusion layer, which shares the two halves, keeps them on
  // the main stream: its stage rewrites dA / dQKV more than once)
  RC(bert_ffn_half_bwd(c, a, pb + 10, a.a, site + 2, 1));
  RC(bert_attn_half_bwd(c, a, pb, site, 1));
  return 0;
}

// ---- LM head (bert_modeling.py:208-217) ---------------------------------------------------------------
template <typename AT>
int lm_transform_fwd(CtxT<AT>* c) {
  const int Mt = c->sh.B * c->sh.T;
  const int pt = param_index("bert_encoder.model.cls.predictions.transform.dense.weight");
  GemmEpilogueT<AT> e;
  e.bias = c->P(pt + 1); e.flags = GEMM_GELU; e.aux_out = c->t_pre; e.ld_aux = 768; e.out_f32 = c->t_act; e.ld_f32 = 768;
  RC(lin_fwd(c, c->layers[BL - 1].h_out, 768, Mt, c->W(pt), 768, 768, e));
  RC(layernorm_fwd(c->t_act, c->P(pt + 2), c->P(pt + 3), 1e-12f, Mt, 768, c->tl, nullptr, c->t_mean, c->t_rstd, c->st));
  return 0;
}
inline int ce_rows_any(CtxT<bf16>* c, int rows, int r0, bool with_grad, float inv_total, float* bias_grad) {
  return ce_chunk(c->logits, VOC, rows, VOC, c->batch.labels + r0, c->batch.weights + r0, c->row_loss + r0,
                  with_grad ? c->g3 + 2 : nullptr, inv_total, with_grad ? 1 : 0, c->st, bias_grad);
}
inline int ce_rows_any(CtxT<float>* c, int rows, int r0, bool with_grad, float inv_total, float*) {
  return ce_chunk(c->logits, VOC, rows, VOC, c->batch.labels + r0, c->batch.weights + r0, c->row_loss + r0,
                  with_grad ? c->g3 + 2 : nullptr, inv_total, with_grad ? 1 : 0, c->st);
}
template <typename AT>
int lm_chunks(CtxT<AT>* c, bool with_grad, bool write_loss) {
  const int Mt = c->sh.B * c->sh.T;
  const int pt = param_index("bert_encoder.model.cls.predictions.transform.dense.weight");
  const int pw = pt + 4, pbias = pt + 5;
  const int R = c->sh.ce_rows < Mt ? c->sh.ce_rows : Mt;
  int acc = c->acc;
  for (int r0 = 0; r0 < Mt; r0 += R) {
    const int rows = Mt - r0 < R ? Mt - r0 : R;
    RC(side_wait_readers(c, c->guard_logits));  // the previous chunk's weight gradient reads the logits buffer
    RC(lin_fwd(c, c->tl + (size_t)r0 * 768, 768, rows, c->W(pw), VOC, 768, ep_bias_bf16(c->P(pbias), c->logits, VOC)));
    RC(ce_rows_any(c, rows, r0, with_grad, 1.0f / (float)Mt, with_grad ? c->Gp(pbias) : nullptr));
    if (with_grad) {
      GemmEpilogueT<AT> e;
      e.out_f32 = c->dTL + (size_t)r0 * 768; e.ld_f32 = 768;
      RC(lin_dgrad(c, c->logits, VOC, rows, c->W(pw), VOC, 768, e));
      // the bias gradient: by the cross-entropy kernel (bf16 mode) or a column-sum pass (fp32-accurate mode)
      RC(lin_wgrad(c, c->logits, VOC, c->tl + (size_t)r0 * 768, 768, rows, VOC, 768, c->Gp(pw), is_hp<AT>::value ? c->Gp(pbias) : nullptr, acc,
                   1, &c->guard_logits));
      acc = 1;
    }
  }
  if (write_loss) RC(sum_to_scalar(c->row_loss, (size_t)Mt, 1.0f / (float)Mt, c->losses + 2, c->st));
  return 0;
}
template <typename AT>
int lm_transform_bwd(CtxT<AT>* c) {  // in: dTL.  out: c->dX = d(bert output)
  const int Mt = c->sh.B * c->sh.T, acc = c->acc;
  const int pt = param_index("bert_encoder.model.cls.predictions.transform.dense.weight");
  RC(layernorm_bwd(c->dTL, c->t_act, c->t_mean, c->t_rstd, c->P(pt + 2), Mt, 768, nullptr, c->dTL, (AT*)nullptr,
                   DropoutCfg(), c->Gp(pt + 2), c->Gp(pt + 3), nullptr, 1, c->st));
  RC(gelu_bwd_bf16(c->dTL, c->t_pre, c->gX, (size_t)Mt * 768, c->st));
  RC(lin_wgrad(c, c->gX, 768, c->layers[BL - 1].h_out, 768, Mt, 768, 768, c->Gp(pt), c->Gp(pt + 1), acc));
  GemmEpilogueT<AT> e;
  e.out_f32 = c->dX; e.ld_f32 = 768;
  RC(lin_dgrad(c, c->gX, 768, Mt, c->W(pt), 768, 768, e));
  return 0;
}

// ---- fusion layer + embeddings + bert_mlp (context_fusion.py:21-67, bert_modeling.py:113-129, model_ecamp.py:267-271)
struct FusionIdx {
  int qkv, cq, ckv, gap, ol, inter;
  FusionIdx() {
    const std::string f = std::string(BERT) + "context_fusion_layer";
    qkv = param_index(f + ".attention.self.query.weight");
    cq = param_index(f + ".cross_self_attention.query.weight");
    ckv = param_index(f + ".cross_self_attention.key.weight");
    gap = param_index(f + ".gap_mlp.weight");
    ol = param_index(f + ".out_layer.dense.weight");
    inter = param_index(f + ".intermediate.dense.weight");
  }
};
const FusionIdx& fidx() {
  static FusionIdx f;
  return f;
}

template <typename AT>
int text_front_fwd(CtxT<AT>* c) {
  const int B = c->sh.B, T = c->sh.T, keep = c->sh.keep, Mt = B * T, Mi = B * keep, Me = B * (keep + 1);
  const FusionIdx& f = fidx();
  const int pm = param_index("bert_mlp.weight");
  const int pe = param_index(std::string(BERT) + "embeddings.word_embeddings.weight");
  // bert_mlp on the latent, drop cls, GAP token
  RC(lin_fwd(c, c->latent, 768, Me, c->W(pm), 768, 768, ep_bias_bf16(c->P(pm + 1), c->lat2, 768)));
  RC(split_latent_gap(c->lat2, B, keep, 768, c->img_tok, c->gap, c->st));
  // embeddings
  RC(bert_embeddings_fwd(c->batch.ids, c->batch.type_ids, c->P(pe), c->P(pe + 2), c->P(pe + 1), c->P(pe + 3),
                         c->P(pe + 4), 1e-12f, B, T, 768, c->drop(1), c->emb_pre, c->emb_mean, c->emb_rstd, c->emb_out,
                         c->hf[0], c->st));
  // text self-attention half
  BertActT<AT>& a = c->fus;
  RC(bert_attn_half_fwd(c, a, f.qkv, c->emb_out, c->hf[0], 10));
  RC(layernorm_fwd(a.s1, c->P(f.qkv + 8), c->P(f.qkv + 9), 1e-12f, Mt, 768, a.a, c->hf[1], a.mean1, a.rstd1, c->st));
  // cross attention: Q from text, K/V from the image tokens, no mask
  RC(lin_fwd(c, a.a, 768, Mt, c->W(f.cq), 768, 768, ep_bias_bf16(c->P(f.cq + 1), c->f_qc, 768)));
  RC(lin_fwd(c, c->img_tok, 768, Mi, c->W(f.ckv), 1536, 768, ep_bias_bf16(c->B32(f.ckv + 2), c->f_kv, 1536)));
  AttnArgsT<AT> at;
  at.q = c->f_qc; at.ldq = 768; at.k = c->f_kv; at.v = c->f_kv + 768; at.ldk = at.ldv = 1536;
  at.o = c->f_oc; at.ldo = 768; at.lse = c->f_lse_c;
  at.B = B; at.H = BH; at.Sq = T; at.Sk = keep; at.D = 128; at.scale = 1.0f / sqrtf(128.f);
  at.drop = c->drop(12);
  RC(attention_fwd(at, c->st));
  // + gap_mlp(gap_token), broadcast over the text positions
  RC(lin_fwd(c, c->gap, 768, B, c->W(f.gap), 768, 768, ep_bias_bf16(c->P(f.gap + 1), c->gp, 768)));
  RC(add_batch_rowvec_oop(c->f_oc, c->gp, B, T, 768, c->f_oc2, c->st));
  // out_layer: dense -> dropout -> LN(. + attention_output)
  GemmEpilogueT<AT> eo;
  eo.bias = c->P(f.ol + 1); eo.residual = c->hf[1]; eo.ld_res = 768; eo.out_f32 = c->f_s_ol; eo.ld_f32 = 768;
  const DropoutCfg d = c->drop(13);
  if (d.p > 0.f) { eo.flags |= GEMM_DROPOUT; eo.drop_p = d.p; eo.seed = d.seed; eo.stream = d.site; }
  RC(lin_fwd(c, c->f_oc2, 768, Mt, c->W(f.ol), 768, 768, eo));
  RC(layernorm_fwd(c->f_s_ol, c->P(f.ol + 2), c->P(f.ol + 3), 1e-12f, Mt, 768, c->f_a2, c->hf[0], c->f_mean_ol,
                   c->f_rstd_ol, c->st));
  // FFN half
  RC(bert_ffn_half_fwd(c, a, f.inter, c->f_a2, c->hf[0], 14));
  RC(layernorm_fwd(a.s2, c->P(f.inter + 4), c->P(f.inter + 5), 1e-12f, Mt, 768, a.h_out, c->hf[1], a.mean2, a.rstd2,
                   c->st));
  return 0;  // fusion output: a.h_out (AT) / hf[1] (fp32)
}

// in: c->dX = d(fusion output).  Produces every gradient of the text front-end and the latent gradient dLat (stored).
template <typename AT>
int text_front_bwd(CtxT<AT>* c) {
  const int B = c->sh.B, T = c->sh.T, keep = c->sh.keep, Mt = B * T, Mi = B * keep, Me = B * (keep + 1);
  const int acc = c->acc;
  const FusionIdx& f = fidx();
  const int pm = param_index("bert_mlp.weight");
  const int pe = param_index(std::string(BERT) + "embeddings.word_embeddings.weight");
  BertActT<AT>& a = c->fus;
  AT* d_kv = c->d_small;                    // [Mi, 1536]
  AT* d_lat2 = d_kv + (size_t)Mi * 1536;    // [Me, 768]
  AT* d_gp = d_lat2 + (size_t)Me * 768;     // [B, 768]
  AT* d_gap = d_gp + (size_t)B * 768;       // [B, 768]
  AT* d_img = c->dAO;                       // [Mi, 768] (dAO is free between its uses)

  RC(bert_ffn_half_bwd(c, a, f.inter, c->f_a2, 14));  // dX = d(a2)
  // out_layer
  RC(layernorm_bwd(c->dX, c->f_s_ol, c->f_mean_ol, c->f_rstd_ol, c->P(f.ol + 2), Mt, 768, nullptr, c->dX, c->gX,
                   c->drop(13), c->Gp(f.ol + 2), c->Gp(f.ol + 3), c->Gp(f.ol + 1) /* out_layer.dense bias */, 1, c->st));
  RC(lin_wgrad(c, c->gX, 768, c->f_oc2, 768, Mt, 768, 768, c->Gp(f.ol), nullptr, acc));
  GemmEpilogueT<AT> eoc;
  eoc.out_bf16 = c->dAO; eoc.ld_bf16 = 768;
  RC(lin_dgrad(c, c->gX, 768, Mt, c->W(f.ol), 768, 768, eoc));  // dAO = d(oc2) = d(oc)
  // gap_mlp path
  RC(batch_colsum(c->dAO, B, T, 768, d_gp, c->st));
  RC(lin_wgrad(c, d_gp, 768, c->gap, 768, B, 768, 768, c->Gp(f.gap), c->Gp(f.gap + 1), acc));
  RC(lin_dgrad(c, d_gp, 768, B, c->W(f.gap), 768, 768, ep_bias_bf16(nullptr, d_gap, 768)));
  // cross attention backward
  AttnArgsT<AT> at;
  at.q = c->f_qc; at.ldq = 768; at.k = c->f_kv; at.v = c->f_kv + 768; at.ldk = at.ldv = 1536;
  at.o = c->f_oc; at.ldo = 768; at.lse = c->f_lse_c;
  at.B = B; at.H = BH; at.Sq = T; at.Sk = keep; at.D = 128; at.scale = 1.0f / sqrtf(128.f);
  at.drop = c->drop(12);
  at.d_o = c->dAO; at.ld_do = 768; at.delta = c->delta;
  at.dq = c->dQKV; at.lddq = 768; at.dk = d_kv; at.dv = d_kv + 768; at.lddk = at.lddv = 1536;
  at.cs_q = c->Gp(f.cq + 1); at.cs_k = c->Gp(f.ckv + 2); at.cs_v = c->Gp(f.ckv + 2) + 768;  // cross q / k | v bias gradients
  attn_scratch(c, at);
  RC(attention_bwd(at, c->st));
  // cross query: d(a1) = dQc Wcq + dX (residual of out_layer)
  RC(lin_wgrad(c, c->dQKV, 768, a.a, 768, Mt, 768, 768, c->Gp(f.cq), nullptr, acc));
  GemmEpilogueT<AT> eq;
  eq.residual = c->dX; eq.ld_res = 768; eq.out_f32 = c->dX; eq.ld_f32 = 768;
  RC(lin_dgrad(c, c->dQKV, 768, Mt, c->W(f.cq), 768, 768, eq));
  // cross key/value -> image tokens
  RC(lin_wgrad(c, d_kv, 1536, c->img_tok, 768, Mi, 1536, 768, c->Gp(f.ckv), nullptr, acc));
  RC(lin_dgrad(c, d_kv, 1536, Mi, c->W(f.ckv), 1536, 768, ep_bias_bf16(nullptr, d_img, 768)));
  // bert_mlp: latent gradient from the text branch (stored into dLat; the image decoder accumulates later)
  RC(split_latent_gap_bwd(d_img, d_gap, B, keep, 768, d_lat2, c->st));
  RC(lin_wgrad(c, d_lat2, 768, c->latent, 768, Me, 768, 768, c->Gp(pm), c->Gp(pm + 1), acc));
  GemmEpilogueT<AT> el;
  el.out_f32 = c->dLat; el.ld_f32 = 768;
  RC(lin_dgrad(c, d_lat2, 768, Me, c->W(pm), 768, 768, el));
  // text self-attention half: dX = d(emb_out)
  RC(bert_attn_half_bwd(c, a, f.qkv, 10));
  // embeddings: dropout -> LN -> tables
  RC(dropout_bwd_f32(c->dX, (size_t)Mt * 768, c->drop(1), c->st));
  RC(layernorm_bwd(c->dX, c->emb_pre, c->emb_mean, c->emb_rstd, c->P(pe + 3), Mt, 768, nullptr, c->dX, (AT*)nullptr,
                   DropoutCfg(), c->Gp(pe + 3), c->Gp(pe + 4), nullptr, 1, c->st));
  RC(bert_embeddings_bwd(c->dX, c->batch.ids, c->batch.type_ids, B, T, 768, c->Gp(pe), c->Gp(pe + 2), c->Gp(pe + 1),
                         acc, c->misc_ws, c->st));
  return 0;
}

// ---- image side ---------------------------------------------------------------------------------------
template <typename AT>
int image_encoder_fwd(CtxT<AT>* c, float* mask_out, int64_t* ids_restore_out, int64_t* ids_keep_out) {
  const int B = c->sh.B, keep = c->sh.keep, Mi = B * keep, Me = B * (keep + 1);
  if (c->sh.has_big) RC(resize_bicubic_patchify(c->batch.image, B, 448, c->tgt, c->st));
  else RC(patchify224(c->batch.image, B, c->tgt, c->st));
  RC(random_masking(c->batch.noise, B, L196, keep, c->ids_restore, c->ids_keep, mask_out ? mask_out : c->maskf,
                    ids_restore_out, ids_keep_out, c->st));
  if (mask_out)
    ECAMP_CUDA_OK(cudaMemcpyAsync(c->maskf, mask_out, (size_t)B * L196 * sizeof(float), cudaMemcpyDeviceToDevice, c->st));
  RC(gather_patches(c->tgt, c->ids_keep, B, L196, keep, PDIM, c->a_pe, c->st));
  GemmEpilogueT<AT> e;
  e.bias = c->P(1); e.out_f32 = c->pe; e.ld_f32 = E;
  RC(lin_fwd(c, c->a_pe, PDIM, Mi, c->W(0), E, PDIM, e));
  RC(assemble_encoder_input(c->pe, c->P(2), c->pos, c->ids_keep, B, keep, E, c->enc.a[0].x_in, c->st));
  for (int l = 0; l < EL; ++l) RC(vit_block_fwd(c, c->enc, l, l + 1 < EL ? c->enc.a[l + 1].x_in : c->enc.x_out));
  const int pn = param_index("norm.weight");
  RC(layernorm_fwd(c->enc.x_out, c->P(pn), c->P(pn + 1), 1e-6f, Me, E, c->latent, nullptr, c->mean_n, c->rstd_n, c->st));
  return 0;
}
template <typename AT>
int image_decoder_fwd(CtxT<AT>* c) {
  const int B = c->sh.B, keep = c->sh.keep, Me = B * (keep + 1), Md = B * 197;
  const int pde = param_index("decoder_embed.weight");
  RC(lin_fwd(c, c->latent, E, Me, c->W(pde), DD, E, ep_bias_bf16(c->P(pde + 1), c->dec_e, DD)));
  RC(assemble_decoder_input(c->dec_e, c->P(pde + 2), c->dpos, c->ids_restore, B, L196, keep, DD, c->dec.a[0].x_in, c->st));
  for (int l = 0; l < DL; ++l) RC(vit_block_fwd(c, c->dec, l, l + 1 < DL ? c->dec.a[l + 1].x_in : c->dec.x_out));
  const int pn = param_index("decoder_norm.weight");
  RC(layernorm_fwd(c->dec.x_out, c->P(pn), c->P(pn + 1), 1e-6f, Md, DD, c->dn, nullptr, c->mean_dn, c->rstd_dn, c->st));
  GemmEpilogueT<AT> e;
  e.bias = c->P(pn + 3); e.out_f32 = c->pred; e.ld_f32 = PDIM;
  RC(lin_fwd(c, c->dn, DD, Md, c->W(pn + 2), PDIM, DD, e));
  return 0;
}
template <typename AT>
int image_losses_fwd(CtxT<AT>* c) {
  const int B = c->sh.B;
  RC(mim_loss_fwd(c->pred, 197, c->tgt, c->maskf, B, L196, PDIM, c->losses + 0, c->loss_ws, c->st));
  if (c->sh.has_big && (c->flags & 2)) {
    // fused training step: sr_bwd_kernel recomputes the head anyway and emits the loss (stage 8)
  } else if (c->sh.has_big) {
    const int ps = param_index("super_res.conv1.weight");
    RC(sr_loss_fwd(c->pred, c->batch.image, c->batch.column, c->batch.row, c->P(ps), c->P(ps + 1), c->P(ps + 2),
                   c->P(ps + 3), B, c->losses + 1, c->loss_ws, c->st));
  } else {
    ECAMP_CUDA_OK(cudaMemsetAsync(c->losses + 1, 0, sizeof(float), c->st));
  }
  return 0;
}

}  // namespace

// =============================================================================================
// fine-tune classification (FT/Classification/models_vit.py:60-98 with global_pool; train.py:438-465): the same
// encoder blocks on the FULL 197-token sequence (no masking), DropPath on both branches, mean over the patch tokens
// -> fc_norm -> head.  Re-uses the context (parameter table entries patch_embed / cls_token / blocks.*, flat gradient
// buffer, bf16 shadows); pos_embed (learnable here), fc_norm, head come in through ClsIO.
// =============================================================================================
namespace {
template <typename AT>
int zero_small_grads(CtxT<AT>* c);  // defined with the backward entry points below
constexpr int CLS_S = 197, CLS_PAD = 16;  // logits are padded to 16 columns (GEMM operand pitch: 16 bytes)

size_t plan_cls(CtxT<bf16>* c, uint8_t* base, int B) {
  Bump bp{base};
  const int M = B * CLS_S, Mp = B * L196;
  c->tgt = bp.take<float>((size_t)Mp * PDIM);
  c->cls_ids = bp.take<int32_t>((size_t)Mp);
  c->a_pe = bp.take<bf16>((size_t)Mp * PDIM + 8);
  c->pe = bp.take<float>((size_t)Mp * E + 4);
  const int enc_base = param_index("blocks.0.norm1.weight");
  plan_vit(bp, c->enc, EL, E, EH, CLS_S, EHID, M, enc_base, B);
  c->cls_pooled = bp.take<float>((size_t)B * E);
  c->cls_mean = bp.take<float>(B); c->cls_rstd = bp.take<float>(B);
  c->cls_feat = bp.take<bf16>((size_t)B * E);
  c->cls_dlogits = bp.take<bf16>((size_t)B * CLS_PAD);
  c->cls_dfeat = bp.take<float>((size_t)B * E);
  c->cls_dpooled = bp.take<float>((size_t)B * E);
  c->dX = bp.take<float>((size_t)M * E);
  c->dH = bp.take<float>((size_t)M * E);
  c->gX = bp.take<bf16>((size_t)M * E);
  c->dA = bp.take<bf16>((size_t)M * EHID);
  c->dAO = bp.take<bf16>((size_t)M * E);
  c->dQKV = bp.take<bf16>((size_t)M * 3 * E);
  c->colsum_ws = bp.take<float>(64);
  c->dw_pe = bp.take<float>((size_t)768 * 768);
  c->delta = bp.take<float>((size_t)B * EH * CLS_S);
  return bp.off + 256;
}
}  // namespace

size_t cls_workspace_bytes(int B) {
  CtxT<bf16> tmp;
  return plan_cls(&tmp, nullptr, B);
}
int t_set_cls_workspace(CtxT<bf16>* c, void* ws, size_t bytes, int B) {
  ECAMP_REQUIRE(B > 0, "cls workspace: bad batch %d", B);
  ECAMP_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "cls workspace: base must be 256-byte aligned");
  const size_t need = plan_cls(c, nullptr, B);
  ECAMP_REQUIRE(bytes >= need, "cls workspace: need %zu bytes, got %zu", need, bytes);
  plan_cls(c, static_cast<uint8_t*>(ws), B);
  c->sh = Shape();
  c->sh.B = B; c->sh.T = 1; c->sh.keep = L196; c->sh.has_big = 0; c->sh.ce_rows = 1;
  c->planned = false;  // the pre-training plan (if any) no longer matches this workspace
  c->cls_planned = true;
  c->cls_B = B;
  return 0;
}

int t_cls_forward(CtxT<bf16>* c, const ClsIO& io, cudaStream_t st) {
  ECAMP_REQUIRE(c->bound && c->cls_planned, "cls forward: context needs bind() and set_cls_workspace() first");
  ECAMP_REQUIRE(io.image && io.pos_embed && io.fc_norm_w && io.fc_norm_b && io.head_w16 && io.head_b && io.logits,
                "cls forward: null argument");
  const int B = c->cls_B, M = B * CLS_S, Mp = B * L196;
  c->st = st;
  c->dp = io.dp_scale;
  RC(patchify224(io.image, B, c->tgt, st));
  RC(iota_mod_i32(c->cls_ids, (size_t)Mp, L196, st));
  RC(gather_patches(c->tgt, c->cls_ids, B, L196, L196, PDIM, c->a_pe, st));
  GemmEpilogue e;
  e.bias = c->P(1); e.out_f32 = c->pe; e.ld_f32 = E;
  RC(lin_fwd(c, c->a_pe, PDIM, Mp, c->W(0), E, PDIM, e));
  RC(assemble_encoder_input(c->pe, c->P(2), io.pos_embed, c->cls_ids, B, L196, E, c->enc.a[0].x_in, st));
  for (int l = 0; l < EL; ++l) RC(vit_block_fwd(c, c->enc, l, l + 1 < EL ? c->enc.a[l + 1].x_in : c->enc.x_out));
  RC(mean_pool_tokens(c->enc.x_out, B, CLS_S, E, c->cls_pooled, st));
  RC(layernorm_fwd(c->cls_pooled, io.fc_norm_w, io.fc_norm_b, 1e-6f, B, E, c->cls_feat, nullptr, c->cls_mean, c->cls_rstd, st));
  GemmEpilogue eh;
  eh.bias = io.head_b; eh.out_f32 = io.logits; eh.ld_f32 = CLS_PAD;
  RC(gemm_bf16(c->cls_feat, E, 0, io.head_w16, E, 0, B, CLS_PAD, E, eh, 0, st));
  (void)M;
  return 0;
}

int t_cls_backward(CtxT<bf16>* c, const ClsIO& io, int accumulate, cudaStream_t st) {
  c->side = nullptr;  // one stream: the blocks run back to back without stage boundaries
  ECAMP_REQUIRE(c->bound && c->cls_planned, "cls backward: no forward has been run");
  ECAMP_REQUIRE(io.d_logits && io.g_pos_embed && io.g_fc_norm_w && io.g_fc_norm_b && io.g_head_w && io.g_head_b,
                "cls backward: null argument");
  const int B = c->cls_B, M = B * CLS_S, Mp = B * L196;
  c->st = st; c->acc = accumulate; c->dp = io.dp_scale;
  if (!accumulate) RC(zero_small_grads(c));
  // head: logits = feat W^T + b
  RC(cast_f32_to_bf16(io.d_logits, c->cls_dlogits, (size_t)B * CLS_PAD, st));
  GemmEpilogue ew;
  ew.out_f32 = io.g_head_w; ew.ld_f32 = E;
  if (accumulate) { ew.residual = io.g_head_w; ew.ld_res = E; }
  RC(gemm_bf16(c->cls_dlogits, CLS_PAD, 1, c->cls_feat, E, 1, CLS_PAD, E, B, ew, 0, st));
  RC(strided_rowsum(io.d_logits, B, CLS_PAD, CLS_PAD, io.g_head_b, accumulate, st));
  GemmEpilogue ed;
  ed.out_f32 = c->cls_dfeat; ed.ld_f32 = E;
  RC(gemm_bf16(c->cls_dlogits, CLS_PAD, 0, io.head_w16, E, 1, B, E, CLS_PAD, ed, 0, st));
  // fc_norm
  RC(layernorm_bwd(c->cls_dfeat, c->cls_pooled, c->cls_mean, c->cls_rstd, io.fc_norm_w, B, E, nullptr, c->cls_dpooled, (bf16*)nullptr,
                   DropoutCfg(), io.g_fc_norm_w, io.g_fc_norm_b, nullptr, accumulate, st));
  // mean over the 196 patch tokens; the bf16 copy is the dY of the last block's fc2 (its DropPath scale, its bias gradient)
  RC(mean_pool_tokens_bwd(c->cls_dpooled, B, CLS_S, E, c->dX, c->gX, c->dps(&c->enc, EL - 1, 1), st));
  RC(colsum_bf16(c->gX, E, M, E, c->Gp(c->enc.pbase + (EL - 1) * VIT_BLOCK_PARAMS + 11), 1, c->colsum_ws, st));
  for (int l = EL - 1; l >= 0; --l) RC(vit_block_bwd(c, c->enc, l));
  // x0 = [cls + pos[0]; patch_embed + pos[1:]]: pos_embed is learnable here
  RC(strided_rowsum(c->dX, B, (size_t)CLS_S * E, CLS_S * E, io.g_pos_embed, accumulate, st));
  bf16* d_pe = c->dAO;  // [Mp, 768]
  RC(assemble_encoder_input_bwd(c->dX, B, L196, E, d_pe, c->Gp(2), accumulate, st));
  GemmEpilogue e;
  e.out_f32 = c->dw_pe; e.ld_f32 = PDIM;
  RC(gemm_bf16(d_pe, E, 1, c->a_pe, PDIM, 1, E, PDIM, Mp, e, 0, st));
  RC(permute_pe_weight_grad(c->dw_pe, c->Gp(0), accumulate, st));
  RC(colsum_bf16(d_pe, E, Mp, E, c->Gp(1), 1, c->colsum_ws, st));
  return 0;
}

// =============================================================================================
// forward / backward entry points
// =============================================================================================
template <typename AT>
int t_forward(CtxT<AT>* c, const Batch& b, int flags, float drop_p, unsigned long long seed, float* losses3,
                float* mask_out, int64_t* ids_restore_out, int64_t* ids_keep_out, cudaStream_t st) {
  ECAMP_REQUIRE(c->bound && c->planned, "forward: context needs bind() and set_workspace() first");
  ECAMP_REQUIRE(b.image && b.ids && b.labels && b.attention_mask && b.type_ids && b.weights && b.noise && losses3,
                "forward: null batch tensor");
  if (c->sh.has_big) ECAMP_REQUIRE(b.column && b.row, "forward: column / row needed with 448-px input");
  c->batch = b; c->flags = flags; c->drop_p = drop_p; c->seed = seed; c->losses = losses3; c->st = st;
  RC(image_encoder_fwd(c, mask_out, ids_restore_out, ids_keep_out));
  RC(image_decoder_fwd(c));
  RC(image_losses_fwd(c));
  RC(text_front_fwd(c));
  const AT* h = c->fus.h_out;
  int hf_in = 1;
  for (int l = 0; l < BL; ++l) {
    RC(bert_layer_fwd(c, l, h, hf_in));
    h = c->layers[l].h_out;
  }
  RC(lm_transform_fwd(c));
  if (!(flags & 2)) RC(lm_chunks(c, false, true));
  return 0;
}

// Visualization/module/model_ecamp.py:308-319 + context_fusion.py:45-57: the heat-map tool returns the probabilities of the
// fusion layer's cross-attention.  They are re-derived from the q / k projections and the log-sum-exp that the forward
// pass left in the workspace (no dropout: the tool runs in eval mode).
template <typename AT>
int t_cross_attention_probs(CtxT<AT>* c, float* probs, cudaStream_t st) {
  ECAMP_REQUIRE(c->bound && c->planned && c->losses, "cross_attention_probs: run ecamp_forward first");
  AttnArgsT<AT> at;
  at.q = c->f_qc; at.ldq = 768; at.k = c->f_kv; at.v = c->f_kv + 768; at.ldk = at.ldv = 1536;
  at.lse = c->f_lse_c;
  at.B = c->sh.B; at.H = BH; at.Sq = c->sh.T; at.Sk = c->sh.keep; at.D = 128; at.scale = 1.0f / sqrtf(128.f);
  return attention_probs(at, probs, st);
}

// stages: 0 LM head | 1..6 BERT layers 5..0 | 7 fusion+embeddings+bert_mlp | 8 losses+decoder head+SR |
//         9..12 decoder blocks 3..0 | 13 decoder_embed+mask_token | 14 final norm | 15..26 encoder blocks 11..0 | 27 patch embed
int backward_stage_count() { return 28; }
int backward_stage_range(int stage, long long* g_begin, long long* g_end) {
  const auto& s = param_specs();
  auto off = [&](const std::string& n) { return s[param_index(n)].g_off; };
  auto blk = [&](const std::string& p, int l) { return off(p + std::to_string(l) + ".norm1.weight"); };
  const std::string b = BERT;
  long long lo, hi;
  if (stage == 0) { lo = off("bert_encoder.model.cls.predictions.transform.dense.weight"); hi = grad_total_floats(); }
  else if (stage <= 6) {
    const int l = 6 - stage;
    lo = off(b + "encoder.layer." + std::to_string(l) + ".attention.self.query.weight");
    hi = l == 5 ? off("bert_encoder.model.cls.predictions.transform.dense.weight")
                : off(b + "encoder.layer." + std::to_string(l + 1) + ".attention.self.query.weight");
  } else if (stage == 7) { lo = off("bert_mlp.weight"); hi = off(b + "encoder.layer.0.attention.self.query.weight"); }
  else if (stage == 8) { lo = off("decoder_norm.weight"); hi = off("bert_mlp.weight"); }
  else if (stage <= 12) {
    const int l = 12 - stage;
    lo = blk("decoder_blocks.", l);
    hi = l == 3 ? off("decoder_norm.weight") : blk("decoder_blocks.", l + 1);
  } else if (stage == 13) { lo = off("decoder_embed.weight"); hi = blk("decoder_blocks.", 0); }
  else if (stage == 14) { lo = off("norm.weight"); hi = off("decoder_embed.weight"); }
  else if (stage <= 26) {
    const int l = 26 - stage;
    lo = blk("blocks.", l);
    hi = l == 11 ? off("norm.weight") : blk("blocks.", l + 1);
  } else if (stage == 27) { lo = 0; hi = blk("blocks.", 0); }
  else { set_last_error("backward_stage_range: bad stage %d", stage); return -1; }
  *g_begin = lo; *g_end = hi;
  return 0;
}

namespace {
// every gradient tensor of at most kSmallGrad elements (all biases, LayerNorm weights, cls / mask tokens, the SR
// convolutions, token-type embeddings, the 30000-entry vocabulary bias): contiguous runs of the flat buffer are merged
constexpr long long kSmallGrad = 32768;
struct ZeroRuns {
  static constexpr int kMax = 160;
  long long off[kMax];
  int n[kMax];
  int count;
};
__global__ void zero_runs_kernel(float* __restrict__ g, const ZeroRuns runs) {
  ECAMP_PDL_ENTRY();
  const int r = blockIdx.x;
  float* p = g + runs.off[r];
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < runs.n[r]; i += gridDim.y * blockDim.x) p[i] = 0.f;
}
template <typename AT>
int zero_small_grads(CtxT<AT>* c) {
  static ZeroRuns runs = [] {
    ZeroRuns z;
    z.count = 0;
    const auto& s = param_specs();
    size_t i = 0;
    while (i < s.size()) {
      if (s[i].numel > kSmallGrad) { ++i; continue; }
      size_t j = i;
      long long n = 0;
      while (j < s.size() && s[j].numel <= kSmallGrad) { n += s[j].numel; ++j; }
      if (z.count < ZeroRuns::kMax) { z.off[z.count] = s[i].g_off; z.n[z.count] = (int)n; }
      ++z.count;
      i = j;
    }
    return z;
  }();
  ECAMP_REQUIRE(runs.count <= ZeroRuns::kMax, "zero_small_grads: %d runs exceed the table", runs.count);
  ECAMP_CUDA_OK(launch_pdl(zero_runs_kernel, dim3(runs.count, 4), 256, 0, c->st, c->G, runs));
  ECAMP_LAUNCHED();
  return 0;
}

template <typename AT>
int run_stage(CtxT<AT>* c, int stage) {
  const int B = c->sh.B, keep = c->sh.keep, Me = B * (keep + 1), Mi = B * keep, Md = B * 197, acc = c->acc;
  if (stage == 0) {
    RC(lm_chunks(c, true, (c->flags & 2) != 0));
    RC(lm_transform_bwd(c));
  } else if (stage <= 6) {
    RC(bert_layer_bwd(c, 6 - stage));
  } else if (stage == 7) {
    RC(text_front_bwd(c));
  } else if (stage == 8) {
    const int pn = param_index("decoder_norm.weight");
    const int ps = param_index("super_res.conv1.weight");
    if (c->sh.has_big)
      RC(sr_loss_bwd(c->pred, c->batch.image, c->batch.column, c->batch.row, c->P(ps), c->P(ps + 1), c->P(ps + 2),
                     c->P(ps + 3), B, c->g3 + 1, c->d_u, c->Gp(ps), acc, c->loss_ws, c->st,
                     (c->flags & 2) ? c->loss_ws + sr_ws_floats(B) : nullptr, (c->flags & 2) ? c->losses + 1 : nullptr));
    else if (!acc) ECAMP_CUDA_OK(cudaMemsetAsync(c->Gp(ps), 0, 168 * sizeof(float), c->st));
    RC(pred_grad(c->pred, c->tgt, c->maskf, c->sh.has_big ? c->d_u : nullptr, c->g3 + 0, B, c->gX, c->st));
    RC(lin_wgrad(c, c->gX, PDIM, c->dn, DD, Md, PDIM, DD, c->Gp(pn + 2), c->Gp(pn + 3), acc));
    GemmEpilogueT<AT> e;
    e.out_f32 = c->dH; e.ld_f32 = DD;
    RC(lin_dgrad(c, c->gX, PDIM, Md, c->W(pn + 2), PDIM, DD, e));
    RC(layernorm_bwd(c->dH, c->dec.x_out, c->mean_dn, c->rstd_dn, c->P(pn), Md, DD, nullptr, c->dX, c->gX,
                     DropoutCfg(), c->Gp(pn), c->Gp(pn + 1),
                     c->Gp(c->dec.pbase + (DL - 1) * VIT_BLOCK_PARAMS + 11) /* last decoder block's fc2 bias */, 1, c->st));
  } else if (stage <= 12) {
    RC(vit_block_bwd(c, c->dec, 12 - stage));
  } else if (stage == 13) {
    const int pde = param_index("decoder_embed.weight");
    AT* d_e = c->dAO;  // [Me, 512]
    ECAMP_CUDA_OK(cudaMemsetAsync(d_e, 0, (size_t)Me * DD * sizeof(AT), c->st));
    RC(assemble_decoder_input_bwd(c->dX, c->ids_restore, B, L196, keep, DD, d_e, c->Gp(pde + 2), acc, c->misc_ws, c->st));
    RC(lin_wgrad(c, d_e, DD, c->latent, E, Me, DD, E, c->Gp(pde), c->Gp(pde + 1), acc));
    GemmEpilogueT<AT> e;
    e.residual = c->dLat; e.ld_res = E; e.out_f32 = c->dLat; e.ld_f32 = E;  // accumulate onto the text-branch gradient
    RC(lin_dgrad(c, d_e, DD, Me, c->W(pde), DD, E, e));
  } else if (stage == 14) {
    const int pn = param_index("norm.weight");
    RC(layernorm_bwd(c->dLat, c->enc.x_out, c->mean_n, c->rstd_n, c->P(pn), Me, E, nullptr, c->dX, c->gX, DropoutCfg(),
                     c->Gp(pn), c->Gp(pn + 1),
                     c->Gp(c->enc.pbase + (EL - 1) * VIT_BLOCK_PARAMS + 11) /* last encoder block's fc2 bias */, 1, c->st));
  } else if (stage <= 26) {
    RC(vit_block_bwd(c, c->enc, 26 - stage));
  } else {
    AT* d_pe = c->dAO;  // [Mi, 768]
    RC(assemble_encoder_input_bwd(c->dX, B, keep, E, d_pe, c->Gp(2), acc, c->st));
    GemmEpilogueT<AT> e;
    e.out_f32 = c->dw_pe; e.ld_f32 = PDIM;
    RC(gemm_any(c, d_pe, E, 1, c->a_pe, PDIM, 1, E, PDIM, Mi, e));
    RC(permute_pe_weight_grad(c->dw_pe, c->Gp(0), acc, c->st));
    RC(colsum_bf16(d_pe, E, Mi, E, c->Gp(1), 1, c->colsum_ws, c->st));
  }
  return 0;
}
}  // namespace

template <typename AT>
int t_backward(CtxT<AT>* c, const float* g3, int accumulate, int stage, int stage_end, cudaStream_t st) {
  ECAMP_REQUIRE(c->bound && c->planned && c->losses, "backward: no forward has been run");
  ECAMP_REQUIRE(g3 != nullptr, "backward: null upstream gradient");
  c->g3 = g3; c->acc = accumulate; c->st = st;
  if (g_side_wgrad < 0) g_side_wgrad = getenv("ECAMP_SIDE_WGRAD") ? atoi(getenv("ECAMP_SIDE_WGRAD")) : 1;
  // (the fp32-accurate mode shares one operand-plane scratch between its GEMMs: one stream only)
  c->side = (g_side_wgrad && !is_hp<AT>::value) ? side_stream_for_device() : nullptr;
  c->side_busy = false;
  // bias / LayerNorm / token gradients are accumulated with atomics by the kernels that produce their operands:
  // a non-accumulating backward zeroes them once, before the first stage
  if (!accumulate && stage <= 0) RC(zero_small_grads(c));
  // [first, last): stage = -1 -> all stages; stage_end <= stage -> the single stage `stage`
  const int first = stage >= 0 ? stage : 0;
  const int last = stage < 0 ? backward_stage_count() : (stage_end > stage ? stage_end : stage + 1);
  ECAMP_REQUIRE(last <= backward_stage_count(), "backward: stage range [%d, %d) out of bounds", first, last);
  for (int s = first; s < last; ++s) {
    const int rc = run_stage(c, s);
    // Every stage ends joined (its gradient slice is final when the call returns / the stage callback fires) - except, when
    // several stages run in one call, between two transformer blocks of the same stack: those only share gX / dA / dQKV with
    // the next block, each protected by its guard, so the last weight gradients of a block may overlap the next block.
    const bool same_stack_next = s + 1 < last && ((s >= 1 && s <= 5) || (s >= 9 && s <= 11) || (s >= 15 && s <= 25));
    const int rj = (rc || !same_stack_next) ? side_join(c) : 0;  // also after a failed stage: nothing stays queued
    if (rc) return rc;
    if (rj) return rj;
  }
  return 0;
}


// =============================================================================================
// public context: one production (bf16 activations) and one fp32-accurate instance of the same schedule
// =============================================================================================
struct Ctx {
  int hp = 0;  // 0 = production (bf16 operands), 1 = fp32-accurate parity mode
  CtxT<bf16> lp;
  CtxT<float> hp32;
};
#define ECAMP_DISPATCH(call_lp, call_hp) (c->hp ? (call_hp) : (call_lp))

void set_side_stream(int on) { g_side_wgrad = on == 2 ? 2 : (on ? 1 : 0); }
Ctx* ctx_new() { return new Ctx(); }
void ctx_free(Ctx* c) { delete c; }
int ctx_set_precision(Ctx* c, int hp) {
  ECAMP_REQUIRE(hp == 0 || hp == 1, "set_precision: 0 = bf16 (production), 1 = fp32-accurate");
  if (c->hp != hp) {  // buffers are laid out per precision: bind() and set_workspace() must follow
    c->lp = CtxT<bf16>();
    c->hp32 = CtxT<float>();
    c->hp = hp;
  }
  return 0;
}
int ctx_precision(Ctx* c) { return c->hp; }
int ctx_bind(Ctx* c, float* const* params, int n, float* G, float* M1, float* M2, void* shadows, const float* pos_embed,
             const float* dec_pos_embed, void* adam_table, void* adam_chunks) {
  return ECAMP_DISPATCH(t_bind(&c->lp, params, n, G, M1, M2, shadows, pos_embed, dec_pos_embed, adam_table, adam_chunks),
                        t_bind(&c->hp32, params, n, G, M1, M2, shadows, pos_embed, dec_pos_embed, adam_table, adam_chunks));
}
size_t workspace_bytes(const Shape& sh, int hp) { return hp ? t_workspace_bytes<float>(sh) : t_workspace_bytes<bf16>(sh); }
int ctx_set_workspace(Ctx* c, void* ws, size_t bytes, const Shape& sh) {
  return ECAMP_DISPATCH(t_set_workspace(&c->lp, ws, bytes, sh), t_set_workspace(&c->hp32, ws, bytes, sh));
}
int ctx_refresh_shadows(Ctx* c, cudaStream_t st) {
  return ECAMP_DISPATCH(t_refresh_shadows(&c->lp, st), t_refresh_shadows(&c->hp32, st));
}
int ctx_forward(Ctx* c, const Batch& b, int flags, float drop_p, unsigned long long seed, float* losses3, float* mask_out,
                int64_t* ids_restore_out, int64_t* ids_keep_out, cudaStream_t st) {
  return ECAMP_DISPATCH(t_forward(&c->lp, b, flags, drop_p, seed, losses3, mask_out, ids_restore_out, ids_keep_out, st),
                        t_forward(&c->hp32, b, flags, drop_p, seed, losses3, mask_out, ids_restore_out, ids_keep_out, st));
}
int ctx_cross_attention_probs(Ctx* c, float* probs, cudaStream_t st) {
  ECAMP_REQUIRE(!c->hp, "cross_attention_probs: available in the production precision only");
  return t_cross_attention_probs(&c->lp, probs, st);
}
int ctx_backward(Ctx* c, const float* g3, int accumulate, int stage, int stage_end, cudaStream_t st) {
  return ECAMP_DISPATCH(t_backward(&c->lp, g3, accumulate, stage, stage_end, st),
                        t_backward(&c->hp32, g3, accumulate, stage, stage_end, st));
}
int ctx_adamw(Ctx* c, float lr, float lr_nodecay, float b1, float b2, float eps, float wd, int step, float grad_scale,
              cudaStream_t st) {
  return ECAMP_DISPATCH(t_adamw(&c->lp, lr, lr_nodecay, b1, b2, eps, wd, step, grad_scale, st),
                        t_adamw(&c->hp32, lr, lr_nodecay, b1, b2, eps, wd, step, grad_scale, st));
}
int ctx_adamw_range(Ctx* c, float lr, float lr_nodecay, float b1, float b2, float eps, float wd, int step, float grad_scale,
                    long long g_lo, long long g_hi, cudaStream_t st) {
  return ECAMP_DISPATCH(t_adamw_range(&c->lp, lr, lr_nodecay, b1, b2, eps, wd, step, grad_scale, g_lo, g_hi, st),
                        t_adamw_range(&c->hp32, lr, lr_nodecay, b1, b2, eps, wd, step, grad_scale, g_lo, g_hi, st));
}
const void* ctx_debug_ptr(Ctx* c, const char* name) {
  return ECAMP_DISPATCH(t_debug_ptr(&c->lp, name), t_debug_ptr(&c->hp32, name));
}
int ctx_set_cls_workspace(Ctx* c, void* ws, size_t bytes, int B) {
  ECAMP_REQUIRE(!c->hp, "fine-tune classification: available in the production precision only");
  return t_set_cls_workspace(&c->lp, ws, bytes, B);
}
int ctx_cls_forward(Ctx* c, const ClsIO& io, cudaStream_t st) {
  ECAMP_REQUIRE(!c->hp, "fine-tune classification: available in the production precision only");
  return t_cls_forward(&c->lp, io, st);
}
int ctx_cls_backward(Ctx* c, const ClsIO& io, int accumulate, cudaStream_t st) {
  ECAMP_REQUIRE(!c->hp, "fine-tune classification: available in the production precision only");
  return t_cls_backward(&c->lp, io, accumulate, st);
}

}  // namespace ecamp
