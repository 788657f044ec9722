// HBM-bound glue kernels of the ECAMP hot path: random masking (stable rank sort), bicubic resize into
// patch layout, kept-patch gather, encoder / decoder sequence assembly and their backward, BERT
// embeddings forward / backward, bias-gradient column sums.  Every kernel cites the reference lines
// it replaces.
#include "kernels.cuh"

namespace ecamp {
namespace {

// ---------------------------------------------------------------------------------------------
// model_ecamp.py:168-193.  rank_i = #{j : n_j < n_i or (n_j == n_i and j < i)} is the position of i in the
// stable ascending argsort, i.e. ids_restore[i]; ids_shuffle[rank_i] = i.  Bit-exact by construction.
// ---------------------------------------------------------------------------------------------
__global__ void random_masking_kernel(const float* __restrict__ noise, int L, int len_keep,
                                      int32_t* __restrict__ ids_restore, int32_t* __restrict__ ids_keep,
                                      float* __restrict__ mask, int64_t* __restrict__ ids_restore64,
                                      int64_t* __restrict__ ids_keep64) {
  ECAMP_PDL_ENTRY();
  extern __shared__ float s_noise[];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < L; i += blockDim.x) s_noise[i] = noise[(size_t)b * L + i];
  __syncthreads();
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float ni = s_noise[i];
    int rank = 0;
    for (int j = 0; j < L; ++j) {
      const float nj = s_noise[j];
      rank += (nj < ni || (nj == ni && j < i)) ? 1 : 0;
    }
    ids_restore[(size_t)b * L + i] = rank;
    if (ids_restore64) ids_restore64[(size_t)b * L + i] = rank;
    mask[(size_t)b * L + i] = rank >= len_keep ? 1.f : 0.f;
    if (rank < len_keep) {
      ids_keep[(size_t)b * len_keep + rank] = i;
      if (ids_keep64) ids_keep64[(size_t)b * len_keep + rank] = i;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// model_ecamp.py:318 — torchvision Resize([224,224], BICUBIC) on a tensor = upsample_bicubic2d,
// align_corners=False, antialias off.  Scale 2 => source centre 2*d + 0.5 => taps 2d-1..2d+2 with the
// fixed weights cubic(A=-0.75, t=0.5) = [-3/32, 19/32, 19/32, -3/32], indices clamped at the border.
// Output goes straight to patch layout tgt[b, hy*14+wx, (p*16+q)*3+c] (the layout of decoder_pred).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(224) resize_patchify_kernel(const float* __restrict__ big, int Hin,
                                                              float* __restrict__ tgt) {
  ECAMP_PDL_ENTRY();
  const int y = blockIdx.x % 224, b = blockIdx.x / 224;
  const int x = threadIdx.x;
  const float w[4] = {-0.09375f, 0.59375f, 0.59375f, -0.09375f};
  int ys[4], xs[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ys[i] = min(max(2 * y - 1 + i, 0), Hin - 1);
    xs[i] = min(max(2 * x - 1 + i, 0), Hin - 1);
  }
  const int hy = y >> 4, p = y & 15, wx = x >> 4, q = x & 15;
  float* dst = tgt + ((size_t)b * 196 + hy * 14 + wx) * 768 + (p * 16 + q) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* src = big + ((size_t)b * 3 + c) * Hin * Hin;
    float rows[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float* r = src + (size_t)ys[i] * Hin;
      rows[i] = r[xs[0]] * w[0] + r[xs[1]] * w[1] + r[xs[2]] * w[2] + r[xs[3]] * w[3];
    }
    dst[c] = rows[0] * w[0] + rows[1] * w[1] + rows[2] * w[2] + rows[3] * w[3];
  }
}

__global__ void __launch_bounds__(224) patchify224_kernel(const float* __restrict__ imgs, float* __restrict__ tgt) {
  ECAMP_PDL_ENTRY();
  const int y = blockIdx.x % 224, b = blockIdx.x / 224;
  const int x = threadIdx.x;
  const int hy = y >> 4, p = y & 15, wx = x >> 4, q = x & 15;
  float* dst = tgt + ((size_t)b * 196 + hy * 14 + wx) * 768 + (p * 16 + q) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) dst[c] = imgs[(((size_t)b * 3 + c) * 224 + y) * 224 + x];
}

// one block per output row, D/4 threads
template <typename AT>
__global__ void gather_patches_kernel(const float* __restrict__ tgt, const int32_t* __restrict__ ids_keep, int L,
                                      int keep, int PD, AT* __restrict__ out) {
  ECAMP_PDL_ENTRY();
  const int r = blockIdx.x, b = r / keep;
  const int src_l = ids_keep[r];
  const float4 v = reinterpret_cast<const float4*>(tgt + ((size_t)b * L + src_l) * PD)[threadIdx.x];
  st4(out + (size_t)r * PD + 4 * threadIdx.x, v);
}

__global__ void assemble_enc_kernel(const float* __restrict__ pe, const float* __restrict__ cls,
                                    const float* __restrict__ pos, const int32_t* __restrict__ ids_keep, int keep,
                                    int D, float* __restrict__ x0) {
  ECAMP_PDL_ENTRY();
  const int r = blockIdx.x, b = r / (keep + 1), s = r % (keep + 1);
  const int c4 = threadIdx.x;
  float4 a, p;
  if (s == 0) {
    a = reinterpret_cast<const float4*>(cls)[c4];
    p = reinterpret_cast<const float4*>(pos)[c4];
  } else {
    a = reinterpret_cast<const float4*>(pe + ((size_t)b * keep + s - 1) * D)[c4];
    p = reinterpret_cast<const float4*>(pos + (size_t)(1 + ids_keep[(size_t)b * keep + s - 1]) * D)[c4];
  }
  reinterpret_cast<float4*>(x0 + (size_t)r * D)[c4] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
}

template <typename AT>
__global__ void assemble_enc_bwd_kernel(const float* __restrict__ dx0, int keep, int D, AT* __restrict__ d_pe) {
  ECAMP_PDL_ENTRY();
  const int r = blockIdx.x, b = r / keep, j = r % keep;
  const float4 v = reinterpret_cast<const float4*>(dx0 + ((size_t)b * (keep + 1) + 1 + j) * D)[threadIdx.x];
  st4(d_pe + (size_t)r * D + 4 * threadIdx.x, v);
}

// out[d] (+)= sum_b x[b * stride + d]
__global__ void strided_rowsum_kernel(const float* __restrict__ x, int B, size_t stride, int D,
                                      float* __restrict__ out, int accumulate) {
  ECAMP_PDL_ENTRY();
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += x[(size_t)b * stride + d];
  out[d] = accumulate ? out[d] + s : s;
}

template <typename AT>
__global__ void assemble_dec_kernel(const AT* __restrict__ e, const float* __restrict__ mask_token,
                                    const float* __restrict__ dpos, const int32_t* __restrict__ ids_restore, int L,
                                    int keep, int D, float* __restrict__ xd) {
  ECAMP_PDL_ENTRY();
  const int r = blockIdx.x, b = r / (L + 1), t = r % (L + 1);
  const int c4 = threadIdx.x;
  float4 a;
  int src = -1;
  if (t == 0) {
    src = 0;
  } else {
    const int rr = ids_restore[(size_t)b * L + t - 1];
    if (rr < keep) src = 1 + rr;
  }
  if (src >= 0) {
    a = ld4(e + ((size_t)b * (keep + 1) + src) * D + 4 * c4);
  } else {
    // torch.cat of the half-precision decoder_embed output with the fp32 mask token promotes to fp32
    // (model_ecamp.py:245-246): the mask token is NOT rounded
    a = reinterpret_cast<const float4*>(mask_token)[c4];
  }
  const float4 p = reinterpret_cast<const float4*>(dpos + (size_t)t * D)[c4];
  reinterpret_cast<float4*>(xd + (size_t)r * D)[c4] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
}

template <typename AT>
__global__ void assemble_dec_bwd_kernel(const float* __restrict__ dxd, const int32_t* __restrict__ ids_restore, int L,
                                        int keep, int D, AT* __restrict__ d_e) {
  ECAMP_PDL_ENTRY();
  const int r = blockIdx.x, b = r / (L + 1), t = r % (L + 1);
  int dst = -1;
  if (t == 0) {
    dst = 0;
  } else {
    const int rr = ids_restore[(size_t)b * L + t - 1];
    if (rr < keep) dst = 1 + rr;
  }
  if (dst < 0) return;
  const float4 v = reinterpret_cast<const float4*>(dxd + (size_t)r * D)[threadIdx.x];
  st4(d_e + ((size_t)b * (keep + 1) + dst) * D + 4 * threadIdx.x, v);
}

// ws[b, d] = sum over masked positions l of dxd[b, 1 + l, d]
__global__ void mask_token_grad_kernel(const float* __restrict__ dxd, const int32_t* __restrict__ ids_restore, int L,
                                       int keep, int D, float* __restrict__ ws) {
  ECAMP_PDL_ENTRY();
  const int b = blockIdx.x;
  const int c4 = threadIdx.x;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int l = 0; l < L; ++l) {
    if (ids_restore[(size_t)b * L + l] >= keep) {
      const float4 v = reinterpret_cast<const float4*>(dxd + ((size_t)b * (L + 1) + 1 + l) * D)[c4];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  reinterpret_cast<float4*>(ws + (size_t)b * D)[c4] = acc;
}

template <typename AT>
__global__ void split_latent_gap_kernel(const AT* __restrict__ lat2, int keep, int D, AT* __restrict__ img_tok,
                                        AT* __restrict__ gap) {
  ECAMP_PDL_ENTRY();
  const int b = blockIdx.x, c4 = threadIdx.x;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = 0; j < keep; ++j) {
    const float4 v = ld4(lat2 + ((size_t)b * (keep + 1) + 1 + j) * D + 4 * c4);
    st4(img_tok + ((size_t)b * keep + j) * D + 4 * c4, v);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const float inv = 1.0f / keep;
  st4(gap + (size_t)b * D + 4 * c4, make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv));
}

template <typename AT>
__global__ void split_latent_gap_bwd_kernel(const AT* __restrict__ d_img_tok, const AT* __restrict__ d_gap,
                                            int keep, int D, AT* __restrict__ d_lat2) {
  ECAMP_PDL_ENTRY();
  const int r = blockIdx.x, b = r / (keep + 1), s = r % (keep + 1);
  const int c4 = threadIdx.x;
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s > 0) {
    const float4 u = ld4(d_img_tok + ((size_t)b * keep + s - 1) * D + 4 * c4);
    const float4 g = ld4(d_gap + (size_t)b * D + 4 * c4);
    const float inv = 1.0f / keep;
    o = make_float4(u.x + g.x * inv, u.y + g.y * inv, u.z + g.z * inv, u.w + g.w * inv);
  }
  st4(d_lat2 + (size_t)r * D + 4 * c4, o);
}

template <typename AT>
__global__ void add_batch_rowvec_kernel(AT* __restrict__ y, const AT* __restrict__ vec, int T, int D) {
  ECAMP_PDL_ENTRY();
  const int r = blockIdx.x, b = r / T, c4 = threadIdx.x;
  const float4 u = ld4(y + (size_t)r * D + 4 * c4), g = ld4(vec + (size_t)b * D + 4 * c4);
  st4(y + (size_t)r * D + 4 * c4, make_float4(u.x + g.x, u.y + g.y, u.z + g.z, u.w + g.w));
}

template <typename AT>
__global__ void add_batch_rowvec_oop_kernel(const AT* __restrict__ x, const AT* __restrict__ vec, int T, int D,
                                            AT* __restrict__ y) {
  ECAMP_PDL_ENTRY();
  const int r = blockIdx.x, b = r / T, c4 = threadIdx.x;
  const float4 u = ld4(x + (size_t)r * D + 4 * c4), g = ld4(vec + (size_t)b * D + 4 * c4);
  st4(y + (size_t)r * D + 4 * c4, make_float4(u.x + g.x, u.y + g.y, u.z + g.z, u.w + g.w));
}

template <typename AT>
__global__ void gelu_bwd_bf16_kernel(const float* __restrict__ d, const AT* __restrict__ pre, AT* __restrict__ out,
                                     size_t n) {
  ECAMP_PDL_ENTRY();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) act_st(out + i, d[i] * (is_hp<AT>::value ? gelu_exact_grad(act_ld(pre + i)) : gelu_erf_grad(act_ld(pre + i))));
}

template <typename AT>
__global__ void batch_colsum_kernel(const AT* __restrict__ x, int T, int D, AT* __restrict__ out) {
  ECAMP_PDL_ENTRY();
  const int b = blockIdx.x, c4 = threadIdx.x;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < T; ++t) {
    const float4 v = ld4(x + ((size_t)b * T + t) * D + 4 * c4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  st4(out + (size_t)b * D + 4 * c4, acc);
}

// ---------------------------------------------------------------------------------------------
// HF BertEmbeddings.forward (called at bert_modeling.py:113-119): (word[id] + type[tt]) + pos[t] -> LN -> dropout
// ---------------------------------------------------------------------------------------------
template <typename AT>
__global__ void __launch_bounds__(256) bert_emb_fwd_kernel(
    const int64_t* __restrict__ ids, const int64_t* __restrict__ type_ids, const float* __restrict__ word,
    const float* __restrict__ type, const float* __restrict__ pos, const float* __restrict__ gamma,
    const float* __restrict__ beta, float eps, int M, int T, DropoutCfg drop, float* __restrict__ pre,
    float* __restrict__ mean_out, float* __restrict__ rstd_out, AT* __restrict__ out_bf16,
    float* __restrict__ out_f32) {
  ECAMP_PDL_ENTRY();
  constexpr int NV = 6, D = 768;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * 8 + warp;
  if (row >= M) return;
  const int t = row % T;
  const long long id = ids[row], tt = type_ids[row];
  const float4* wr = reinterpret_cast<const float4*>(word + (size_t)id * D);
  const float4* tr = reinterpret_cast<const float4*>(type + (size_t)tt * D);
  const float4* pr = reinterpret_cast<const float4*>(pos + (size_t)t * D);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = i * 32 + lane;
    const float4 a = __ldg(wr + c4), b = __ldg(tr + c4), c = __ldg(pr + c4);
    v[i] = make_float4((a.x + b.x) + c.x, (a.y + b.y) + c.y, (a.z + b.z) + c.z, (a.w + b.w) + c.w);
    reinterpret_cast<float4*>(pre + (size_t)row * D)[c4] = v[i];
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += a * a + b * b + c * c + d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
  const Philox ph(drop.seed);
  const uint32_t thr = dropout_threshold(drop.p);
  const float ks = drop.p > 0.f ? 1.0f / (1.0f - drop.p) : 1.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = i * 32 + lane;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + b.x;
    y.y = (v[i].y - mean) * rstd * g.y + b.y;
    y.z = (v[i].z - mean) * rstd * g.z + b.z;
    y.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (drop.p > 0.f) {
      const uint4 rnd = ph(((uint64_t)row * D + (uint64_t)c4 * 4) >> 2, drop.site);
      y.x = rnd.x >= thr ? y.x * ks : 0.f;
      y.y = rnd.y >= thr ? y.y * ks : 0.f;
      y.z = rnd.z >= thr ? y.z * ks : 0.f;
      y.w = rnd.w >= thr ? y.w * ks : 0.f;
    }
    if (out_f32) reinterpret_cast<float4*>(out_f32 + (size_t)row * D)[c4] = y;
    st4(out_bf16 + (size_t)row * D + 4 * c4, y);
  }
}

// in-place dropout backward on an fp32 gradient (embedding-output dropout site)
__global__ void dropout_bwd_f32_kernel(float* __restrict__ g, size_t n4, DropoutCfg drop) {
  ECAMP_PDL_ENTRY();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const Philox ph(drop.seed);
  const uint32_t thr = dropout_threshold(drop.p);
  const float ks = 1.0f / (1.0f - drop.p);
  const uint4 rnd = ph(i, drop.site);
  float4 v = reinterpret_cast<float4*>(g)[i];
  v.x = rnd.x >= thr ? v.x * ks : 0.f;
  v.y = rnd.y >= thr ? v.y * ks : 0.f;
  v.z = rnd.z >= thr ? v.z * ks : 0.f;
  v.w = rnd.w >= thr ? v.w * ks : 0.f;
  reinterpret_cast<float4*>(g)[i] = v;
}

// word-embedding gradient: scatter-add rows (padding_idx = 0 receives nothing).  The special tokens [CLS] = 2,
// [MASK] = 3, [SEP] = 4 (ECAMP/Pre-training/module/pretrain_datasets.py:23-24) appear in every report - [MASK] on ~45 % of the
// positions - and their rows serialise in the L2 atomic unit, so each CTA first folds its kEmbRows rows of those ids in
// registers and adds them once; all other rows go out as 16-byte vector reductions.
constexpr int kEmbRows = 64;
__global__ void __launch_bounds__(192) emb_word_bwd_kernel(const float* __restrict__ d_pre, const int64_t* __restrict__ ids,
                                                           int M, int D, float* __restrict__ d_word) {
  ECAMP_PDL_ENTRY();
  const int r0 = blockIdx.x * kEmbRows, r1 = min(M, r0 + kEmbRows);
  float4 hot[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) hot[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int row = r0; row < r1; ++row) {
    const long long id = ids[row];
    if (id == 0) continue;
    const float4 v = reinterpret_cast<const float4*>(d_pre + (size_t)row * D)[threadIdx.x];
    if (id >= 2 && id <= 4) {
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (id == 2 + k) { hot[k].x += v.x; hot[k].y += v.y; hot[k].z += v.z; hot[k].w += v.w; }
    } else {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d_word + (size_t)id * D + threadIdx.x * 4), "f"(v.x),
                   "f"(v.y), "f"(v.z), "f"(v.w)
                   : "memory");
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (hot[k].x != 0.f || hot[k].y != 0.f || hot[k].z != 0.f || hot[k].w != 0.f)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d_word + (size_t)(2 + k) * D + threadIdx.x * 4),
                   "f"(hot[k].x), "f"(hot[k].y), "f"(hot[k].z), "f"(hot[k].w)
                   : "memory");
}

// one block per position t: position gradient (sum over batch) and per-type partial sums
__global__ void emb_pos_type_bwd_kernel(const float* __restrict__ d_pre, const int64_t* __restrict__ type_ids, int B,
                                        int T, int D, float* __restrict__ d_pos, int accumulate,
                                        float* __restrict__ type_ws) {
  ECAMP_PDL_ENTRY();
  const int t = blockIdx.x, c4 = threadIdx.x;
  float4 ap = make_float4(0.f, 0.f, 0.f, 0.f), a1 = ap;
  for (int b = 0; b < B; ++b) {
    const size_t row = (size_t)b * T + t;
    const float4 v = reinterpret_cast<const float4*>(d_pre + row * D)[c4];
    ap.x += v.x; ap.y += v.y; ap.z += v.z; ap.w += v.w;
    if (type_ids[row] != 0) { a1.x += v.x; a1.y += v.y; a1.z += v.z; a1.w += v.w; }
  }
  float4* dp = reinterpret_cast<float4*>(d_pos + (size_t)t * D) + c4;
  if (accumulate) {
    const float4 o = *dp;
    *dp = make_float4(o.x + ap.x, o.y + ap.y, o.z + ap.z, o.w + ap.w);
  } else {
    *dp = ap;
  }
  // type 0 = total - type 1
  reinterpret_cast<float4*>(type_ws + ((size_t)t * 2 + 0) * D)[c4] =
      make_float4(ap.x - a1.x, ap.y - a1.y, ap.z - a1.z, ap.w - a1.w);
  reinterpret_cast<float4*>(type_ws + ((size_t)t * 2 + 1) * D)[c4] = a1;
}

__global__ void emb_type_finalize_kernel(const float* __restrict__ type_ws, int T, int D, float* __restrict__ d_type,
                                         int accumulate) {
  ECAMP_PDL_ENTRY();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over 2*D
  if (i >= 2 * D) return;
  float s = 0.f;
  for (int t = 0; t < T; ++t) s += type_ws[(size_t)t * 2 * D + i];
  d_type[i] = accumulate ? d_type[i] + s : s;
}

// ---------------------------------------------------------------------------------------------
// bias gradients
// ---------------------------------------------------------------------------------------------
// out[n] = sum_m x[m, n]: CTA = 256 columns x one row chunk; each lane owns 8 consecutive columns (128-bit loads),
// the 8 warps stride over the rows of the chunk; per-chunk partials are reduced by a second tiny kernel.
ECAMP_DEVINL void ld8(const bf16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
ECAMP_DEVINL void ld8(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <typename AT>
__global__ void __launch_bounds__(256) colsum_kernel(const AT* __restrict__ x, int ld, int M, int N,
                                                     float* __restrict__ out) {
  ECAMP_PDL_ENTRY();
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (col < N) {  // N % 8 == 0 checked on the host
    const AT* base = x + col;
    int r = r0 + warp;
    for (; r + 24 < r1; r += 32) {  // four independent (128-bit) loads in flight per lane
      float f[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i) ld8(base + (size_t)(r + 8 * i) * ld, f[i]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += f[i][k];
    }
    for (; r < r1; r += 8) {
      float f[8];
      ld8(base + (size_t)r * ld, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[warp][lane * 8 + k] = acc[k];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    atomicAdd(out + c, s);  // one add per row chunk and column (the destination starts from zero / the running gradient)
  }
}
__global__ void iota_mod_kernel(int32_t* __restrict__ out, size_t n, int mod) {
  ECAMP_PDL_ENTRY();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)(i % (size_t)mod);
}
// x[b, 1:, :].mean(dim=1)   (models_vit.py:92); one CTA per sample, thread = 4 columns
__global__ void mean_pool_kernel(const float* __restrict__ x, int S, int D, float* __restrict__ out) {
  ECAMP_PDL_ENTRY();
  const int b = blockIdx.x, c4 = threadIdx.x;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 1; t < S; ++t) {
    const float4 v = reinterpret_cast<const float4*>(x + ((size_t)b * S + t) * D)[c4];
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  const float inv = 1.0f / (float)(S - 1);
  reinterpret_cast<float4*>(out + (size_t)b * D)[c4] = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
}
__global__ void mean_pool_bwd_kernel(const float* __restrict__ d_pooled, int S, int D, float* __restrict__ dx,
                                     bf16* __restrict__ gx, const float* __restrict__ scale) {
  ECAMP_PDL_ENTRY();
  const int r = blockIdx.x, b = r / S, t = r % S, c4 = threadIdx.x;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t > 0) {
    const float inv = 1.0f / (float)(S - 1);
    const float4 d = reinterpret_cast<const float4*>(d_pooled + (size_t)b * D)[c4];
    v = make_float4(d.x * inv, d.y * inv, d.z * inv, d.w * inv);
  }
  reinterpret_cast<float4*>(dx + (size_t)r * D)[c4] = v;
  const float s = scale ? scale[b] : 1.0f;
  uint2 u;
  u.x = pack_bf16x2(v.x * s, v.y * s);
  u.y = pack_bf16x2(v.z * s, v.w * s);
  reinterpret_cast<uint2*>(gx + (size_t)r * D)[c4] = u;
}
__global__ void scale_f32_kernel(float* __restrict__ x, const float* __restrict__ scale, size_t n) {
  ECAMP_PDL_ENTRY();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= *scale;
}
__global__ void cast_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, size_t n) {
  ECAMP_PDL_ENTRY();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = f2bf(x[i]);
}
__global__ void permute_pe_grad_kernel(const float* __restrict__ dw_pqc, float* __restrict__ grad_cpq, int accumulate) {
  ECAMP_PDL_ENTRY();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // canonical index n*768 + c*256 + pq
  if (i >= 768 * 768) return;
  const int n = i / 768, k = i % 768, c = k / 256, pq = k % 256;
  const float v = dw_pqc[(size_t)n * 768 + pq * 3 + c];
  grad_cpq[i] = accumulate ? grad_cpq[i] + v : v;
}

}  // namespace

#define LAUNCH_OK() ECAMP_LAUNCHED()

int random_masking(const float* noise, int B, int L, int len_keep, int32_t* ids_restore, int32_t* ids_keep,
                   float* mask, int64_t* ids_restore64, int64_t* ids_keep64, cudaStream_t st) {
  ECAMP_REQUIRE(L > 0 && L <= 4096 && len_keep >= 0 && len_keep <= L, "random_masking: bad L %d / len_keep %d", L,
                len_keep);
  if (B <= 0) return 0;
  ECAMP_CUDA_OK(launch_pdl(random_masking_kernel, B, 256, L * sizeof(float), st, noise, L, len_keep, ids_restore, ids_keep, mask,
                                                           ids_restore64, ids_keep64));
  LAUNCH_OK();
  return 0;
}

int resize_bicubic_patchify(const float* big, int B, int Hin, float* tgt, cudaStream_t st) {
  ECAMP_REQUIRE(Hin == 448, "resize: only 448 -> 224 is on the reference path (got %d)", Hin);
  if (B <= 0) return 0;
  ECAMP_CUDA_OK(launch_pdl(resize_patchify_kernel, B * 224, 224, 0, st, big, Hin, tgt));
  LAUNCH_OK();
  return 0;
}
int patchify224(const float* imgs, int B, float* tgt, cudaStream_t st) {
  if (B <= 0) return 0;
  ECAMP_CUDA_OK(launch_pdl(patchify224_kernel, B * 224, 224, 0, st, imgs, tgt));
  LAUNCH_OK();
  return 0;
}
template <typename AT>
int gather_patches(const float* tgt, const int32_t* ids_keep, int B, int L, int keep, int PD, AT* out,
                   cudaStream_t st) {
  if (B * keep <= 0) return 0;
  ECAMP_CUDA_OK(launch_pdl(gather_patches_kernel<AT>, B * keep, PD / 4, 0, st, tgt, ids_keep, L, keep, PD, out));
  LAUNCH_OK();
  return 0;
}
int assemble_encoder_input(const float* pe, const float* cls, const float* pos, const int32_t* ids_keep, int B,
                           int keep, int D, float* x0, cudaStream_t st) {
  ECAMP_CUDA_OK(launch_pdl(assemble_enc_kernel, B * (keep + 1), D / 4, 0, st, pe, cls, pos, ids_keep, keep, D, x0));
  LAUNCH_OK();
  return 0;
}
template <typename AT>
int assemble_encoder_input_bwd(const float* dx0, int B, int keep, int D, AT* d_pe, float* d_cls, int accumulate,
                               cudaStream_t st) {
  if (keep > 0) {
    ECAMP_CUDA_OK(launch_pdl(assemble_enc_bwd_kernel<AT>, B * keep, D / 4, 0, st, dx0, keep, D, d_pe));
    LAUNCH_OK();
  }
  ECAMP_CUDA_OK(launch_pdl(strided_rowsum_kernel, (D + 127) / 128, 128, 0, st, dx0, B, (size_t)(keep + 1) * D, D, d_cls, accumulate));
  LAUNCH_OK();
  return 0;
}
template <typename AT>
int assemble_decoder_input(const AT* e, const float* mask_token, const float* dpos, const int32_t* ids_restore,
                           int B, int L, int keep, int D, float* xd, cudaStream_t st) {
  ECAMP_CUDA_OK(launch_pdl(assemble_dec_kernel<AT>, B * (L + 1), D / 4, 0, st, e, mask_token, dpos, ids_restore, L, keep, D, xd));
  LAUNCH_OK();
  return 0;
}
template <typename AT>
int assemble_decoder_input_bwd(const float* dxd, const int32_t* ids_restore, int B, int L, int keep, int D, AT* d_e,
                               float* d_mask_token, int accumulate, float* ws, cudaStream_t st) {
  ECAMP_CUDA_OK(launch_pdl(assemble_dec_bwd_kernel<AT>, B * (L + 1), D / 4, 0, st, dxd, ids_restore, L, keep, D, d_e));
  LAUNCH_OK();
  ECAMP_CUDA_OK(launch_pdl(mask_token_grad_kernel, B, D / 4, 0, st, dxd, ids_restore, L, keep, D, ws));
  LAUNCH_OK();
  ECAMP_CUDA_OK(launch_pdl(strided_rowsum_kernel, (D + 127) / 128, 128, 0, st, ws, B, (size_t)D, D, d_mask_token, accumulate));
  LAUNCH_OK();
  return 0;
}
template <typename AT>
int split_latent_gap(const AT* lat2, int B, int keep, int D, AT* img_tok, AT* gap, cudaStream_t st) {
  ECAMP_CUDA_OK(launch_pdl(split_latent_gap_kernel<AT>, B, D / 4, 0, st, lat2, keep, D, img_tok, gap));
  LAUNCH_OK();
  return 0;
}
template <typename AT>
int split_latent_gap_bwd(const AT* d_img_tok, const AT* d_gap, int B, int keep, int D, AT* d_lat2,
                         cudaStream_t st) {
  ECAMP_CUDA_OK(launch_pdl(split_latent_gap_bwd_kernel<AT>, B * (keep + 1), D / 4, 0, st, d_img_tok, d_gap, keep, D, d_lat2));
  LAUNCH_OK();
  return 0;
}
template <typename AT>
int add_batch_rowvec(AT* y, const AT* vec, int B, int T, int D, cudaStream_t st) {
  ECAMP_CUDA_OK(launch_pdl(add_batch_rowvec_kernel<AT>, B * T, D / 4, 0, st, y, vec, T, D));
  LAUNCH_OK();
  return 0;
}
template <typename AT>
int add_batch_rowvec_oop(const AT* x, const AT* vec, int B, int T, int D, AT* y, cudaStream_t st) {
  ECAMP_CUDA_OK(launch_pdl(add_batch_rowvec_oop_kernel<AT>, B * T, D / 4, 0, st, x, vec, T, D, y));
  LAUNCH_OK();
  return 0;
}
template <typename AT>
int gelu_bwd_bf16(const float* d, const AT* pre, AT* out, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  ECAMP_CUDA_OK(launch_pdl(gelu_bwd_bf16_kernel<AT>, (unsigned)((n + 255) / 256), 256, 0, st, d, pre, out, n));
  LAUNCH_OK();
  return 0;
}
template <typename AT>
int batch_colsum(const AT* x, int B, int T, int D, AT* out, cudaStream_t st) {
  ECAMP_CUDA_OK(launch_pdl(batch_colsum_kernel<AT>, B, D / 4, 0, st, x, T, D, out));
  LAUNCH_OK();
  return 0;
}
template <typename AT>
int bert_embeddings_fwd(const int64_t* ids, const int64_t* type_ids, const float* word, const float* type,
                        const float* pos, const float* gamma, const float* beta, float eps, int B, int T, int D,
                        DropoutCfg drop, float* pre, float* mean, float* rstd, AT* out_bf16, float* out_f32,
                        cudaStream_t st) {
  ECAMP_REQUIRE(D == 768, "bert embeddings: hidden size must be 768");
  const int M = B * T;
  ECAMP_CUDA_OK(launch_pdl(bert_emb_fwd_kernel<AT>, (M + 7) / 8, 256, 0, st, ids, type_ids, word, type, pos, gamma, beta, eps, M, T, drop, pre,
                                                   mean, rstd, out_bf16, out_f32));
  LAUNCH_OK();
  return 0;
}
int dropout_bwd_f32(float* g, size_t n, DropoutCfg drop, cudaStream_t st) {
  if (drop.p <= 0.f || n == 0) return 0;
  const size_t n4 = n / 4;
  ECAMP_CUDA_OK(launch_pdl(dropout_bwd_f32_kernel, (unsigned)((n4 + 255) / 256), 256, 0, st, g, n4, drop));
  LAUNCH_OK();
  return 0;
}
int bert_embeddings_bwd(const float* d_pre, const int64_t* ids, const int64_t* type_ids, int B, int T, int D,
                        float* d_word, float* d_type, float* d_pos, int accumulate, float* ws, cudaStream_t st) {
  const int M = B * T;
  if (!accumulate) {
    // rows that no token of this batch touches (and positions >= T) must read as zero, not as stale gradients
    ECAMP_CUDA_OK(cudaMemsetAsync(d_word, 0, (size_t)30000 * D * sizeof(float), st));
    ECAMP_CUDA_OK(cudaMemsetAsync(d_pos, 0, (size_t)256 * D * sizeof(float), st));
  }
  ECAMP_CUDA_OK(launch_pdl(emb_word_bwd_kernel, (M + kEmbRows - 1) / kEmbRows, D / 4, 0, st, d_pre, ids, M, D, d_word));
  LAUNCH_OK();
  ECAMP_CUDA_OK(launch_pdl(emb_pos_type_bwd_kernel, T, D / 4, 0, st, d_pre, type_ids, B, T, D, d_pos, accumulate, ws));
  LAUNCH_OK();
  ECAMP_CUDA_OK(launch_pdl(emb_type_finalize_kernel, (2 * D + 255) / 256, 256, 0, st, ws, T, D, d_type, accumulate));
  LAUNCH_OK();
  return 0;
}
size_t colsum_ws_floats(int) { return 64; }  // the column sums are added with atomics: no partials workspace any more
template <typename AT>
int colsum_bf16(const AT* x, int ld, int M, int N, float* out, int accumulate, float* /*ws*/, cudaStream_t st) {
  ECAMP_REQUIRE(ld % 8 == 0 && N % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
                "colsum: pitch and width must be multiples of 8 elements, base 16-byte aligned");
  int chunks = (M + 127) / 128;  // 128 rows per CTA: ~600-1500 CTAs for the shapes of the step
  if (chunks > 512) chunks = 512;
  if (chunks < 1) chunks = 1;
  if (!accumulate) ECAMP_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), st));
  dim3 grid((N + 255) / 256, chunks);
  ECAMP_CUDA_OK(launch_pdl(colsum_kernel<AT>, grid, 256, 0, st, x, ld, M, N, out));
  LAUNCH_OK();
  return 0;
}
int iota_mod_i32(int32_t* out, size_t n, int mod, cudaStream_t st) {
  if (n == 0) return 0;
  ECAMP_CUDA_OK(launch_pdl(iota_mod_kernel, (unsigned)((n + 255) / 256), 256, 0, st, out, n, mod));
  LAUNCH_OK();
  return 0;
}
int mean_pool_tokens(const float* x, int B, int S, int D, float* out, cudaStream_t st) {
  ECAMP_REQUIRE(D % 4 == 0 && D / 4 <= 1024 && S > 1, "mean_pool: unsupported shape");
  ECAMP_CUDA_OK(launch_pdl(mean_pool_kernel, B, D / 4, 0, st, x, S, D, out));
  LAUNCH_OK();
  return 0;
}
int mean_pool_tokens_bwd(const float* d_pooled, int B, int S, int D, float* dx, bf16* gx, const float* scale, cudaStream_t st) {
  ECAMP_REQUIRE(D % 4 == 0 && D / 4 <= 1024 && S > 1, "mean_pool_bwd: unsupported shape");
  ECAMP_CUDA_OK(launch_pdl(mean_pool_bwd_kernel, B * S, D / 4, 0, st, d_pooled, S, D, dx, gx, scale));
  LAUNCH_OK();
  return 0;
}
int strided_rowsum(const float* x, int B, size_t stride, int D, float* out, int accumulate, cudaStream_t st) {
  ECAMP_CUDA_OK(launch_pdl(strided_rowsum_kernel, (D + 127) / 128, 128, 0, st, x, B, stride, D, out, accumulate));
  LAUNCH_OK();
  return 0;
}
int scale_f32(float* x, const float* scale_dev, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  ECAMP_CUDA_OK(launch_pdl(scale_f32_kernel, (unsigned)((n + 255) / 256), 256, 0, st, x, scale_dev, n));
  LAUNCH_OK();
  return 0;
}
int cast_f32_to_bf16(const float* x, bf16* y, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  ECAMP_CUDA_OK(launch_pdl(cast_bf16_kernel, (unsigned)((n + 255) / 256), 256, 0, st, x, y, n));
  LAUNCH_OK();
  return 0;
}
int permute_pe_weight_grad(const float* dw_pqc, float* grad_cpq, int accumulate, cudaStream_t st) {
  ECAMP_CUDA_OK(launch_pdl(permute_pe_grad_kernel, (768 * 768 + 255) / 256, 256, 0, st, dw_pqc, grad_cpq, accumulate));
  LAUNCH_OK();
  return 0;
}


// ---- explicit instantiations: AT = bf16 (production) and float (fp32-accurate parity mode) ---------------------------
#define ECAMP_INST_ELEMENTWISE(AT)                                                                                        \
  template int gather_patches<AT>(const float*, const int32_t*, int, int, int, int, AT*, cudaStream_t);                   \
  template int assemble_encoder_input_bwd<AT>(const float*, int, int, int, AT*, float*, int, cudaStream_t);               \
  template int assemble_decoder_input<AT>(const AT*, const float*, const float*, const int32_t*, int, int, int, int,      \
                                          float*, cudaStream_t);                                                          \
  template int assemble_decoder_input_bwd<AT>(const float*, const int32_t*, int, int, int, int, AT*, float*, int, float*, \
                                              cudaStream_t);                                                              \
  template int split_latent_gap<AT>(const AT*, int, int, int, AT*, AT*, cudaStream_t);                                    \
  template int split_latent_gap_bwd<AT>(const AT*, const AT*, int, int, int, AT*, cudaStream_t);                          \
  template int add_batch_rowvec<AT>(AT*, const AT*, int, int, int, cudaStream_t);                                         \
  template int add_batch_rowvec_oop<AT>(const AT*, const AT*, int, int, int, AT*, cudaStream_t);                          \
  template int gelu_bwd_bf16<AT>(const float*, const AT*, AT*, size_t, cudaStream_t);                                     \
  template int batch_colsum<AT>(const AT*, int, int, int, AT*, cudaStream_t);                                             \
  template int bert_embeddings_fwd<AT>(const int64_t*, const int64_t*, const float*, const float*, const float*,          \
                                       const float*, const float*, float, int, int, int, DropoutCfg, float*, float*,      \
                                       float*, AT*, float*, cudaStream_t);                                                \
  template int colsum_bf16<AT>(const AT*, int, int, int, float*, int, float*, cudaStream_t);
ECAMP_INST_ELEMENTWISE(bf16)
ECAMP_INST_ELEMENTWISE(float)
#undef ECAMP_INST_ELEMENTWISE

// ---------------------------------------------------------------------------------------------
// Tail of the image loader on the GPU: Grayscale(3) + ToTensor + Normalize of pretrain_datasets.py:47-52 applied to the
// loader's 8-bit grayscale crop, so that a step ships 1 byte per pixel over PCIe instead of 12 (3 identical fp32
// channels: 617 MB per 256-pair batch, which can take longer than the step itself).  Bit-exact with the CPU transform:
// ToTensor is u8 -> fp32 divided by 255 (correctly rounded division), Normalize is (x - mean) / std in fp32 - the
// same three IEEE operations, no FMA contraction, no approximate division.  HBM-bound: 1 B read + 12 B written / pixel.
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) u8_gray_normalize_kernel(const uint8_t* __restrict__ in, long long pixels_per_image,
                                                                long long n_images, float mean, float stdv,
                                                                float* __restrict__ out) {
  const long long groups_per_image = pixels_per_image / 16;  // 16 pixels per thread (host checks divisibility)
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= groups_per_image * n_images) return;
  const long long img = gid / groups_per_image, grp = gid - img * groups_per_image;
  const uint4 raw = __ldg(reinterpret_cast<const uint4*>(in + img * pixels_per_image) + grp);
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
  float4 v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float f[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const float x = __fdiv_rn((float)((w[k] >> (8 * b)) & 0xffu), 255.0f);  // ToTensor
      f[b] = __fdiv_rn(__fsub_rn(x, mean), stdv);                                // Normalize
    }
    v[k] = make_float4(f[0], f[1], f[2], f[3]);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float4* dst = reinterpret_cast<float4*>(out + (img * 3 + c) * pixels_per_image) + grp * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[k] = v[k];
  }
}
}  // namespace

int image_u8_normalize(const uint8_t* gray, long long n_images, long long pixels_per_image, float mean, float stdv,
                       float* out, cudaStream_t st) {
  ECAMP_REQUIRE(gray && out, "image_u8_normalize: null pointer");
  ECAMP_REQUIRE(n_images > 0 && pixels_per_image > 0 && pixels_per_image % 16 == 0,
                "image_u8_normalize: pixels per image must be a positive multiple of 16 (got %lld)", pixels_per_image);
  ECAMP_REQUIRE((reinterpret_cast<uintptr_t>(gray) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "image_u8_normalize: buffers must be 16-byte aligned");
  ECAMP_REQUIRE(stdv != 0.f, "image_u8_normalize: std must not be zero");
  const long long groups = n_images * (pixels_per_image / 16);
  u8_gray_normalize_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(gray, pixels_per_image, n_images, mean, stdv, out);
  ECAMP_CUDA_OK(cudaGetLastError());
  ECAMP_LAUNCHED();
  return 0;
}

}  // namespace ecamp
