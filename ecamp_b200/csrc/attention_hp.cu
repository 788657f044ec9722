// Attention for the fp32-accurate parity mode: the SAME function as attention.cu / attention_tc.cu (softmax(scale q k^T +
// key mask) -> dropout -> @ v, per-row log-sum-exp saved; backward with delta_i = sum_j P_ij dPeff_ij) on fp32 operands with
// fp32 CUDA-core arithmetic and the same counter-based dropout masks.  Not a performance path: one warp per query row (forward,
// dQ) or per key row (dK / dV); the backward hands dS and the dropped probabilities over through a global scratch.
// Mirrors timm Attention (model_ecamp.py:66-68,80-82) and HF BertSelfAttention (context_fusion.py:32-53).
#include "kernels.cuh"

namespace ecamp {
namespace {

constexpr int kWarps = 4;
constexpr int kMaxSk = 256;

// scores of query row (b, h, i) against all keys -> s[j] (scaled, masked: -inf); returns the row maximum
ECAMP_DEVINL float score_row(const AttnArgsT<float>& a, int b, int h, int i, int lane, const float* sq, float* s) {
  float mx = -INFINITY;
  for (int j = lane; j < a.Sk; j += 32) {
    const float* kr = a.k + ((size_t)b * a.Sk + j) * a.ldk + h * a.D;
    float acc = 0.f;
    for (int d = 0; d < a.D; ++d) acc = fmaf(sq[d], kr[d], acc);
    const bool ok = a.key_mask == nullptr || a.key_mask[(size_t)b * a.Sk + j] != 0;
    const float v = ok ? acc * a.scale : -INFINITY;
    s[j] = v;
    mx = fmaxf(mx, v);
  }
  return warp_max(mx);
}

__global__ void __launch_bounds__(kWarps * 32) attn_hp_fwd_kernel(AttnArgsT<float> a) {
  __shared__ float sq[kWarps][128];
  __shared__ float ss[kWarps][kMaxSk];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kWarps + warp, h = blockIdx.y, b = blockIdx.z;
  if (i >= a.Sq) return;
  const uint64_t bh = (uint64_t)b * a.H + h;
  for (int d = lane; d < a.D; d += 32) sq[warp][d] = a.q[((size_t)b * a.Sq + i) * a.ldq + h * a.D + d];
  __syncwarp();
  float* s = ss[warp];
  const float mx = score_row(a, b, h, i, lane, sq[warp], s);
  const float m_use = mx == -INFINITY ? 0.f : mx;
  float l = 0.f;
  for (int j = lane; j < a.Sk; j += 32) {
    const float p = expf(s[j] - m_use);  // exp(-inf) = 0 for masked keys
    s[j] = p;
    l += p;
  }
  l = warp_sum(l);
  __syncwarp();
  const float inv = l > 0.f ? 1.0f / l : 0.f;
  const Philox ph(a.drop.seed);
  const uint32_t thr = dropout_threshold16(a.drop.p);
  const float ks = dropout_keep_scale16(thr);
  const bool use_drop = a.drop.p > 0.f;
  const int kgroups = (a.Sk + 7) >> 3;
  if (use_drop) {
    for (int g = lane; g < kgroups; g += 32) {
      const uint32_t keep = philox_keep8(ph, bh * a.Sq + i, kgroups, g, a.drop.site, thr);
      for (int t = 0; t < 8 && g * 8 + t < a.Sk; ++t) s[g * 8 + t] = ((keep >> t) & 1u) ? s[g * 8 + t] * ks : 0.f;
    }
    __syncwarp();
  }
  for (int d = lane; d < a.D; d += 32) {
    float acc = 0.f;
    for (int j = 0; j < a.Sk; ++j) acc = fmaf(s[j], a.v[((size_t)b * a.Sk + j) * a.ldv + h * a.D + d], acc);
    a.o[((size_t)b * a.Sq + i) * a.ldo + h * a.D + d] = acc * inv;
  }
  if (lane == 0 && a.lse) a.lse[bh * a.Sq + i] = l > 0.f ? mx + logf(l) : -INFINITY;
}

// per query row: P, dP, delta, dS -> scratch; dQ
__global__ void __launch_bounds__(kWarps * 32) attn_hp_bwd_q_kernel(AttnArgsT<float> a) {
  __shared__ float sq[kWarps][128];
  __shared__ float sdo[kWarps][128];
  __shared__ float ss[kWarps][kMaxSk];
  __shared__ float sp[kWarps][kMaxSk];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kWarps + warp, h = blockIdx.y, b = blockIdx.z;
  if (i >= a.Sq) return;
  const uint64_t bh = (uint64_t)b * a.H + h;
  for (int d = lane; d < a.D; d += 32) {
    sq[warp][d] = a.q[((size_t)b * a.Sq + i) * a.ldq + h * a.D + d];
    sdo[warp][d] = a.d_o[((size_t)b * a.Sq + i) * a.ld_do + h * a.D + d];
  }
  __syncwarp();
  float* s = ss[warp];
  float* pd = sp[warp];
  score_row(a, b, h, i, lane, sq[warp], s);
  const float lse = a.lse[bh * a.Sq + i];
  const Philox ph(a.drop.seed);
  const uint32_t thr = dropout_threshold16(a.drop.p);
  const bool use_drop = a.drop.p > 0.f;
  const float ks = use_drop ? dropout_keep_scale16(thr) : 1.0f;
  const int kgroups = (a.Sk + 7) >> 3;
  float delta = 0.f;
  for (int j = lane; j < a.Sk; j += 32) {
    const float p = lse == -INFINITY ? 0.f : expf(s[j] - lse);
    const float* vr = a.v + ((size_t)b * a.Sk + j) * a.ldv + h * a.D;
    float dp = 0.f;
    for (int d = 0; d < a.D; ++d) dp = fmaf(sdo[warp][d], vr[d], dp);
    bool kp = true;
    if (use_drop) kp = (philox_keep8(ph, bh * a.Sq + i, kgroups, j >> 3, a.drop.site, thr) >> (j & 7)) & 1u;
    const float dpe = kp ? dp * ks : 0.f;
    delta += p * dpe;
    s[j] = p;      // P
    pd[j] = dpe;   // dPeff
  }
  delta = warp_sum(delta);
  __syncwarp();
  float* ws_ds = a.hp_ws + (bh * a.Sq + i) * (uint64_t)a.Sk;
  float* ws_pd = a.hp_ws + (uint64_t)a.B * a.H * a.Sq * a.Sk + (bh * a.Sq + i) * (uint64_t)a.Sk;
  for (int j = lane; j < a.Sk; j += 32) {
    const float p = s[j];
    const float ds = p * (pd[j] - delta) * a.scale;
    bool kp = true;
    if (use_drop) kp = (philox_keep8(ph, bh * a.Sq + i, kgroups, j >> 3, a.drop.site, thr) >> (j & 7)) & 1u;
    ws_ds[j] = ds;
    ws_pd[j] = kp ? p * ks : 0.f;
    s[j] = ds;
  }
  __syncwarp();
  if (lane == 0 && a.delta) a.delta[bh * a.Sq + i] = delta;
  for (int d = lane; d < a.D; d += 32) {
    float acc = 0.f;
    for (int j = 0; j < a.Sk; ++j) acc = fmaf(s[j], a.k[((size_t)b * a.Sk + j) * a.ldk + h * a.D + d], acc);
    a.dq[((size_t)b * a.Sq + i) * a.lddq + h * a.D + d] = acc;
    if (a.cs_q) atomicAdd(a.cs_q + h * a.D + d, acc);
  }
}

// per key row: dK_j = sum_i dS_ij q_i, dV_j = sum_i Pdrop_ij dO_i
__global__ void __launch_bounds__(kWarps * 32) attn_hp_bwd_kv_kernel(AttnArgsT<float> a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * kWarps + warp, h = blockIdx.y, b = blockIdx.z;
  if (j >= a.Sk) return;
  const uint64_t bh = (uint64_t)b * a.H + h;
  const float* ws_ds = a.hp_ws + bh * a.Sq * (uint64_t)a.Sk + j;
  const float* ws_pd = ws_ds + (uint64_t)a.B * a.H * a.Sq * a.Sk;
  for (int d = lane; d < a.D; d += 32) {
    float dk = 0.f, dv = 0.f;
    for (int i = 0; i < a.Sq; ++i) {
      dk = fmaf(ws_ds[(uint64_t)i * a.Sk], a.q[((size_t)b * a.Sq + i) * a.ldq + h * a.D + d], dk);
      dv = fmaf(ws_pd[(uint64_t)i * a.Sk], a.d_o[((size_t)b * a.Sq + i) * a.ld_do + h * a.D + d], dv);
    }
    a.dk[((size_t)b * a.Sk + j) * a.lddk + h * a.D + d] = dk;
    a.dv[((size_t)b * a.Sk + j) * a.lddv + h * a.D + d] = dv;
    if (a.cs_k) atomicAdd(a.cs_k + h * a.D + d, dk);
    if (a.cs_v) atomicAdd(a.cs_v + h * a.D + d, dv);
  }
}

int check_hp(const AttnArgsT<float>& a, bool bwd) {
  ECAMP_REQUIRE(a.q && a.k && a.v && a.o && a.lse, "attention (fp32): null pointer");
  ECAMP_REQUIRE(a.B > 0 && a.H > 0 && a.Sq > 0 && a.Sk > 0 && a.Sk <= kMaxSk && a.D > 0 && a.D <= 128,
                "attention (fp32): unsupported shape B=%d H=%d Sq=%d Sk=%d D=%d", a.B, a.H, a.Sq, a.Sk, a.D);
  if (bwd) ECAMP_REQUIRE(a.d_o && a.dq && a.dk && a.dv && a.hp_ws, "attention (fp32) backward: null pointer");
  return 0;
}

}  // namespace

int attention_fwd(const AttnArgsT<float>& a, cudaStream_t st) {
  if (int rc = check_hp(a, false)) return rc;
  dim3 grid((a.Sq + kWarps - 1) / kWarps, a.H, a.B);
  attn_hp_fwd_kernel<<<grid, kWarps * 32, 0, st>>>(a);
  ECAMP_LAUNCHED();
  return 0;
}

int attention_bwd(const AttnArgsT<float>& a, cudaStream_t st) {
  if (int rc = check_hp(a, true)) return rc;
  dim3 gq((a.Sq + kWarps - 1) / kWarps, a.H, a.B), gk((a.Sk + kWarps - 1) / kWarps, a.H, a.B);
  attn_hp_bwd_q_kernel<<<gq, kWarps * 32, 0, st>>>(a);
  ECAMP_LAUNCHED();
  attn_hp_bwd_kv_kernel<<<gk, kWarps * 32, 0, st>>>(a);
  ECAMP_LAUNCHED();
  return 0;
}

}  // namespace ecamp
