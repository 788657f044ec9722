// LayerNorm forward / backward (HBM-bound; one warp per row, 128-bit accesses).
// Replaces ATen layer_norm behind timm Block.norm1/norm2, ECAMP.norm / decoder_norm
// (model_ecamp.py:66-69,80-84) and every HF BertSelfOutput / BertOutput / BertEmbeddings LayerNorm.
#include "kernels.cuh"

namespace ecamp {
namespace {

constexpr int kRowsPerBlock = 8;
constexpr int kBwdBlocks = 592;  // 148 SMs x 4

template <int NV>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps, int M,
                                                     bf16* __restrict__ out_bf16, float* __restrict__ out_f32,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * kRowsPerBlock + warp;
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = xr[i * 32 + lane];
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
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
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
    if (out_f32) reinterpret_cast<float4*>(out_f32 + (size_t)row * D)[c4] = y;
    if (out_bf16) {
      uint2 u;
      u.x = pack_bf16x2(y.x, y.y);
      u.y = pack_bf16x2(y.z, y.w);
      reinterpret_cast<uint2*>(out_bf16 + (size_t)row * D)[c4] = u;
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, int M,
                                                     const float* __restrict__ addend, float* __restrict__ dx_f32,
                                                     bf16* __restrict__ dx_bf16, DropoutCfg drop,
                                                     float* __restrict__ partial) {
  constexpr int D = NV * 128;
  __shared__ float red[kRowsPerBlock][D];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 dg[NV], db[NV], g[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    g[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
  }
  const Philox ph(drop.seed);
  const uint32_t thr = dropout_threshold(drop.p);
  const float keep_scale = drop.p > 0.f ? 1.0f / (1.0f - drop.p) : 1.0f;

  for (int row = blockIdx.x * kRowsPerBlock + warp; row < M; row += gridDim.x * kRowsPerBlock) {
    const float mu = mean[row], rs = rstd[row];
    const float4* dyr = reinterpret_cast<const float4*>(dy + (size_t)row * D);
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
    float4 dv[NV], xh[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 d = dyr[i * 32 + lane];
      const float4 xv = xr[i * 32 + lane];
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      dg[i].x += d.x * xh[i].x; dg[i].y += d.y * xh[i].y; dg[i].z += d.z * xh[i].z; dg[i].w += d.w * xh[i].w;
      db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
      dv[i] = make_float4(d.x * g[i].x, d.y * g[i].y, d.z * g[i].z, d.w * g[i].w);
      s1 += dv[i].x + dv[i].y + dv[i].z + dv[i].w;
      s2 += dv[i].x * xh[i].x + dv[i].y * xh[i].y + dv[i].z * xh[i].z + dv[i].w * xh[i].w;
    }
    const float c1 = warp_sum(s1) * (1.0f / D);
    const float c2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = i * 32 + lane;
      float4 r;
      r.x = rs * (dv[i].x - c1 - xh[i].x * c2);
      r.y = rs * (dv[i].y - c1 - xh[i].y * c2);
      r.z = rs * (dv[i].z - c1 - xh[i].z * c2);
      r.w = rs * (dv[i].w - c1 - xh[i].w * c2);
      if (addend) {
        const float4 a = reinterpret_cast<const float4*>(addend + (size_t)row * D)[c4];
        r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
      }
      if (dx_f32) reinterpret_cast<float4*>(dx_f32 + (size_t)row * D)[c4] = r;
      if (dx_bf16) {
        if (drop.p > 0.f) {
          const uint4 rnd = ph(((uint64_t)row * D + (uint64_t)c4 * 4) >> 2, drop.site);
          r.x = rnd.x >= thr ? r.x * keep_scale : 0.f;
          r.y = rnd.y >= thr ? r.y * keep_scale : 0.f;
          r.z = rnd.z >= thr ? r.z * keep_scale : 0.f;
          r.w = rnd.w >= thr ? r.w * keep_scale : 0.f;
        }
        uint2 u;
        u.x = pack_bf16x2(r.x, r.y);
        u.y = pack_bf16x2(r.z, r.w);
        reinterpret_cast<uint2*>(dx_bf16 + (size_t)row * D)[c4] = u;
      }
    }
  }
  // block-level reduction of the per-warp dgamma / dbeta partials, two passes through shared memory
  float* pg = partial + (size_t)blockIdx.x * 2 * D;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i)
      reinterpret_cast<float4*>(&red[warp][0])[i * 32 + lane] = pass == 0 ? dg[i] : db[i];
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kRowsPerBlock; ++w) s += red[w][c];
      pg[pass * D + c] = s;
    }
  }
}

// one CTA per 32 columns of [dgamma | dbeta]; 8 warps stride over the per-CTA partial rows
__global__ void __launch_bounds__(256) ln_bwd_finalize(const float* __restrict__ partial, int nblocks, int D,
                                                       float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                       int accumulate) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;  // 2 * D is a multiple of 32
  float s = 0.f;
  for (int b = warp; b < nblocks; b += 8) s += partial[(size_t)b * 2 * D + c];
  red[warp][lane] = s;
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) s += red[w][lane];
    float* dst = c < D ? dgamma + c : dbeta + (c - D);
    *dst = accumulate ? *dst + s : s;
  }
}

int bwd_blocks(int M) {
  const int need = (M + kRowsPerBlock - 1) / kRowsPerBlock;
  return need < kBwdBlocks ? need : kBwdBlocks;
}

}  // namespace

size_t layernorm_bwd_ws_floats(int D) { return (size_t)kBwdBlocks * 2 * D; }

int layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int M, int D, bf16* out_bf16,
                  float* out_f32, float* mean, float* rstd, cudaStream_t st) {
  ECAMP_REQUIRE(D == 768 || D == 512, "layernorm: D must be 512 or 768 (got %d)", D);
  if (M <= 0) return 0;
  const int grid = (M + kRowsPerBlock - 1) / kRowsPerBlock;
  if (D == 768)
    ln_fwd_kernel<6><<<grid, 256, 0, st>>>(x, gamma, beta, eps, M, out_bf16, out_f32, mean, rstd);
  else
    ln_fwd_kernel<4><<<grid, 256, 0, st>>>(x, gamma, beta, eps, M, out_bf16, out_f32, mean, rstd);
  ECAMP_LAUNCHED();
  return 0;
}

int layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, int M,
                  int D, const float* addend, float* dx_f32, bf16* dx_bf16, DropoutCfg drop, float* dgamma,
                  float* dbeta, int accumulate, float* partial_ws, cudaStream_t st) {
  ECAMP_REQUIRE(D == 768 || D == 512, "layernorm: D must be 512 or 768 (got %d)", D);
  if (M <= 0) return 0;
  const int grid = bwd_blocks(M);
  if (D == 768)
    ln_bwd_kernel<6><<<grid, 256, 0, st>>>(dy, x, mean, rstd, gamma, M, addend, dx_f32, dx_bf16, drop, partial_ws);
  else
    ln_bwd_kernel<4><<<grid, 256, 0, st>>>(dy, x, mean, rstd, gamma, M, addend, dx_f32, dx_bf16, drop, partial_ws);
  ECAMP_LAUNCHED();
  if (dgamma && dbeta) {
    ln_bwd_finalize<<<2 * D / 32, 256, 0, st>>>(partial_ws, grid, D, dgamma, dbeta, accumulate);
    ECAMP_LAUNCHED();
  }
  return 0;
}

}  // namespace ecamp
