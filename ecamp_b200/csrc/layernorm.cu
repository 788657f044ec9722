// LayerNorm forward / backward (HBM-bound; one warp per row, 128-bit accesses).
// Replaces ATen layer_norm behind timm Block.norm1/norm2, ECAMP.norm / decoder_norm
// (model_ecamp.py:66-69,80-84) and every HF BertSelfOutput / BertOutput / BertEmbeddings LayerNorm.
#include "kernels.cuh"

#include <cstdlib>

namespace ecamp {
namespace {

constexpr int kRowsPerBlock = 8;
constexpr int kBwdWarps = 4;     // rows in flight per backward CTA (one shared-memory slab of column sums per warp)
constexpr int kBwdBlocks = 740;  // 148 SMs x 5 resident CTAs (slabs + gamma: 27-40 KB per CTA)

template <int NV, typename AT>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps, int M,
                                                     AT* __restrict__ out_bf16, float* __restrict__ out_f32,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  ECAMP_PDL_ENTRY();
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
    if (out_bf16) st4(out_bf16 + (size_t)row * D + 4 * c4, y);
  }
}

// Backward.  One warp per row; lane l owns float4 columns {32 i + l}.  The per-column sums (dgamma, dbeta and,
// optionally, the column sum of the emitted gradient = the bias gradient of the Linear that consumes it) are
// accumulated in SHARED memory, one private [3][D] slab per warp (each lane only ever touches its own columns, so
// there are no conflicts and no atomics), instead of 12-18 float4 registers per thread: the register version ran
// at 151 registers / one CTA per SM and 2.9 TB/s.  At the end the eight slabs are summed and added to global memory
// with fp32 atomics (the destinations are zeroed at the start of the backward pass), which also removes the
// separate finalize kernel.
template <int NV, bool COLSUM, typename AT>
__global__ void __launch_bounds__(kBwdWarps * 32, 6) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, int M,
                                                     const float* __restrict__ addend, float* __restrict__ dx_f32,
                                                     AT* __restrict__ dx_bf16, DropoutCfg drop,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                     float* __restrict__ colsum_out,
                                                     const float* __restrict__ out_row_scale, int rows_per_scale) {
  ECAMP_PDL_ENTRY();
  constexpr int D = NV * 128;
  constexpr int NA = COLSUM ? 3 : 2;
  extern __shared__ __align__(16) float sm_ln[];
  float* s_gamma = sm_ln;                       // [D]
  float* slab = sm_ln + D;                      // [8 warps][NA][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x)
    reinterpret_cast<float4*>(s_gamma)[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i);
  for (int i = threadIdx.x; i < kBwdWarps * NA * D / 4; i += blockDim.x)
    reinterpret_cast<float4*>(slab)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  float4* my_dg = reinterpret_cast<float4*>(slab + (size_t)warp * NA * D);
  float4* my_db = my_dg + D / 4;
  float4* my_cs = my_db + D / 4;
  const float4* g4 = reinterpret_cast<const float4*>(s_gamma);
  const Philox ph(drop.seed);
  const uint32_t thr = dropout_threshold(drop.p);
  const float keep_scale = drop.p > 0.f ? 1.0f / (1.0f - drop.p) : 1.0f;

  for (int row = blockIdx.x * kBwdWarps + warp; row < M; row += gridDim.x * kBwdWarps) {
    const float mu = mean[row], rs = rstd[row];
    const float4* dyr = reinterpret_cast<const float4*>(dy + (size_t)row * D);
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
    float4 dv[NV], xh[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {  // all global loads of the row in flight before the first use
      dv[i] = dyr[i * 32 + lane];
      xh[i] = xr[i * 32 + lane];
    }
    // the addend row is only needed after the two row reductions: pull it into L2 now (no registers held)
    if (addend && lane < D / 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(addend + (size_t)row * D + lane * 32));
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = i * 32 + lane;
      const float4 d = dv[i];
      xh[i] = make_float4((xh[i].x - mu) * rs, (xh[i].y - mu) * rs, (xh[i].z - mu) * rs, (xh[i].w - mu) * rs);
      float4 a = my_dg[c4];
      a.x += d.x * xh[i].x; a.y += d.y * xh[i].y; a.z += d.z * xh[i].z; a.w += d.w * xh[i].w;
      my_dg[c4] = a;
      float4 b = my_db[c4];
      b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
      my_db[c4] = b;
      const float4 g = g4[c4];
      dv[i] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
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
        if (out_row_scale) {  // DropPath of the branch this gradient feeds (per-sample mask / keep_prob)
          const float rs_ = __ldg(out_row_scale + row / rows_per_scale);
          r.x *= rs_; r.y *= rs_; r.z *= rs_; r.w *= rs_;
        }
        if (drop.p > 0.f) {
          const uint4 rnd = ph(((uint64_t)row * D + (uint64_t)c4 * 4) >> 2, drop.site);
          r.x = rnd.x >= thr ? r.x * keep_scale : 0.f;
          r.y = rnd.y >= thr ? r.y * keep_scale : 0.f;
          r.z = rnd.z >= thr ? r.z * keep_scale : 0.f;
          r.w = rnd.w >= thr ? r.w * keep_scale : 0.f;
        }
        st4(dx_bf16 + (size_t)row * D + 4 * c4, r);
        if (COLSUM) {  // bias gradient of the Linear fed by dx_bf16: column sum of what that GEMM reads
          const float4 rr = act_round4(dx_bf16, r);
          float4 cs = my_cs[c4];
          cs.x += rr.x; cs.y += rr.y; cs.z += rr.z; cs.w += rr.w;
          my_cs[c4] = cs;
        }
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < NA * D; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kBwdWarps; ++w) s += slab[(size_t)w * NA * D + c];
    const int which = c / D, col = c - which * D;
    float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : colsum_out);
    if (dst) atomicAdd(dst + col, s);
  }
}

// Backward, staged form (default).  The slab kernel above streams at the HBM rate once it is going, but every launch pays
// a fixed ~15 us (measured: 41 us at M = 12800 against 79 us at M = 32768) for 740 CTAs that each zero a slab, own four rows or so
// per warp and then push 2304 atomics; with 42 launches per step that was 0.7 ms.  Here ONE CTA per SM keeps
// kStagedWarps warps busy; every warp owns a ring of whole-row buffers in shared memory that it fills itself with bulk
// asynchronous copies (cp.async.bulk -> mbarrier complete_tx), so the bytes in flight no longer depend on resident
// warps or on registers, the row is read twice from shared memory instead of being held in 48 registers, and the
// per-column sums live in registers (lane l only ever touches its own columns).  One reduction over the warps and 148
// x 3 D atomics end the launch.
constexpr int kStagedWarps = 8;

ECAMP_DEVINL void bulk_load_row(float* smem_dst, const float* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int NV, bool COLSUM, bool ADDEND, typename AT>
__global__ void __launch_bounds__(kStagedWarps * 32, 1)
    ln_bwd_staged_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                         const float* __restrict__ rstd, const float* __restrict__ gamma, int M,
                         const float* __restrict__ addend, float* __restrict__ dx_f32, AT* __restrict__ dx_bf16,
                         DropoutCfg drop, float* __restrict__ dgamma, float* __restrict__ dbeta,
                         float* __restrict__ colsum_out, const float* __restrict__ out_row_scale, int rows_per_scale) {
  constexpr int D = NV * 128;
  constexpr int W = kStagedWarps;
  constexpr int NROW = ADDEND ? 3 : 2;           // arrays staged per row: dy, x (, addend)
  constexpr int STAGES = ADDEND ? 3 : 4;
  constexpr int NA = COLSUM ? 3 : 2;
  constexpr uint32_t kRowBytes = D * sizeof(float);
  extern __shared__ __align__(128) float sm_ln[];
  float* s_gamma = sm_ln;                                                   // [D]
  float* ring = sm_ln + D;                                                  // [W][STAGES][NROW][D]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)W * STAGES * NROW * D);  // [W][STAGES]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* my_ring = ring + (size_t)warp * STAGES * NROW * D;
  uint64_t* my_bar = bars + warp * STAGES;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(my_bar + s, 1);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  ECAMP_PDL_ENTRY();
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x)
    reinterpret_cast<float4*>(s_gamma)[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i);
  __syncthreads();

  const int row0 = blockIdx.x * W + warp, stride = gridDim.x * W;
  auto issue = [&](int s, int row) {  // lane 0 only
    float* dst = my_ring + (size_t)s * NROW * D;
    mbar_arrive_expect_tx(my_bar + s, NROW * kRowBytes);
    bulk_load_row(dst, dy + (size_t)row * D, kRowBytes, my_bar + s);
    bulk_load_row(dst + D, x + (size_t)row * D, kRowBytes, my_bar + s);
    if (ADDEND) bulk_load_row(dst + 2 * D, addend + (size_t)row * D, kRowBytes, my_bar + s);
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s)
      if (row0 + s * stride < M) issue(s, row0 + s * stride);
  }
  const float4* g4 = reinterpret_cast<const float4*>(s_gamma);
  const Philox ph(drop.seed);
  const uint32_t thr = dropout_threshold(drop.p);
  const float keep_scale = drop.p > 0.f ? 1.0f / (1.0f - drop.p) : 1.0f;
  float4 dg[NV], db[NV], cs[COLSUM ? NV : 1];
#pragma unroll
  for (int i = 0; i < NV; ++i) dg[i] = db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < (COLSUM ? NV : 1); ++i) cs[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  float mu_n = 0.f, rs_n = 0.f;
  if (row0 < M) { mu_n = __ldg(mean + row0); rs_n = __ldg(rstd + row0); }
  int s = 0;
  uint32_t parity = 0;
  for (int row = row0; row < M; row += stride) {
    const float mu = mu_n, rs = rs_n;
    if (row + stride < M) { mu_n = __ldg(mean + row + stride); rs_n = __ldg(rstd + row + stride); }
    mbar_wait(my_bar + s, parity);
    const float4* dy4 = reinterpret_cast<const float4*>(my_ring + (size_t)s * NROW * D);
    const float4* x4 = dy4 + D / 4;
    const float4* a4 = x4 + D / 4;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = i * 32 + lane;
      const float4 d = dy4[c4], xv = x4[c4], g = g4[c4];
      const float4 xh = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      dg[i].x += d.x * xh.x; dg[i].y += d.y * xh.y; dg[i].z += d.z * xh.z; dg[i].w += d.w * xh.w;
      db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
      const float4 t = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
      s1 += t.x + t.y + t.z + t.w;
      s2 += t.x * xh.x + t.y * xh.y + t.z * xh.z + t.w * xh.w;
    }
    const float c1 = warp_sum(s1) * (1.0f / D);
    const float c2 = warp_sum(s2) * (1.0f / D);
    float row_scale = 1.f;
    if (dx_bf16 && out_row_scale) row_scale = __ldg(out_row_scale + row / rows_per_scale);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = i * 32 + lane;
      const float4 d = dy4[c4], xv = x4[c4], g = g4[c4];
      const float4 xh = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      float4 r;
      r.x = rs * (d.x * g.x - c1 - xh.x * c2);
      r.y = rs * (d.y * g.y - c1 - xh.y * c2);
      r.z = rs * (d.z * g.z - c1 - xh.z * c2);
      r.w = rs * (d.w * g.w - c1 - xh.w * c2);
      if (ADDEND) {
        const float4 a = a4[c4];
        r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
      }
      if (dx_f32) reinterpret_cast<float4*>(dx_f32 + (size_t)row * D)[c4] = r;
      if (dx_bf16) {
        if (out_row_scale) { r.x *= row_scale; r.y *= row_scale; r.z *= row_scale; r.w *= row_scale; }
        if (drop.p > 0.f) {
          const uint4 rnd = ph(((uint64_t)row * D + (uint64_t)c4 * 4) >> 2, drop.site);
          r.x = rnd.x >= thr ? r.x * keep_scale : 0.f;
          r.y = rnd.y >= thr ? r.y * keep_scale : 0.f;
          r.z = rnd.z >= thr ? r.z * keep_scale : 0.f;
          r.w = rnd.w >= thr ? r.w * keep_scale : 0.f;
        }
        st4(dx_bf16 + (size_t)row * D + 4 * c4, r);
        if (COLSUM) {
          const float4 rr = act_round4(dx_bf16, r);
          cs[i].x += rr.x; cs[i].y += rr.y; cs[i].z += rr.z; cs[i].w += rr.w;
        }
      }
    }
    __syncwarp();  // every lane is done reading this stage before it is refilled
    if (lane == 0 && row + STAGES * stride < M) issue(s, row + STAGES * stride);
    if (++s == STAGES) { s = 0; parity ^= 1u; }
  }
  // every copy this warp issued has been waited for; the ring is free: reuse it for the reduction over the warps
  __syncthreads();
  float4* red = reinterpret_cast<float4*>(ring) + (size_t)warp * NA * (D / 4);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    red[i * 32 + lane] = dg[i];
    red[D / 4 + i * 32 + lane] = db[i];
    if (COLSUM) red[2 * (D / 4) + i * 32 + lane] = cs[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < NA * D; c += blockDim.x) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < W; ++w) sum += ring[(size_t)w * NA * D + c];
    const int which = c / D, col = c - which * D;
    float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : colsum_out);
    if (dst) atomicAdd(dst + col, sum);
  }
}

int num_sms_ln() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int bwd_blocks(int M) {
  const int need = (M + kBwdWarps - 1) / kBwdWarps;
  static const int cap = getenv("ECAMP_LN_BWD_BLOCKS") ? atoi(getenv("ECAMP_LN_BWD_BLOCKS")) : kBwdBlocks;  // tuning knob
  return need < cap ? need : cap;
}

int g_ln_bwd_slab = -1;  // 1: the slab kernel (kept for A/B measurements: ECAMP_LN_BWD_SLAB=1 or ecamp_layernorm_set_bwd_slab)
bool use_slab() {
  if (g_ln_bwd_slab < 0) g_ln_bwd_slab = getenv("ECAMP_LN_BWD_SLAB") ? atoi(getenv("ECAMP_LN_BWD_SLAB")) : 0;
  return g_ln_bwd_slab != 0;
}

template <int NV, bool COLSUM, bool ADDEND, typename AT>
int launch_ln_bwd_staged(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, int M,
                         const float* addend, float* dx_f32, AT* dx_bf16, DropoutCfg drop, float* dgamma, float* dbeta,
                         float* colsum_out, cudaStream_t st, const float* out_row_scale, int rows_per_scale) {
  constexpr int D = NV * 128;
  constexpr int W = kStagedWarps, NROW = ADDEND ? 3 : 2, STAGES = ADDEND ? 3 : 4;
  constexpr size_t smem = (size_t)(D + W * STAGES * NROW * D) * sizeof(float) + W * STAGES * sizeof(uint64_t);
  static_assert(smem <= 227 * 1024, "staged LayerNorm backward: ring does not fit");
  auto kfn = ln_bwd_staged_kernel<NV, COLSUM, ADDEND, AT>;
  static bool attr = false;
  if (!attr) {
    ECAMP_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int need = (M + W - 1) / W;
  const int grid = need < num_sms_ln() ? need : num_sms_ln();
  ECAMP_CUDA_OK(launch_pdl(kfn, grid, W * 32, smem, st, dy, x, mean, rstd, gamma, M, addend, dx_f32, dx_bf16, drop, dgamma, dbeta, colsum_out, out_row_scale, rows_per_scale));
  ECAMP_LAUNCHED();
  return 0;
}

template <int NV, bool COLSUM, typename AT>
int launch_ln_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, int M,
                  const float* addend, float* dx_f32, AT* dx_bf16, DropoutCfg drop, float* dgamma, float* dbeta,
                  float* colsum_out, cudaStream_t st, const float* out_row_scale, int rows_per_scale) {
  constexpr int D = NV * 128;
  if (!use_slab()) {
    if (addend) return launch_ln_bwd_staged<NV, COLSUM, true, AT>(dy, x, mean, rstd, gamma, M, addend, dx_f32, dx_bf16, drop, dgamma, dbeta, colsum_out, st, out_row_scale, rows_per_scale);
    return launch_ln_bwd_staged<NV, COLSUM, false, AT>(dy, x, mean, rstd, gamma, M, addend, dx_f32, dx_bf16, drop, dgamma, dbeta, colsum_out, st, out_row_scale, rows_per_scale);
  }
  constexpr size_t smem = (size_t)(D + kBwdWarps * (COLSUM ? 3 : 2) * D) * sizeof(float);
  auto kfn = ln_bwd_kernel<NV, COLSUM, AT>;
  static bool attr = false;
  if (!attr) {
    ECAMP_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  ECAMP_CUDA_OK(launch_pdl(kfn, bwd_blocks(M), kBwdWarps * 32, smem, st, dy, x, mean, rstd, gamma, M, addend, dx_f32, dx_bf16, drop, dgamma, dbeta, colsum_out, out_row_scale, rows_per_scale));
  ECAMP_LAUNCHED();
  return 0;
}

}  // namespace

void layernorm_set_bwd_slab(int on) { g_ln_bwd_slab = on ? 1 : 0; }
size_t layernorm_bwd_ws_floats(int) { return 0; }  // kept for the C ABI: the backward no longer needs a workspace

template <typename AT>
int layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int M, int D, AT* out_bf16,
                  float* out_f32, float* mean, float* rstd, cudaStream_t st) {
  ECAMP_REQUIRE(D == 768 || D == 512, "layernorm: D must be 512 or 768 (got %d)", D);
  if (M <= 0) return 0;
  const int grid = (M + kRowsPerBlock - 1) / kRowsPerBlock;
  if (D == 768)
    ECAMP_CUDA_OK(launch_pdl(ln_fwd_kernel<6, AT>, grid, 256, 0, st, x, gamma, beta, eps, M, out_bf16, out_f32, mean, rstd));
  else
    ECAMP_CUDA_OK(launch_pdl(ln_fwd_kernel<4, AT>, grid, 256, 0, st, x, gamma, beta, eps, M, out_bf16, out_f32, mean, rstd));
  ECAMP_LAUNCHED();
  return 0;
}

template <typename AT>
int layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, int M,
                  int D, const float* addend, float* dx_f32, AT* dx_bf16, DropoutCfg drop, float* dgamma,
                  float* dbeta, float* colsum_out, int accumulate, cudaStream_t st, const float* out_row_scale,
                  int rows_per_scale) {
  ECAMP_REQUIRE(D == 768 || D == 512, "layernorm: D must be 512 or 768 (got %d)", D);
  ECAMP_REQUIRE(!colsum_out || dx_bf16, "layernorm_bwd: the column sum is taken over the bf16 output");
  if (M <= 0) return 0;
  if (!accumulate) {  // the kernel adds with atomics: start from zero
    if (dgamma) ECAMP_CUDA_OK(cudaMemsetAsync(dgamma, 0, (size_t)D * sizeof(float), st));
    if (dbeta) ECAMP_CUDA_OK(cudaMemsetAsync(dbeta, 0, (size_t)D * sizeof(float), st));
    if (colsum_out) ECAMP_CUDA_OK(cudaMemsetAsync(colsum_out, 0, (size_t)D * sizeof(float), st));
  }
  if (D == 768) {
    if (colsum_out) return launch_ln_bwd<6, true, AT>(dy, x, mean, rstd, gamma, M, addend, dx_f32, dx_bf16, drop, dgamma, dbeta, colsum_out, st, out_row_scale, rows_per_scale);
    return launch_ln_bwd<6, false, AT>(dy, x, mean, rstd, gamma, M, addend, dx_f32, dx_bf16, drop, dgamma, dbeta, nullptr, st, out_row_scale, rows_per_scale);
  }
  if (colsum_out) return launch_ln_bwd<4, true, AT>(dy, x, mean, rstd, gamma, M, addend, dx_f32, dx_bf16, drop, dgamma, dbeta, colsum_out, st, out_row_scale, rows_per_scale);
  return launch_ln_bwd<4, false, AT>(dy, x, mean, rstd, gamma, M, addend, dx_f32, dx_bf16, drop, dgamma, dbeta, nullptr, st, out_row_scale, rows_per_scale);
}

#define ECAMP_INST_LN(AT)                                                                                              \
  template int layernorm_fwd<AT>(const float*, const float*, const float*, float, int, int, AT*, float*, float*, float*, \
                                 cudaStream_t);                                                                          \
  template int layernorm_bwd<AT>(const float*, const float*, const float*, const float*, const float*, int, int,         \
                                 const float*, float*, AT*, DropoutCfg, float*, float*, float*, int, cudaStream_t,       \
                                 const float*, int);
ECAMP_INST_LN(bf16)
ECAMP_INST_LN(float)
#undef ECAMP_INST_LN

}  // namespace ecamp
