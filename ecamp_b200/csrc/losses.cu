// Loss kernels of the ECAMP step (all HBM / CUDA-core bound):
//   * masked-pixel MSE computed in patch space, no mask tensors materialised      (model_ecamp.py:196-215,276-300)
//   * super-resolution branch: bilinear x2 -> conv3x3 -> ReLU -> conv3x3 -> +skip -> ReLU -> windowed MSE,
//     one fused stencil kernel forward, one backward                                (model_ecamp.py:28-46,286,291-299)
//   * re-weighted cross-entropy over the 30 000-word vocabulary, per row chunk     (bert_modeling.py:211-217)
#include "kernels.cuh"

namespace ecamp {
namespace {

constexpr int IMG = 224, BIG = 448, GRID = 14, PATCH = 16, PD = 768;

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(192) mim_rows_kernel(const float* __restrict__ pred, int rows_per_batch,
                                                       const float* __restrict__ tgt, const float* __restrict__ mask,
                                                       int L, float* __restrict__ ws) {
  ECAMP_PDL_ENTRY();
  __shared__ float red[32];
  const int r = blockIdx.x, b = r / L, l = r % L;
  float s = 0.f;
  if (mask[r] != 0.f) {
    const float4 p = reinterpret_cast<const float4*>(pred + ((size_t)b * rows_per_batch + 1 + l) * PD)[threadIdx.x];
    const float4 t = reinterpret_cast<const float4*>(tgt + (size_t)r * PD)[threadIdx.x];
    const float a = p.x - t.x, c = p.y - t.y, d = p.z - t.z, e = p.w - t.w;
    s = a * a + c * c + d * d + e * e;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) ws[r] = s;
}

__global__ void __launch_bounds__(1024) sum_to_scalar_kernel(const float* __restrict__ x, size_t n, float scale,
                                                             float* __restrict__ out) {
  ECAMP_PDL_ENTRY();
  __shared__ float red[32];
  float s = 0.f;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) *out = s * scale;
}

// ---------------------------------------------------------------------------------------------
// super-resolution branch
// ---------------------------------------------------------------------------------------------
struct SrWeights {
  float w1[81], b1[3], w2[81], b2[3];
};
// The 168 conv parameters live in constant memory (copied device-to-device on the stream before each launch) so
// that every unrolled FMA takes its weight straight from the constant bank instead of a shared-memory load.
__constant__ SrWeights c_sr;

ECAMP_DEVINL float pred_pixel(const float* __restrict__ pred_b, int c, int y, int x) {
  // pred_b: [197, 768] of one sample, row 0 = cls; unpatchify 'nhwpqc->nchpwq' (model_ecamp.py:153-165)
  return pred_b[(size_t)(1 + (y >> 4) * GRID + (x >> 4)) * PD + (((y & 15) << 4) + (x & 15)) * 3 + c];
}
// F.interpolate(scale_factor=2, mode='bilinear', align_corners=False): source index and weight of output index Y
ECAMP_DEVINL void bilinear_src(int Y, int& y0, int& y1, float& lam) {
  const float src = fmaxf(0.f, (float)Y * 0.5f - 0.25f);
  y0 = (int)src;
  y1 = min(y0 + 1, IMG - 1);
  lam = src - (float)y0;
}

// 3x3 convolution over the 3 planes of a shared-memory region for a strip of 4 horizontally adjacent outputs:
// 6 shared loads feed 36 FMAs (weights come from the constant bank).
//   FLIP = false (forward):    acc[o][p] += w[o][i][ky][kx] * in[i][y0 + ky][x0 + p + kx]
//   FLIP = true  (transposed): acc[i][p] += w[o][i][ky][kx] * in[o][y0 + 2 - ky][x0 + p + 2 - kx]
template <int W, int PLANE, bool FLIP>
ECAMP_DEVINL void conv3_strip4(const float* in, int y0, int x0, const float* wt, float (&acc)[3][4]) {
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const float* r = in + a * PLANE + (y0 + (FLIP ? 2 - ky : ky)) * W + x0;
      float v[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) v[j] = r[j];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float w = FLIP ? wt[(a * 3 + c) * 9 + ky * 3 + kx] : wt[(c * 3 + a) * 9 + ky * 3 + kx];
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[c][q] += w * v[q + (FLIP ? 2 - kx : kx)];
        }
    }
}

// Forward of the SR head on an OUT x OUT output region whose top-left output pixel is (Y0, X0):
//   sU: (OUT+4)^2 x 3 up-sampled input, origin (Y0-2, X0-2), zero outside the image (conv zero padding)
//   sH: (OUT+2)^2 x 3 relu(conv1), origin (Y0-1, X0-1), zero outside the image
// (ny0, ny1, nx0, nx1): the pixels whose OUTPUT the caller will use (the loss window); u is only needed within 2 pixels
// and relu(conv1) within 1 pixel of them - everything else is written as zero without being computed, which is what
// makes the tiles that merely border the window cheap.
template <int OUT>
ECAMP_DEVINL void sr_forward_region(const float* __restrict__ pred_b, int Y0, int X0, float* sU, float* sH,
                                    int ny0 = -(1 << 20), int ny1 = 1 << 20, int nx0 = -(1 << 20), int nx1 = 1 << 20) {
  const SrWeights& w = c_sr;
  constexpr int UW = OUT + 4, HW = OUT + 2;
#pragma unroll 2  // two positions' 24 gathers in flight (the loop was exposed to the latency of its global loads)
  for (int i = threadIdx.x; i < UW * UW; i += blockDim.x) {
    const int uy = i / UW, ux = i % UW;
    const int Y = Y0 - 2 + uy, X = X0 - 2 + ux;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
    if (Y >= 0 && Y < BIG && X >= 0 && X < BIG && Y >= ny0 - 2 && Y < ny1 + 2 && X >= nx0 - 2 && X < nx1 + 2) {
      int y0, y1, x0, x1;
      float ly, lx;
      bilinear_src(Y, y0, y1, ly);
      bilinear_src(X, x0, x1, lx);
      const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
      v0 = w00 * pred_pixel(pred_b, 0, y0, x0) + w01 * pred_pixel(pred_b, 0, y0, x1) +
           w10 * pred_pixel(pred_b, 0, y1, x0) + w11 * pred_pixel(pred_b, 0, y1, x1);
      v1 = w00 * pred_pixel(pred_b, 1, y0, x0) + w01 * pred_pixel(pred_b, 1, y0, x1) +
           w10 * pred_pixel(pred_b, 1, y1, x0) + w11 * pred_pixel(pred_b, 1, y1, x1);
      v2 = w00 * pred_pixel(pred_b, 2, y0, x0) + w01 * pred_pixel(pred_b, 2, y0, x1) +
           w10 * pred_pixel(pred_b, 2, y1, x0) + w11 * pred_pixel(pred_b, 2, y1, x1);
    }
    sU[i] = v0;
    sU[UW * UW + i] = v1;
    sU[2 * UW * UW + i] = v2;
  }
  __syncthreads();
  constexpr int HS = (HW + 3) / 4;
  for (int t = threadIdx.x; t < HW * HS; t += blockDim.x) {
    const int hy = t / HS, hx0 = (t % HS) * 4;
    float acc[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[c][q] = w.b1[c];
    const int Y = Y0 - 1 + hy;
    const bool need = Y >= ny0 - 1 && Y < ny1 + 1 && X0 - 1 + hx0 + 3 >= nx0 - 1 && X0 - 1 + hx0 < nx1 + 1;
    if (need) conv3_strip4<UW, UW * UW, false>(sU, hy, hx0, w.w1, acc);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int hx = hx0 + q, X = X0 - 1 + hx;
      if (hx < HW) {
        const bool in_img = need && Y >= 0 && Y < BIG && X >= 0 && X < BIG;  // outside: conv2's zero padding
#pragma unroll
        for (int c = 0; c < 3; ++c) sH[c * HW * HW + hy * HW + hx] = in_img ? fmaxf(acc[c][q], 0.f) : 0.f;
      }
    }
  }
  __syncthreads();
}

// pre-activation outputs of the head for the strip (oy, ox0 .. ox0+3): conv2(h1) + b2 + u
template <int OUT>
ECAMP_DEVINL void sr_out_strip(const float* sU, const float* sH, int oy, int ox0, float (&o)[3][4]) {
  const SrWeights& w = c_sr;
  constexpr int UW = OUT + 4, HW = OUT + 2;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int q = 0; q < 4; ++q) o[c][q] = w.b2[c] + sU[c * UW * UW + (oy + 2) * UW + ox0 + q + 2];
  conv3_strip4<HW, HW * HW, false>(sH, oy, ox0, w.w2, o);
}

// window of sample b in 32-px tiles: rows [c0, c1), cols [r0, r1) (model_ecamp.py:207-208: slices clip at 14)
ECAMP_DEVINL void sr_window(const int64_t* column, const int64_t* row, int b, int& c0, int& c1, int& r0, int& r1) {
  const long long c = column[b], r = row[b];
  c0 = (int)max(0LL, min((long long)GRID, c));
  c1 = (int)max(0LL, min((long long)GRID, c + 12));
  r0 = (int)max(0LL, min((long long)GRID, r));
  r1 = (int)max(0LL, min((long long)GRID, r + 12));
}

__global__ void __launch_bounds__(256) sr_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ big,
                                                     const int64_t* __restrict__ column,
                                                     const int64_t* __restrict__ row, float* __restrict__ ws) {
  ECAMP_PDL_ENTRY();
  constexpr int OUT = 32, UW = OUT + 4, HW = OUT + 2;
  __shared__ float sU[3 * UW * UW + 8];  // +8: strips may read (and discard) a few floats past a row end
  __shared__ float sH[3 * HW * HW + 8];
  __shared__ float red[32];
  const int tile = blockIdx.x % (GRID * GRID), b = blockIdx.x / (GRID * GRID);
  const int ty = tile / GRID, tx = tile % GRID;
  int c0, c1, r0, r1;
  sr_window(column, row, b, c0, c1, r0, r1);
  if (ty < c0 || ty >= c1 || tx < r0 || tx >= r1) {
    if (threadIdx.x == 0) ws[blockIdx.x] = 0.f;
    return;
  }
  const float* pred_b = pred + (size_t)b * 197 * PD;
  const int Y0 = ty * 32, X0 = tx * 32;
  sr_forward_region<OUT>(pred_b, Y0, X0, sU, sH);
  float s = 0.f;
  for (int t = threadIdx.x; t < OUT * (OUT / 4); t += blockDim.x) {
    const int oy = t / (OUT / 4), ox0 = (t % (OUT / 4)) * 4;
    float o[3][4];
    sr_out_strip<OUT>(sU, sH, oy, ox0, o);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 tg = *reinterpret_cast<const float4*>(big + (((size_t)b * 3 + c) * BIG + Y0 + oy) * BIG + X0 + ox0);
      const float d0 = fmaxf(o[c][0], 0.f) - tg.x, d1 = fmaxf(o[c][1], 0.f) - tg.y;
      const float d2 = fmaxf(o[c][2], 0.f) - tg.z, d3 = fmaxf(o[c][3], 0.f) - tg.w;
      s += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) ws[blockIdx.x] = s;
}

// backward: one CTA per 32x32 tile of d_u.  Dynamic smem layout: U 40^2x3 | H 38^2x3 | dOut 36^2x3 | dH 34^2x3
constexpr int SR_BWD_SMEM_FLOATS = 3 * (40 * 40 + 38 * 38 + 36 * 36 + 34 * 34) + 8;
__global__ void __launch_bounds__(256, 3) sr_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ big,
                                                     const int64_t* __restrict__ column,
                                                     const int64_t* __restrict__ row, int B,
                                                     const float* __restrict__ g_res, float* __restrict__ d_u,
                                                     float* __restrict__ ws, float* __restrict__ loss_tiles, int skip) {
  ECAMP_PDL_ENTRY();
  constexpr int OUT = 36, UW = 40, HW = 38, OW = 36, GW = 34;
  extern __shared__ float sm[];
  float* sU = sm;
  float* sH = sU + 3 * UW * UW;
  float* sDO = sH + 3 * HW * HW;
  float* sDH = sDO + 3 * OW * OW;
  __shared__ float redw[8][84];
  const SrWeights& sw = c_sr;
  const int tile = blockIdx.x % (GRID * GRID), b = blockIdx.x / (GRID * GRID);
  const int ty = tile / GRID, tx = tile % GRID;
  const int Y0 = ty * 32, X0 = tx * 32;  // origin of the owned 32x32 region
  int c0, c1, r0, r1;
  sr_window(column, row, b, c0, c1, r0, r1);
  const int wy0 = c0 * 32, wy1 = c1 * 32, wx0 = r0 * 32, wx1 = r1 * 32;  // window in pixels
  float* ws_t = ws + (size_t)blockIdx.x * 168;
  // does the 36x36 d_out neighbourhood touch the window at all?
  const bool touches = (Y0 - 2 < wy1) && (Y0 + 34 > wy0) && (X0 - 2 < wx1) && (X0 + 34 > wx0);
  if (!touches) {
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
      const int oy = i >> 5, ox = i & 31;
#pragma unroll
      for (int c = 0; c < 3; ++c) d_u[(((size_t)b * 3 + c) * BIG + Y0 + oy) * BIG + X0 + ox] = 0.f;
    }
    for (int i = threadIdx.x; i < 168; i += blockDim.x) ws_t[i] = 0.f;
    if (loss_tiles && threadIdx.x == 0) loss_tiles[blockIdx.x] = 0.f;
    return;
  }
  const float* pred_b = pred + (size_t)b * 197 * PD;
  // work is limited to what the window can reach (skip == 0: the round-1 behaviour, every stage on the whole tile)
  const int ky0 = skip ? wy0 : -(1 << 20), ky1 = skip ? wy1 : 1 << 20, kx0 = skip ? wx0 : -(1 << 20), kx1 = skip ? wx1 : 1 << 20;
  sr_forward_region<OUT>(pred_b, Y0 - 2, X0 - 2, sU, sH, ky0, ky1, kx0, kx1);
  const float gscale = 2.0f * (*g_res) / ((float)B * 3.f * BIG * BIG);

  // d_out (pre-ReLU) on the 36x36 region with origin (Y0-2, X0-2); the squared error of the OWNED in-window pixels
  // is the tile's share of the loss (training steps skip sr_fwd_kernel and take the loss from here)
  float lsum = 0.f;
  for (int t = threadIdx.x; t < OW * (OW / 4); t += blockDim.x) {
    const int oy = t / (OW / 4), ox0 = (t % (OW / 4)) * 4;
    const int Y = Y0 - 2 + oy;
    // the strip starts at an even X and windows / image edges are even: each PAIR of pixels is inside or outside as a whole.
    // The targets are fetched BEFORE the convolution of the strip so that their latency hides behind its 324 FMAs (ncu:
    // a quarter of the kernel's stall samples sat on these loads when they were issued at the point of use).
    const bool row_in = Y >= wy0 && Y < wy1;
    bool in_win[2];
    float2 tg[2][3];
#pragma unroll
    for (int pr = 0; pr < 2; ++pr) {
      const int X = X0 - 2 + ox0 + 2 * pr;
      in_win[pr] = row_in && X >= wx0 && X < wx1;  // inside the window (hence inside the image)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        tg[pr][c] = in_win[pr] ? __ldg(reinterpret_cast<const float2*>(big + (((size_t)b * 3 + c) * BIG + Y) * BIG + X))
                               : make_float2(0.f, 0.f);
    }
    if (skip && !in_win[0] && !in_win[1]) {  // no pixel of the strip is in the window: d_out = 0, nothing to convolve
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        *reinterpret_cast<float2*>(sDO + c * OW * OW + oy * OW + ox0) = make_float2(0.f, 0.f);
        *reinterpret_cast<float2*>(sDO + c * OW * OW + oy * OW + ox0 + 2) = make_float2(0.f, 0.f);
      }
      continue;
    }
    float o[3][4];
    sr_out_strip<OUT>(sU, sH, oy, ox0, o);
#pragma unroll
    for (int pr = 0; pr < 2; ++pr) {
      const bool owned = oy >= 2 && oy < 34 && ox0 + 2 * pr >= 2 && ox0 + 2 * pr < 34;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float d0 = 0.f, d1 = 0.f;
        if (in_win[pr]) {
          const float e0 = fmaxf(o[c][2 * pr], 0.f) - tg[pr][c].x, e1 = fmaxf(o[c][2 * pr + 1], 0.f) - tg[pr][c].y;
          if (o[c][2 * pr] > 0.f) d0 = gscale * e0;
          if (o[c][2 * pr + 1] > 0.f) d1 = gscale * e1;
          if (owned) lsum += e0 * e0 + e1 * e1;
        }
        *reinterpret_cast<float2*>(sDO + c * OW * OW + oy * OW + ox0 + 2 * pr) = make_float2(d0, d1);
      }
    }
  }
  __syncthreads();
  // d_h1 (pre-ReLU) on the 34x34 region with origin (Y0-1, X0-1):
  //   dH(Yh, Xh)[ci] = relu'(h1) * sum w2[co][ci][ky][kx] * dOut(Yh - ky + 1, Xh - kx + 1)[co]; region coords of dOut: (gy + 2 - ky, gx + 2 - kx)
  constexpr int GS = (GW + 3) / 4;
  for (int t = threadIdx.x; t < GW * GS; t += blockDim.x) {
    const int gy = t / GS, gx0 = (t % GS) * 4;
    float d[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int q = 0; q < 4; ++q) d[c][q] = 0.f;
    const int Y = Y0 - 1 + gy;
    // d_h1 is zero further than one pixel from the window (d_out is zero outside it)
    const bool need = Y >= ky0 - 1 && Y < ky1 + 1 && X0 - 1 + gx0 + 3 >= kx0 - 1 && X0 - 1 + gx0 < kx1 + 1;
    if (need) conv3_strip4<OW, OW * OW, true>(sDO, gy, gx0, sw.w2, d);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int gx = gx0 + q, X = X0 - 1 + gx;
      if (gx < GW) {
        const bool in_img = need && Y >= 0 && Y < BIG && X >= 0 && X < BIG;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          sDH[c * GW * GW + gy * GW + gx] = (in_img && sH[c * HW * HW + (gy + 2) * HW + gx + 2] > 0.f) ? d[c][q] : 0.f;
      }
    }
  }
  __syncthreads();
  // d_u on the owned 32x32 region: skip path + conv1^T
  for (int t = threadIdx.x; t < 32 * 8; t += blockDim.x) {
    const int oy = t >> 3, ox0 = (t & 7) * 4;
    float d[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int q = 0; q < 4; ++q) d[c][q] = sDO[c * OW * OW + (oy + 2) * OW + ox0 + q + 2];
    // dH region coords of (Y - ky + 1, X - kx + 1) with (Y, X) = (Y0 + oy, X0 + ox): (oy + 2 - ky, ox + 2 - kx)
    // (d_u is zero further than two pixels from the window: d_out and d_h1 are zero there)
    if (Y0 + oy >= ky0 - 2 && Y0 + oy < ky1 + 2 && X0 + ox0 + 3 >= kx0 - 2 && X0 + ox0 < kx1 + 2)
      conv3_strip4<GW, GW * GW, true>(sDH, oy, ox0, sw.w1, d);
#pragma unroll
    for (int c = 0; c < 3; ++c)
      *reinterpret_cast<float4*>(d_u + (((size_t)b * 3 + c) * BIG + Y0 + oy) * BIG + X0 + ox0) =
          make_float4(d[c][0], d[c][1], d[c][2], d[c][3]);
  }
  // conv weight gradients over the OWNED 32x32 pixels only (each pixel is owned by exactly one tile): every thread
  // takes one strip of 4 horizontally adjacent pixels (6 shared loads feed the 36 products of a (ci, ky) row), then
  // the 84 per-thread sums are reduced across the warp with a halving butterfly (31 shuffles per 32 values instead
  // of 5 per value) and across the 8 warps through shared memory.
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // warps 0-3 take conv2, warps 4-7 conv1 (warp-uniform); every thread covers TWO strips (rows soy and soy + 16) before
  // the warp reduction, so a warp runs one butterfly per 2 x 324 products instead of one per 324
  const int which = warp >> 2;
  const int soy = (threadIdx.x & 127) >> 3, sox0 = (threadIdx.x & 7) * 4;
  {
    float acc[96];
#pragma unroll
    for (int k = 0; k < 96; ++k) acc[k] = 0.f;
    // which == 0: conv2: d_w2[co][ci][ky][kx] += dOut(Y, X)[co] * H(Y + ky - 1, X + kx - 1)[ci]
    // which == 1: conv1: d_w1[co][ci][ky][kx] += dH(Y, X)[co]   * U(Y + ky - 1, X + kx - 1)[ci]
    const float* gsrc = which == 0 ? sDO : sDH;
    const int gW = which == 0 ? OW : GW, goff = which == 0 ? 2 : 1;
    const float* isrc = which == 0 ? sH : sU;
    const int iW = which == 0 ? HW : UW, ioff = which == 0 ? 2 : 3;
    float g[2][3][4];  // the thread's two strips: rows soy and soy + 16
#pragma unroll
    for (int half = 0; half < 2; ++half)
#pragma unroll
      for (int co = 0; co < 3; ++co)
#pragma unroll
        for (int q = 0; q < 4; ++q) g[half][co][q] = gsrc[co * gW * gW + (soy + 16 * half + goff) * gW + sox0 + goff + q];
    bool nz = false;  // a tile that only borders the window has d_out = 0 on its own pixels and d_h1 != 0 on a 1-pixel band
#pragma unroll
    for (int half = 0; half < 2; ++half)
#pragma unroll
      for (int co = 0; co < 3; ++co)
#pragma unroll
        for (int q = 0; q < 4; ++q) nz = nz || g[half][co][q] != 0.f;
    const bool warp_has_work = __any_sync(0xffffffffu, nz || !skip);
#pragma unroll
    for (int co = 0; co < 3; ++co)
      acc[81 + co] = ((g[0][co][0] + g[0][co][1]) + (g[0][co][2] + g[0][co][3])) +
                     ((g[1][co][0] + g[1][co][1]) + (g[1][co][2] + g[1][co][3]));
    if (warp_has_work)
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const float* r = isrc + ci * iW * iW + (soy + ioff + ky) * iW + sox0 + ioff;
        float v[2][6];
#pragma unroll
        for (int j = 0; j < 6; ++j) { v[0][j] = r[j]; v[1][j] = r[16 * iW + j]; }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int co = 0; co < 3; ++co) {
            float s = g[0][co][0] * v[0][kx];
            s = fmaf(g[0][co][1], v[0][kx + 1], s);
            s = fmaf(g[0][co][2], v[0][kx + 2], s);
            s = fmaf(g[0][co][3], v[0][kx + 3], s);
            s = fmaf(g[1][co][0], v[1][kx], s);
            s = fmaf(g[1][co][1], v[1][kx + 1], s);
            s = fmaf(g[1][co][2], v[1][kx + 2], s);
            s = fmaf(g[1][co][3], v[1][kx + 3], s);
            acc[(co * 3 + ci) * 9 + ky * 3 + kx] = s;
          }
      }
    // butterfly: after the step with distance d a lane keeps the half of its values selected by (lane & d), so
    // that in the end lane l holds the warp total of value l (three groups of 32 values)
#pragma unroll
    for (int grp = 0; grp < 3; ++grp) {
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) {
        const bool up = (lane & d) != 0;
#pragma unroll
        for (int i = 0; i < d; ++i) {
          const float lo = acc[grp * 32 + i], hi = acc[grp * 32 + i + d];
          const float send = up ? lo : hi, keep = up ? hi : lo;
          acc[grp * 32 + i] = keep + __shfl_xor_sync(0xffffffffu, send, d);
        }
      }
      if (grp * 32 + lane < 84) redw[warp][grp * 32 + lane] = acc[grp * 32];
    }
    __syncthreads();
    if (threadIdx.x < 168) {
      const int cv = threadIdx.x / 84, idx = threadIdx.x % 84;
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) s += redw[cv * 4 + w][idx];
      // ws layout per tile: [w1 81 | b1 3 | w2 81 | b2 3]
      ws_t[(cv == 0 ? 84 : 0) + idx] = s;
    }
    __syncthreads();
  }
  if (loss_tiles) {
    lsum = block_sum(lsum, &redw[0][0]);
    if (threadIdx.x == 0) loss_tiles[blockIdx.x] = lsum;
  }
}

__global__ void __launch_bounds__(256) sr_wgrad_finalize_kernel(const float* __restrict__ ws, int ntiles,
                                                                float* __restrict__ d_conv, int accumulate) {
  ECAMP_PDL_ENTRY();
  __shared__ float red[32];
  const int k = blockIdx.x;  // 0..167
  float s = 0.f;
  for (int t = threadIdx.x; t < ntiles; t += blockDim.x) s += ws[(size_t)t * 168 + k];
  s = block_sum(s, red);
  if (threadIdx.x == 0) d_conv[k] = accumulate ? d_conv[k] + s : s;
}

// ---------------------------------------------------------------------------------------------
// d_pred (bf16, the dY operand of decoder_pred's dgrad / wgrad)
// ---------------------------------------------------------------------------------------------
// weights of the transposed x2 bilinear up-sampling (align_corners = False) for source index y: the taps are the
// up-sampled rows 2y-1, 2y, 2y+1, 2y+2 with weights .25 .75 .75 .25; at the image border the up-sampling clamps its
// source index, which moves the missing tap's weight onto the border pixel itself
ECAMP_DEVINL void bilinear_t_weights(int y, float (&w)[4]) {
  w[0] = 0.25f; w[1] = 0.75f; w[2] = 0.75f; w[3] = 0.25f;
  if (y == 0) { w[0] = 0.f; w[1] = 1.0f; }
  if (y == IMG - 1) { w[2] = 1.0f; w[3] = 0.f; }
}
template <typename AT>
__global__ void __launch_bounds__(256) pred_grad_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                                        const float* __restrict__ mask, const float* __restrict__ d_u,
                                                        const float* __restrict__ g_mim, int B,
                                                        AT* __restrict__ d_pred) {
  ECAMP_PDL_ENTRY();
  constexpr int RW = 34;   // rows: 16 source pixels x 2 + one halo pixel on each side, in the up-sampled grid
  constexpr int RWP = 40;  // columns staged: the 16-byte aligned window [32 wx - 4, 32 wx + 36) that contains the 34 needed ones
  __shared__ __align__(16) float sdu[3 * RW * RWP];
  const int r = blockIdx.x, b = r / 197, t = r % 197;
  AT* out = d_pred + (size_t)r * PD;
  if (t == 0) {
    for (int e = threadIdx.x; e < PD; e += blockDim.x) act_st(out + e, 0.f);
    return;
  }
  const int l = t - 1, hy = l / GRID, wx = l % GRID;
  const float m = mask[(size_t)b * 196 + l];
  const float gm = m != 0.f ? 2.0f * (*g_mim) / ((float)B * 3.f * IMG * IMG) : 0.f;
  if (d_u) {
    // stage the (34 rows x 40 columns) x 3 neighbourhood of this patch of d_u with 128-bit loads (zero outside the image: those
    // taps have weight 0); the image width and the window origin are multiples of 4, so a float4 is inside or outside as a whole
    // (the scalar version issued 3468 dependent-index loads per block and ran at 1.8 TB/s)
    const int Y0 = 2 * hy * PATCH - 1, X0a = 2 * wx * PATCH - 4;
    for (int i = threadIdx.x; i < 3 * RW * (RWP / 4); i += blockDim.x) {
      const int c = i / (RW * (RWP / 4)), rem = i % (RW * (RWP / 4)), yy = rem / (RWP / 4), x4 = rem % (RWP / 4);
      const int Y = Y0 + yy, X = X0a + 4 * x4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (Y >= 0 && Y < BIG && X >= 0 && X < BIG) v = *reinterpret_cast<const float4*>(d_u + (((size_t)b * 3 + c) * BIG + Y) * BIG + X);
      *reinterpret_cast<float4*>(sdu + (c * RW + yy) * RWP + 4 * x4) = v;
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < PD; e += blockDim.x) {
    float g = 0.f;
    if (gm != 0.f) g = gm * (pred[(size_t)r * PD + e] - tgt[((size_t)b * 196 + l) * PD + e]);
    if (d_u) {
      const int c = e % 3, pq = e / 3, p = pq >> 4, q = pq & 15;
      float wy[4], wxx[4];
      bilinear_t_weights(hy * PATCH + p, wy);
      bilinear_t_weights(wx * PATCH + q, wxx);
      const float* s0 = sdu + (c * RW + 2 * p) * RWP + 2 * q + 3;  // local row of up-sampled row 2y-1, column 2x-1
      float acc = 0.f;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float* sr = s0 + a * RWP;
        acc = fmaf(wy[a], fmaf(wxx[0], sr[0], fmaf(wxx[1], sr[1], fmaf(wxx[2], sr[2], wxx[3] * sr[3]))), acc);
      }
      g += acc;
    }
    act_st(out + e, g);
  }
}

// ---------------------------------------------------------------------------------------------
// cross-entropy over one chunk of rows; the row is staged in shared memory so logits are read once
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ce_rows_kernel(bf16* __restrict__ logits, int ldl, int V,
                                                      const int64_t* __restrict__ labels,
                                                      const float* __restrict__ weights, float* __restrict__ row_loss,
                                                      const float* __restrict__ g_mlm, float inv_total,
                                                      int write_grad) {
  ECAMP_PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t ce_smem[];
  bf16* srow = reinterpret_cast<bf16*>(ce_smem);
  __shared__ float red[32];
  const int r = blockIdx.x;
  bf16* grow = logits + (size_t)r * ldl;
  const int nv = V / 8;  // V % 8 == 0 checked on the host
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const uint4 u = reinterpret_cast<const uint4*>(grow)[i];
    reinterpret_cast<uint4*>(srow)[i] = u;
    float2 f;
    f = unpack_bf16x2(u.x); mx = fmaxf(mx, fmaxf(f.x, f.y));
    f = unpack_bf16x2(u.y); mx = fmaxf(mx, fmaxf(f.x, f.y));
    f = unpack_bf16x2(u.z); mx = fmaxf(mx, fmaxf(f.x, f.y));
    f = unpack_bf16x2(u.w); mx = fmaxf(mx, fmaxf(f.x, f.y));
  }
  // block max
  mx = warp_max(mx);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  // e_j = exp(z_j - max) is evaluated ONCE: summed in fp32 and kept (bf16, over the staged logit) for the gradient
  // pass, which then is a scale and a pack (the kernel was bound by the two exps per logit; the gradient is stored
  // in bf16 anyway)
  const long long label = labels[r];
  const bool valid = label >= 0 && label < V;  // CrossEntropyLoss ignore_index (-100) -> no loss, no gradient
  const float z_label = valid ? bf2f(srow[label]) : 0.f;
  __syncthreads();  // every thread has read its label logit before the row is overwritten
  float se = 0.f;
  const float mx2 = mx * 1.4426950408889634f;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const uint4 u = reinterpret_cast<const uint4*>(srow)[i];
    float v[8];
    float2 f;
    f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
    f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
    f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
    f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[k] = ex2_approx(fmaf(v[k], 1.4426950408889634f, -mx2));
      se += v[k];
    }
    if (write_grad) {
      uint4 o;
      o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
      o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
      reinterpret_cast<uint4*>(srow)[i] = o;
    }
  }
  se = block_sum(se, red);
  const float lse = mx + __logf(se);
  const float w = weights[r];
  if (threadIdx.x == 0) row_loss[r] = valid ? (lse - z_label) * w : 0.f;
  if (!write_grad) return;
  const float coef = valid ? w * (*g_mlm) * inv_total : 0.f;
  const float pscale = coef / se;  // softmax_j * coef = e_j * coef / sum
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const uint4 u = reinterpret_cast<const uint4*>(srow)[i];
    float v[8];
    float2 f;
    f = unpack_bf16x2(u.x); v[0] = f.x * pscale; v[1] = f.y * pscale;
    f = unpack_bf16x2(u.y); v[2] = f.x * pscale; v[3] = f.y * pscale;
    f = unpack_bf16x2(u.z); v[4] = f.x * pscale; v[5] = f.y * pscale;
    f = unpack_bf16x2(u.w); v[6] = f.x * pscale; v[7] = f.y * pscale;
    if ((long long)i == (label >> 3) && valid) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k == (int)(label & 7)) v[k] -= coef;
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    reinterpret_cast<uint4*>(grow)[i] = o;
  }
}

// Persistent form with the bias gradient folded in (default for the training step).  One CTA per SM walks rows
// r, r + grid, ...; the NEXT two rows are in flight as bulk asynchronous copies into a two-slot shared-memory ring while
// the current one is reduced, so the loads no longer wait behind the three block-wide phases of a row.  Every thread
// owns the same columns in every row (uint4 groups tid, tid + 512, ...), so the column sums of the gradient it writes
// (= the gradient of cls.predictions.bias) stay in registers for the whole launch and leave as one vector atomic per
// four columns and CTA: the separate column-sum pass re-read all V x rows gradients (245 MB per 4096-row chunk, 0.39 ms
// per step).  The sums are taken over the fp32 values before they are rounded to bf16 for the GEMMs.
int g_ce_fused = -1;  // 1 (default): persistent kernel with the bias gradient folded in; 0: one CTA per row + column-sum pass
constexpr int kCeThreads = 512;
constexpr int kCeGroups = 8;  // uint4 groups per thread: V <= 8 * 512 * 8 = 32768

ECAMP_DEVINL void ce_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(kCeThreads, 1)
    ce_rows_fused_kernel(bf16* __restrict__ logits, int ldl, int V, int rows, const int64_t* __restrict__ labels,
                         const float* __restrict__ weights, float* __restrict__ row_loss, const float* __restrict__ g_mlm,
                         float inv_total, float* __restrict__ colsum) {
  extern __shared__ __align__(128) uint8_t ce_smem[];
  __shared__ __align__(8) uint64_t full[2];
  __shared__ float red[32];
  __shared__ float s_zlabel;
  const int nv = V / 8;
  const uint32_t row_bytes = (uint32_t)V * 2u;
  const uint32_t slot_bytes = (row_bytes + 127u) & ~127u;
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  __syncthreads();
  ECAMP_PDL_ENTRY();
  const int stride = gridDim.x;
  auto issue = [&](int slot, int r) {  // thread 0 only
    mbar_arrive_expect_tx(&full[slot], row_bytes);
    ce_bulk_load(ce_smem + (size_t)slot * slot_bytes, logits + (size_t)r * ldl, row_bytes, &full[slot]);
  };
  if (threadIdx.x == 0) {
    if ((int)blockIdx.x < rows) issue(0, blockIdx.x);
    if ((int)blockIdx.x + stride < rows) issue(1, blockIdx.x + stride);
  }
  float cs[kCeGroups][8];
#pragma unroll
  for (int j = 0; j < kCeGroups; ++j)
#pragma unroll
    for (int k = 0; k < 8; ++k) cs[j][k] = 0.f;
  const float gm = *g_mlm;
  int it = 0;
  for (int r = blockIdx.x; r < rows; r += stride, ++it) {
    const int slot = it & 1;
    mbar_wait(&full[slot], (uint32_t)(it >> 1) & 1u);
    uint4* srow = reinterpret_cast<uint4*>(ce_smem + (size_t)slot * slot_bytes);
    const long long label = labels[r];
    const bool valid = label >= 0 && label < V;  // CrossEntropyLoss ignore_index (-100) -> no loss, no gradient
    const int lgroup = valid ? (int)(label >> 3) : -1, lsub = (int)(label & 7);
    const float w = weights[r];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kCeGroups; ++j) {
      const int i = threadIdx.x + j * kCeThreads;
      if (i < nv) {
        const uint4 u = srow[i];
        float2 f;
        f = unpack_bf16x2(u.x); mx = fmaxf(mx, fmaxf(f.x, f.y));
        f = unpack_bf16x2(u.y); mx = fmaxf(mx, fmaxf(f.x, f.y));
        f = unpack_bf16x2(u.z); mx = fmaxf(mx, fmaxf(f.x, f.y));
        f = unpack_bf16x2(u.w); mx = fmaxf(mx, fmaxf(f.x, f.y));
      }
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int wi = 1; wi < kCeThreads / 32; ++wi) mx = fmaxf(mx, red[wi]);
    __syncthreads();  // red is reused by the sum below
    // e_j = exp(z_j - max), once: summed in fp32 and kept (bf16, over the staged logit) for the gradient pass
    const float mx2 = mx * 1.4426950408889634f;
    float se = 0.f;
#pragma unroll
    for (int j = 0; j < kCeGroups; ++j) {
      const int i = threadIdx.x + j * kCeThreads;
      if (i < nv) {
        const uint4 u = srow[i];
        float v[8];
        float2 f;
        f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
        f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
        f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
        f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
        if (i == lgroup) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k == lsub) s_zlabel = v[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v[k] = ex2_approx(fmaf(v[k], 1.4426950408889634f, -mx2));
          se += v[k];
        }
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
        o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
        srow[i] = o;
      }
    }
    se = block_sum(se, red);
    if (threadIdx.x == 0) row_loss[r] = valid ? (mx + __logf(se) - s_zlabel) * w : 0.f;
    const float coef = valid ? w * gm * inv_total : 0.f;
    const float pscale = coef / se;  // softmax_j * coef = e_j * coef / sum
    uint4* grow = reinterpret_cast<uint4*>(logits + (size_t)r * ldl);
#pragma unroll
    for (int j = 0; j < kCeGroups; ++j) {
      const int i = threadIdx.x + j * kCeThreads;
      if (i < nv) {
        const uint4 u = srow[i];
        float v[8];
        float2 f;
        f = unpack_bf16x2(u.x); v[0] = f.x * pscale; v[1] = f.y * pscale;
        f = unpack_bf16x2(u.y); v[2] = f.x * pscale; v[3] = f.y * pscale;
        f = unpack_bf16x2(u.z); v[4] = f.x * pscale; v[5] = f.y * pscale;
        f = unpack_bf16x2(u.w); v[6] = f.x * pscale; v[7] = f.y * pscale;
        if (i == lgroup) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k == lsub) v[k] -= coef;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) cs[j][k] += v[k];
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
        o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
        grow[i] = o;
      }
    }
    __syncthreads();  // every thread is done with this slot (and with s_zlabel / red) before the slot is refilled
    if (threadIdx.x == 0 && r + 2 * stride < rows) issue(slot, r + 2 * stride);
  }
  if (colsum) {
    const bool vec = (reinterpret_cast<uintptr_t>(colsum) & 15) == 0;  // the flat gradient buffer packs tensors without padding
#pragma unroll
    for (int j = 0; j < kCeGroups; ++j) {
      const int i = threadIdx.x + j * kCeThreads;
      if (i < nv) {
        if (vec) {
          atomicAdd(reinterpret_cast<float4*>(colsum + 8 * (size_t)i), make_float4(cs[j][0], cs[j][1], cs[j][2], cs[j][3]));
          atomicAdd(reinterpret_cast<float4*>(colsum + 8 * (size_t)i + 4), make_float4(cs[j][4], cs[j][5], cs[j][6], cs[j][7]));
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) atomicAdd(colsum + 8 * (size_t)i + k, cs[j][k]);
        }
      }
    }
  }
}

// fp32-accurate mode: the same cross-entropy on fp32 logits (three passes over the row in global / L2: max, sum of
// exponentials with expf / logf, gradient written over the logits)
__global__ void __launch_bounds__(256) ce_rows_f32_kernel(float* __restrict__ logits, int ldl, int V,
                                                          const int64_t* __restrict__ labels,
                                                          const float* __restrict__ weights, float* __restrict__ row_loss,
                                                          const float* __restrict__ g_mlm, float inv_total,
                                                          int write_grad) {
  ECAMP_PDL_ENTRY();
  __shared__ float red[32];
  const int r = blockIdx.x;
  float* grow = logits + (size_t)r * ldl;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < V; i += blockDim.x) mx = fmaxf(mx, grow[i]);
  mx = warp_max(mx);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  const long long label = labels[r];
  const bool valid = label >= 0 && label < V;
  const float z_label = valid ? grow[label] : 0.f;
  float se = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) se += expf(grow[i] - mx);
  se = block_sum(se, red);
  const float lse = mx + logf(se);
  const float w = weights[r];
  if (threadIdx.x == 0) row_loss[r] = valid ? (lse - z_label) * w : 0.f;
  if (!write_grad) return;
  __syncthreads();  // every thread has read the label logit before the row is overwritten
  const float coef = valid ? w * (*g_mlm) * inv_total : 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    float g = expf(grow[i] - lse) * coef;
    if (valid && i == (int)label) g -= coef;
    grow[i] = g;
  }
}

}  // namespace

#define LAUNCH_OK() ECAMP_LAUNCHED()

int sum_to_scalar(const float* x, size_t n, float scale, float* out, cudaStream_t st) {
  ECAMP_CUDA_OK(launch_pdl(sum_to_scalar_kernel, 1, 1024, 0, st, x, n, scale, out));
  LAUNCH_OK();
  return 0;
}

int mim_loss_fwd(const float* pred, int rows_per_batch, const float* tgt, const float* mask, int B, int L, int pd,
                 float* loss_out, float* ws, cudaStream_t st) {
  ECAMP_REQUIRE(pd == PD && L == 196, "mim loss: only 196 patches x 768 supported");
  ECAMP_CUDA_OK(launch_pdl(mim_rows_kernel, B * L, 192, 0, st, pred, rows_per_batch, tgt, mask, L, ws));
  LAUNCH_OK();
  return sum_to_scalar(ws, (size_t)B * L, 1.0f / ((float)B * 3.f * IMG * IMG), loss_out, st);
}

size_t sr_ws_floats(int B) { return (size_t)B * GRID * GRID * 168; }

static int upload_sr_weights(const float* w1, const float* b1, const float* w2, const float* b2, cudaStream_t st) {
  static SrWeights* dev = nullptr;
  if (!dev) ECAMP_CUDA_OK(cudaGetSymbolAddress(reinterpret_cast<void**>(&dev), c_sr));
  ECAMP_CUDA_OK(cudaMemcpyAsync(dev->w1, w1, 81 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  ECAMP_CUDA_OK(cudaMemcpyAsync(dev->b1, b1, 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  ECAMP_CUDA_OK(cudaMemcpyAsync(dev->w2, w2, 81 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  ECAMP_CUDA_OK(cudaMemcpyAsync(dev->b2, b2, 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int sr_loss_fwd(const float* pred, const float* big, const int64_t* column, const int64_t* row, const float* w1,
                const float* b1, const float* w2, const float* b2, int B, float* loss_out, float* ws,
                cudaStream_t st) {
  if (int rc = upload_sr_weights(w1, b1, w2, b2, st)) return rc;
  ECAMP_CUDA_OK(launch_pdl(sr_fwd_kernel, B * GRID * GRID, 256, 0, st, pred, big, column, row, ws));
  LAUNCH_OK();
  return sum_to_scalar(ws, (size_t)B * GRID * GRID, 1.0f / ((float)B * 3.f * BIG * BIG), loss_out, st);
}

int g_sr_window_skip = 1;  // measurement switch (ecamp_sr_set_window_skip): 0 = every stage on the whole tile
void sr_set_window_skip(int on) { g_sr_window_skip = on ? 1 : 0; }

int sr_loss_bwd(const float* pred, const float* big, const int64_t* column, const int64_t* row, const float* w1,
                const float* b1, const float* w2, const float* b2, int B, const float* g_res, float* d_u,
                float* d_conv, int accumulate, float* ws, cudaStream_t st, float* loss_tiles, float* loss_out) {
  static bool attr = false;
  if (!attr) {
    ECAMP_CUDA_OK(cudaFuncSetAttribute(sr_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       SR_BWD_SMEM_FLOATS * (int)sizeof(float)));
    attr = true;
  }
  if (int rc = upload_sr_weights(w1, b1, w2, b2, st)) return rc;
  ECAMP_CUDA_OK(launch_pdl(sr_bwd_kernel, B * GRID * GRID, 256, SR_BWD_SMEM_FLOATS * sizeof(float), st, pred, big, column, row, B, g_res,
                                                                                   d_u, ws, loss_tiles, g_sr_window_skip));
  LAUNCH_OK();
  ECAMP_CUDA_OK(launch_pdl(sr_wgrad_finalize_kernel, 168, 256, 0, st, ws, B * GRID * GRID, d_conv, accumulate));
  LAUNCH_OK();
  if (loss_tiles && loss_out)
    return sum_to_scalar(loss_tiles, (size_t)B * GRID * GRID, 1.0f / ((float)B * 3.f * BIG * BIG), loss_out, st);
  return 0;
}

template <typename AT>
int pred_grad(const float* pred, const float* tgt, const float* mask, const float* d_u, const float* g_mim, int B,
              AT* d_pred, cudaStream_t st) {
  ECAMP_CUDA_OK(launch_pdl(pred_grad_kernel<AT>, B * 197, 256, 0, st, pred, tgt, mask, d_u, g_mim, B, d_pred));
  LAUNCH_OK();
  return 0;
}

template int pred_grad<bf16>(const float*, const float*, const float*, const float*, const float*, int, bf16*, cudaStream_t);
template int pred_grad<float>(const float*, const float*, const float*, const float*, const float*, int, float*, cudaStream_t);

int ce_chunk(float* logits, int ldl, int rows, int V, const int64_t* labels, const float* weights, float* row_loss,
             const float* g_mlm, float inv_total, int write_grad, cudaStream_t st) {
  if (rows <= 0) return 0;
  ECAMP_CUDA_OK(launch_pdl(ce_rows_f32_kernel, rows, 256, 0, st, logits, ldl, V, labels, weights, row_loss, g_mlm, inv_total,
                           write_grad));
  LAUNCH_OK();
  return 0;
}

void ce_set_fused(int on) { g_ce_fused = on ? 1 : 0; }

int ce_chunk(bf16* logits, int ldl, int rows, int V, const int64_t* labels, const float* weights, float* row_loss,
             const float* g_mlm, float inv_total, int write_grad, cudaStream_t st, float* bias_grad) {
  ECAMP_REQUIRE(V % 8 == 0 && ldl % 8 == 0, "cross-entropy: vocabulary / pitch must be multiples of 8");
  ECAMP_REQUIRE((size_t)V * 2 <= 200 * 1024, "cross-entropy: vocabulary row does not fit in shared memory");
  ECAMP_REQUIRE(!bias_grad || write_grad, "cross-entropy: the bias gradient needs the gradient pass");
  if (rows <= 0) return 0;
  if (g_ce_fused < 0) g_ce_fused = getenv("ECAMP_CE_FUSED") ? atoi(getenv("ECAMP_CE_FUSED")) : 1;
  const int fused_on = g_ce_fused;
  if (write_grad && fused_on && V <= kCeGroups * kCeThreads * 8) {
    const size_t slot = ((size_t)V * 2 + 127) & ~(size_t)127;
    static bool attr2 = false;
    if (!attr2) {
      ECAMP_CUDA_OK(cudaFuncSetAttribute(ce_rows_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr2 = true;
    }
    static int sms = 0;
    if (sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (sms <= 0) sms = 148;
    }
    ECAMP_CUDA_OK(launch_pdl(ce_rows_fused_kernel, rows < sms ? rows : sms, kCeThreads, 2 * slot, st, logits, ldl, V, rows, labels, weights,
                             row_loss, g_mlm, inv_total, bias_grad));
    LAUNCH_OK();
    return 0;
  }
  static bool attr = false;
  if (!attr) {
    ECAMP_CUDA_OK(cudaFuncSetAttribute(ce_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  ECAMP_CUDA_OK(launch_pdl(ce_rows_kernel, rows, 256, (size_t)V * 2, st, logits, ldl, V, labels, weights, row_loss, g_mlm, inv_total,
                                                   write_grad));
  LAUNCH_OK();
  if (bias_grad) return colsum_bf16(logits, ldl, rows, V, bias_grad, 1, nullptr, st);
  return 0;
}

}  // namespace ecamp
