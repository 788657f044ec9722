// tcgen05 / TMEM / TMA attention for head_dim 64 and 128 (BERT self-attention, text->image cross-attention,
// ViT encoder): one CTA per (batch, head, 128-query tile).  Q K^T and dO V^T accumulate in TMEM; one thread owns
// one query row (= one TMEM lane), so softmax, its backward and delta_i = sum_j P_ij dP_ij need no cross-thread
// reduction; P / dS are written back to shared memory as bf16 in the 128-byte-swizzled layout that tcgen05.mma
// reads, and the second set of MMAs (P V; dS K, P^T dO, dS^T Q) accumulates again in TMEM.  Every shared-memory
// tile is loaded once by TMA and serves two MMAs through two views: a K-major tile X[rows, d] is at the same time
// the MN-major operand with `rows` as the contraction index (K and V in forward / backward, Q and dO in backward).
//
// Shapes: forward  any Sq (tiles of 128), Sk <= 256;  backward Sq <= 128 and Sk <= 128 (one tile per head).
// Everything else (head_dim 32, longer sequences) stays on the mma.sync kernels of attention.cu.
#include <cuda.h>

#include "gemm.cuh"
#include "kernels.cuh"

namespace ecamp {
namespace {

constexpr int kRows = 128;            // query rows per CTA = TMEM lanes
constexpr int kAtomBytes = 128 * 128;  // one 64-column atom of a 128-row K-major tile

// store 8 consecutive bf16 (columns c8*8 .. c8*8+7 of row r) into a K-major, 128B-swizzled [128 x ncols] tile
ECAMP_DEVINL void st_swizzled8(uint8_t* tile, int r, int c8, const float (&v)[8]) {
  const int atom = c8 >> 3, chunk = c8 & 7;
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(tile + atom * kAtomBytes + r * 128 + ((chunk ^ (r & 7)) << 4)) = u;
}

// Coalesced store of a [32 rows x NC columns] bf16 block that a warp holds ONE ROW PER THREAD (the layout tcgen05.ld
// delivers): written that way every store instruction touches 32 different lines; instead the block goes through a
// swizzled shared-memory tile and leaves as whole row segments (NC = 64: 4 rows x 128 B per instruction).
// `packed`: the thread's NC bf16 as NC / 2 words; `g0`: global address of (row 0 of the warp's 32 rows, column 0 of
// the block); rows >= rows_valid are not written.  `stage`: 32 * NC * 2 bytes private to the warp.
// `colsum` (optional): += the column sums of the valid rows (fp32 atomics; 8 columns per lane group).
template <int NC>
ECAMP_DEVINL void store_rows_coalesced(uint8_t* stage, const uint32_t (&packed)[NC / 2], int lane, bf16* g0, size_t ld,
                                       int rows_valid, float* colsum = nullptr) {
  constexpr int NU = NC / 8;       // 16-byte units per row (8 or 4)
  constexpr int RB = NC * 2;       // row bytes (128 or 64)
  constexpr int RPS = 32 / NU;     // rows per step of the coalesced read-back
  const uint32_t sbase = smem_u32(stage);
  const int wswz = NC == 64 ? (lane & 7) : ((lane >> 1) & 3);
#pragma unroll
  for (int j = 0; j < NU; ++j)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + (uint32_t)(lane * RB + ((j ^ wswz) << 4))),
                 "r"(packed[4 * j]), "r"(packed[4 * j + 1]), "r"(packed[4 * j + 2]), "r"(packed[4 * j + 3])
                 : "memory");
  __syncwarp();
  const int sub = lane / NU, u = lane % NU;
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < NU; ++i) {
    const int r = RPS * i + sub;
    const int rswz = NC == 64 ? (r & 7) : ((r >> 1) & 3);
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(sbase + (uint32_t)(r * RB + ((u ^ rswz) << 4)))
                 : "memory");
    if (r < rows_valid) {
      *reinterpret_cast<uint4*>(g0 + (size_t)r * ld + u * 8) = v;
      float2 f;
      f = unpack_bf16x2(v.x); cs[0] += f.x; cs[1] += f.y;
      f = unpack_bf16x2(v.y); cs[2] += f.x; cs[3] += f.y;
      f = unpack_bf16x2(v.z); cs[4] += f.x; cs[5] += f.y;
      f = unpack_bf16x2(v.w); cs[6] += f.x; cs[7] += f.y;
    }
  }
  if (colsum) {  // lanes that differ only in `sub` hold the same 8 columns
#pragma unroll
    for (int o = NU; o < 32; o <<= 1)
#pragma unroll
      for (int k = 0; k < 8; ++k) cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], o);
    if (sub == 0) {
      float* dst = colsum + u * 8;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(cs[0]), "f"(cs[1]), "f"(cs[2]), "f"(cs[3]) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(cs[4]), "f"(cs[5]), "f"(cs[6]), "f"(cs[7]) : "memory");
    }
  }
  __syncwarp();
}

struct TcMaps {
  CUtensorMap q, k, v, d_o;
};

// =============================================================================================
// forward
// =============================================================================================
template <int D>
__global__ void __launch_bounds__(160) attn_tc_fwd_kernel(const __grid_constant__ TcMaps maps, AttnArgs a, int tmem_cols) {
  ECAMP_PDL_ENTRY();
  constexpr int ATOMS = D / 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int Skp = (a.Sk + 15) & ~15;
  uint8_t* sQ = smem;                               // ATOMS x [128 x 128 B]
  uint8_t* sK = sQ + ATOMS * kAtomBytes;            // ATOMS x [Skp x 128 B]
  uint8_t* sV = sK + ATOMS * Skp * 128;             // ATOMS x [Skp x 128 B]
  uint8_t* sP = smem;                               // aliases Q | K once S = Q K^T has completed
  float* sBias = reinterpret_cast<float*>(sV + ATOMS * Skp * 128);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 256);
  uint64_t *bar_qk = bars, *bar_v = bars + 1, *bar_s = bars + 2, *bar_p = bars + 3, *bar_o = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kRows, h = blockIdx.y, b = blockIdx.z;
  const uint64_t bh = (uint64_t)b * a.H + h;

  if (tid == 0) {
    mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 128); mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_slot, (uint32_t)tmem_cols);
    tmem_relinquish();
  }
  for (int j = tid; j < 256; j += blockDim.x) {
    bool ok = j < a.Sk;
    if (ok && a.key_mask) ok = a.key_mask[(size_t)b * a.Sk + j] != 0;
    sBias[j] = ok ? 0.f : -INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tO = tmem + (uint32_t)((Skp + 31) & ~31);

  if (warp == 4) {
    if (lane == 0) {
      // ---- TMA: Q and K behind one barrier, V behind another ----
      mbar_arrive_expect_tx(bar_qk, (uint32_t)(ATOMS * kAtomBytes + ATOMS * Skp * 128));
#pragma unroll
      for (int at = 0; at < ATOMS; ++at) {
        tma_load_2d(sQ + at * kAtomBytes, &maps.q, bar_qk, h * D + at * 64, b * a.Sq + q0);
        tma_load_2d(sK + at * Skp * 128, &maps.k, bar_qk, h * D + at * 64, b * a.Sk);
      }
      mbar_arrive_expect_tx(bar_v, (uint32_t)(ATOMS * Skp * 128));
#pragma unroll
      for (int at = 0; at < ATOMS; ++at) tma_load_2d(sV + at * Skp * 128, &maps.v, bar_v, h * D + at * 64, b * a.Sk);
      // ---- S = Q K^T ----
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      const uint32_t idesc_s = umma_idesc_bf16(kRows, Skp, false, false);
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks) {
        const uint64_t ad = umma_smem_desc_sw128(smem_u32(sQ + (ks >> 2) * kAtomBytes) + (ks & 3) * 32, 16, 1024);
        const uint64_t bd = umma_smem_desc_sw128(smem_u32(sK + (ks >> 2) * Skp * 128) + (ks & 3) * 32, 16, 1024);
        umma_f16(tS, ad, bd, idesc_s, ks > 0 ? 1u : 0u);
      }
      umma_commit(bar_s);
      // ---- O = P V (P written by the softmax warps; V is read MN-major: contraction over its rows) ----
      mbar_wait(bar_v, 0);
      mbar_wait(bar_p, 0);
      tc_fence_after();
      const uint32_t idesc_o = umma_idesc_bf16(kRows, D, false, true);
      for (int ks = 0; ks < Skp / 16; ++ks) {
        const uint64_t ad = umma_smem_desc_sw128(smem_u32(sP + (ks >> 2) * kAtomBytes) + (ks & 3) * 32, 16, 1024);
        const uint64_t bd = umma_smem_desc_sw128(smem_u32(sV) + ks * 2048, (uint32_t)(Skp * 128), 1024);
        umma_f16(tO, ad, bd, idesc_o, ks > 0 ? 1u : 0u);
      }
      umma_commit(bar_o);
    }
  } else {
    // ---- softmax: thread = query row ----
    const int r = tid;  // 0..127 == TMEM lane
    const int qi = q0 + r;
    const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
    const float sl2 = a.scale * 1.4426950408889634f;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    const int nchunk = (Skp + 31) / 32;
    float m = -INFINITY;
    for (int c = 0; c < nchunk; ++c) {
      uint32_t raw[32];
      tmem_ld_32x32(tS + lane_base + c * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int col = c * 32 + i;
        const float bias = col < Skp ? sBias[col] : -INFINITY;
        m = fmaxf(m, __uint_as_float(raw[i]) * sl2 + bias);
      }
    }
    const float m_use = (m == -INFINITY) ? 0.f : m;
    const Philox ph(a.drop.seed);
    const uint32_t thr = dropout_threshold16(a.drop.p);
    const float keep_scale = dropout_keep_scale16(thr);
    const bool use_drop = a.drop.p > 0.f;
    const int kgroups = (a.Sk + 7) >> 3;
    float l = 0.f;
    for (int c = 0; c < nchunk; ++c) {
      uint32_t raw[32];
      tmem_ld_32x32(tS + lane_base + c * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        const int col0 = c * 32 + g8 * 8;
        if (col0 < Skp) {
          float pv[8];
          const uint32_t keep = use_drop ? philox_keep8(ph, bh * a.Sq + qi, kgroups, col0 >> 3, a.drop.site, thr) : 0xFFu;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int col = col0 + i;
            const float p = exp2f(__uint_as_float(raw[g8 * 8 + i]) * sl2 + sBias[col] - m_use);
            l += p;
            pv[i] = ((keep >> i) & 1u) ? (use_drop ? p * keep_scale : p) : 0.f;
          }
          st_swizzled8(sP, r, col0 >> 3, pv);
        }
      }
    }
    fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
    tc_fence_before();
    mbar_arrive(bar_p);
    // ---- epilogue: O / l ----
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    // all MMAs have completed: the operand tiles are dead, their space stages the output (4 KB per warp)
    uint8_t* stage = smem + warp * 4096;
    const int rows_valid = a.Sq - (q0 + warp * 32);
    bf16* g0 = a.o + ((size_t)b * a.Sq + q0 + warp * 32) * a.ldo + h * D;
#pragma unroll 1
    for (int c2 = 0; c2 < D / 64; ++c2) {
      uint32_t packed[32];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t raw[32];
        tmem_ld_32x32(tO + lane_base + (c2 * 2 + cc) * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          packed[cc * 16 + i] = pack_bf16x2(__uint_as_float(raw[2 * i]) * inv, __uint_as_float(raw[2 * i + 1]) * inv);
      }
      store_rows_coalesced<64>(stage, packed, lane, g0 + c2 * 64, (size_t)a.ldo, rows_valid);
    }
    if (qi < a.Sq && a.lse) a.lse[bh * a.Sq + qi] = l > 0.f ? (m + log2f(l)) * 0.6931471805599453f : -INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, (uint32_t)tmem_cols);
  }
}

// =============================================================================================
// backward (Sq <= 128, Sk <= 128): one CTA per (batch, head) produces dQ, dK and dV
// =============================================================================================
template <int D>
__global__ void __launch_bounds__(288) attn_tc_bwd_kernel(const __grid_constant__ TcMaps maps, AttnArgs a) {
  ECAMP_PDL_ENTRY();
  constexpr int ATOMS = D / 64;
  constexpr int TILE = ATOMS * kAtomBytes;  // a [128 x D] bf16 tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + TILE;
  uint8_t* sK = sdO + TILE;
  uint8_t* sV = sK + TILE;
  uint8_t* sP = sV + TILE;                  // [128 q x 128 keys] bf16 = 2 atoms (dropped probabilities)
  uint8_t* sdS = sP + 2 * kAtomBytes;       // [128 q x 128 keys] bf16
  float* sBias = reinterpret_cast<float*>(sdS + 2 * kAtomBytes);
  float* sDelta = sBias + 128;  // [2][128] partial row sums of the two column halves
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + 256);
  uint64_t *bar_ld = bars, *bar_s = bars + 1, *bar_p = bars + 2, *bar_g = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const uint64_t bh = (uint64_t)b * a.H + h;

  if (tid == 0) {
    mbar_init(bar_ld, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 256); mbar_init(bar_g, 1);
    fence_mbar_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int j = tid; j < 128; j += blockDim.x) {
    bool ok = j < a.Sk;
    if (ok && a.key_mask) ok = a.key_mask[(size_t)b * a.Sk + j] != 0;
    sBias[j] = ok ? 0.f : -INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tdP = tmem + 128, tdK = tmem + 256;  // dQ re-uses tS, dV re-uses tdP

  if (warp == 8) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_ld, 4u * TILE);
#pragma unroll
      for (int at = 0; at < ATOMS; ++at) {
        tma_load_2d(sQ + at * kAtomBytes, &maps.q, bar_ld, h * D + at * 64, b * a.Sq);
        tma_load_2d(sdO + at * kAtomBytes, &maps.d_o, bar_ld, h * D + at * 64, b * a.Sq);
        tma_load_2d(sK + at * kAtomBytes, &maps.k, bar_ld, h * D + at * 64, b * a.Sk);
        tma_load_2d(sV + at * kAtomBytes, &maps.v, bar_ld, h * D + at * 64, b * a.Sk);
      }
      mbar_wait(bar_ld, 0);
      tc_fence_after();
      // S = Q K^T and dP = dO V^T: all four tiles K-major, N = 128 keys
      const uint32_t idesc1 = umma_idesc_bf16(kRows, 128, false, false);
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks) {
        const uint32_t off = (ks >> 2) * kAtomBytes + (ks & 3) * 32;
        umma_f16(tS, umma_smem_desc_sw128(smem_u32(sQ) + off, 16, 1024), umma_smem_desc_sw128(smem_u32(sK) + off, 16, 1024),
                 idesc1, ks > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks) {
        const uint32_t off = (ks >> 2) * kAtomBytes + (ks & 3) * 32;
        umma_f16(tdP, umma_smem_desc_sw128(smem_u32(sdO) + off, 16, 1024),
                 umma_smem_desc_sw128(smem_u32(sV) + off, 16, 1024), idesc1, ks > 0 ? 1u : 0u);
      }
      umma_commit(bar_s);
      // second set, contraction over 128 keys (dQ) or 128 queries (dK, dV)
      mbar_wait(bar_p, 0);
      tc_fence_after();
      const uint32_t idesc_q = umma_idesc_bf16(kRows, D, false, true);  // A K-major (dS), B MN-major (K)
      const uint32_t idesc_t = umma_idesc_bf16(kRows, D, true, true);   // A MN-major (P^T / dS^T), B MN-major (dO / Q)
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {  // dQ = dS K
        const uint64_t ad = umma_smem_desc_sw128(smem_u32(sdS) + (ks >> 2) * kAtomBytes + (ks & 3) * 32, 16, 1024);
        const uint64_t bd = umma_smem_desc_sw128(smem_u32(sK) + ks * 2048, kAtomBytes, 1024);
        umma_f16(tS, ad, bd, idesc_q, ks > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {  // dV = Pdrop^T dO
        const uint64_t ad = umma_smem_desc_sw128(smem_u32(sP) + ks * 2048, kAtomBytes, 1024);
        const uint64_t bd = umma_smem_desc_sw128(smem_u32(sdO) + ks * 2048, kAtomBytes, 1024);
        umma_f16(tdP, ad, bd, idesc_t, ks > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {  // dK = dS^T Q
        const uint64_t ad = umma_smem_desc_sw128(smem_u32(sdS) + ks * 2048, kAtomBytes, 1024);
        const uint64_t bd = umma_smem_desc_sw128(smem_u32(sQ) + ks * 2048, kAtomBytes, 1024);
        umma_f16(tdK, ad, bd, idesc_t, ks > 0 ? 1u : 0u);
      }
      umma_commit(bar_g);
    }
  } else {
    // 8 warps: warp w serves TMEM lane quarter (w & 3) and column half (w >> 2), i.e. two threads share a row
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;  // query row for the softmax part, key row for the dK / dV epilogue
    const uint32_t lane_base = ((uint32_t)(quarter * 32)) << 16;
    const bool qvalid = r < a.Sq;
    const float lse = qvalid ? a.lse[bh * a.Sq + r] : INFINITY;  // +inf -> P = 0 for padded query rows
    const Philox ph(a.drop.seed);
    const uint32_t thr = dropout_threshold16(a.drop.p);
    const bool use_drop = a.drop.p > 0.f;
    const float keep_scale = use_drop ? dropout_keep_scale16(thr) : 1.0f;
    const int kgroups = (a.Sk + 7) >> 3;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    // pass 1: delta = sum_j P_ij dPeff_ij (exact fp32, same P / dP as pass 2)
    float delta = 0.f;
#pragma unroll 1
    for (int c = 2 * half; c < 2 * half + 2; ++c) {
      uint32_t rs[32], rp[32];
      tmem_ld_32x32(tS + lane_base + c * 32, rs);
      tmem_ld_32x32(tdP + lane_base + c * 32, rp);
      tmem_ld_wait();
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        const uint32_t keep = use_drop ? philox_keep8(ph, bh * a.Sq + r, kgroups, c * 4 + g8, a.drop.site, thr) : 0xFFu;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int col = c * 32 + g8 * 8 + i;
          const float p = __expf(__uint_as_float(rs[g8 * 8 + i]) * a.scale + sBias[col] - lse);
          const float dpe = ((keep >> i) & 1u) ? __uint_as_float(rp[g8 * 8 + i]) * keep_scale : 0.f;
          delta += p * dpe;
        }
      }
    }
    sDelta[half * 128 + r] = delta;
    asm volatile("bar.sync 1, 256;" ::: "memory");  // the 256 softmax threads only
    delta = sDelta[r] + sDelta[128 + r];
    if (half == 0 && qvalid && a.delta) a.delta[bh * a.Sq + r] = delta;
    // pass 2: Pdrop and dS = P (dPeff - delta) scale -> shared memory (bf16, swizzled K-major [q, key])
#pragma unroll 1
    for (int c = 2 * half; c < 2 * half + 2; ++c) {
      uint32_t rs[32], rp[32];
      tmem_ld_32x32(tS + lane_base + c * 32, rs);
      tmem_ld_32x32(tdP + lane_base + c * 32, rp);
      tmem_ld_wait();
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        float pv[8], dv[8];
        const uint32_t keep = use_drop ? philox_keep8(ph, bh * a.Sq + r, kgroups, c * 4 + g8, a.drop.site, thr) : 0xFFu;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int col = c * 32 + g8 * 8 + i;
          const float p = __expf(__uint_as_float(rs[g8 * 8 + i]) * a.scale + sBias[col] - lse);
          const bool kp = (keep >> i) & 1u;
          const float dpe = kp ? __uint_as_float(rp[g8 * 8 + i]) * keep_scale : 0.f;
          pv[i] = kp ? p * keep_scale : 0.f;
          dv[i] = p * (dpe - delta) * a.scale;
        }
        st_swizzled8(sP, r, c * 4 + g8, pv);
        st_swizzled8(sdS, r, c * 4 + g8, dv);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(bar_p);
    // epilogue: dQ rows are queries, dK / dV rows are keys
    mbar_wait(bar_g, 0);
    tc_fence_after();
    // all MMAs have completed (bar_g): the operand tiles are dead, their space stages the outputs (4 KB per warp)
    uint8_t* stage = smem + warp * 4096;
#pragma unroll 1
    for (int which = 0; which < 3; ++which) {
      const uint32_t tsrc = which == 0 ? tS : (which == 1 ? tdK : tdP);
      const int nrows = which == 0 ? a.Sq : a.Sk;
      bf16* base = which == 0 ? a.dq + (size_t)b * a.Sq * a.lddq
                              : (which == 1 ? a.dk + (size_t)b * a.Sk * a.lddk : a.dv + (size_t)b * a.Sk * a.lddv);
      const size_t ld = which == 0 ? (size_t)a.lddq : (which == 1 ? (size_t)a.lddk : (size_t)a.lddv);
      constexpr int NC = D / 2;  // columns per warp (two warps share a row): 64 or 32
      uint32_t packed[NC / 2];
#pragma unroll
      for (int cc = 0; cc < NC / 32; ++cc) {
        uint32_t raw[32];
        tmem_ld_32x32(tsrc + lane_base + (half * (NC / 32) + cc) * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) packed[cc * 16 + i] = pack_bf16x2(__uint_as_float(raw[2 * i]), __uint_as_float(raw[2 * i + 1]));
      }
      float* cs = which == 0 ? a.cs_q : (which == 1 ? a.cs_k : a.cs_v);
      store_rows_coalesced<NC>(stage, packed, lane, base + (size_t)(quarter * 32) * ld + h * D + half * NC, ld,
                               nrows - quarter * 32, cs ? cs + h * D + half * NC : nullptr);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int build_maps(const AttnArgs& a, bool bwd, int k_box_rows, TcMaps* m) {
  const unsigned long long rq = (unsigned long long)a.B * a.Sq, rk = (unsigned long long)a.B * a.Sk;
  const unsigned long long w = (unsigned long long)a.H * a.D;
  int rc;
  if ((rc = make_tmap_bf16(&m->q, a.q, w, rq, (unsigned long long)a.ldq, kRows))) return rc;
  if ((rc = make_tmap_bf16(&m->k, a.k, w, rk, (unsigned long long)a.ldk, (unsigned)k_box_rows))) return rc;
  if ((rc = make_tmap_bf16(&m->v, a.v, w, rk, (unsigned long long)a.ldv, (unsigned)k_box_rows))) return rc;
  if (bwd) {
    if ((rc = make_tmap_bf16(&m->d_o, a.d_o, w, rq, (unsigned long long)a.ld_do, kRows))) return rc;
  } else {
    m->d_o = m->q;
  }
  return 0;
}

template <int D>
int launch_tc_fwd(const AttnArgs& a, cudaStream_t st) {
  const int Skp = (a.Sk + 15) & ~15;
  TcMaps maps;
  if (int rc = build_maps(a, false, Skp, &maps)) return rc;
  constexpr int ATOMS = D / 64;
  const size_t smem = (size_t)ATOMS * kAtomBytes + 2 * (size_t)ATOMS * Skp * 128 + 256 * 4 + 64 + 1024;
  const int need = ((Skp + 31) & ~31) + D;
  const int tmem_cols = need <= 128 ? 128 : (need <= 256 ? 256 : 512);
  static bool attr = false;
  if (!attr) {
    ECAMP_CUDA_OK(cudaFuncSetAttribute(attn_tc_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr = true;
  }
  ECAMP_REQUIRE(smem <= 226 * 1024, "attention (tcgen05): shared memory %zu B too large", smem);
  dim3 grid((a.Sq + kRows - 1) / kRows, a.H, a.B);
  ECAMP_CUDA_OK(launch_pdl(attn_tc_fwd_kernel<D>, grid, 160, smem, st, maps, a, tmem_cols));
  ECAMP_LAUNCHED();
  return 0;
}

template <int D>
int launch_tc_bwd(const AttnArgs& a, cudaStream_t st) {
  TcMaps maps;
  if (int rc = build_maps(a, true, kRows, &maps)) return rc;
  constexpr int ATOMS = D / 64;
  const size_t smem = 4 * (size_t)ATOMS * kAtomBytes + 4 * (size_t)kAtomBytes + 384 * 4 + 64 + 1024;
  static bool attr = false;
  if (!attr) {
    ECAMP_CUDA_OK(cudaFuncSetAttribute(attn_tc_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr = true;
  }
  dim3 grid(a.H, a.B);
  ECAMP_CUDA_OK(launch_pdl(attn_tc_bwd_kernel<D>, grid, 288, smem, st, maps, a));
  ECAMP_LAUNCHED();
  return 0;
}

bool tc_layout_ok(const AttnArgs& a) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return (a.D == 64 || a.D == 128) && a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 8 == 0 && al(a.q) &&
         al(a.k) && al(a.v) && al(a.o);
}

}  // namespace

bool attention_tc_fwd_supported(const AttnArgs& a) {
  if (!tc_layout_ok(a) || a.Sk > 256 || a.Sk < 1) return false;
  const int Skp = (a.Sk + 15) & ~15, atoms = a.D / 64;
  // P (bf16 [128, Skp] in 64-column atoms) re-uses the Q | K staging area
  return (size_t)((Skp + 63) / 64) * kAtomBytes <= (size_t)atoms * (kAtomBytes + Skp * 128);
}
bool attention_tc_bwd_supported(const AttnArgs& a) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  // the ViT encoder (50 x 50, head_dim 64) wastes most of a 128 x 128 tile: measured slower than the mma.sync kernels
  if (a.D == 64 && a.Sq <= 64 && a.Sk <= 64) return false;
  return tc_layout_ok(a) && a.Sq <= 128 && a.Sk <= 128 && a.ld_do % 8 == 0 && a.lddq % 8 == 0 && a.lddk % 8 == 0 &&
         a.lddv % 8 == 0 && al(a.d_o) && al(a.dq) && al(a.dk) && al(a.dv);
}
int attention_tc_fwd(const AttnArgs& a, cudaStream_t st) {
  return a.D == 64 ? launch_tc_fwd<64>(a, st) : launch_tc_fwd<128>(a, st);
}
int attention_tc_bwd(const AttnArgs& a, cudaStream_t st) {
  return a.D == 64 ? launch_tc_bwd<64>(a, st) : launch_tc_bwd<128>(a, st);
}

}  // namespace ecamp
