// Fused attention forward / backward for the four shapes of the ECAMP step:
//   ViT encoder 12 heads x 64, S = 50;  ViT decoder 16 heads x 32, S = 197            (timm Attention)
//   BERT self-attention 6 heads x 128, S = T, additive key-padding mask, prob-dropout  (HF BertSelfAttention)
//   cross-attention text -> 49 image tokens, 6 heads x 128                             (context_fusion.py:45-53)
// Scores are never written to memory: one CTA owns 64 rows of one (batch, head), keeps the whole
// K/V (or Q/dO) of that head in shared memory and runs an online softmax over 64-column blocks.
// Tensor-core path: mma.sync m16n8k16 bf16 (legacy HMMA issue path; attention is 2.4 % of the step's
// FLOPs — moving it onto tcgen05 is a later-round item).
//
// Backward is two passes of ONE kernel template: pass "dQ" walks key blocks for a query tile, pass
// "dK/dV" walks query blocks for a key tile (the transposed problem); no atomics, deterministic.
#include "kernels.cuh"

#include <cstdlib>
#include <type_traits>

namespace ecamp {
namespace {

ECAMP_DEVINL void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// first k-step of an accumulation: C = 0 comes from the zero register instead of 4 cleared accumulator registers
ECAMP_DEVINL void mma_bf16_16816_z(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%10, %10, %10, %10};"
      : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
ECAMP_DEVINL void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
ECAMP_DEVINL void ldsm_x4_t(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}

// A fragment (16 rows x 16 k) of a row-major [rows][LDS] bf16 tile: row0 = first row, k0 = first column.
template <int LDS>
ECAMP_DEVINL void load_a_frag(uint32_t (&a)[4], const bf16* tile, int row0, int k0, int lane) {
  const bf16* p = tile + (size_t)(row0 + (lane & 15)) * LDS + k0 + ((lane >> 4) << 3);
  ldsm_x4(a, smem_u32(p));
}
// B fragments for two adjacent n-tiles from a tile stored [n][k] (k contiguous): r[0..1] -> n-tile 0, r[2..3] -> n-tile 1.
template <int LDS>
ECAMP_DEVINL void load_b_frag_nk(uint32_t (&r)[4], const bf16* tile, int n0, int k0, int lane) {
  const bf16* p = tile + (size_t)(n0 + (lane & 7) + ((lane >> 4) << 3)) * LDS + k0 + (((lane >> 3) & 1) << 3);
  ldsm_x4(r, smem_u32(p));
}
// B fragments for two adjacent n-tiles from a tile stored [k][n] (n contiguous): transposing load.
template <int LDS>
ECAMP_DEVINL void load_b_frag_kn(uint32_t (&r)[4], const bf16* tile, int k0, int n0, int lane) {
  const bf16* p = tile + (size_t)(k0 + (lane & 15)) * LDS + n0 + ((lane >> 4) << 3);
  ldsm_x4_t(r, smem_u32(p));
}

// Lane-dependent part of the three fragment addresses, computed ONCE per kernel: a fragment load inside the unrolled
// loops is then `base + block_offset + constant` (one integer add; the per-call index arithmetic was ~15 % of the
// executed instructions of the head_dim-32 kernels).
template <int LDS>
ECAMP_DEVINL uint32_t frag_base_a(const bf16* tile, int row0, int lane) {  // + k0 * 2
  return smem_u32(tile + (size_t)(row0 + (lane & 15)) * LDS + ((lane >> 4) << 3));
}
template <int LDS>
ECAMP_DEVINL uint32_t frag_base_nk(const bf16* tile, int lane) {  // + (n0 * LDS + k0) * 2
  return smem_u32(tile + (size_t)((lane & 7) + ((lane >> 4) << 3)) * LDS + (((lane >> 3) & 1) << 3));
}
template <int LDS>
ECAMP_DEVINL uint32_t frag_base_kn(const bf16* tile, int lane) {  // + (k0 * LDS + n0) * 2
  return smem_u32(tile + (size_t)(lane & 15) * LDS + ((lane >> 4) << 3));
}

// cooperative asynchronous copy of `rows` rows (zero-filled from `valid` on) of width D from global (row pitch ld)
// to shared memory: every 16-byte chunk is one cp.async, so all of a thread's loads are in flight at once
// (a plain load/store loop kept one load in flight per thread and dominated the kernels: long-scoreboard stalls).
template <int D, int LDS>
ECAMP_DEVINL void load_tile(bf16* dst, const bf16* src, int ld, int rows, int valid, int tid, int nthreads) {
  constexpr int CPR = D / 8;
  for (int idx = tid; idx < rows * CPR; idx += nthreads) {
    const int r = idx / CPR, c = idx % CPR;
    const bool ok = r < valid;
    const bf16* g = src + (size_t)(ok ? r : 0) * ld + c * 8;
    const uint32_t sz = ok ? 16u : 0u;  // src-size 0 -> the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst + (size_t)r * LDS + c * 8)), "l"(g),
                 "r"(sz)
                 : "memory");
  }
}
ECAMP_DEVINL void load_tile_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

ECAMP_DEVINL float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
ECAMP_DEVINL float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// keep-decision of attention-probability dropout for (query i, key j) of head-instance bh: the same 16-bit-per-
// element Philox stream as the tcgen05 kernels (attention_tc.cu), so forward and backward may use either family
ECAMP_DEVINL bool attn_keep(const Philox& ph, uint32_t thr16, uint64_t site, uint64_t bh, int Sq, int Sk, int i, int j) {
  const uint32_t m = philox_keep8(ph, bh * (uint64_t)Sq + (uint64_t)i, (Sk + 7) >> 3, j >> 3, site, thr16);
  return (m >> (j & 7)) & 1u;
}

// =============================================================================================
// forward
// =============================================================================================
// The kernels are bound by the element-wise work on the score tile, not by the tensor pipe (ncu: issue slots 75-80 %
// busy, tensor pipe 26 %), so that part is kept to ~5 instructions per score: one FMNMX for the running maximum on
// the RAW score (the scale is positive), one FFMA that folds scale * log2(e) and the maximum into the ex2 argument,
// ex2, one FADD for the row sum and half a pack.  Key masks / the ragged tail are only tested in the 16-column group
// that contains them, fully padded groups and fully padded 16-row warp slices are skipped.
ECAMP_DEVINL float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// NW warps of 16 query rows per CTA; the whole K / V of the head is staged once per CTA, so NW is chosen to cover all
// queries of a head when they fit (decoder: 13 warps for S = 197) - with 64-row CTAs K / V were re-read 4 times
// through L2, which bounded the kernel.
template <int D, int NW, bool DROP>
__global__ void __launch_bounds__(NW * 32) attn_fwd_kernel(AttnArgs a) {
  ECAMP_PDL_ENTRY();
  constexpr int RT = NW * 16, NT_ = NW * 32;
  constexpr int LDS = D + 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int Skp = (a.Sk + 15) & ~15;
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sK = sQ + RT * LDS;
  bf16* sV = sK + (size_t)Skp * LDS;
  float* sBias = reinterpret_cast<float*>(sV + (size_t)Skp * LDS);  // additive key mask: 0 or -inf

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.x * RT, h = blockIdx.y, b = blockIdx.z;
  const int g = lane >> 2, t4 = lane & 3;

  load_tile<D, LDS>(sQ, a.q + ((size_t)b * a.Sq + q0) * a.ldq + h * D, a.ldq, RT, min(RT, a.Sq - q0), tid, NT_);
  load_tile<D, LDS>(sK, a.k + (size_t)b * a.Sk * a.ldk + h * D, a.ldk, Skp, a.Sk, tid, NT_);
  load_tile<D, LDS>(sV, a.v + (size_t)b * a.Sk * a.ldv + h * D, a.ldv, Skp, a.Sk, tid, NT_);
  // keys after the last attendable one contribute exactly zero: the key loop stops there (padding is a suffix)
  __shared__ int s_kend, s_kfull;
  if (tid == 0) { s_kend = 0; s_kfull = Skp; }
  __syncthreads();
  for (int j = tid; j < Skp; j += NT_) {
    bool ok = j < a.Sk;
    if (ok && a.key_mask) ok = a.key_mask[(size_t)b * a.Sk + j] != 0;
    sBias[j] = ok ? 0.f : -INFINITY;
    if (ok) atomicMax(&s_kend, j + 1);
    else atomicMin(&s_kfull, j);
  }
  load_tile_wait();
  __syncthreads();
  const int kend = (s_kend + 15) & ~15;
  const int kfull = s_kfull & ~15;  // every key below this index is attendable: no per-element test needed there
  if (q0 + warp * 16 >= a.Sq) return;  // this warp's 16 query rows are all padding (no block-wide barrier follows)

  uint32_t qf[D / 16][4];
#pragma unroll
  for (int kt = 0; kt < D / 16; ++kt) load_a_frag<LDS>(qf[kt], sQ, warp * 16, kt * 16, lane);
  const uint32_t kbase = frag_base_nk<LDS>(sK, lane), vbase = frag_base_kn<LDS>(sV, lane);

  float o[D / 8][4];
#pragma unroll
  for (int j = 0; j < D / 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};  // running maximum of the RAW scores, row sums
  const float sl2 = a.scale * 1.4426950408889634f;

  const Philox ph(a.drop.seed);
  const uint32_t thr = dropout_threshold16(a.drop.p);
  const float keep_scale = a.drop.p > 0.f ? dropout_keep_scale16(thr) : 1.0f;
  const uint64_t bh = (uint64_t)b * a.H + h;
  const int row_g = q0 + warp * 16 + g;  // this thread's rows: row_g and row_g + 8

  // One 64-key block.  ALL: all four 16-key groups lie below kend - no per-group tests, first k-step from the zero
  // register; NOBIAS: every key of the block is attendable - no mask bias either (the tests and the accumulator clears
  // were ~25 % of the executed instructions).
  auto key_block = [&](const int kb, auto all_tag, auto nobias_tag) {
    constexpr bool ALL = decltype(all_tag)::value, NOBIAS = decltype(nobias_tag)::value;
    float s[8][4];
    if (!ALL) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
    }
#pragma unroll
    for (int kt = 0; kt < D / 16; ++kt) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        if (ALL || kb + jp * 16 < kend) {
          uint32_t bf[4];
          ldsm_x4(bf, kbase + (uint32_t)(kb * LDS * 2) + (uint32_t)((jp * 16 * LDS + kt * 16) * 2));
          if (ALL && kt == 0) {
            mma_bf16_16816_z(s[2 * jp], qf[kt], bf[0], bf[1]);
            mma_bf16_16816_z(s[2 * jp + 1], qf[kt], bf[2], bf[3]);
          } else {
            mma_bf16_16816(s[2 * jp], qf[kt], bf[0], bf[1]);
            mma_bf16_16816(s[2 * jp + 1], qf[kt], bf[2], bf[3]);
          }
        }
      }
    }
    // mask (only in 16-column groups that contain a masked / padded key), block row-max of the raw scores
    float bm[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      const int c0 = kb + jp * 16;
      if (ALL || c0 < kend) {
        if (!NOBIAS && c0 + 16 > kfull) {
#pragma unroll
          for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[2 * jp + jj][e] += sBias[c0 + jj * 8 + t4 * 2 + (e & 1)];
        }
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          bm[0] = fmaxf(bm[0], fmaxf(s[2 * jp + jj][0], s[2 * jp + jj][1]));
          bm[1] = fmaxf(bm[1], fmaxf(s[2 * jp + jj][2], s[2 * jp + jj][3]));
        }
      }
    }
    float corr[2], m_sc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float m_new = fmaxf(m_run[r], quad_max(bm[r]));
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      corr[r] = ex2f((m_run[r] - m_use) * sl2);  // m_run = -inf -> 0
      m_sc[r] = m_use * sl2;
      m_run[r] = m_new;
      l_run[r] *= corr[r];
    }
#pragma unroll
    for (int j = 0; j < D / 8; ++j) {
      o[j][0] *= corr[0]; o[j][1] *= corr[0];
      o[j][2] *= corr[1]; o[j][3] *= corr[1];
    }
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      if (ALL || kb + jp * 16 < kend) {
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int j = 2 * jp + jj;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float p = ex2f(fmaf(s[j][e], sl2, -m_sc[e >> 1]));  // masked: -inf -> 0
            l_run[e >> 1] += p;
            if (DROP) {
              const int col = kb + j * 8 + t4 * 2 + (e & 1);
              const int row = row_g + (e >> 1) * 8;
              p = attn_keep(ph, thr, a.drop.site, bh, a.Sq, a.Sk, row, col) ? p * keep_scale : 0.f;
            }
            s[j][e] = p;
          }
        }
      }
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (ALL || kb + kk * 16 < kend) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int dp = 0; dp < D / 16; ++dp) {
          uint32_t vf[4];
          ldsm_x4_t(vf, vbase + (uint32_t)(kb * LDS * 2) + (uint32_t)((kk * 16 * LDS + dp * 16) * 2));
          mma_bf16_16816(o[2 * dp], pa, vf[0], vf[1]);
          mma_bf16_16816(o[2 * dp + 1], pa, vf[2], vf[3]);
        }
      }
    }
  };
  {
    const int kall = kend & ~63;               // keys [0, kall): whole blocks
    const int kfast = min(kall, kfull & ~63);  // keys [0, kfast): whole blocks without a masked or padded key
    int kb = 0;
    for (; kb < kfast; kb += 64) key_block(kb, std::true_type{}, std::true_type{});
    for (; kb < kall; kb += 64) key_block(kb, std::true_type{}, std::false_type{});
    for (; kb < kend; kb += 64) key_block(kb, std::false_type{}, std::false_type{});
  }

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const float l = quad_sum(l_run[r]);
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    const int row = row_g + r * 8;
    if (row < a.Sq) {
      bf16* orow = a.o + ((size_t)b * a.Sq + row) * a.ldo + h * D;
#pragma unroll
      for (int j = 0; j < D / 8; ++j)
        *reinterpret_cast<uint32_t*>(orow + j * 8 + t4 * 2) = pack_bf16x2(o[j][2 * r] * inv, o[j][2 * r + 1] * inv);
      if (t4 == 0 && a.lse) a.lse[(bh * a.Sq) + row] = (l > 0.f) ? m_run[r] * a.scale + __logf(l) : -INFINITY;
    }
  }
}

// delta[b, h, i] = sum_d dO[i, d] * O[i, d]; one warp per (b, h, i)
__global__ void attn_delta_kernel(AttnArgs a) {
  ECAMP_PDL_ENTRY();
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int total = a.B * a.H * a.Sq;
  if (gw >= total) return;
  const int i = gw % a.Sq, h = (gw / a.Sq) % a.H, b = gw / (a.Sq * a.H);
  const bf16* orow = a.o + ((size_t)b * a.Sq + i) * a.ldo + h * a.D;
  const bf16* grow = a.d_o + ((size_t)b * a.Sq + i) * a.ld_do + h * a.D;
  float s = 0.f;
  for (int d = lane * 2; d < a.D; d += 64) {
    const float2 x = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(orow + d));
    const float2 y = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(grow + d));
    s += x.x * y.x + x.y * y.y;
  }
  s = warp_sum(s);
  if (lane == 0) a.delta[gw] = s;
}

// =============================================================================================
// backward.  TR = false: rows = queries (R1 = Q, R2 = dO), columns = keys (C1 = K, C2 = V); out = dQ = dS K.
//            TR = true : rows = keys    (R1 = K, R2 = V),  columns = queries (C1 = Q, C2 = dO);
//                        out1 = dK = dS^T Q, out2 = dV = Pdrop^T dO.
// =============================================================================================
// The backward of one row tile (NW * 16 rows starting at r0) against all columns, operands already in shared memory.
// Returns (block-uniformly) true when the tile was fully masked and has been zero-filled.
template <int D, bool TR, bool DROP, int NW>
ECAMP_DEVINL bool attn_bwd_compute(const AttnArgs& a, const bf16* sR1, const bf16* sR2, const bf16* sC1, const bf16* sC2,
                                   const float* sColA, const float* sColB, int r0, int cend, int cfull, int b, int h) {
  constexpr int LDS = D + 8;
  constexpr int CB = TR ? 32 : 64;  // column block
  constexpr int NT = CB / 8;
  const int Sr = TR ? a.Sk : a.Sq;  // row-side length
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const uint64_t bh = (uint64_t)b * a.H + h;
  const float sl2 = a.scale * 1.4426950408889634f;
  const int rvalid = min(NW * 16, Sr - r0);
  // per-row statistics / validity for this thread's two rows
  const int row_g = r0 + warp * 16 + g;
  float row_a[2], row_b[2];  // !TR: -lse * log2(e), delta;  TR: key bias (0 / -inf), unused
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row_g + r * 8;
    if (!TR) {
      row_a[r] = row < Sr ? -a.lse[bh * a.Sq + row] * 1.4426950408889634f : -INFINITY;  // lse = -inf (fully masked row) -> +inf
      if (row < Sr && a.lse[bh * a.Sq + row] == -INFINITY) row_a[r] = -INFINITY;          // ... which must still give p = 0
      row_b[r] = 0.f;  // delta: produced by pass 0 below
    } else {
      bool ok = row < Sr;
      if (ok && a.key_mask) ok = a.key_mask[(size_t)b * a.Sk + row] != 0;
      row_a[r] = ok ? 0.f : -INFINITY;  // key bias
      row_b[r] = 0.f;
    }
  }

  if (TR) {
    const int any_valid = __syncthreads_or((row_a[0] == 0.f || row_a[1] == 0.f) ? 1 : 0);
    if (!any_valid) {
      for (int idx = tid; idx < rvalid * (D / 8); idx += NW * 32) {
        const int r = idx / (D / 8), c8 = idx % (D / 8);
        const size_t row = (size_t)b * a.Sk + r0 + r;
        *reinterpret_cast<uint4*>(a.dk + row * a.lddk + h * D + c8 * 8) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(a.dv + row * a.lddv + h * D + c8 * 8) = make_uint4(0u, 0u, 0u, 0u);
      }
      return true;  // block-uniform: nothing else to do for this row tile
    }
  }
  if (r0 + warp * 16 >= Sr) return false;  // this warp's 16 rows are all padding

  float acc1[D / 8][4];
  float acc2[TR ? D / 8 : 1][4];
#pragma unroll
  for (int j = 0; j < D / 8; ++j) acc1[j][0] = acc1[j][1] = acc1[j][2] = acc1[j][3] = 0.f;
#pragma unroll
  for (int j = 0; j < (TR ? D / 8 : 1); ++j) acc2[j][0] = acc2[j][1] = acc2[j][2] = acc2[j][3] = 0.f;

  const Philox ph(a.drop.seed);
  const uint32_t thr = dropout_threshold16(a.drop.p);
  const float keep_scale = a.drop.p > 0.f ? dropout_keep_scale16(thr) : 1.0f;

  // The dQ kernel runs two passes over the key blocks: pass 0 accumulates delta_i = sum_j P_ij dP_ij in fp32 from
  // the SAME P / dP that pass 1 uses for dS = P (dP - delta), so the cancellation inside (dP - delta) is exact
  // (computing delta from the bf16-rounded O, FlashAttention-style, loses it when attention is near-uniform).
  // delta is also published for the dK/dV kernel, which runs afterwards on the same stream.
  // Element-wise work per score: p = ex2(s * scale*log2e + (bias) - lse*log2e) (one FFMA + ex2; the bias only in
  // 16-column groups that contain a masked / padded key), then pass 0: one FFMA; pass 1: FADD + 2 FMUL + packs.
  float dsum[2] = {0.f, 0.f};
#pragma unroll 1
  for (int pass = TR ? 1 : 0; pass < 2; ++pass) {
  const uint32_t r1base = frag_base_a<LDS>(sR1, warp * 16, lane), r2base = frag_base_a<LDS>(sR2, warp * 16, lane);
  const uint32_t c1nk = frag_base_nk<LDS>(sC1, lane), c2nk = frag_base_nk<LDS>(sC2, lane);
  const uint32_t c1kn = frag_base_kn<LDS>(sC1, lane), c2kn = frag_base_kn<LDS>(sC2, lane);
  // One block of CB columns.  ALL: every 16-column group lies below cend - no per-group tests, first k-step from the
  // zero register; NOBIAS (dQ pass): every key of the block is attendable - no mask bias either.
  // FUSED (dQ problem with a single column block, e.g. the encoder's 50 keys): the delta step and the dS step share one
  // evaluation of S, dP and the exponentials - they stay in registers across the row reduction.
  auto col_block = [&](const int cb, auto all_tag, auto nobias_tag, auto fused_tag) {
    constexpr bool ALL = decltype(all_tag)::value, NOBIAS = decltype(nobias_tag)::value, FUSED = decltype(fused_tag)::value;
    float s[NT][4], dp[NT][4];
    if (!ALL) {
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
      }
    }
#pragma unroll
    for (int kt = 0; kt < D / 16; ++kt) {
      uint32_t a1[4], a2[4];
      ldsm_x4(a1, r1base + (uint32_t)(kt * 16 * 2));
      ldsm_x4(a2, r2base + (uint32_t)(kt * 16 * 2));
#pragma unroll
      for (int jp = 0; jp < NT / 2; ++jp) {
        if (ALL || cb + jp * 16 < cend) {
          uint32_t bf[4];
          ldsm_x4(bf, c1nk + (uint32_t)(cb * LDS * 2) + (uint32_t)((jp * 16 * LDS + kt * 16) * 2));
          if (ALL && kt == 0) {
            mma_bf16_16816_z(s[2 * jp], a1, bf[0], bf[1]);
            mma_bf16_16816_z(s[2 * jp + 1], a1, bf[2], bf[3]);
          } else {
            mma_bf16_16816(s[2 * jp], a1, bf[0], bf[1]);
            mma_bf16_16816(s[2 * jp + 1], a1, bf[2], bf[3]);
          }
          ldsm_x4(bf, c2nk + (uint32_t)(cb * LDS * 2) + (uint32_t)((jp * 16 * LDS + kt * 16) * 2));
          if (ALL && kt == 0) {
            mma_bf16_16816_z(dp[2 * jp], a2, bf[0], bf[1]);
            mma_bf16_16816_z(dp[2 * jp + 1], a2, bf[2], bf[3]);
          } else {
            mma_bf16_16816(dp[2 * jp], a2, bf[0], bf[1]);
            mma_bf16_16816(dp[2 * jp + 1], a2, bf[2], bf[3]);
          }
        }
      }
    }
#pragma unroll
    for (int jp = 0; jp < NT / 2; ++jp) {
      const int c0 = cb + jp * 16;
      if (ALL || c0 < cend) {
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int j = 2 * jp + jj;
          const int colb = c0 + jj * 8 + t4 * 2;  // this thread's two columns of the 8-wide n-tile: colb, colb + 1
          float ce[2], cd[2];                     // per-column exponent offset / delta
          if (!TR) {
            ce[0] = ce[1] = 0.f;
            if (!NOBIAS && c0 + 16 > cfull) { ce[0] = sColA[colb]; ce[1] = sColA[colb + 1]; }
            cd[0] = cd[1] = 0.f;
          } else {
            const float2 la = *reinterpret_cast<const float2*>(sColA + colb);
            const float2 de = *reinterpret_cast<const float2*>(sColB + colb);
            ce[0] = la.x; ce[1] = la.y; cd[0] = de.x; cd[1] = de.y;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r = e >> 1, cc = e & 1;
            // !TR: exponent = s*sl2 + key_bias - lse2(row);  TR: s*sl2 + key_bias(row) - lse2(col)
            const float p = ex2f(fmaf(s[j][e], sl2, ce[cc] + row_a[r]));  // masked / padded -> ex2(-inf) = 0
            float dpe = dp[j][e];
            float pd = p;
            if (DROP) {
              const int row = row_g + r * 8, col = colb + cc;
              const int qi = TR ? col : row, kj = TR ? row : col;
              const bool keep = attn_keep(ph, thr, a.drop.site, bh, a.Sq, a.Sk, qi, kj);
              dpe = keep ? dpe * keep_scale : 0.f;
              pd = keep ? p * keep_scale : 0.f;
            }
            if (FUSED) {
              dsum[r] = fmaf(p, dpe, dsum[r]);
              s[j][e] = p;
              dp[j][e] = dpe;
            } else if (!TR && pass == 0) {
              dsum[r] = fmaf(p, dpe, dsum[r]);
            } else {
              s[j][e] = (p * a.scale) * (dpe - (TR ? cd[cc] : row_b[r]));  // dS
              dp[j][e] = pd;                                                // dropped probabilities (only used when TR)
            }
          }
        }
      }
    }
    if (FUSED) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        row_b[r] = quad_sum(dsum[r]);
        const int row = row_g + r * 8;
        if (t4 == 0 && row < Sr) a.delta[bh * a.Sq + row] = row_b[r];
      }
#pragma unroll
      for (int jp = 0; jp < NT / 2; ++jp) {
        if (ALL || cb + jp * 16 < cend) {
#pragma unroll
          for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int e = 0; e < 4; ++e)
              s[2 * jp + jj][e] = (s[2 * jp + jj][e] * a.scale) * (dp[2 * jp + jj][e] - row_b[e >> 1]);  // dS
        }
      }
    } else if (pass == 0) {
      return;
    }
    // out1 += dS . C1 ; (TR) out2 += Pdrop . C2
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
      if (ALL || cb + kk * 16 < cend) {
        uint32_t da[4], pa[4];
        da[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        da[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        da[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        da[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
        if (TR) {
          pa[0] = pack_bf16x2(dp[2 * kk][0], dp[2 * kk][1]);
          pa[1] = pack_bf16x2(dp[2 * kk][2], dp[2 * kk][3]);
          pa[2] = pack_bf16x2(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
          pa[3] = pack_bf16x2(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
        }
#pragma unroll
        for (int dd = 0; dd < D / 16; ++dd) {
          uint32_t cf[4];
          ldsm_x4_t(cf, c1kn + (uint32_t)(cb * LDS * 2) + (uint32_t)((kk * 16 * LDS + dd * 16) * 2));
          mma_bf16_16816(acc1[2 * dd], da, cf[0], cf[1]);
          mma_bf16_16816(acc1[2 * dd + 1], da, cf[2], cf[3]);
          if (TR) {
            ldsm_x4_t(cf, c2kn + (uint32_t)(cb * LDS * 2) + (uint32_t)((kk * 16 * LDS + dd * 16) * 2));
            mma_bf16_16816(acc2[2 * dd], pa, cf[0], cf[1]);
            mma_bf16_16816(acc2[2 * dd + 1], pa, cf[2], cf[3]);
          }
        }
      }
    }
  };
  if constexpr (!TR) {
    if (cend <= CB) {  // a single column block: one fused evaluation instead of two passes
      if (cend == CB && cfull >= CB) col_block(0, std::true_type{}, std::true_type{}, std::true_type{});
      else if (cend == CB) col_block(0, std::true_type{}, std::false_type{}, std::true_type{});
      else col_block(0, std::false_type{}, std::false_type{}, std::true_type{});
      break;
    }
  }
  {
    const int call = cend & ~(CB - 1);                          // whole blocks
    const int cfast = TR ? call : min(call, cfull & ~(CB - 1));  // whole blocks that need no test at all
    int cb = 0;
    for (; cb < cfast; cb += CB) col_block(cb, std::true_type{}, std::true_type{}, std::false_type{});
    for (; cb < call; cb += CB) col_block(cb, std::true_type{}, std::false_type{}, std::false_type{});
    for (; cb < cend; cb += CB) col_block(cb, std::false_type{}, std::false_type{}, std::false_type{});
  }
    if (pass == 0) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        row_b[r] = quad_sum(dsum[r]);
        const int row = row_g + r * 8;
        if (t4 == 0 && row < Sr) a.delta[bh * a.Sq + row] = row_b[r];
      }
    }
  }  // passes

  // bias gradients of the q / k / v projections: column sums of this warp's 16 rows (fp32 accumulators), one atomic
  // per column and warp
  {
    float* cs1 = TR ? a.cs_k : a.cs_q;
    float* cs2 = TR ? a.cs_v : nullptr;
    if (cs1 || cs2) {
      const bool v0 = row_g < Sr, v1 = row_g + 8 < Sr;
#pragma unroll
      for (int j = 0; j < D / 8; ++j) {
        float c1a = (v0 ? acc1[j][0] : 0.f) + (v1 ? acc1[j][2] : 0.f), c1b = (v0 ? acc1[j][1] : 0.f) + (v1 ? acc1[j][3] : 0.f);
        float c2a = 0.f, c2b = 0.f;
        if (TR) {
          c2a = (v0 ? acc2[TR ? j : 0][0] : 0.f) + (v1 ? acc2[TR ? j : 0][2] : 0.f);
          c2b = (v0 ? acc2[TR ? j : 0][1] : 0.f) + (v1 ? acc2[TR ? j : 0][3] : 0.f);
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          c1a += __shfl_xor_sync(0xffffffffu, c1a, o); c1b += __shfl_xor_sync(0xffffffffu, c1b, o);
          if (TR) { c2a += __shfl_xor_sync(0xffffffffu, c2a, o); c2b += __shfl_xor_sync(0xffffffffu, c2b, o); }
        }
        if (g == 0) {
          const int col = h * D + j * 8 + t4 * 2;
          if (cs1) { atomicAdd(cs1 + col, c1a); atomicAdd(cs1 + col + 1, c1b); }
          if (TR && cs2) { atomicAdd(cs2 + col, c2a); atomicAdd(cs2 + col + 1, c2b); }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row_g + r * 8;
    if (row < Sr) {
      if (!TR) {
        bf16* orow = a.dq + ((size_t)b * a.Sq + row) * a.lddq + h * D;
#pragma unroll
        for (int j = 0; j < D / 8; ++j)
          *reinterpret_cast<uint32_t*>(orow + j * 8 + t4 * 2) = pack_bf16x2(acc1[j][2 * r], acc1[j][2 * r + 1]);
      } else {
        bf16* krow = a.dk + ((size_t)b * a.Sk + row) * a.lddk + h * D;
        bf16* vrow = a.dv + ((size_t)b * a.Sk + row) * a.lddv + h * D;
#pragma unroll
        for (int j = 0; j < D / 8; ++j) {
          *reinterpret_cast<uint32_t*>(krow + j * 8 + t4 * 2) = pack_bf16x2(acc1[j][2 * r], acc1[j][2 * r + 1]);
          *reinterpret_cast<uint32_t*>(vrow + j * 8 + t4 * 2) =
              pack_bf16x2(acc2[TR ? j : 0][2 * r], acc2[TR ? j : 0][2 * r + 1]);
        }
      }
    }
  }
  return false;
}

// stand-alone kernels: TR = false produces dQ (+ delta), TR = true produces dK / dV; one CTA per 64-row tile
template <int D, bool TR, bool DROP>
__global__ void __launch_bounds__(128) attn_bwd_kernel(AttnArgs a) {
  ECAMP_PDL_ENTRY();
  constexpr int LDS = D + 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int Sr = TR ? a.Sk : a.Sq;  // row-side length
  const int Sc = TR ? a.Sq : a.Sk;  // column-side length
  const int Scp = (Sc + 15) & ~15;
  bf16* sR1 = reinterpret_cast<bf16*>(smem_raw);
  bf16* sR2 = sR1 + 64 * LDS;
  bf16* sC1 = sR2 + 64 * LDS;
  bf16* sC2 = sC1 + (size_t)Scp * LDS;
  float* sColA = reinterpret_cast<float*>(sC2 + (size_t)Scp * LDS);  // !TR: key bias (0/-inf); TR: -lse * log2(e) per query
  float* sColB = sColA + Scp;                                        // TR: delta per query
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  const uint64_t bh = (uint64_t)b * a.H + h;
  const bf16* gQ = a.q + (size_t)b * a.Sq * a.ldq + h * D;
  const bf16* gK = a.k + (size_t)b * a.Sk * a.ldk + h * D;
  const bf16* gV = a.v + (size_t)b * a.Sk * a.ldv + h * D;
  const bf16* gdO = a.d_o + (size_t)b * a.Sq * a.ld_do + h * D;
  const int rvalid = min(64, Sr - r0);
  __shared__ int s_cend, s_cfull;
  if (tid == 0) { s_cend = TR ? Scp : 0; s_cfull = Scp; }
  __syncthreads();
  if (!TR) {
    load_tile<D, LDS>(sR1, gQ + (size_t)r0 * a.ldq, a.ldq, 64, rvalid, tid, 128);
    load_tile<D, LDS>(sR2, gdO + (size_t)r0 * a.ld_do, a.ld_do, 64, rvalid, tid, 128);
    load_tile<D, LDS>(sC1, gK, a.ldk, Scp, Sc, tid, 128);
    load_tile<D, LDS>(sC2, gV, a.ldv, Scp, Sc, tid, 128);
    for (int j = tid; j < Scp; j += 128) {
      bool ok = j < Sc;
      if (ok && a.key_mask) ok = a.key_mask[(size_t)b * a.Sk + j] != 0;
      sColA[j] = ok ? 0.f : -INFINITY;
      if (ok) atomicMax(&s_cend, j + 1);
      else atomicMin(&s_cfull, j);
    }
  } else {
    load_tile<D, LDS>(sR1, gK + (size_t)r0 * a.ldk, a.ldk, 64, rvalid, tid, 128);
    load_tile<D, LDS>(sR2, gV + (size_t)r0 * a.ldv, a.ldv, 64, rvalid, tid, 128);
    load_tile<D, LDS>(sC1, gQ, a.ldq, Scp, Sc, tid, 128);
    load_tile<D, LDS>(sC2, gdO, a.ld_do, Scp, Sc, tid, 128);
    for (int i = tid; i < Scp; i += 128) {
      const float l = i < Sc ? a.lse[bh * a.Sq + i] : -INFINITY;  // -inf: padded / fully masked query row -> p = 0
      sColA[i] = l == -INFINITY ? -INFINITY : -l * 1.4426950408889634f;
      sColB[i] = i < Sc ? a.delta[bh * a.Sq + i] : 0.f;
    }
  }
  load_tile_wait();
  __syncthreads();
  const int cend = TR ? Scp : ((s_cend + 15) & ~15);  // dQ pass: keys after the last attendable one are skipped
  const int cfull = TR ? Scp : (s_cfull & ~15);       // dQ pass: columns below this need no mask test
  attn_bwd_compute<D, TR, DROP, 4>(a, sR1, sR2, sC1, sC2, sColA, sColB, r0, cend, cfull, b, h);
}

// merged kernel: ONE CTA per (batch, head) with all of Q, dO, K, V (<= NW * 16 rows each) staged once; first the dQ
// problem (rows = queries), then - delta handed over through global memory, block-local - the transposed dK / dV
// problem (rows = keys) on the same tiles.  The two stand-alone kernels re-read K, V / Q, dO for every 64-row tile
// (8 x the bytes through L2 for the decoder's 197 x 197 heads), which is what bounded them.
template <int D, int NW, bool DROP>
__global__ void __launch_bounds__(NW * 32) attn_bwd_merged_kernel(AttnArgs a) {
  ECAMP_PDL_ENTRY();
  constexpr int LDS = D + 8, RT = NW * 16, NT_ = NW * 32;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sdO = sQ + RT * LDS;
  bf16* sK = sdO + RT * LDS;
  bf16* sV = sK + RT * LDS;
  float* sColA = reinterpret_cast<float*>(sV + RT * LDS);
  float* sColB = sColA + RT;
  const int tid = threadIdx.x;
  const int h = blockIdx.x, b = blockIdx.y;
  const uint64_t bh = (uint64_t)b * a.H + h;
  const int Skp = (a.Sk + 15) & ~15, Sqp = (a.Sq + 15) & ~15;
  __shared__ int s_cend, s_cfull;
  if (tid == 0) { s_cend = 0; s_cfull = Skp; }
  __syncthreads();
  load_tile<D, LDS>(sQ, a.q + (size_t)b * a.Sq * a.ldq + h * D, a.ldq, RT, a.Sq, tid, NT_);
  load_tile<D, LDS>(sdO, a.d_o + (size_t)b * a.Sq * a.ld_do + h * D, a.ld_do, RT, a.Sq, tid, NT_);
  load_tile<D, LDS>(sK, a.k + (size_t)b * a.Sk * a.ldk + h * D, a.ldk, RT, a.Sk, tid, NT_);
  load_tile<D, LDS>(sV, a.v + (size_t)b * a.Sk * a.ldv + h * D, a.ldv, RT, a.Sk, tid, NT_);
  for (int j = tid; j < Skp; j += NT_) {
    bool ok = j < a.Sk;
    if (ok && a.key_mask) ok = a.key_mask[(size_t)b * a.Sk + j] != 0;
    sColA[j] = ok ? 0.f : -INFINITY;
    if (ok) atomicMax(&s_cend, j + 1);
    else atomicMin(&s_cfull, j);
  }
  load_tile_wait();
  __syncthreads();
  attn_bwd_compute<D, false, DROP, NW>(a, sQ, sdO, sK, sV, sColA, sColB, 0, (s_cend + 15) & ~15, s_cfull & ~15, b, h);
  __syncthreads();  // delta of every query row of this head is in global memory (written by this CTA); sColA is free
  for (int i = tid; i < Sqp; i += NT_) {
    const float l = i < a.Sq ? a.lse[bh * a.Sq + i] : -INFINITY;
    sColA[i] = l == -INFINITY ? -INFINITY : -l * 1.4426950408889634f;
    sColB[i] = i < a.Sq ? a.delta[bh * a.Sq + i] : 0.f;
  }
  __syncthreads();
  attn_bwd_compute<D, true, DROP, NW>(a, sK, sV, sQ, sdO, sColA, sColB, 0, Sqp, Sqp, b, h);
}

constexpr int kMaxDynSmem = 226 * 1024;  // 227 KB per CTA minus the kernels' static shared variables
size_t fwd_smem(int D, int Sk, int rows = 64) {
  const int LDS = D + 8, Skp = (Sk + 15) & ~15;
  return (size_t)(rows + 2 * Skp) * LDS * 2 + (size_t)Skp * 4;
}
size_t bwd_smem(int D, int Sc) {
  const int LDS = D + 8, Scp = (Sc + 15) & ~15;
  return (size_t)(128 + 2 * Scp) * LDS * 2 + (size_t)Scp * 8;
}
constexpr int kBigNW = 13;  // 13 warps x 16 rows = 208 rows: one CTA covers a whole decoder head (S = 197)
size_t merged_smem(int D) { return (size_t)4 * kBigNW * 16 * (D + 8) * 2 + (size_t)kBigNW * 16 * 8; }

int check_args(const AttnArgs& a, bool bwd) {
  ECAMP_REQUIRE(a.D == 32 || a.D == 64 || a.D == 128, "attention: head_dim must be 32, 64 or 128 (got %d)", a.D);
  ECAMP_REQUIRE(a.B > 0 && a.H > 0 && a.Sq > 0 && a.Sk > 0, "attention: empty problem");
  ECAMP_REQUIRE(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 8 == 0,
                "attention: row pitches must be multiples of 8 elements");
  ECAMP_REQUIRE(a.drop.p >= 0.f && a.drop.p < 1.f, "attention: dropout p out of range");
  const size_t sm = bwd ? (bwd_smem(a.D, a.Sk) > bwd_smem(a.D, a.Sq) ? bwd_smem(a.D, a.Sk) : bwd_smem(a.D, a.Sq))
                        : fwd_smem(a.D, a.Sk);
  ECAMP_REQUIRE(sm <= (size_t)kMaxDynSmem, "attention: sequence too long for the single-pass kernel (smem %zu B)", sm);
  if (bwd) {
    ECAMP_REQUIRE(a.d_o && a.delta && a.lse && a.dq && a.dk && a.dv, "attention_bwd: missing pointers");
    ECAMP_REQUIRE(a.ld_do % 8 == 0 && a.lddq % 8 == 0 && a.lddk % 8 == 0 && a.lddv % 8 == 0,
                  "attention_bwd: row pitches must be multiples of 8 elements");
  }
  return 0;
}

template <int D, int NW>
int launch_fwd_nw(const AttnArgs& a, cudaStream_t st) {
  const size_t sm = fwd_smem(D, a.Sk, NW * 16);
  dim3 grid((a.Sq + NW * 16 - 1) / (NW * 16), a.H, a.B);
  if (a.drop.p > 0.f) {
    ECAMP_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel<D, NW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    ECAMP_CUDA_OK(launch_pdl(attn_fwd_kernel<D, NW, true>, grid, NW * 32, sm, st, a));
  } else {
    ECAMP_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel<D, NW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    ECAMP_CUDA_OK(launch_pdl(attn_fwd_kernel<D, NW, false>, grid, NW * 32, sm, st, a));
  }
  ECAMP_LAUNCHED();
  return 0;
}
template <int D>
int launch_fwd(const AttnArgs& a, cudaStream_t st) {
  // head_dim 32 (ViT decoder): one CTA per head when all its queries fit in 13 warps
  static const int big = getenv("ECAMP_ATTN_BIG") ? atoi(getenv("ECAMP_ATTN_BIG")) : 0;  // measured slower (1 CTA / SM)
  if (big && D == 32 && a.Sq > 64 && a.Sq <= kBigNW * 16 && fwd_smem(D, a.Sk, kBigNW * 16) <= (size_t)kMaxDynSmem)
    return launch_fwd_nw<32, kBigNW>(a, st);
  // (7 warps = two CTAs per 197-query decoder head, K / V staged twice instead of 4 times: measured slower, 171 vs 143 us)
  return launch_fwd_nw<D, 4>(a, st);
}
template <int D>
int launch_bwd(const AttnArgs& a, cudaStream_t st) {
  static const int small_merged = getenv("ECAMP_ATTN_SMALL_MERGED") ? atoi(getenv("ECAMP_ATTN_SMALL_MERGED")) : 1;
  if (small_merged && D <= 64 && a.Sq <= 64 && a.Sk <= 64) {
    // ViT encoder heads (50 x 50): dQ and dK/dV in ONE 128-thread CTA per head - one staging of Q, dO, K, V, one launch
    dim3 grid(a.H, a.B);
    const size_t sm = (size_t)4 * 64 * (D + 8) * 2 + 64 * 8;
    if (a.drop.p > 0.f) {
      ECAMP_CUDA_OK(cudaFuncSetAttribute(attn_bwd_merged_kernel<D, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
      ECAMP_CUDA_OK(launch_pdl(attn_bwd_merged_kernel<D, 4, true>, grid, 128, sm, st, a));
    } else {
      ECAMP_CUDA_OK(cudaFuncSetAttribute(attn_bwd_merged_kernel<D, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
      ECAMP_CUDA_OK(launch_pdl(attn_bwd_merged_kernel<D, 4, false>, grid, 128, sm, st, a));
    }
    ECAMP_LAUNCHED();
    return 0;
  }
  static const int big = getenv("ECAMP_ATTN_BIG") ? atoi(getenv("ECAMP_ATTN_BIG")) : 0;  // measured slower (1 CTA / SM)
  if (big && D == 32 && a.Sq <= kBigNW * 16 && a.Sk <= kBigNW * 16 && (a.Sq > 64 || a.Sk > 64)) {
    dim3 grid(a.H, a.B);
    if (a.drop.p > 0.f) {
      ECAMP_CUDA_OK(cudaFuncSetAttribute(attn_bwd_merged_kernel<32, kBigNW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
      ECAMP_CUDA_OK(launch_pdl(attn_bwd_merged_kernel<32, kBigNW, true>, grid, kBigNW * 32, merged_smem(32), st, a));
    } else {
      ECAMP_CUDA_OK(cudaFuncSetAttribute(attn_bwd_merged_kernel<32, kBigNW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
      ECAMP_CUDA_OK(launch_pdl(attn_bwd_merged_kernel<32, kBigNW, false>, grid, kBigNW * 32, merged_smem(32), st, a));
    }
    ECAMP_LAUNCHED();
    return 0;
  }
  dim3 gq((a.Sq + 63) / 64, a.H, a.B);
  dim3 gk((a.Sk + 63) / 64, a.H, a.B);
  if (a.drop.p > 0.f) {
    ECAMP_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel<D, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    ECAMP_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel<D, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    ECAMP_CUDA_OK(launch_pdl(attn_bwd_kernel<D, false, true>, gq, 128, bwd_smem(D, a.Sk), st, a));
    ECAMP_LAUNCHED();
    ECAMP_CUDA_OK(launch_pdl(attn_bwd_kernel<D, true, true>, gk, 128, bwd_smem(D, a.Sq), st, a));
    ECAMP_LAUNCHED();
  } else {
    ECAMP_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel<D, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    ECAMP_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel<D, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    ECAMP_CUDA_OK(launch_pdl(attn_bwd_kernel<D, false, false>, gq, 128, bwd_smem(D, a.Sk), st, a));
    ECAMP_LAUNCHED();
    ECAMP_CUDA_OK(launch_pdl(attn_bwd_kernel<D, true, false>, gk, 128, bwd_smem(D, a.Sq), st, a));
    ECAMP_LAUNCHED();
  }
  return 0;
}

}  // namespace

// tcgen05 kernels (attention_tc.cu)
bool attention_tc_fwd_supported(const AttnArgs& a);
bool attention_tc_bwd_supported(const AttnArgs& a);
int attention_tc_fwd(const AttnArgs& a, cudaStream_t st);
int attention_tc_bwd(const AttnArgs& a, cudaStream_t st);
static int g_attn_tc = [] {
  const char* e = getenv("ECAMP_ATTN_TC");
  return e ? atoi(e) : 1;
}();
void set_attention_tc(int on) { g_attn_tc = on; }

// =============================================================================================
// attention probabilities (inference tool): probs[b, h, i, j] = exp(scale * q_i . k_j - lse[b, h, i]), masked keys 0.
// Replaces `output_attentions=True` of the cross-attention in Visualization/module/context_fusion.py:45-57 (the
// heat-map tool reads cross_self_outputs[1]); the training kernels never materialise the probabilities.
// One CTA = 8 query rows of one (batch, head); the K tile of the head is staged once in shared memory.
// =============================================================================================
namespace {
constexpr int kProbRows = 8;
__global__ void __launch_bounds__(256) attn_probs_kernel(AttnArgs a, float* __restrict__ probs) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int D = a.D, LDK = D + 8;
  bf16* sK = reinterpret_cast<bf16*>(smem_raw);                       // [Sk][D + 8]
  float* sQ = reinterpret_cast<float*>(sK + (size_t)a.Sk * LDK);      // [8][D]
  float* sL = sQ + kProbRows * D;                                     // [8] lse
  const int tid = threadIdx.x, q0 = blockIdx.x * kProbRows, h = blockIdx.y, b = blockIdx.z;
  const uint64_t bh = (uint64_t)b * a.H + h;
  const int cpr = D / 8;
  for (int idx = tid; idx < a.Sk * cpr; idx += blockDim.x) {
    const int r = idx / cpr, c = idx % cpr;
    *reinterpret_cast<uint4*>(sK + (size_t)r * LDK + c * 8) =
        *reinterpret_cast<const uint4*>(a.k + ((size_t)b * a.Sk + r) * a.ldk + h * D + c * 8);
  }
  for (int idx = tid; idx < kProbRows * D; idx += blockDim.x) {
    const int r = idx / D, d = idx % D;
    sQ[idx] = q0 + r < a.Sq ? __bfloat162float(a.q[((size_t)b * a.Sq + q0 + r) * a.ldq + h * D + d]) : 0.f;
  }
  if (tid < kProbRows) sL[tid] = q0 + tid < a.Sq ? a.lse[bh * a.Sq + q0 + tid] : INFINITY;
  __syncthreads();
  for (int j = tid; j < a.Sk; j += blockDim.x) {
    float acc[kProbRows];
#pragma unroll
    for (int r = 0; r < kProbRows; ++r) acc[r] = 0.f;
    for (int c = 0; c < cpr; ++c) {
      const uint4 kv = *reinterpret_cast<const uint4*>(sK + (size_t)j * LDK + c * 8);
      const float2 k0 = unpack_bf16x2(kv.x), k1 = unpack_bf16x2(kv.y), k2 = unpack_bf16x2(kv.z), k3 = unpack_bf16x2(kv.w);
#pragma unroll
      for (int r = 0; r < kProbRows; ++r) {
        const float4 qa = *reinterpret_cast<const float4*>(sQ + r * D + c * 8);
        const float4 qb = *reinterpret_cast<const float4*>(sQ + r * D + c * 8 + 4);
        acc[r] += k0.x * qa.x + k0.y * qa.y + k1.x * qa.z + k1.y * qa.w + k2.x * qb.x + k2.y * qb.y + k3.x * qb.z + k3.y * qb.w;
      }
    }
    const bool ok = a.key_mask == nullptr || a.key_mask[(size_t)b * a.Sk + j] != 0;
#pragma unroll
    for (int r = 0; r < kProbRows; ++r) {
      if (q0 + r < a.Sq) {
        const float l = sL[r];
        probs[(bh * a.Sq + q0 + r) * (uint64_t)a.Sk + j] = (ok && l != -INFINITY) ? __expf(acc[r] * a.scale - l) : 0.f;
      }
    }
  }
}
}  // namespace

int attention_probs(const AttnArgs& a, float* probs, cudaStream_t st) {
  ECAMP_REQUIRE(a.q && a.k && a.lse && probs, "attention_probs: null pointer");
  ECAMP_REQUIRE(a.B > 0 && a.H > 0 && a.Sq > 0 && a.Sk > 0 && a.D % 8 == 0 && a.D >= 8 && a.D <= 128,
                "attention_probs: unsupported shape B=%d H=%d Sq=%d Sk=%d D=%d", a.B, a.H, a.Sq, a.Sk, a.D);
  ECAMP_REQUIRE(a.ldk % 8 == 0 && (reinterpret_cast<uintptr_t>(a.k) & 15) == 0, "attention_probs: K must be 16-byte aligned");
  const size_t sm = (size_t)a.Sk * (a.D + 8) * 2 + (size_t)kProbRows * a.D * 4 + kProbRows * 4;
  ECAMP_REQUIRE(sm <= (size_t)kMaxDynSmem, "attention_probs: %d keys do not fit in shared memory", a.Sk);
  ECAMP_CUDA_OK(cudaFuncSetAttribute(attn_probs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  dim3 grid((a.Sq + kProbRows - 1) / kProbRows, a.H, a.B);
  attn_probs_kernel<<<grid, 256, sm, st>>>(a, probs);
  ECAMP_CUDA_OK(cudaGetLastError());
  ECAMP_LAUNCHED();
  return 0;
}

int attention_fwd(const AttnArgs& a, cudaStream_t st) {
  if (int rc = check_args(a, false)) return rc;
  // measured (scripts/attn_time.py): ViT encoder heads (50 x 50, head_dim 64) 27 us on the mma.sync kernel, 39 us on tcgen05
  const bool small64 = a.D == 64 && a.Sq <= 64 && a.Sk <= 64;
  if (g_attn_tc && !small64 && attention_tc_fwd_supported(a)) return attention_tc_fwd(a, st);
  if (a.D == 32) return launch_fwd<32>(a, st);
  if (a.D == 64) return launch_fwd<64>(a, st);
  return launch_fwd<128>(a, st);
}
int attention_bwd(const AttnArgs& a, cudaStream_t st) {
  if (int rc = check_args(a, true)) return rc;
  if (g_attn_tc && attention_tc_bwd_supported(a)) return attention_tc_bwd(a, st);
  if (a.D == 32) return launch_bwd<32>(a, st);
  if (a.D == 64) return launch_bwd<64>(a, st);
  return launch_bwd<128>(a, st);
}

}  // namespace ecamp
