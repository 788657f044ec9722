// tcgen05 / TMEM / TMA GEMM for sm_100a.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128B swizzle, STAGES-deep mbarrier ring)
//   warp 1      MMA issuer     (one lane issues tcgen05.mma 128 x BN x 16, fp32 accumulators in TMEM)
//   warp 2      TMEM allocator (512 columns = two accumulator stages, so the epilogue of tile i
//                               overlaps the main loop of tile i+1)
//   warps 4-11  epilogue       (tcgen05.ld -> bias / GELU / dGELU / dropout / residual -> global)
//
// Operands may be K-major (the contraction index is contiguous in memory) or MN-major (the
// output index is contiguous); the latter is what dgrad (weights) and wgrad (both operands)
// need, so no transposed copies of activations or weights are ever made.
//
// Replaces the cuBLAS calls behind every nn.Linear of the reference hot path
// (ECAMP/Pre-training/module/model_ecamp.py:60-90, timm Block, HF BertLayer).
#include "gemm.cuh"

#include <cuda.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

namespace ecamp {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle span
// Warps 0-2: TMA producer, MMA issuer, TMEM allocator; then kNumEpiWarps epilogue warps (8: two per TMEM lane quarter,
// each draining half of the tile's columns in 32-column chunks; or 16: four per quarter, 16-column chunks, <= 96
// registers).  Measured on B200 (profiles/r01c_gemm_epilogue_study.md): both configurations run the GELU / dGELU /
// residual epilogues at the same speed, and so does a build whose epilogue skips its global stores - the fused
// epilogues are bound by SHARED-MEMORY bandwidth (operand reads of the MMAs + TMA writes + the transpose tile), not
// by issue slots or by the stores, so the 8-warp configuration (fewer instructions per element) is kept.
#ifndef ECAMP_NEPI
#define ECAMP_NEPI 8
#endif
constexpr int kNumEpiWarps = ECAMP_NEPI;           // 8 or 16
constexpr int kEpiWarp0 = 3;
constexpr int kThreads = (kEpiWarp0 + kNumEpiWarps) * 32;
constexpr int kCW = kNumEpiWarps == 16 ? 16 : 32;  // chunk width (columns per tcgen05.ld / transpose)
constexpr int kSlices = kNumEpiWarps / 4;          // column slices of a tile (one per group of 4 epilogue warps)
constexpr int kLPR = kCW / 4;                      // lanes per row in the coalesced layout (16 bytes each)
constexpr int kRPS = 32 / kLPR;                    // rows per step
constexpr int kSteps = 32 / kRPS;                  // steps per chunk (= kLPR)
constexpr int kStageFloats = 32 * kCW;             // transpose tile of one warp


constexpr int kBarBytes = 512;      // mbarriers + TMEM slot
constexpr int kTmaSlots = 3;        // TMA epilogue: per epilogue warp a ring of 32 x 32 fp32 tiles (4 KB each)
constexpr int kTmaSlotBytes = 4096;

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 128) ? 6 : 4;
  static constexpr int EPI_BYTES = kNumEpiWarps * kStageFloats * 4;  // one 32 x kCW fp32 transpose tile per epilogue warp
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + kBarBytes /*barriers*/ + 1024 /*alignment slack*/;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget of one CTA exceeded");
};

struct EpiArgs {
  GemmEpilogue ep;
  int vec_ok;
  int mode;     // EpiMode
  int split_k;  // > 1: the contraction is split over `split_k` work units per tile, partials added with fp32 atomics
  int kb_per;   // k-blocks per split
  int dbg;      // development: 1 = the epilogue only drains TMEM (no transpose / math / stores)
  int n_fast;   // tile order: 1 = consecutive work units walk the N tiles of one row block (A is the big operand:
                // its tile is fetched from DRAM once and served from L2 to the other column tiles), 0 = walk M
  int n_tiles;
  // fp32-accurate mode (gemm_hp): the contraction runs over `nterms` (A-plane, B-plane) pairs of bf16 x 3 split operands;
  // k-blocks are numbered virtually, v = term * num_kb + kb, and split over work units like a plain split-K.  1 otherwise.
  int nterms;
  unsigned char term_a[6], term_b[6];
  int stages;     // depth of the operand ring (what the shared memory left by the epilogue region holds)
  int epi_bytes;  // shared memory of the epilogue region
  int tma_epi;    // 1: fp32-output epilogue through TMA (residual tiles loaded, results stored by cp.async.bulk.tensor)
  int direct_epi;  // 1: bf16-output epilogue straight from the TMEM row-per-thread layout with 256-bit stores (no transpose)
};
struct TmaSet {  // one tensor map per operand plane (production: plane 0 only) + the fp32 epilogue operands
  CUtensorMap a[3], b[3];
  CUtensorMap res, out;  // tma_epi: fp32 [M, N] residual / output, box 32 x 32, 128-byte swizzle
};


// work unit -> (row block, column block)
ECAMP_DEVINL void decode_tile(const EpiArgs& ea, int tile, int m_tiles, int& m_blk, int& n_blk) {
  if (ea.n_fast) { n_blk = tile % ea.n_tiles; m_blk = tile / ea.n_tiles; }
  else { m_blk = tile % m_tiles; n_blk = tile / m_tiles; }
}

// ---------------------------------------------------------------------------------------------
// epilogue.  The accumulator arrives from TMEM with one ROW per thread (tcgen05.ld 32x32b); writing global memory in
// that layout touches 32 different 128-byte lines per warp instruction, and the LSU wavefront rate (one line per
// ~2 cycles) then bounds the whole GEMM (measured: 550-750 TFLOP/s on the fp32-residual shapes).  Each epilogue warp
// therefore transposes its 32 x 32 chunk through a private 4 KB swizzled shared-memory tile and works in the
// COALESCED layout: in step i (0..7) lane l owns row 4 i + l / 8, columns 4 (l % 8) .. +3, so a warp instruction
// covers 4 rows x 128 contiguous bytes (fp32) or 4 x 64 bytes (bf16).
// ---------------------------------------------------------------------------------------------
// The epilogue is specialised (inside one kernel, selected once per tile) for the operator combinations the step
// uses; anything else takes the generic path, which tests every flag per element group.  With only two epilogue
// warps per scheduler the per-element branches of the generic path are latency, not throughput, so they matter.
enum EpiMode : int {
  EM_GENERIC = 0,
  EM_BF16,          // (+bias) -> bf16
  EM_GELU,          // +bias -> bf16 pre-activation to aux_out -> GELU -> bf16
  EM_DGELU,         // * GELU'(aux_in) -> bf16
  EM_GELU_G,        // EM_GELU, but aux_out receives bf16(GELU'(pre-activation)) (GEMM_AUX_GRAD)
  EM_DGELU_G,       // * aux_in (the stored GELU') -> bf16 (GEMM_AUX_GRAD)
  EM_F32,           // (+bias) -> fp32
  EM_F32_RES,       // (+bias) + fp32 residual -> fp32
  EM_F32_RES_DROP,  // (+bias) -> dropout -> + fp32 residual -> fp32
  EM_ATOMIC,        // split-K partial: fp32 vector reduction into out_f32
};
__host__ __device__ constexpr bool is_gelu(int m) { return m == EM_GELU || m == EM_GELU_G; }
__host__ __device__ constexpr bool is_dgelu(int m) { return m == EM_DGELU || m == EM_DGELU_G; }

// scalar fall-back for one element (unaligned operands or a column tail that is not a multiple of 4)
ECAMP_DEVINL void epi_scalar(const EpiArgs& ea, float v, int row, int col, int N) {
  const GemmEpilogue& ep = ea.ep;
  if (ea.split_k != 1) { atomicAdd(ep.out_f32 + (size_t)row * ep.ld_f32 + col, v); return; }
  if (ep.bias) v += __ldg(ep.bias + col);
  if (ep.flags & GEMM_GELU) {
    v = bf2f(f2bf(v));
    float gr = v;
    if (ep.flags & GEMM_AUX_GRAD) gelu_erf_both(v, v, gr);
    else v = gelu_erf(v);
    if (ep.aux_out) ep.aux_out[(size_t)row * ep.ld_aux + col] = f2bf(gr);
  }
  if (ep.flags & GEMM_DGELU) {
    const float t = bf2f(ep.aux_in[(size_t)row * ep.ld_aux + col]);
    v *= (ep.flags & GEMM_AUX_GRAD) ? t : gelu_erf_grad(t);
  }
  if (ep.flags & GEMM_DROPOUT) {
    const Philox ph(ep.seed);
    const uint32_t w = philox_word(ph, (uint64_t)row * (uint64_t)N + (uint64_t)col, ep.stream);
    v = (w >= dropout_threshold(ep.drop_p)) ? v * (1.0f / (1.0f - ep.drop_p)) : 0.f;
  }
  if (ep.row_scale) v *= __ldg(ep.row_scale + row / ep.rows_per_scale);
  if (ep.residual) v += ep.residual[(size_t)row * ep.ld_res + col];
  if (ep.out_f32) ep.out_f32[(size_t)row * ep.ld_f32 + col] = v;
  if (ep.out_bf16) ep.out_bf16[(size_t)row * ep.ld_bf16 + col] = f2bf(v);
  if (ep.colsum_out) atomicAdd(ep.colsum_out + col, bf2f(f2bf(v)));
}

ECAMP_DEVINL float4 dropout4(const GemmEpilogue& ep, float4 v, int row, int col, int N) {
  const Philox ph(ep.seed);
  const uint32_t thr = dropout_threshold(ep.drop_p);
  const float scale = 1.0f / (1.0f - ep.drop_p);
  const uint64_t base = (uint64_t)row * (uint64_t)N + (uint64_t)col;  // N % 4 == 0 checked on the host
  const uint4 r = ph(base >> 2, ep.stream);
  v.x = (r.x >= thr) ? v.x * scale : 0.f;
  v.y = (r.y >= thr) ? v.y * scale : 0.f;
  v.z = (r.z >= thr) ? v.z * scale : 0.f;
  v.w = (r.w >= thr) ? v.w * scale : 0.f;
  return v;
}
ECAMP_DEVINL uint2 pack4(const float4& v) {
  uint2 u;
  u.x = pack_bf16x2(v.x, v.y); u.y = pack_bf16x2(v.z, v.w);
  return u;
}

// generic: one thread, 4 consecutive columns of one row, every operator tested at run time
ECAMP_DEVINL void epi_vec4_generic(const EpiArgs& ea, float4 v, int row, int col, int N, const float4& bias4) {
  const GemmEpilogue& ep = ea.ep;
  if (!ea.vec_ok || col + 4 > N) {
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (col + j < N) epi_scalar(ea, e[j], row, col + j, N);
    return;
  }
  if (ea.split_k != 1) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ep.out_f32 + (size_t)row * ep.ld_f32 + col),
                 "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
    return;
  }
  v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
  if (ep.flags & GEMM_GELU) {
    const uint2 u = pack4(v);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    if (ep.flags & GEMM_AUX_GRAD) {
      float4 gr;
      gelu_erf_both(a.x, v.x, gr.x); gelu_erf_both(a.y, v.y, gr.y); gelu_erf_both(b.x, v.z, gr.z); gelu_erf_both(b.y, v.w, gr.w);
      if (ep.aux_out) *reinterpret_cast<uint2*>(ep.aux_out + (size_t)row * ep.ld_aux + col) = pack4(gr);
    } else {
      if (ep.aux_out) *reinterpret_cast<uint2*>(ep.aux_out + (size_t)row * ep.ld_aux + col) = u;
      v.x = gelu_erf(a.x); v.y = gelu_erf(a.y); v.z = gelu_erf(b.x); v.w = gelu_erf(b.y);
    }
  }
  if (ep.flags & GEMM_DGELU) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(ep.aux_in + (size_t)row * ep.ld_aux + col));
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    if (ep.flags & GEMM_AUX_GRAD) { v.x *= a.x; v.y *= a.y; v.z *= b.x; v.w *= b.y; }
    else { v.x *= gelu_erf_grad(a.x); v.y *= gelu_erf_grad(a.y); v.z *= gelu_erf_grad(b.x); v.w *= gelu_erf_grad(b.y); }
  }
  if (ep.flags & GEMM_DROPOUT) v = dropout4(ep, v, row, col, N);
  if (ep.row_scale) {
    const float rs = __ldg(ep.row_scale + row / ep.rows_per_scale);
    v.x *= rs; v.y *= rs; v.z *= rs; v.w *= rs;
  }
  if (ep.residual) {
    const float4 r = *reinterpret_cast<const float4*>(ep.residual + (size_t)row * ep.ld_res + col);
    v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
  }
  if (ep.out_f32) *reinterpret_cast<float4*>(ep.out_f32 + (size_t)row * ep.ld_f32 + col) = v;
  if (ep.out_bf16) *reinterpret_cast<uint2*>(ep.out_bf16 + (size_t)row * ep.ld_bf16 + col) = pack4(v);
  if (ep.colsum_out) {
    const uint2 pk = pack4(v);
    const float2 a = unpack_bf16x2(pk.x), b = unpack_bf16x2(pk.y);
    atomicAdd(ep.colsum_out + col, a.x); atomicAdd(ep.colsum_out + col + 1, a.y);
    atomicAdd(ep.colsum_out + col + 2, b.x); atomicAdd(ep.colsum_out + col + 3, b.y);
  }
}

// what a specialised mode prefetches one chunk ahead (one 16-byte register slot per step)
// (EM_DGELU needs 8 bytes per step, so a slot holds TWO chunks - `hi` selects the half - and the pre-activation is
//  fetched two chunks ahead: that operand comes from HBM, one chunk of lead did not cover its latency)
template <int MODE>
ECAMP_DEVINL void epi_prefetch(const GemmEpilogue& ep, int row0, int col, int M, int N, uint4 (&p)[kSteps], bool hi = false) {
  if (MODE != EM_F32_RES && MODE != EM_F32_RES_DROP && !is_dgelu(MODE)) return;
  if (col >= N) return;
#pragma unroll
  for (int i = 0; i < kSteps; ++i) {
    const int row = row0 + kRPS * i;
    if (row < M) {
      if (is_dgelu(MODE)) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(ep.aux_in + (size_t)row * ep.ld_aux + col));
        if (hi) { p[i].z = u.x; p[i].w = u.y; }
        else { p[i].x = u.x; p[i].y = u.y; }
      } else {
        p[i] = *reinterpret_cast<const uint4*>(ep.residual + (size_t)row * ep.ld_res + col);
      }
    }
  }
}

ECAMP_DEVINL void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
ECAMP_DEVINL float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
// swizzle of the transpose tile: 16-byte unit `unit` of row `r` lives at unit ^ swz(r), which makes both the
// row-per-thread stores and the coalesced-layout loads conflict-free (rows are kCW * 4 = 128 or 64 bytes)
ECAMP_DEVINL int epi_swz(int r) { return kCW == 32 ? (r & 7) : ((r >> 1) & 3); }

// Pull the epilogue operand (fp32 residual or bf16 dGELU pre-activation) of a 32-row x COLS-column region into L2:
// issued one tile ahead, so that the register prefetch of epi_prefetch (one chunk ahead) only ever pays L2 latency.
template <int COLS, int MODE>
ECAMP_DEVINL void epi_prefetch_l2(const GemmEpilogue& ep, int m0, int ncol0, int M, int N, int lane) {
  if (MODE != EM_F32_RES && MODE != EM_F32_RES_DROP && !is_dgelu(MODE)) return;
  constexpr int ROW_BYTES = COLS * (is_dgelu(MODE) ? 2 : 4);
  constexpr int LINES = (ROW_BYTES + 127) / 128;  // 128-byte lines per row of the region
#pragma unroll
  for (int j = 0; j < LINES; ++j) {
    const int row = m0 + lane;
    const int col = ncol0 + j * (is_dgelu(MODE) ? 64 : 32);
    if (row < M && col < N && col < ncol0 + COLS) {
      const void* p = is_dgelu(MODE) ? static_cast<const void*>(ep.aux_in + (size_t)row * ep.ld_aux + col)
                                       : static_cast<const void*>(ep.residual + (size_t)row * ep.ld_res + col);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
  }
}

// The epilogue of one warp for one accumulator stage: TMEM lane quarter at `taddr`, columns [ncol0, ncol0 + COLS) of
// the tile whose first output row (of this quarter) is m0.  `stage` is the warp's private transpose tile.
// Arrives on `tmem_empty` as soon as the last TMEM read has landed.
template <int COLS, int MODE>
ECAMP_DEVINL void epilogue_warp(const EpiArgs& ea, uint32_t taddr, int m0, int ncol0, int M, int N, float* stage,
                                int lane, uint64_t* tmem_full, uint32_t full_phase, uint64_t* tmem_empty,
                                bool remote_empty, int next_m0, int next_ncol0) {
  constexpr int NCH = COLS / kCW;
  static_assert(COLS % kCW == 0, "column slice must be whole chunks");
  const GemmEpilogue& ep = ea.ep;
  const int sub = lane / kLPR, u = lane % kLPR;  // coalesced layout: step i -> row kRPS i + sub, 16-byte unit u
  const bool vbias = MODE != EM_ATOMIC && !is_dgelu(MODE) && ea.split_k == 1 && ep.bias && ea.vec_ok;
  const int row0 = m0 + sub;
  const uint32_t stage_addr = smem_u32(stage);  // explicit shared-space accesses (a generic pointer compiled to LD.E / ST.E)
  uint4 pcur[kSteps], pnext[kSteps];
  epi_prefetch<MODE>(ep, row0, ncol0 + 4 * u, M, N, pcur);  // in flight while the accumulator is still being produced
  if (is_dgelu(MODE) && NCH > 1) epi_prefetch<MODE>(ep, row0, ncol0 + kCW + 4 * u, M, N, pcur, true);
  if (next_m0 >= 0) epi_prefetch_l2<COLS, MODE>(ep, next_m0, next_ncol0, M, N, lane);
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), bias_next = bias4;
  if (vbias && ncol0 + 4 * u + 4 <= N) bias4 = __ldg(reinterpret_cast<const float4*>(ep.bias + ncol0 + 4 * u));
  mbar_wait(tmem_full, full_phase);
  tc_fence_after();
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    uint32_t raw[kCW];
    if (kCW == 32) tmem_ld_32x32(taddr + (uint32_t)(c * kCW), reinterpret_cast<uint32_t(&)[32]>(raw));
    else tmem_ld_32x16(taddr + (uint32_t)(c * kCW), reinterpret_cast<uint32_t(&)[16]>(raw));
    const int col = ncol0 + c * kCW + 4 * u;
    if (vbias && c + 1 < NCH && col + kCW + 4 <= N) bias_next = __ldg(reinterpret_cast<const float4*>(ep.bias + col + kCW));
    tmem_ld_wait();
    if (c == NCH - 1) {
      // all of this warp's TMEM reads for the tile are done: hand the accumulator stage back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (remote_empty) mbar_arrive_cluster(tmem_empty, 0);
        else mbar_arrive(tmem_empty);
      }
    }
    if (ea.dbg == 1) continue;
    // transpose: thread = row `lane` writes its kLPR 16-byte units, unit j at physical slot j ^ swz(lane)
#pragma unroll
    for (int j = 0; j < kLPR; ++j)
      sts128(stage_addr + (uint32_t)(lane * (kCW * 4) + ((j ^ epi_swz(lane)) << 4)), raw[4 * j], raw[4 * j + 1],
             raw[4 * j + 2], raw[4 * j + 3]);
    if (is_dgelu(MODE)) {
      if (c + 2 < NCH) epi_prefetch<MODE>(ep, row0, col + 2 * kCW, M, N, pnext);  // two chunks ahead
    } else if (c + 1 < NCH) {
      epi_prefetch<MODE>(ep, row0, col + kCW, M, N, pnext);  // next chunk's operand in flight from here
    }
    __syncwarp();
    float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);  // EM_DGELU: column sums of this chunk's emitted values
    constexpr int BATCH = kSteps < 4 ? kSteps : 4;   // steps per batch: keeps the live registers bounded
#pragma unroll
    for (int hb = 0; hb < kSteps / BATCH; ++hb) {
      float4 v[BATCH];
#pragma unroll
      for (int k = 0; k < BATCH; ++k) {
        const int r = kRPS * (BATCH * hb + k) + sub;
        v[k] = lds128(stage_addr + (uint32_t)(r * (kCW * 4) + ((u ^ epi_swz(r)) << 4)));
      }
      if (MODE == EM_GENERIC) {
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
          const int row = row0 + kRPS * (BATCH * hb + k);
          if (row < M && col < N) epi_vec4_generic(ea, v[k], row, col, N, bias4);
        }
      } else if (col < N) {  // specialised modes require N % 4 == 0 and 16-byte aligned operands (host-checked)
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
          const int i = BATCH * hb + k;
          const int row = row0 + kRPS * i;
          if (row < M) {
            float4 x = v[k];
            if (MODE == EM_ATOMIC) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ep.out_f32 + (size_t)row * ep.ld_f32 + col),
                           "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w)
                           : "memory");
              continue;
            }
            if (!is_dgelu(MODE)) { x.x += bias4.x; x.y += bias4.y; x.z += bias4.z; x.w += bias4.w; }
            if (is_gelu(MODE)) {
              const uint2 pk = pack4(x);
              const float2 a = unpack_bf16x2(pk.x), b = unpack_bf16x2(pk.y);
              if (MODE == EM_GELU_G) {  // backward only ever needs GELU'(pre-activation): store that instead
                float4 gr;
                gelu_erf_both(a.x, x.x, gr.x); gelu_erf_both(a.y, x.y, gr.y);
                gelu_erf_both(b.x, x.z, gr.z); gelu_erf_both(b.y, x.w, gr.w);
                *reinterpret_cast<uint2*>(ep.aux_out + (size_t)row * ep.ld_aux + col) = pack4(gr);
              } else {
                *reinterpret_cast<uint2*>(ep.aux_out + (size_t)row * ep.ld_aux + col) = pk;
                x.x = gelu_erf(a.x); x.y = gelu_erf(a.y); x.z = gelu_erf(b.x); x.w = gelu_erf(b.y);
              }
            }
            if (is_dgelu(MODE)) {
              const float2 a = unpack_bf16x2(pcur[i].x), b = unpack_bf16x2(pcur[i].y);
              if (MODE == EM_DGELU_G) { x.x *= a.x; x.y *= a.y; x.z *= b.x; x.w *= b.y; }
              else { x.x *= gelu_erf_grad(a.x); x.y *= gelu_erf_grad(a.y); x.z *= gelu_erf_grad(b.x); x.w *= gelu_erf_grad(b.y); }
            }
            if (MODE == EM_F32_RES_DROP) x = dropout4(ep, x, row, col, N);
            if (MODE == EM_F32_RES && ep.row_scale) {  // DropPath (fine-tune path only)
              const float rs = __ldg(ep.row_scale + row / ep.rows_per_scale);
              x.x *= rs; x.y *= rs; x.z *= rs; x.w *= rs;
            }
            if (MODE == EM_F32_RES || MODE == EM_F32_RES_DROP) {
              x.x += __uint_as_float(pcur[i].x); x.y += __uint_as_float(pcur[i].y);
              x.z += __uint_as_float(pcur[i].z); x.w += __uint_as_float(pcur[i].w);
            }
            if (MODE == EM_F32 || MODE == EM_F32_RES || MODE == EM_F32_RES_DROP) {
              *reinterpret_cast<float4*>(ep.out_f32 + (size_t)row * ep.ld_f32 + col) = x;
            } else {
              const uint2 pk = pack4(x);
              *reinterpret_cast<uint2*>(ep.out_bf16 + (size_t)row * ep.ld_bf16 + col) = pk;
              if (is_dgelu(MODE)) { csum.x += x.x; csum.y += x.y; csum.z += x.z; csum.w += x.w; }  // fp32, before rounding
            }
          }
        }
      }
    }
    if (is_dgelu(MODE) && ep.colsum_out) {
      // lanes that differ only in `sub` hold the same four columns (different rows): fold them, then one vector
      // reduction per 4 columns and 32-row slab
#pragma unroll
      for (int o = kLPR; o <= 16; o <<= 1) {
        csum.x += __shfl_xor_sync(0xffffffffu, csum.x, o);
        csum.y += __shfl_xor_sync(0xffffffffu, csum.y, o);
        csum.z += __shfl_xor_sync(0xffffffffu, csum.z, o);
        csum.w += __shfl_xor_sync(0xffffffffu, csum.w, o);
      }
      if (sub == 0 && col < N)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ep.colsum_out + col), "f"(csum.x), "f"(csum.y),
                     "f"(csum.z), "f"(csum.w)
                     : "memory");
    }
    __syncwarp();
    if (MODE == EM_F32_RES || MODE == EM_F32_RES_DROP) {
#pragma unroll
      for (int i = 0; i < kSteps; ++i) pcur[i] = pnext[i];
    }
    if (is_dgelu(MODE)) {
#pragma unroll
      for (int i = 0; i < kSteps; ++i) {
        pcur[i].x = pcur[i].z; pcur[i].y = pcur[i].w;    // chunk c + 1 becomes current
        pcur[i].z = pnext[i].x; pcur[i].w = pnext[i].y;  // chunk c + 2 moves up
      }
    }
    bias4 = bias_next;
  }
}

// The whole persistent epilogue loop of one warp, specialised per mode (the switch sits OUTSIDE the tile loop so that
// every mode is an independent region for the register allocator).  Work unit u covers tile u % num_tiles; the first
// output row of the CTA's 128-row slab is (tile % m_tiles) * m_stride + m_off.
template <int BN, int MODE>
ECAMP_DEVINL void epilogue_loop(const EpiArgs& ea, uint32_t tmem_base, int q, int slice, int unit0, int unit_step,
                                int num_units, int num_tiles, int m_tiles, int m_stride, int m_off, int M, int N,
                                float* stage, int lane, uint64_t* tmem_full, uint64_t* tmem_empty, bool remote_empty) {
  constexpr int COLS = BN / kSlices;
  int acc = 0;
  uint32_t acc_phase = 0;
  if (unit0 < num_units) {  // the first tile's operand: nobody prefetched it one tile ahead
    const int tile = unit0 % num_tiles;
    int mb, nb;
    decode_tile(ea, tile, m_tiles, mb, nb);
    epi_prefetch_l2<COLS, MODE>(ea.ep, mb * m_stride + m_off + q * 32, nb * BN + slice * COLS, M, N, lane);
  }
  for (int unit = unit0; unit < num_units; unit += unit_step) {
    const int tile = unit % num_tiles;
    int m_blk, n_blk;
    decode_tile(ea, tile, m_tiles, m_blk, n_blk);
    int next_m0 = -1, next_ncol0 = 0;
    if (unit + unit_step < num_units) {
      int mb, nb;
      decode_tile(ea, (unit + unit_step) % num_tiles, m_tiles, mb, nb);
      next_m0 = mb * m_stride + m_off + q * 32;
      next_ncol0 = nb * BN + slice * COLS;
    }
    epilogue_warp<COLS, MODE>(ea, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + slice * COLS),
                              m_blk * m_stride + m_off + q * 32, n_blk * BN + slice * COLS, M, N, stage, lane,
                              &tmem_full[acc], acc_phase, &tmem_empty[acc], remote_empty, next_m0, next_ncol0);
    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
  }
}
// ---------------------------------------------------------------------------------------------
// bf16-output epilogue WITHOUT the shared-memory transpose (EM_BF16 / EM_GELU_G / EM_DGELU_G).  tcgen05.ld hands every
// thread one row of the accumulator; 32 bf16 columns of a row are 64 contiguous bytes = two 256-bit stores (STG.256, new on
// sm_100).  A warp store then touches 32 lines, but with one FULL 32-byte sector each, so L2 / DRAM see only whole sectors;
// per 128 x 256 tile that is 64 store instructions (~4.2 k LSU cycles, under the 6.1 k cycles of the tile's MMAs at K = 768),
// while the transpose tile cost 262 KB of shared-memory traffic per tile on top of the 768 KB of operand traffic - and
// shared-memory bandwidth is what bounds these kernels.  dGELU reads its operand the same way (two 256-bit loads per row
// and chunk, fetched one chunk ahead) and reduces the column sums with a halving butterfly across the warp.
// ---------------------------------------------------------------------------------------------
ECAMP_DEVINL void stg256(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
ECAMP_DEVINL void ldg256(const void* p, uint32_t (&v)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}
template <int BN, int MODE>
ECAMP_DEVINL void epilogue_direct_loop(const EpiArgs& ea, uint32_t tmem_base, int q, int slice, int unit0, int unit_step,
                                       int num_units, int num_tiles, int m_tiles, int m_stride, int m_off, int M, int N, int lane,
                                       uint64_t* tmem_full, uint64_t* tmem_empty, bool remote_empty) {
  constexpr int COLS = BN / kSlices, NCH = COLS / 32;
  static_assert(kCW == 32, "the direct epilogue works on 32-column chunks");
  const GemmEpilogue& ep = ea.ep;
  int acc = 0;
  uint32_t acc_phase = 0;
  for (int unit = unit0; unit < num_units; unit += unit_step) {
    int m_blk, n_blk;
    decode_tile(ea, unit % num_tiles, m_tiles, m_blk, n_blk);
    const int row = m_blk * m_stride + m_off + q * 32 + lane;
    const int ncol0 = n_blk * BN + slice * COLS;
    const bool rvalid = row < M;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + slice * COLS);
    uint32_t aux_cur[16], aux_next[16];
    if (MODE == EM_DGELU_G) {  // operand of the first chunk in flight while the accumulator is still being produced
      if (rvalid && ncol0 + 32 <= N) {
        const bf16* ap = ep.aux_in + (size_t)row * ep.ld_aux + ncol0;
        ldg256(ap, reinterpret_cast<uint32_t(&)[8]>(aux_cur[0]));
        ldg256(ap + 16, reinterpret_cast<uint32_t(&)[8]>(aux_cur[8]));
      }
    }
    mbar_wait(&tmem_full[acc], acc_phase);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      uint32_t raw[32];
      tmem_ld_32x32(taddr + (uint32_t)(c * 32), raw);
      const int col0 = ncol0 + c * 32;
      const bool full = col0 + 32 <= N;  // N % 8 == 0 (host-checked): a partial chunk is handled in 8-column pieces
      if (MODE == EM_DGELU_G && c + 1 < NCH && rvalid && col0 + 64 <= N) {
        const bf16* ap = ep.aux_in + (size_t)row * ep.ld_aux + col0 + 32;
        ldg256(ap, reinterpret_cast<uint32_t(&)[8]>(aux_next[0]));
        ldg256(ap + 16, reinterpret_cast<uint32_t(&)[8]>(aux_next[8]));
      }
      tmem_ld_wait();
      if (c == NCH - 1) {  // all of this warp's TMEM reads for the tile are done: hand the accumulator stage back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (remote_empty) mbar_arrive_cluster(&tmem_empty[acc], 0);
          else mbar_arrive(&tmem_empty[acc]);
        }
      }
      if (col0 >= N) continue;
      float x[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(raw[j]);
      if (MODE != EM_DGELU_G && ep.bias) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (col0 + 4 * j < N) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + 4 * j));
            x[4 * j] += b4.x; x[4 * j + 1] += b4.y; x[4 * j + 2] += b4.z; x[4 * j + 3] += b4.w;
          }
        }
      }
      uint32_t out[16], aux[16];
      if (MODE == EM_GELU_G) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 a = unpack_bf16x2(pack_bf16x2(x[2 * j], x[2 * j + 1]));  // GELU of the bf16-rounded pre-activation
          float y0, y1, g0, g1;
          gelu_erf_both(a.x, y0, g0);
          gelu_erf_both(a.y, y1, g1);
          out[j] = pack_bf16x2(y0, y1);
          aux[j] = pack_bf16x2(g0, g1);
        }
      } else if (MODE == EM_DGELU_G) {
        if (!full && rvalid) {  // partial chunk at the right edge: fetch the operand in 16-byte pieces
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (col0 + 8 * j < N) {
              const uint4 u = __ldg(reinterpret_cast<const uint4*>(ep.aux_in + (size_t)row * ep.ld_aux + col0 + 8 * j));
              aux_cur[4 * j] = u.x; aux_cur[4 * j + 1] = u.y; aux_cur[4 * j + 2] = u.z; aux_cur[4 * j + 3] = u.w;
            }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 a = unpack_bf16x2(aux_cur[j]);
          x[2 * j] *= a.x;
          x[2 * j + 1] *= a.y;
          out[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) out[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
      }
      if (rvalid) {
        bf16* op = ep.out_bf16 + (size_t)row * ep.ld_bf16 + col0;
        if (full) {
          stg256(op, reinterpret_cast<uint32_t(&)[8]>(out[0]));
          stg256(op + 16, reinterpret_cast<uint32_t(&)[8]>(out[8]));
          if (MODE == EM_GELU_G) {
            bf16* xp = ep.aux_out + (size_t)row * ep.ld_aux + col0;
            stg256(xp, reinterpret_cast<uint32_t(&)[8]>(aux[0]));
            stg256(xp + 16, reinterpret_cast<uint32_t(&)[8]>(aux[8]));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (col0 + 8 * j < N) {
              *reinterpret_cast<uint4*>(op + 8 * j) = make_uint4(out[4 * j], out[4 * j + 1], out[4 * j + 2], out[4 * j + 3]);
              if (MODE == EM_GELU_G)
                *reinterpret_cast<uint4*>(ep.aux_out + (size_t)row * ep.ld_aux + col0 + 8 * j) =
                    make_uint4(aux[4 * j], aux[4 * j + 1], aux[4 * j + 2], aux[4 * j + 3]);
            }
        }
      }
      if (MODE == EM_DGELU_G && ep.colsum_out) {
        // column sums over the warp's 32 rows (fp32, before rounding): after the step with distance d a lane keeps the half of
        // its values selected by (lane & d); in the end lane l holds the total of column l
        if (!rvalid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = 0.f;
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
          const bool up = (lane & d) != 0;
#pragma unroll
          for (int i = 0; i < d; ++i) {
            const float lo = x[i], hi = x[i + d];
            const float send = up ? lo : hi, keep = up ? hi : lo;
            x[i] = keep + __shfl_xor_sync(0xffffffffu, send, d);
          }
        }
        if (col0 + lane < N) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(ep.colsum_out + col0 + lane), "f"(x[0]) : "memory");
      }
      if (MODE == EM_DGELU_G) {
#pragma unroll
        for (int j = 0; j < 16; ++j) aux_cur[j] = aux_next[j];
      }
    }
    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
  }
}

// ---------------------------------------------------------------------------------------------
// fp32-output epilogue through TMA (EM_F32 / EM_F32_RES / EM_F32_RES_DROP on the CTA-pair kernel).  tcgen05.ld hands every
// thread one ROW of the accumulator; instead of transposing to a coalesced layout, each epilogue warp keeps a ring of three
// 32 x 32 fp32 tiles in shared memory with the 128-byte swizzle of the tensor maps: the residual tile arrives by TMA (two
// chunks ahead, across tile boundaries - nothing waits in registers), thread r combines its row with row r of the tile in
// place (16-byte unit j of row r lives at j ^ (r & 7): conflict-free for row-per-thread access), and one lane sends the
// tile to global memory with a TMA store.  No global loads / stores go through the LSU any more.
// ---------------------------------------------------------------------------------------------
template <int BN, int MODE>
ECAMP_DEVINL void epilogue_tma_loop(const EpiArgs& ea, const TmaSet& maps, uint32_t tmem_base, int q, int slice, int unit0,
                                    int unit_step, int num_units, int num_tiles, int m_tiles, int m_stride, int m_off, int M,
                                    int N, uint8_t* slots, uint64_t* res_bar, int lane, uint64_t* tmem_full,
                                    uint64_t* tmem_empty, bool remote_empty) {
  constexpr bool RES = MODE != EM_F32;
  constexpr int COLS = BN / kSlices, NCH = COLS / 32;
  static_assert(kCW == 32, "the TMA epilogue works on 32-column chunks");
  const GemmEpilogue& ep = ea.ep;
  const uint32_t slot_base = smem_u32(slots);
  // chunk stream of this warp: (unit, c) for unit = unit0, unit0 + unit_step, ... and c = 0 .. NCH-1; chunks whose rows or
  // columns lie completely outside the matrix are skipped by both cursors
  auto coords = [&](int unit, int c, int& row0, int& col0) {
    int mb, nb;
    decode_tile(ea, unit % num_tiles, m_tiles, mb, nb);
    row0 = mb * m_stride + m_off + q * 32;
    col0 = nb * BN + slice * COLS + c * 32;
    return row0 < M && col0 < N;
  };
  uint32_t n_issued = 0, n_done = 0;  // residual loads issued / chunks completed
  int lu = unit0, lc = 0;            // load cursor
  auto issue_next_load = [&]() {     // lane 0 only
    while (lu < num_units) {
      int r0, c0;
      const bool act = coords(lu, lc, r0, c0);
      if (++lc == NCH) { lc = 0; lu += unit_step; }
      if (act) {
        const uint32_t s = n_issued % kTmaSlots;
        mbar_arrive_expect_tx(&res_bar[s], kTmaSlotBytes);
        tma_load_2d(slots + s * kTmaSlotBytes, &maps.res, &res_bar[s], c0, r0);
        ++n_issued;
        return;
      }
    }
  };
  if (RES && lane == 0) { issue_next_load(); issue_next_load(); }
  int acc = 0;
  uint32_t acc_phase = 0;
  const Philox ph(ep.seed);
  const uint32_t thr = dropout_threshold(ep.drop_p);
  const float keep_scale = 1.0f / (1.0f - ep.drop_p);
  for (int unit = unit0; unit < num_units; unit += unit_step) {
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + slice * COLS);
    mbar_wait(&tmem_full[acc], acc_phase);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      uint32_t raw[32];
      tmem_ld_32x32(taddr + (uint32_t)(c * 32), raw);
      int row0, col0;
      const bool act = coords(unit, c, row0, col0);
      tmem_ld_wait();
      if (c == NCH - 1) {  // all of this warp's TMEM reads for the tile are done: hand the accumulator stage back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (remote_empty) mbar_arrive_cluster(&tmem_empty[acc], 0);
          else mbar_arrive(&tmem_empty[acc]);
        }
      }
      if (!act) continue;
      const uint32_t s = n_done % kTmaSlots;
      const uint32_t sa = slot_base + s * kTmaSlotBytes + (uint32_t)lane * 128u;
      if (RES) {
        mbar_wait(&res_bar[s], (n_done / kTmaSlots) & 1u);
      } else {
        // the slot was the source of the store two chunks ago: that store must have read it
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
      }
      const int row = row0 + lane;
      float rs = 1.0f;
      if (MODE == EM_F32_RES && ep.row_scale && row < M) rs = __ldg(ep.row_scale + row / ep.rows_per_scale);  // DropPath
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t a = sa + (uint32_t)((j ^ (lane & 7)) << 4);
        float4 x = make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]), __uint_as_float(raw[4 * j + 2]),
                               __uint_as_float(raw[4 * j + 3]));
        const int col = col0 + 4 * j;
        if (ep.bias && col < N) {  // N % 4 == 0 (host-checked): a 4-column group is inside or outside as a whole
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
          x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
        }
        if (MODE == EM_F32_RES_DROP) {
          const uint4 r = ph((((uint64_t)row * (uint64_t)N + (uint64_t)col)) >> 2, ep.stream);
          x.x = (r.x >= thr) ? x.x * keep_scale : 0.f;
          x.y = (r.y >= thr) ? x.y * keep_scale : 0.f;
          x.z = (r.z >= thr) ? x.z * keep_scale : 0.f;
          x.w = (r.w >= thr) ? x.w * keep_scale : 0.f;
        }
        if (RES) {
          const float4 r4 = lds128(a);
          x.x = fmaf(x.x, rs, r4.x); x.y = fmaf(x.y, rs, r4.y); x.z = fmaf(x.z, rs, r4.z); x.w = fmaf(x.w, rs, r4.w);
        }
        sts128(a, __float_as_uint(x.x), __float_as_uint(x.y), __float_as_uint(x.z), __float_as_uint(x.w));
      }
      fence_proxy_async_smem();  // generic-proxy writes of the tile -> visible to the TMA store
      __syncwarp();
      ++n_done;
      if (lane == 0) {
        tma_store_2d(&maps.out, slots + s * kTmaSlotBytes, col0, row0);
        tma_store_commit();
        if (RES) {
          // the next load goes to the slot of the chunk BEFORE this one: its store must have read the tile
          tma_store_wait_read<1>();
          issue_next_load();
        }
      }
    }
    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
  }
  if (lane == 0) tma_store_wait_all();  // the stores read shared memory that dies with the CTA
}

template <int BN>
ECAMP_DEVINL void epilogue_tma_dispatch(const EpiArgs& ea, const TmaSet& maps, uint32_t tmem_base, int q, int slice, int unit0,
                                        int unit_step, int num_units, int num_tiles, int m_tiles, int m_stride, int m_off,
                                        int M, int N, uint8_t* slots, uint64_t* res_bar, int lane, uint64_t* tmem_full,
                                        uint64_t* tmem_empty, bool remote_empty) {
  if (ea.mode == EM_F32)
    epilogue_tma_loop<BN, EM_F32>(ea, maps, tmem_base, q, slice, unit0, unit_step, num_units, num_tiles, m_tiles, m_stride, m_off,
                                  M, N, slots, res_bar, lane, tmem_full, tmem_empty, remote_empty);
  else if (ea.mode == EM_F32_RES)
    epilogue_tma_loop<BN, EM_F32_RES>(ea, maps, tmem_base, q, slice, unit0, unit_step, num_units, num_tiles, m_tiles, m_stride,
                                      m_off, M, N, slots, res_bar, lane, tmem_full, tmem_empty, remote_empty);
  else
    epilogue_tma_loop<BN, EM_F32_RES_DROP>(ea, maps, tmem_base, q, slice, unit0, unit_step, num_units, num_tiles, m_tiles,
                                           m_stride, m_off, M, N, slots, res_bar, lane, tmem_full, tmem_empty, remote_empty);
}

template <int BN>
ECAMP_DEVINL void epilogue_dispatch(const EpiArgs& ea, uint32_t tmem_base, int q, int slice, int unit0, int unit_step,
                                    int num_units, int num_tiles, int m_tiles, int m_stride, int m_off, int M, int N,
                                    float* stage, int lane, uint64_t* tmem_full, uint64_t* tmem_empty,
                                    bool remote_empty) {
  if (ea.direct_epi) {
    if (ea.mode == EM_BF16)
      epilogue_direct_loop<BN, EM_BF16>(ea, tmem_base, q, slice, unit0, unit_step, num_units, num_tiles, m_tiles, m_stride, m_off, M, N,
                                        lane, tmem_full, tmem_empty, remote_empty);
    else if (ea.mode == EM_GELU_G)
      epilogue_direct_loop<BN, EM_GELU_G>(ea, tmem_base, q, slice, unit0, unit_step, num_units, num_tiles, m_tiles, m_stride, m_off, M,
                                          N, lane, tmem_full, tmem_empty, remote_empty);
    else
      epilogue_direct_loop<BN, EM_DGELU_G>(ea, tmem_base, q, slice, unit0, unit_step, num_units, num_tiles, m_tiles, m_stride, m_off, M,
                                           N, lane, tmem_full, tmem_empty, remote_empty);
    return;
  }
#define ECAMP_EPI_CASE(MODE_)                                                                                        \
  case MODE_:                                                                                                         \
    epilogue_loop<BN, MODE_>(ea, tmem_base, q, slice, unit0, unit_step, num_units, num_tiles, m_tiles, m_stride, m_off, \
                             M, N, stage, lane, tmem_full, tmem_empty, remote_empty);                               \
    break;
  switch (ea.mode) {
    ECAMP_EPI_CASE(EM_BF16)
    ECAMP_EPI_CASE(EM_GELU)
    ECAMP_EPI_CASE(EM_DGELU)
    ECAMP_EPI_CASE(EM_GELU_G)
    ECAMP_EPI_CASE(EM_DGELU_G)
    ECAMP_EPI_CASE(EM_F32)
    ECAMP_EPI_CASE(EM_F32_RES)
    ECAMP_EPI_CASE(EM_F32_RES_DROP)
    ECAMP_EPI_CASE(EM_ATOMIC)
    default:
      epilogue_loop<BN, EM_GENERIC>(ea, tmem_base, q, slice, unit0, unit_step, num_units, num_tiles, m_tiles, m_stride,
                                    m_off, M, N, stage, lane, tmem_full, tmem_empty, remote_empty);
  }
#undef ECAMP_EPI_CASE
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ TmaSet maps, int M, int N, int K, EpiArgs ea) {
  using C = Cfg<BN>;
  const int STAGES = ea.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * C::A_BYTES;
  float* s_epi = reinterpret_cast<float*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES + ea.epi_bytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + BM - 1) / BM;
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + BK - 1) / BK;
  const int num_vkb = num_kb * ea.nterms;         // virtual k-blocks (= num_kb in production)
  const int num_units = num_tiles * ea.split_k;  // work unit = (tile, k-split); unit % num_tiles = tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.b[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kNumEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped with the previous kernel's tail
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const int tile = unit % num_tiles, ks = unit / num_tiles;
        int m_blk, n_blk;
        decode_tile(ea, tile, m_tiles, m_blk, n_blk);
        const int kb0 = ks * ea.kb_per, kb1 = ea.dbg == 2 ? kb0 + 1 : min(num_vkb, kb0 + ea.kb_per);
        for (int vkb = kb0; vkb < kb1; ++vkb) {
          const int term = ea.nterms > 1 ? vkb / num_kb : 0, kb = vkb - term * num_kb;
          const CUtensorMap* tma_a = &maps.a[ea.term_a[term]];
          const CUtensorMap* tma_b = &maps.b[ea.term_b[term]];
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          uint8_t* a_dst = sA + stage * C::A_BYTES;
          uint8_t* b_dst = sB + stage * C::B_BYTES;
          if (!A_MN) {
            tma_load_2d(a_dst, tma_a, &full_bar[stage], kb * BK, m_blk * BM);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(a_dst + j * (BK * 128), tma_a, &full_bar[stage], m_blk * BM + j * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(b_dst, tma_b, &full_bar[stage], kb * BK, n_blk * BN);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(b_dst + j * (BK * 128), tma_b, &full_bar[stage], n_blk * BN + j * 64, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      if (lane == 0) {
        const int ks = unit / num_tiles;
        const int kb0 = ks * ea.kb_per, kb1 = ea.dbg == 2 ? kb0 + 1 : min(num_vkb, kb0 + ea.kb_per);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + stage * C::A_BYTES);
          const uint32_t b_base = smem_u32(sB + stage * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major:  rows of 128 B, 8-row groups 1024 B apart; 16 k-elements = 32 B inside the swizzle span.
            // MN-major: 64 mn-elements per 128 B row, one row per k, 8-k groups 1024 B apart (SBO),
            //           64-wide mn groups BK*128 B apart (LBO); 16 k-elements = 16 rows = 2048 B.
            const uint64_t adesc = A_MN ? umma_smem_desc_sw128(a_base + k * 2048, BK * 128, 1024)
                                        : umma_smem_desc_sw128(a_base + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? umma_smem_desc_sw128(b_base + k * 2048, BK * 128, 1024)
                                        : umma_smem_desc_sw128(b_base + k * 32, 16, 1024);
            umma_f16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (kb == kb1 - 1) umma_commit(&tmem_full[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue =====================
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const int half = (warp - kEpiWarp0) >> 2;  // which column slice of the tile
    float* stage4k = s_epi + (warp - kEpiWarp0) * kStageFloats;
    epilogue_dispatch<BN>(ea, tmem_base, q, half, blockIdx.x, gridDim.x, num_units, num_tiles, m_tiles, BM, 0, M, N, stage4k,
                          lane, tmem_full, tmem_empty, false);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): the two CTAs of a cluster own a 256 x BN tile.  Each CTA stages its own
// 128 rows of A and HALF of the B tile, so per MMA every SM pulls two thirds of the bytes of the 1-CTA kernel through
// L2 and the shared-memory ring holds 6-8 stages instead of 4-6 (the 1-CTA kernel measured ~43 % tensor-pipe activity,
// limited by operand delivery).  The leader (even) CTA issues every MMA; TMA completions of both CTAs are credited
// to the leader's "full" barrier; MMA commits are multicast to both CTAs' "empty" / "accumulator full" barriers; the
// epilogue warps of both CTAs (each draining its own 128 TMEM lanes) arrive on the leader's "accumulator empty".
// ---------------------------------------------------------------------------------------------
template <int BN>
struct Cfg2 {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 128) ? 8 : 6;
  static constexpr int EPI_BYTES = kNumEpiWarps * kStageFloats * 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + kBarBytes + 1024;
  static constexpr int TMA_EPI_BYTES = kNumEpiWarps * kTmaSlots * kTmaSlotBytes;  // 96 KB
  // operand ring of the TMA-epilogue variant: whatever fits next to the 96 KB of residual / output tiles
  static constexpr int TMA_STAGES = (227 * 1024 - 1024 - kBarBytes - TMA_EPI_BYTES) / STAGE_BYTES;
  static_assert(TMA_STAGES >= 3, "operand ring too shallow");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget of one CTA exceeded");
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ TmaSet maps, int M, int N, int K, EpiArgs ea) {
  using C = Cfg2<BN>;
  const int STAGES = ea.stages;
  constexpr int HB = BN / 2;  // B rows staged by each CTA
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * C::A_BYTES;
  float* s_epi = reinterpret_cast<float*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES + ea.epi_bytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* res_bar = reinterpret_cast<uint64_t*>(tmem_slot + 2);  // [kNumEpiWarps][kTmaSlots] (tma_epi)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int pair = (int)cluster_id_x();
  const int num_pairs = (int)(gridDim.x >> 1);
  const int m_tiles = (M + 2 * BM - 1) / (2 * BM);
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + BK - 1) / BK;
  const int num_vkb = num_kb * ea.nterms;  // virtual k-blocks (= num_kb in production)
  const int num_units = num_tiles * ea.split_k;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.b[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * kNumEpiWarps);
    }
    if (ea.tma_epi)
      for (int i = 0; i < kNumEpiWarps * kTmaSlots; ++i) mbar_init(&res_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();  // the peer's barriers must be initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped with the previous kernel's tail
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = pair; unit < num_units; unit += num_pairs) {
        const int tile = unit % num_tiles, ks = unit / num_tiles;
        int m_blk, n_blk;
        decode_tile(ea, tile, m_tiles, m_blk, n_blk);
        const int kb0 = ks * ea.kb_per, kb1 = ea.dbg == 2 ? kb0 + 1 : min(num_vkb, kb0 + ea.kb_per);
        const int m0 = m_blk * 2 * BM + (int)rank * BM;
        const int n0 = n_blk * BN + (int)rank * HB;
        for (int vkb = kb0; vkb < kb1; ++vkb) {
          const int term = ea.nterms > 1 ? vkb / num_kb : 0, kb = vkb - term * num_kb;
          const CUtensorMap* tma_a = &maps.a[ea.term_a[term]];
          const CUtensorMap* tma_b = &maps.b[ea.term_b[term]];
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
          uint8_t* a_dst = sA + stage * C::A_BYTES;
          uint8_t* b_dst = sB + stage * C::B_BYTES;
          if (!A_MN) {
            tma_load_2d_2sm(a_dst, tma_a, &full_bar[stage], kb * BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d_2sm(a_dst + j * (BK * 128), tma_a, &full_bar[stage], m0 + j * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d_2sm(b_dst, tma_b, &full_bar[stage], kb * BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < HB / 64; ++j)
              tma_load_2d_2sm(b_dst + j * (BK * 128), tma_b, &full_bar[stage], n0 + j * 64, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = pair; unit < num_units; unit += num_pairs) {
        const int ks = unit / num_tiles;
        const int kb0 = ks * ea.kb_per, kb1 = ea.dbg == 2 ? kb0 + 1 : min(num_vkb, kb0 + ea.kb_per);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + stage * C::A_BYTES);
          const uint32_t b_base = smem_u32(sB + stage * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = A_MN ? umma_smem_desc_sw128(a_base + k * 2048, BK * 128, 1024)
                                        : umma_smem_desc_sw128(a_base + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? umma_smem_desc_sw128(b_base + k * 2048, BK * 128, 1024)
                                        : umma_smem_desc_sw128(b_base + k * 32, 16, 1024);
            umma_f16_2sm(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage]);
          if (kb == kb1 - 1) umma_commit_2sm(&tmem_full[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int q = warp & 3;
    const int half = (warp - kEpiWarp0) >> 2;
    // the leader's MMA warp owns the accumulator hand-back: both CTAs arrive on ITS barrier
    if (ea.tma_epi) {
      const int ew = warp - kEpiWarp0;
      epilogue_tma_dispatch<BN>(ea, maps, tmem_base, q, half, pair, num_pairs, num_units, num_tiles, m_tiles, 2 * BM,
                                (int)rank * BM, M, N, reinterpret_cast<uint8_t*>(s_epi) + ew * kTmaSlots * kTmaSlotBytes,
                                res_bar + ew * kTmaSlots, lane, tmem_full, tmem_empty, true);
    } else {
      float* stage4k = s_epi + (warp - kEpiWarp0) * kStageFloats;
      epilogue_dispatch<BN>(ea, tmem_base, q, half, pair, num_pairs, num_units, num_tiles, m_tiles, 2 * BM, (int)rank * BM, M,
                            N, stage4k, lane, tmem_full, tmem_empty, true);
    }
  }

  tc_fence_before();
  cluster_sync_all();  // nobody may exit (or free TMEM) while the peer can still signal / read it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// bf16 matrix stored row-major [outer, inner] with a row pitch; box = [box_outer, box_inner = 64].
int make_tmap(CUtensorMap* map, const bf16* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
              uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  ECAMP_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  ECAMP_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "GEMM operand base must be 16-byte aligned");
  ECAMP_REQUIRE((pitch_elems * 2) % 16 == 0, "GEMM operand pitch must be a multiple of 16 bytes (got %llu elements)",
                (unsigned long long)pitch_elems);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {64, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ECAMP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (inner %llu outer %llu pitch %llu box %u)",
                (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_elems,
                box_outer);
  return 0;
}

// fp32 matrix [outer, inner] row-major with a row pitch; box = 32 x 32 elements (128-byte rows), 128-byte swizzle
int make_tmap_f32(CUtensorMap* map, const float* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems) {
  EncodeTiledFn fn = get_encode_fn();
  ECAMP_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ECAMP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp32) failed with %d (inner %llu outer %llu pitch %llu)", (int)r,
                (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_elems);
  return 0;
}

// 0 = automatic (CTA pairs whenever the problem has more than one 128-row tile), 1 = single-CTA kernel only,
// 2 = CTA-pair kernel always.  ECAMP_GEMM_CTA_PAIR overrides the default; ecamp_gemm_set_cta_pair() at run time.
// Default 0: once the epilogue stopped being the bound (coalesced stores through the shared-memory transpose), the
// pair kernel - each SM pulls 2/3 of the operand bytes of the single-CTA kernel through L2 - is faster on every shape
// of the step (profiles/r01c_gemm_tile_sweep.log: 8192^3 1411 vs 1289 TFLOP/s, BERT qkv 1222 vs 1020).
int g_cta_pair_mode = [] {
  const char* e = getenv("ECAMP_GEMM_CTA_PAIR");
  return e ? atoi(e) : 0;
}();

int g_direct_epilogue = [] {
  const char* e = getenv("ECAMP_GEMM_DIRECT_EPI");
  return e ? atoi(e) : 1;
}();
int g_tma_epilogue = [] {
  const char* e = getenv("ECAMP_GEMM_TMA_EPI");
  return e ? atoi(e) : 0;
}();

// tests: route every GEMM through the generic (run-time checked) epilogue
int g_force_generic_epilogue = [] {
  const char* e = getenv("ECAMP_GEMM_GENERIC_EPILOGUE");
  return e ? atoi(e) : 0;
}();

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Tile width and k-split minimising (waves) x (k-blocks per unit + fixed per-unit cost) x (cost per k-block).
// split_k > 1 is only offered to plain fp32-output GEMMs (the weight gradients), whose output tiles are few
// (e.g. 768 x 768 -> 18 tiles on 148 SMs) while the contraction (all rows of the batch) is long.
void pick_config(int M, int N, int K, bool splittable, int force_bn, bool cta2, bool b_mn, int* bn_out,
                 int* split_out) {
  const int sms = cta2 ? num_sms() / 2 : num_sms();            // schedulable units: CTA pairs or CTAs
  const int m_tiles = cta2 ? (M + 2 * BM - 1) / (2 * BM) : (M + BM - 1) / BM;
  const int num_kb = (K + BK - 1) / BK;
  const int cand[3] = {256, 192, 128};
  // narrower tiles move more operand bytes through L2 and shared memory per MMA; measured on B200 at 8192^3
  // (profiles/r01c_gemm_tile_sweep.log): pair kernel 1411 / 1152 / 754 TFLOP/s, single-CTA kernel 1289 / 1061 / ~800
  const float eff1[3] = {1.00f, 0.85f, 0.65f}, eff2[3] = {1.00f, 0.82f, 0.55f};
  const float* eff = cta2 ? eff2 : eff1;
  const int splits[9] = {1, 2, 3, 4, 6, 8, 12, 16, 24};
  float best_cost = 1e30f;
  *bn_out = 256;
  *split_out = 1;
  for (int i = 0; i < 3; ++i) {
    const int bn = cand[i];
    if (force_bn && bn != force_bn) continue;
    if (cta2 && b_mn && bn == 192) continue;  // each CTA stages BN/2 columns of an MN-major B in 64-wide groups
    const int tiles = m_tiles * ((N + bn - 1) / bn);
    for (int j = 0; j < (splittable ? 9 : 1); ++j) {
      const int sp = splits[j];
      const int kb_per = (num_kb + sp - 1) / sp;
      if (sp > 1 && (kb_per < 8 || (sp - 1) * kb_per >= num_kb)) continue;
      const int waves = (tiles * sp + sms - 1) / sms;
      const float fixed = sp > 1 ? 14.f : 6.f;  // pipeline fill + epilogue, in k-block units (atomics cost more)
      const float cost = (float)waves * ((float)kb_per + fixed) * (float)bn / eff[i];
      if (cost < best_cost) { best_cost = cost; *bn_out = bn; *split_out = sp; }
    }
  }
}

template <int BN, bool A_MN, bool B_MN>
int launch(const TmaSet& maps, int M, int N, int K, const EpiArgs& ea, cudaStream_t st) {
  auto kfn = gemm_tcgen05_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    ECAMP_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES));
    attr_set = true;
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN) * ea.split_k;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  EpiArgs e2 = ea;
  e2.stages = Cfg<BN>::STAGES; e2.epi_bytes = Cfg<BN>::EPI_BYTES; e2.tma_epi = 0;
  ECAMP_CUDA_OK(launch_pdl(kfn, grid, kThreads, Cfg<BN>::SMEM_BYTES, st, maps, M, N, K, e2));
  ECAMP_LAUNCHED();
  return 0;
}

template <int BN, bool A_MN, bool B_MN>
int launch2(const TmaSet& maps, int M, int N, int K, const EpiArgs& ea, cudaStream_t st) {
  auto kfn = gemm_tcgen05_2cta_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    ECAMP_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  EpiArgs e2 = ea;
  e2.stages = ea.tma_epi ? Cfg2<BN>::TMA_STAGES : Cfg2<BN>::STAGES;
  e2.epi_bytes = ea.tma_epi ? Cfg2<BN>::TMA_EPI_BYTES : Cfg2<BN>::EPI_BYTES;
  const int smem_bytes = e2.stages * Cfg2<BN>::STAGE_BYTES + e2.epi_bytes + kBarBytes + 1024;
  const int units = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN) * ea.split_k;
  // ECAMP_GEMM_GRID_MULT (measurement knob, default 1 = one persistent CTA pair per two SMs): with k > 1 the tile list is
  // dealt to k x 74 pairs, of which 74 are resident at a time - the hardware scheduler then hands the later pairs to whichever
  // SMs come free first, a coarse dynamic schedule for when another kernel (an all-reduce) holds some of the SMs
  static const int grid_mult = getenv("ECAMP_GEMM_GRID_MULT") ? atoi(getenv("ECAMP_GEMM_GRID_MULT")) : 1;
  const int max_pairs = (num_sms() / 2) * (grid_mult > 0 ? grid_mult : 1);
  const int pairs = units < max_pairs ? units : max_pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled();
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  ECAMP_CUDA_OK(cudaLaunchKernelEx(&cfg, kfn, maps, M, N, K, e2));
  ECAMP_LAUNCHED();
  return 0;
}

template <int BN>
int launch_major2(int a_mn, int b_mn, const TmaSet& maps, int M, int N, int K, const EpiArgs& ea, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch2<BN, false, false>(maps, M, N, K, ea, st);
  if (!a_mn && b_mn) return launch2<BN, false, true>(maps, M, N, K, ea, st);
  if (a_mn && b_mn) return launch2<BN, true, true>(maps, M, N, K, ea, st);
  return launch2<BN, true, false>(maps, M, N, K, ea, st);
}

template <int BN>
int launch_major(int a_mn, int b_mn, const TmaSet& maps, int M, int N, int K, const EpiArgs& ea, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch<BN, false, false>(maps, M, N, K, ea, st);
  if (!a_mn && b_mn) return launch<BN, false, true>(maps, M, N, K, ea, st);
  if (a_mn && b_mn) return launch<BN, true, true>(maps, M, N, K, ea, st);
  return launch<BN, true, false>(maps, M, N, K, ea, st);
}

}  // namespace

int make_tmap_bf16(void* map, const bf16* ptr, unsigned long long inner, unsigned long long outer,
                   unsigned long long pitch_elems, unsigned box_outer) {
  return make_tmap(static_cast<CUtensorMap*>(map), ptr, inner, outer, pitch_elems, box_outer);
}

namespace {
// The launch shared by the production GEMM (one plane per operand) and the fp32-accurate one (three bf16 planes per
// operand, six plane pairs accumulated; `plane_a` / `plane_b` = elements between consecutive planes, 0 = production).
int gemm_launch(const bf16* A, size_t plane_a, int lda, int a_mn, const bf16* B, size_t plane_b, int ldb, int b_mn, int M,
                int N, int K, const GemmEpilogue& ep, int force_bn, bool hp, cudaStream_t stream) {
  ECAMP_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem %d x %d x %d", M, N, K);
  ECAMP_REQUIRE(ep.out_f32 || ep.out_bf16, "gemm: no output given");
  if (ep.flags & GEMM_DROPOUT)
    ECAMP_REQUIRE(N % 4 == 0 && ep.drop_p >= 0.f && ep.drop_p < 1.f, "gemm: dropout needs N %% 4 == 0, 0 <= p < 1");
  if (ep.flags & GEMM_DGELU) ECAMP_REQUIRE(ep.aux_in != nullptr, "gemm: dGELU needs aux_in");
  if (ep.colsum_out)
    ECAMP_REQUIRE(ep.out_bf16 && (reinterpret_cast<uintptr_t>(ep.colsum_out) & 15) == 0,
                  "gemm: colsum_out needs a bf16 output and a 16-byte aligned destination");
  ECAMP_REQUIRE(force_bn == 0 || force_bn == 128 || force_bn == 192 || force_bn == 256, "gemm: unsupported tile N %d",
                force_bn);
  const bool accumulate = ep.residual != nullptr && ep.residual == ep.out_f32 && ep.ld_res == ep.ld_f32;
  const bool splittable = ep.out_f32 && !ep.out_bf16 && !ep.bias && ep.flags == 0 && !ep.aux_out && !ep.colsum_out && !ep.row_scale &&
                          (ep.residual == nullptr || accumulate);
  int bn = 256, split_k = 1;
  const bool cta2 = g_cta_pair_mode == 2 || (g_cta_pair_mode == 0 && M > BM);
  const int num_kb = (K + BK - 1) / BK;
  const int nterms = hp ? 6 : 1;
  if (!hp) {
    pick_config(M, N, K, splittable, force_bn, cta2, b_mn != 0, &bn, &split_k);
  } else {
    // fp32-accurate mode: the tensor core accumulates with truncation, which shows as a bias that grows with the length
    // of the accumulation (measured, scripts/tc_accum_probe.py: 7e-5 relative at 6 x 32768 all-positive products,
    // 1e-6 when every <= 1024 products are summed outside in fp32).  So the 6 x K contraction is cut into units of at
    // most 16 k-blocks whose partial sums are added with fp32 atomics (round to nearest) into a zeroed accumulator.
    ECAMP_REQUIRE(splittable && !accumulate, "gemm_hp: the tensor-core pass takes a plain fp32 accumulator");
    bn = force_bn ? force_bn : 256;
    split_k = (nterms * num_kb + 15) / 16;
  }
  ECAMP_REQUIRE(!(cta2 && b_mn && bn == 192), "gemm: tile N 192 is not available to the CTA-pair kernel with an MN-major B");

  TmaSet maps;
  int rc;
  // K-major operand: inner = contraction, outer = rows, box [rows_tile, 64]
  // MN-major operand: inner = output index, outer = contraction, box [64 (k), 64]
  for (int pl = 0; pl < (hp ? 3 : 1); ++pl) {
    const bf16* Ap = A + pl * plane_a;
    const bf16* Bp = B + pl * plane_b;
    rc = a_mn ? make_tmap(&maps.a[pl], Ap, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BK)
              : make_tmap(&maps.a[pl], Ap, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BM);
    if (rc) return rc;
    rc = b_mn ? make_tmap(&maps.b[pl], Bp, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BK)
              : make_tmap(&maps.b[pl], Bp, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, (uint32_t)(cta2 ? bn / 2 : bn));
    if (rc) return rc;
  }
  if (!hp) { maps.a[1] = maps.a[2] = maps.a[0]; maps.b[1] = maps.b[2] = maps.b[0]; }

  EpiArgs ea;
  ea.ep = ep;
  ea.split_k = split_k;
  ea.nterms = nterms;
  // plane pairs of x = hi + mid + lo (8 mantissa bits each): everything down to 2^-24 relative
  static const unsigned char TA[6] = {0, 0, 1, 1, 0, 2}, TB[6] = {0, 1, 0, 1, 2, 0};
  for (int t = 0; t < 6; ++t) { ea.term_a[t] = hp ? TA[t] : 0; ea.term_b[t] = hp ? TB[t] : 0; }
  ea.kb_per = hp ? 16 : (num_kb + split_k - 1) / split_k;
  if (split_k > 1 || hp) {
    ea.ep.residual = nullptr;  // partial sums are added atomically on top of the running gradient / zeros
    if (!accumulate)
      ECAMP_CUDA_OK(cudaMemset2DAsync(ep.out_f32, (size_t)ep.ld_f32 * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M,
                                      stream));
  }
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  ea.vec_ok = 1;
  if (ep.bias && !al16(ep.bias)) ea.vec_ok = 0;
  if ((ep.aux_in || ep.aux_out) && (ep.ld_aux % 8 != 0 || !al16(ep.aux_in) || !al16(ep.aux_out))) ea.vec_ok = 0;
  if (ep.residual && (ep.ld_res % 4 != 0 || !al16(ep.residual))) ea.vec_ok = 0;
  if (ep.out_f32 && (ep.ld_f32 % 4 != 0 || !al16(ep.out_f32))) ea.vec_ok = 0;
  if (ep.out_bf16 && (ep.ld_bf16 % 8 != 0 || !al16(ep.out_bf16))) ea.vec_ok = 0;

  ea.mode = EM_GENERIC;
  if (ea.vec_ok && N % 4 == 0) {
    const bool only_bf16 = ep.out_bf16 && !ep.out_f32, only_f32 = ep.out_f32 && !ep.out_bf16;
    if (split_k > 1 || hp) ea.mode = EM_ATOMIC;
    else if (ep.flags == 0 && !ep.aux_out && !ep.residual && only_bf16) ea.mode = EM_BF16;
    else if (ep.flags == GEMM_GELU && ep.aux_out && !ep.residual && only_bf16) ea.mode = EM_GELU;
    else if (ep.flags == GEMM_DGELU && !ep.bias && !ep.aux_out && !ep.residual && only_bf16) ea.mode = EM_DGELU;
    else if (ep.flags == (GEMM_GELU | GEMM_AUX_GRAD) && ep.aux_out && !ep.residual && only_bf16) ea.mode = EM_GELU_G;
    else if (ep.flags == (GEMM_DGELU | GEMM_AUX_GRAD) && !ep.bias && !ep.aux_out && !ep.residual && only_bf16) ea.mode = EM_DGELU_G;
    else if (ep.flags == 0 && !ep.aux_out && !ep.residual && only_f32) ea.mode = EM_F32;
    else if (ep.flags == 0 && !ep.aux_out && ep.residual && only_f32) ea.mode = EM_F32_RES;
    else if (ep.flags == GEMM_DROPOUT && !ep.aux_out && ep.residual && only_f32) ea.mode = EM_F32_RES_DROP;
    if (ep.colsum_out && !is_dgelu(ea.mode)) ea.mode = EM_GENERIC;  // only the dGELU mode folds the column sums in
    if (ep.row_scale && ea.mode != EM_F32_RES) ea.mode = EM_GENERIC;
  }
  if (g_force_generic_epilogue) ea.mode = EM_GENERIC;
  // fp32 outputs of the CTA-pair kernel can leave through TMA (ecamp_gemm_set_tma_epilogue / ECAMP_GEMM_TMA_EPI=1).  OFF by
  // default: measured on B200 (profiles/r02c_gemm_tma_epilogue_ab.md) the residual tiles' extra trips through shared
  // memory (TMA write, row read, row write, TMA read - the transpose path makes two) and the operand ring shrinking from
  // 6 to 4 stages cost more than the LSU traffic they remove on every shape but the dropout+residual ones.
  // bf16 outputs straight from the TMEM row-per-thread layout with 256-bit stores (ecamp_gemm_set_direct_epilogue /
  // ECAMP_GEMM_DIRECT_EPI): needs 32-byte aligned rows
  ea.direct_epi = 0;
  // Measured (profiles/r02h_gemm_direct_epilogue_ab.md): the plain bf16 epilogue gains 2 - 5 % (vocabulary projection 1298 ->
  // 1365 TFLOP/s), the GELU (two outputs: four 256-bit stores per row and chunk) and dGELU ones lose 15 - 25 % - their LSU
  // wavefronts then outweigh the shared-memory traffic saved.  Mode 1 (default) = plain bf16 only, 2 = all three.
  if (g_direct_epilogue && kCW == 32 && N % 8 == 0 &&
      (ea.mode == EM_BF16 || (g_direct_epilogue >= 2 && (ea.mode == EM_GELU_G || ea.mode == EM_DGELU_G)))) {
    auto al32 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; };
    bool ok = al32(ep.out_bf16) && ep.ld_bf16 % 16 == 0;
    if (ea.mode != EM_BF16) ok = ok && ep.ld_aux % 16 == 0 && al32(ea.mode == EM_GELU_G ? (const void*)ep.aux_out : (const void*)ep.aux_in);
    ea.direct_epi = ok ? 1 : 0;
  }
  ea.tma_epi = 0;
  if (g_tma_epilogue && cta2 && kCW == 32 && (ea.mode == EM_F32 || ea.mode == EM_F32_RES || ea.mode == EM_F32_RES_DROP)) {
    ea.tma_epi = 1;
    rc = make_tmap_f32(&maps.out, ep.out_f32, (uint64_t)N, (uint64_t)M, (uint64_t)ep.ld_f32);
    if (rc) return rc;
    if (ea.mode != EM_F32) {
      rc = make_tmap_f32(&maps.res, ep.residual, (uint64_t)N, (uint64_t)M, (uint64_t)ep.ld_res);
      if (rc) return rc;
    } else {
      maps.res = maps.out;
    }
  } else {
    maps.res = maps.a[0]; maps.out = maps.a[0];
  }
  // (hp with split_k == 1, i.e. K <= 170: one unit per tile, the generic / scalar paths store into the zeroed accumulator)
  static const int dbg = getenv("ECAMP_GEMM_DBG") ? atoi(getenv("ECAMP_GEMM_DBG")) : 0;
  ea.dbg = dbg;
  // Tile order.  The operand that is larger than L2 can hold next to the output stream should be fetched from DRAM
  // once: walking the column tiles of one row block first keeps the A tile in L2 for its other column tiles
  // (ncu on 32768 x 768 x 768 + fp32 residual, row-block order: 268 MB read for 150 MB of operands - A came in once
  // per column tile).  Split-K units keep the row-block order.
  static const int order = getenv("ECAMP_GEMM_ORDER") ? atoi(getenv("ECAMP_GEMM_ORDER")) : -1;
  ea.n_tiles = (N + bn - 1) / bn;
  ea.n_fast = order >= 0 ? order : ((split_k == 1 && (long long)M > (long long)N) ? 1 : 0);

#ifdef ECAMP_EPI_ONLY_MODE
  return launch<256, false, false>(maps, M, N, K, ea, stream);
#endif
  if (cta2) {
    if (bn == 256) return launch_major2<256>(a_mn, b_mn, maps, M, N, K, ea, stream);
    if (bn == 192) return launch_major2<192>(a_mn, b_mn, maps, M, N, K, ea, stream);
    return launch_major2<128>(a_mn, b_mn, maps, M, N, K, ea, stream);
  }
  if (bn == 256) return launch_major<256>(a_mn, b_mn, maps, M, N, K, ea, stream);
  if (bn == 192) return launch_major<192>(a_mn, b_mn, maps, M, N, K, ea, stream);
  return launch_major<128>(a_mn, b_mn, maps, M, N, K, ea, stream);
}

// ---------------------------------------------------------------------------------------------
// fp32-accurate mode: operand split and fp32 epilogue
// ---------------------------------------------------------------------------------------------
// x (fp32 [R, C], pitch ld) -> three bf16 planes [3][R][Cp]: hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid);
// hi + mid + lo reproduces x to 2^-24 relative (the subtractions are exact in fp32).  Columns C..Cp-1 are zero.
__global__ void split3_kernel(const float* __restrict__ x, int R, int C, int ld, int Cp, bf16* __restrict__ planes) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)R * Cp) return;
  const int r = (int)(i / Cp), c = (int)(i % Cp);
  const float v = c < C ? x[(size_t)r * ld + c] : 0.f;
  const bf16 h = f2bf(v);
  const float r1 = v - bf2f(h);
  const bf16 m = f2bf(r1);
  const bf16 l = f2bf(r1 - bf2f(m));
  const size_t plane = (size_t)R * Cp;
  planes[i] = h; planes[plane + i] = m; planes[2 * plane + i] = l;
}

// The epilogue operators of the production kernel (epi_scalar above) on the fp32 accumulator, without any 16-bit rounding
// and with the exact erf GELU: bias -> GELU (aux_out = pre-activation or GELU') -> x GELU'(aux_in) -> dropout -> row scale
// -> + residual -> outputs (+ column sums of the emitted activation).
__global__ void hp_epilogue_kernel(const float* __restrict__ acc, int ldacc, int M, int N, GemmEpilogueT<float> ep) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)M * N) return;
  const int row = (int)(i / N), col = (int)(i % N);
  float v = acc[(size_t)row * ldacc + col];
  if (ep.bias) v += ep.bias[col];
  if (ep.flags & GEMM_GELU) {
    const float pre = v;
    v = gelu_exact(pre);
    if (ep.aux_out) ep.aux_out[(size_t)row * ep.ld_aux + col] = (ep.flags & GEMM_AUX_GRAD) ? gelu_exact_grad(pre) : pre;
  }
  if (ep.flags & GEMM_DGELU) {
    const float t = ep.aux_in[(size_t)row * ep.ld_aux + col];
    v *= (ep.flags & GEMM_AUX_GRAD) ? t : gelu_exact_grad(t);
  }
  if (ep.flags & GEMM_DROPOUT) {
    const Philox ph(ep.seed);
    const uint32_t w = philox_word(ph, (uint64_t)row * (uint64_t)N + (uint64_t)col, ep.stream);
    v = (w >= dropout_threshold(ep.drop_p)) ? v * (1.0f / (1.0f - ep.drop_p)) : 0.f;
  }
  if (ep.row_scale) v *= ep.row_scale[row / ep.rows_per_scale];
  if (ep.residual) v += ep.residual[(size_t)row * ep.ld_res + col];
  if (ep.out_f32) ep.out_f32[(size_t)row * ep.ld_f32 + col] = v;
  if (ep.out_bf16) ep.out_bf16[(size_t)row * ep.ld_bf16 + col] = v;
  if (ep.colsum_out) atomicAdd(ep.colsum_out + col, v);
}
}  // namespace

int gemm_bf16(const bf16* A, int lda, int a_mn, const bf16* B, int ldb, int b_mn, int M, int N, int K,
              const GemmEpilogue& ep, int force_bn, cudaStream_t stream) {
  return gemm_launch(A, 0, lda, a_mn, B, 0, ldb, b_mn, M, N, K, ep, force_bn, false, stream);
}

size_t gemm_hp_ws_bytes(int M, int N, int K) {
  auto pad8 = [](size_t c) { return (c + 7) & ~(size_t)7; };
  // K-major operand [rows, K]; MN-major operand [K, rows]: either way rows * K elements up to the pitch padding
  const size_t a = 3 * (pad8((size_t)K) * M > pad8((size_t)M) * K ? pad8((size_t)K) * M : pad8((size_t)M) * K) * sizeof(bf16);
  const size_t b = 3 * (pad8((size_t)K) * N > pad8((size_t)N) * K ? pad8((size_t)K) * N : pad8((size_t)N) * K) * sizeof(bf16);
  const size_t acc = (size_t)M * (((size_t)N + 3) & ~(size_t)3) * sizeof(float);
  return ((a + 255) & ~(size_t)255) + ((b + 255) & ~(size_t)255) + acc + 1024;
}

int gemm_hp(const float* A, int lda, int a_mn, const float* B, int ldb, int b_mn, int M, int N, int K,
            const GemmEpilogueT<float>& ep, void* ws, size_t ws_bytes, cudaStream_t stream) {
  ECAMP_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_hp: empty problem %d x %d x %d", M, N, K);
  ECAMP_REQUIRE(ws && ws_bytes >= gemm_hp_ws_bytes(M, N, K), "gemm_hp: workspace of %zu bytes needed, %zu given",
                gemm_hp_ws_bytes(M, N, K), ws_bytes);
  ECAMP_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "gemm_hp: workspace must be 256-byte aligned");
  // stored shape of each operand: K-major [rows, K], MN-major [K, rows]
  const int Ra = a_mn ? K : M, Ca = a_mn ? M : K, Rb = b_mn ? K : N, Cb = b_mn ? N : K;
  const int Cpa = (Ca + 7) & ~7, Cpb = (Cb + 7) & ~7;
  uint8_t* w = static_cast<uint8_t*>(ws);
  bf16* pa = reinterpret_cast<bf16*>(w);
  w += (3 * (size_t)Ra * Cpa * sizeof(bf16) + 255) & ~(size_t)255;
  bf16* pb = reinterpret_cast<bf16*>(w);
  w += (3 * (size_t)Rb * Cpb * sizeof(bf16) + 255) & ~(size_t)255;
  float* acc = reinterpret_cast<float*>(w);
  const int ldacc = (N + 3) & ~3;
  {
    const size_t na = (size_t)Ra * Cpa, nb = (size_t)Rb * Cpb;
    split3_kernel<<<(unsigned)((na + 255) / 256), 256, 0, stream>>>(A, Ra, Ca, lda, Cpa, pa);
    ECAMP_LAUNCHED();
    split3_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, stream>>>(B, Rb, Cb, ldb, Cpb, pb);
    ECAMP_LAUNCHED();
  }
  GemmEpilogue raw;
  raw.out_f32 = acc;
  raw.ld_f32 = ldacc;
  if (int rc = gemm_launch(pa, (size_t)Ra * Cpa, Cpa, a_mn, pb, (size_t)Rb * Cpb, Cpb, b_mn, M, N, K, raw, 0, true, stream)) return rc;
  const size_t n = (size_t)M * N;
  hp_epilogue_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(acc, ldacc, M, N, ep);
  ECAMP_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------
int pdl_enabled() {
  static const int on = [] {
    const char* e = getenv("ECAMP_PDL");
    return e ? atoi(e) : 0;  // measured on B200: 47.12 ms/step with, 47.11 without - the step is not launch-latency bound
  }();
  return on;
}
void set_cta_pair_mode(int mode) { g_cta_pair_mode = mode; }
void set_tma_epilogue(int on) { g_tma_epilogue = on; }
void set_direct_epilogue(int on) { g_direct_epilogue = on; }
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }
static thread_local char g_err[1024] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error_cstr() { return g_err; }

}  // namespace ecamp
