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
constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kNumEpiWarps = 8;

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : ((BN == 192) ? 5 : 6);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 /*barriers*/ + 2048 /*bias x2*/ + 1024 /*alignment slack*/;
};

struct EpiArgs {
  GemmEpilogue ep;
  int vec_ok;
  int split_k;  // > 1: the contraction is split over `split_k` work units per tile, partials added with fp32 atomics
  int kb_per;   // k-blocks per split
};

// split-K epilogue: out_f32[row, col0..col0+31] += v (the buffer was zeroed, or holds the running gradient)
ECAMP_DEVINL void epilogue_atomic_row32(const EpiArgs& ea, const float (&v)[32], int row, int col0, int N) {
  float* op = ea.ep.out_f32 + (size_t)row * ea.ep.ld_f32 + col0;
  const int nvalid = min(32, N - col0);
  if (ea.vec_ok && nvalid == 32) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(op + 4 * i), "f"(v[4 * i]), "f"(v[4 * i + 1]),
                   "f"(v[4 * i + 2]), "f"(v[4 * i + 3])
                   : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < nvalid) atomicAdd(op + i, v[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// epilogue for one thread: one output row, 32 consecutive columns
// ---------------------------------------------------------------------------------------------
// `sbias` (optional): this tile's bias staged in shared memory, pointing at column col0; `pre_res` (optional): the
// residual of this row / chunk already fetched into registers (both hide global-load latency from the epilogue).
ECAMP_DEVINL void epilogue_row32(const EpiArgs& ea, float (&v)[32], int row, int col0, int N,
                                 const float* sbias = nullptr, const float4* pre_res = nullptr) {
  const GemmEpilogue& ep = ea.ep;
  const bool full = ea.vec_ok && (col0 + 32 <= N);
  const int nvalid = min(32, N - col0);

  if (ep.bias && sbias) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = reinterpret_cast<const float4*>(sbias)[i];
      v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
    }
  } else if (ep.bias) {
    if (full) {
      const float4* bp = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b = __ldg(bp + i);
        v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nvalid) v[i] += __ldg(ep.bias + col0 + i);
    }
  }
  if (ep.flags & GEMM_GELU) {
    // the reference applies GELU to the half-precision Linear output; keep forward and backward
    // consistent by activating the rounded value that is saved for backward
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = bf2f(f2bf(v[i]));
    if (ep.aux_out) {
      bf16* ap = ep.aux_out + (size_t)row * ep.ld_aux + col0;
      if (full) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]); u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
          u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]); u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
          reinterpret_cast<uint4*>(ap)[i] = u;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < nvalid) ap[i] = f2bf(v[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
  }
  if (ep.flags & GEMM_DGELU) {
    const bf16* ap = ep.aux_in + (size_t)row * ep.ld_aux + col0;
    if (full) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(ap) + i);
        float2 f;
        f = unpack_bf16x2(u.x); v[8 * i + 0] *= gelu_erf_grad(f.x); v[8 * i + 1] *= gelu_erf_grad(f.y);
        f = unpack_bf16x2(u.y); v[8 * i + 2] *= gelu_erf_grad(f.x); v[8 * i + 3] *= gelu_erf_grad(f.y);
        f = unpack_bf16x2(u.z); v[8 * i + 4] *= gelu_erf_grad(f.x); v[8 * i + 5] *= gelu_erf_grad(f.y);
        f = unpack_bf16x2(u.w); v[8 * i + 6] *= gelu_erf_grad(f.x); v[8 * i + 7] *= gelu_erf_grad(f.y);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nvalid) v[i] *= gelu_erf_grad(bf2f(ap[i]));
    }
  }
  if (ep.flags & GEMM_DROPOUT) {
    const Philox ph(ep.seed);
    const uint32_t thr = dropout_threshold(ep.drop_p);
    const float scale = 1.0f / (1.0f - ep.drop_p);
    const uint64_t base = (uint64_t)row * (uint64_t)N + (uint64_t)col0;  // N % 4 == 0 checked on the host
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 r = ph((base >> 2) + i, ep.stream);
      v[4 * i + 0] = (r.x >= thr) ? v[4 * i + 0] * scale : 0.f;
      v[4 * i + 1] = (r.y >= thr) ? v[4 * i + 1] * scale : 0.f;
      v[4 * i + 2] = (r.z >= thr) ? v[4 * i + 2] * scale : 0.f;
      v[4 * i + 3] = (r.w >= thr) ? v[4 * i + 3] * scale : 0.f;
    }
  }
  if (ep.residual && pre_res && full) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[4 * i + 0] += pre_res[i].x; v[4 * i + 1] += pre_res[i].y; v[4 * i + 2] += pre_res[i].z; v[4 * i + 3] += pre_res[i].w;
    }
  } else if (ep.residual) {
    const float* rp = ep.residual + (size_t)row * ep.ld_res + col0;
    if (full) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 r = reinterpret_cast<const float4*>(rp)[i];
        v[4 * i + 0] += r.x; v[4 * i + 1] += r.y; v[4 * i + 2] += r.z; v[4 * i + 3] += r.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nvalid) v[i] += rp[i];
    }
  }
  if (ep.out_f32) {
    float* op = ep.out_f32 + (size_t)row * ep.ld_f32 + col0;
    if (full) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        reinterpret_cast<float4*>(op)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nvalid) op[i] = v[i];
    }
  }
  if (ep.out_bf16) {
    bf16* op = ep.out_bf16 + (size_t)row * ep.ld_bf16 + col0;
    if (full) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]); u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
        u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]); u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
        reinterpret_cast<uint4*>(op)[i] = u;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nvalid) op[i] = f2bf(v[i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, int M,
                    int N, int K, EpiArgs ea) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * C::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_bias = reinterpret_cast<float*>(smem + STAGES * C::STAGE_BYTES + 256);  // [2][256], one per accumulator stage

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + BM - 1) / BM;
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + BK - 1) / BK;
  const int num_units = num_tiles * ea.split_k;  // work unit = (tile, k-split); unit % num_tiles = tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const int tile = unit % num_tiles, ks = unit / num_tiles;
        const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
        const int kb0 = ks * ea.kb_per, kb1 = min(num_kb, kb0 + ea.kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          uint8_t* a_dst = sA + stage * C::A_BYTES;
          uint8_t* b_dst = sB + stage * C::B_BYTES;
          if (!A_MN) {
            tma_load_2d(a_dst, &tma_a, &full_bar[stage], kb * BK, m_blk * BM);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(a_dst + j * (BK * 128), &tma_a, &full_bar[stage], m_blk * BM + j * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(b_dst, &tma_b, &full_bar[stage], kb * BK, n_blk * BN);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(b_dst + j * (BK * 128), &tma_b, &full_bar[stage], n_blk * BN + j * 64, kb * BK);
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
        const int kb0 = ks * ea.kb_per, kb1 = min(num_kb, kb0 + ea.kb_per);
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
    const int half = (warp - kEpiWarp0) >> 2;  // which half of the tile's columns
    constexpr int HALF_N = BN / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      const int tile = unit % num_tiles;
      const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
      const int row = m_blk * BM + q * 32 + lane;
      const bool plain = ea.split_k == 1;
      // stage this tile's bias in shared memory while the accumulator is still being produced
      float* sb = s_bias + acc * 256;
      if (plain && ea.ep.bias) {
        const int t = threadIdx.x - kEpiWarp0 * 32;
        const int col = n_blk * BN + t;
        if (t < BN) sb[t] = col < N ? __ldg(ea.ep.bias + col) : 0.f;
      }
      // residual of the first chunk, fetched before waiting for the accumulator
      const bool pre = plain && ea.ep.residual != nullptr && ea.vec_ok && row < M;
      float4 rcur[8];
      if (pre) {
        const int col0 = n_blk * BN + half * HALF_N;
        if (col0 + 32 <= N) {
          const float4* rp = reinterpret_cast<const float4*>(ea.ep.residual + (size_t)row * ea.ep.ld_res + col0);
#pragma unroll
          for (int i = 0; i < 8; ++i) rcur[i] = rp[i];
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      asm volatile("bar.sync 1, 256;" ::: "memory");  // epilogue warps only: s_bias[acc] is complete
#pragma unroll 1
      for (int c = 0; c < HALF_N / 32; ++c) {
        const int tcol = acc * BN + half * HALF_N + c * 32;
        const int col0 = n_blk * BN + half * HALF_N + c * 32;
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)tcol, raw);
        float4 rnext[8];
        if (pre && c + 1 < HALF_N / 32 && col0 + 64 <= N) {  // next chunk's residual in flight during this chunk
          const float4* rp = reinterpret_cast<const float4*>(ea.ep.residual + (size_t)row * ea.ep.ld_res + col0 + 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) rnext[i] = rp[i];
        }
        tmem_ld_wait();
        if (c == HALF_N / 32 - 1) {
          // all of this warp's TMEM reads for the tile are done: hand the accumulator stage back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (row < M && col0 < N) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
          if (!plain) epilogue_atomic_row32(ea, v, row, col0, N);
          else epilogue_row32(ea, v, row, col0, N, ea.ep.bias ? sb + half * HALF_N + c * 32 : nullptr, pre ? rcur : nullptr);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) rcur[i] = rnext[i];
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
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
  static constexpr int STAGES = (BN == 256) ? 6 : ((BN == 192) ? 7 : 8);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, int M,
                         int N, int K, EpiArgs ea) {
  using C = Cfg2<BN>;
  constexpr int STAGES = C::STAGES;
  constexpr int HB = BN / 2;  // B rows staged by each CTA
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * C::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int pair = (int)cluster_id_x();
  const int num_pairs = (int)(gridDim.x >> 1);
  const int m_tiles = (M + 2 * BM - 1) / (2 * BM);
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + BK - 1) / BK;
  const int num_units = num_tiles * ea.split_k;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
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

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = pair; unit < num_units; unit += num_pairs) {
        const int tile = unit % num_tiles, ks = unit / num_tiles;
        const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
        const int kb0 = ks * ea.kb_per, kb1 = min(num_kb, kb0 + ea.kb_per);
        const int m0 = m_blk * 2 * BM + (int)rank * BM;
        const int n0 = n_blk * BN + (int)rank * HB;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
          uint8_t* a_dst = sA + stage * C::A_BYTES;
          uint8_t* b_dst = sB + stage * C::B_BYTES;
          if (!A_MN) {
            tma_load_2d_2sm(a_dst, &tma_a, &full_bar[stage], kb * BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d_2sm(a_dst + j * (BK * 128), &tma_a, &full_bar[stage], m0 + j * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d_2sm(b_dst, &tma_b, &full_bar[stage], kb * BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < HB / 64; ++j)
              tma_load_2d_2sm(b_dst + j * (BK * 128), &tma_b, &full_bar[stage], n0 + j * 64, kb * BK);
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
        const int kb0 = ks * ea.kb_per, kb1 = min(num_kb, kb0 + ea.kb_per);
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
    constexpr int HALF_N = BN / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = pair; unit < num_units; unit += num_pairs) {
      const int tile = unit % num_tiles;
      const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * 2 * BM + (int)rank * BM + q * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < HALF_N / 32; ++c) {
        const int tcol = acc * BN + half * HALF_N + c * 32;
        const int col0 = n_blk * BN + half * HALF_N + c * 32;
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)tcol, raw);
        tmem_ld_wait();
        if (c == HALF_N / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);  // the leader's MMA warp owns the hand-back
        }
        if (row < M && col0 < N) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
          if (ea.split_k > 1) epilogue_atomic_row32(ea, v, row, col0, N);
          else epilogue_row32(ea, v, row, col0, N);
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
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

// 0 = automatic (CTA pairs whenever the problem has more than one 128-row tile), 1 = single-CTA kernel only,
// 2 = CTA-pair kernel always.  ECAMP_GEMM_CTA_PAIR overrides the default; ecamp_gemm_set_cta_pair() at run time.
// Default 1: measured on B200 (profiles/r01_gemm_pair_vs_single.log) the pair kernel is bit-correct but not faster
// than the single-CTA kernel on this step's shapes (-3 % .. +1 % at K = 768, +5 % at 8192^3, slower for MN-major B),
// so operand delivery through L2 is not what limits the single-CTA kernel; it stays selectable for further tuning.
int g_cta_pair_mode = [] {
  const char* e = getenv("ECAMP_GEMM_CTA_PAIR");
  return e ? atoi(e) : 1;
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
  const float eff[3] = {1.00f, 0.97f, 0.88f};  // smaller tiles put more shared-memory traffic behind each MMA
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
int launch(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EpiArgs& ea, cudaStream_t st) {
  auto kfn = gemm_tcgen05_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    ECAMP_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES));
    attr_set = true;
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN) * ea.split_k;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kfn<<<grid, kThreads, Cfg<BN>::SMEM_BYTES, st>>>(ta, tb, M, N, K, ea);
  ECAMP_LAUNCHED();
  return 0;
}

template <int BN, bool A_MN, bool B_MN>
int launch2(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EpiArgs& ea, cudaStream_t st) {
  auto kfn = gemm_tcgen05_2cta_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    ECAMP_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<BN>::SMEM_BYTES));
    attr_set = true;
  }
  const int units = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN) * ea.split_k;
  const int max_pairs = num_sms() / 2;
  const int pairs = units < max_pairs ? units : max_pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg2<BN>::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ECAMP_CUDA_OK(cudaLaunchKernelEx(&cfg, kfn, ta, tb, M, N, K, ea));
  ECAMP_LAUNCHED();
  return 0;
}

template <int BN>
int launch_major2(int a_mn, int b_mn, const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K,
                  const EpiArgs& ea, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch2<BN, false, false>(ta, tb, M, N, K, ea, st);
  if (!a_mn && b_mn) return launch2<BN, false, true>(ta, tb, M, N, K, ea, st);
  if (a_mn && b_mn) return launch2<BN, true, true>(ta, tb, M, N, K, ea, st);
  return launch2<BN, true, false>(ta, tb, M, N, K, ea, st);
}

template <int BN>
int launch_major(int a_mn, int b_mn, const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K,
                 const EpiArgs& ea, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch<BN, false, false>(ta, tb, M, N, K, ea, st);
  if (!a_mn && b_mn) return launch<BN, false, true>(ta, tb, M, N, K, ea, st);
  if (a_mn && b_mn) return launch<BN, true, true>(ta, tb, M, N, K, ea, st);
  return launch<BN, true, false>(ta, tb, M, N, K, ea, st);
}

}  // namespace

int make_tmap_bf16(void* map, const bf16* ptr, unsigned long long inner, unsigned long long outer,
                   unsigned long long pitch_elems, unsigned box_outer) {
  return make_tmap(static_cast<CUtensorMap*>(map), ptr, inner, outer, pitch_elems, box_outer);
}

int gemm_bf16(const bf16* A, int lda, int a_mn, const bf16* B, int ldb, int b_mn, int M, int N, int K,
              const GemmEpilogue& ep, int force_bn, cudaStream_t stream) {
  ECAMP_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem %d x %d x %d", M, N, K);
  ECAMP_REQUIRE(ep.out_f32 || ep.out_bf16, "gemm: no output given");
  if (ep.flags & GEMM_DROPOUT)
    ECAMP_REQUIRE(N % 4 == 0 && ep.drop_p >= 0.f && ep.drop_p < 1.f, "gemm: dropout needs N %% 4 == 0, 0 <= p < 1");
  if (ep.flags & GEMM_DGELU) ECAMP_REQUIRE(ep.aux_in != nullptr, "gemm: dGELU needs aux_in");
  ECAMP_REQUIRE(force_bn == 0 || force_bn == 128 || force_bn == 192 || force_bn == 256, "gemm: unsupported tile N %d",
                force_bn);
  const bool accumulate = ep.residual != nullptr && ep.residual == ep.out_f32 && ep.ld_res == ep.ld_f32;
  const bool splittable = ep.out_f32 && !ep.out_bf16 && !ep.bias && ep.flags == 0 && !ep.aux_out &&
                          (ep.residual == nullptr || accumulate);
  int bn = 256, split_k = 1;
  const bool cta2 = g_cta_pair_mode == 2 || (g_cta_pair_mode == 0 && M > BM);
  pick_config(M, N, K, splittable, force_bn, cta2, b_mn != 0, &bn, &split_k);
  ECAMP_REQUIRE(!(cta2 && b_mn && bn == 192), "gemm: tile N 192 is not available to the CTA-pair kernel with an MN-major B");

  CUtensorMap ta, tb;
  int rc;
  // K-major operand: inner = contraction, outer = rows, box [rows_tile, 64]
  // MN-major operand: inner = output index, outer = contraction, box [64 (k), 64]
  rc = a_mn ? make_tmap(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BK)
            : make_tmap(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BM);
  if (rc) return rc;
  rc = b_mn ? make_tmap(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BK)
            : make_tmap(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, (uint32_t)(cta2 ? bn / 2 : bn));
  if (rc) return rc;

  EpiArgs ea;
  ea.ep = ep;
  ea.split_k = split_k;
  ea.kb_per = (((K + BK - 1) / BK) + split_k - 1) / split_k;
  if (split_k > 1) {
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

  if (cta2) {
    if (bn == 256) return launch_major2<256>(a_mn, b_mn, ta, tb, M, N, K, ea, stream);
    if (bn == 192) return launch_major2<192>(a_mn, b_mn, ta, tb, M, N, K, ea, stream);
    return launch_major2<128>(a_mn, b_mn, ta, tb, M, N, K, ea, stream);
  }
  if (bn == 256) return launch_major<256>(a_mn, b_mn, ta, tb, M, N, K, ea, stream);
  if (bn == 192) return launch_major<192>(a_mn, b_mn, ta, tb, M, N, K, ea, stream);
  return launch_major<128>(a_mn, b_mn, ta, tb, M, N, K, ea, stream);
}

// ---------------------------------------------------------------------------------------------
void set_cta_pair_mode(int mode) { g_cta_pair_mode = mode; }
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
