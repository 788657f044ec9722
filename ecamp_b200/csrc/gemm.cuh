// Internal (C++) interface of the tcgen05 GEMM used by every Linear on the ECAMP hot path.
#pragma once
#include "common.cuh"

namespace ecamp {

enum GemmFlags : int {
  GEMM_GELU = 1,     // v = gelu(bf16(v)); the rounded pre-activation is stored to aux_out (bf16) if given
  GEMM_DGELU = 2,    // v *= gelu'(aux_in[m, n])   (aux_in = bf16 pre-activation saved by the forward)
  GEMM_DROPOUT = 4,  // v = keep(m, n) ? v / (1 - p) : 0   (Philox keyed by seed / stream / m * N + n)
  GEMM_AUX_GRAD = 8,  // with GEMM_GELU: aux_out receives bf16(gelu'(pre-activation)) instead of the pre-activation;
                      // with GEMM_DGELU: aux_in IS that stored derivative (v *= aux_in)
};

// AT = activation element type: bf16 in production, float in the fp32-accurate parity mode (gemm_hp)
template <typename AT>
struct GemmEpilogueT {
  const float* bias = nullptr;      // [N], added first
  const AT* aux_in = nullptr;       // [M, ld_aux]
  AT* aux_out = nullptr;            // [M, ld_aux]
  int ld_aux = 0;
  const float* residual = nullptr;  // [M, ld_res] fp32, added last (may alias out_f32: accumulate)
  int ld_res = 0;
  float* out_f32 = nullptr;
  int ld_f32 = 0;
  AT* out_bf16 = nullptr;           // activation-typed output
  int ld_bf16 = 0;
  float* colsum_out = nullptr;      // [N] fp32: += column sums of the emitted values (atomics; the bias gradient of the
                                    // Linear that consumes this output).  Only with out_bf16, not with split-K.
  const float* row_scale = nullptr;  // optional [M / rows_per_scale] fp32: v *= row_scale[m / rows_per_scale] before the
  int rows_per_scale = 1;            // residual is added (timm DropPath: per-sample mask / keep_prob on the branch)
  int flags = 0;
  float drop_p = 0.f;
  unsigned long long seed = 0, stream = 0;
};
typedef GemmEpilogueT<bf16> GemmEpilogue;

// D[M, N] = A . B^T over the contraction dimension K, bf16 inputs, fp32 accumulation in TMEM.
//   a_mn == 0: A is stored [M, K] row-major with pitch lda (K-major);  a_mn == 1: stored [K, M] with pitch lda.
//   b_mn == 0: B is stored [N, K] row-major with pitch ldb (K-major);  b_mn == 1: stored [K, N] with pitch ldb.
// Forward  y = x W^T      : A = x [M,K] (a_mn 0), B = W [N,K] (b_mn 0)
// dgrad    dx = dy W      : A = dy [M,N'] (a_mn 0), B = W [N',K'] stored [contraction, out] (b_mn 1)
// wgrad    dW = dy^T x    : A = dy stored [rows, N'] (a_mn 1), B = x stored [rows, K'] (b_mn 1)
int gemm_bf16(const bf16* A, int lda, int a_mn, const bf16* B, int ldb, int b_mn, int M, int N, int K,
              const GemmEpilogue& ep, int force_bn, cudaStream_t stream);

// fp32-accurate mode: the same product from fp32 operands.  Each operand is split into three bf16 planes (hi + mid + lo,
// 24 mantissa bits), the six significant plane pairs are accumulated by the SAME tcgen05 kernel (fp32 in TMEM, units of
// <= 1024 products summed outside with fp32 atomics), then the epilogue operators run in fp32 with the exact erf GELU.
// `ws`: scratch of gemm_hp_ws_bytes(M, N, K) bytes, 256-byte aligned.
size_t gemm_hp_ws_bytes(int M, int N, int K);
int gemm_hp(const float* A, int lda, int a_mn, const float* B, int ldb, int b_mn, int M, int N, int K,
            const GemmEpilogueT<float>& ep, void* ws, size_t ws_bytes, cudaStream_t stream);

// 2-D TMA descriptor (128-byte swizzle, box = [box_outer rows, 64 elements]) over a row-major bf16 matrix;
// `map` points to a CUtensorMap (kept opaque here so that callers need not include <cuda.h>).
int make_tmap_bf16(void* map, const bf16* ptr, unsigned long long inner, unsigned long long outer,
                   unsigned long long pitch_elems, unsigned box_outer);

}  // namespace ecamp
