/*
 * ecamp_b200 — C ABI of the B200-native ECAMP pre-training hot path.
 *
 * The reference (ToniChopp/ECAMP) has no FFI or operator registry for this path: the boundary
 * is the nn.Module returned by `module.model_ecamp.ecamp(**kwargs)`
 * (ECAMP/Pre-training/module/model_ecamp.py:328-333, called at
 * ECAMP/Pre-training/main_pretrain.py:141,233).  The Python mirror of that module
 * (ecamp_b200/model_ecamp.py) binds the entry points below with ctypes; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.  Each entry point cites the reference code whose
 * library kernels it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - every function returns 0 on success, a negative code on failure, never throws and never
 *     exits; `ecamp_last_error()` returns a thread-local description of the last failure;
 *   - all work is enqueued asynchronously on `stream`; asynchronous CUDA errors surface at the
 *     caller's next synchronisation exactly as they do for torch ops;
 *   - nothing allocated by the library is returned to the caller and no argument is retained
 *     past the call.
 */
#ifndef ECAMP_B200_H_
#define ECAMP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECAMP_ABI_VERSION 1
#if defined(__GNUC__)
#define ECAMP_API __attribute__((visibility("default")))
#else
#define ECAMP_API
#endif

ECAMP_API int ecamp_abi_version(void);
ECAMP_API const char* ecamp_last_error(void);

/* ------------------------------------------------------------------------------------------
 * GEMM (tcgen05 / TMEM / TMA): replaces every nn.Linear on the path — timm Block qkv/proj/fc1/fc2
 * (model_ecamp.py:66-68,80-82), decoder_embed / decoder_pred / bert_mlp (model_ecamp.py:73,85,100),
 * HF BertSelfAttention/BertSelfOutput/BertIntermediate/BertOutput (context_fusion.py:12-19),
 * the LM head (bert_modeling.py:209) — forward, dgrad and wgrad.
 * ------------------------------------------------------------------------------------------ */
enum {
  ECAMP_GEMM_GELU = 1,    /* v = gelu(bf16(v)), rounded pre-activation stored to aux_out          */
  ECAMP_GEMM_DGELU = 2,   /* v *= gelu'(aux_in[m, n])                                              */
  ECAMP_GEMM_DROPOUT = 4  /* inverted dropout with a Philox mask keyed by (seed, site, m * N + n)  */
};

typedef struct ecamp_epilogue {
  const float* bias;     /* [N] or NULL                                   */
  const void* aux_in;    /* bf16 [M, ld_aux] or NULL                      */
  void* aux_out;         /* bf16 [M, ld_aux] or NULL                      */
  int32_t ld_aux;
  const float* residual; /* fp32 [M, ld_res] or NULL; may alias out_f32   */
  int32_t ld_res;
  float* out_f32;        /* fp32 [M, ld_f32] or NULL                      */
  int32_t ld_f32;
  void* out_bf16;        /* bf16 [M, ld_bf16] or NULL                     */
  int32_t ld_bf16;
  int32_t flags;
  float drop_p;
  uint64_t seed;
  uint64_t site;
} ecamp_epilogue;

/* D[M,N] = epilogue(A . B^T).  a_mn / b_mn = 0: operand stored [rows, contraction] (contraction
 * contiguous); = 1: stored [contraction, rows].  tile_n = 0 lets the library choose. */
ECAMP_API int ecamp_gemm_bf16(const void* A, int32_t lda, int32_t a_mn, const void* B, int32_t ldb, int32_t b_mn, int32_t M,
                    int32_t N, int32_t K, const ecamp_epilogue* ep, int32_t tile_n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ECAMP_B200_H_ */
