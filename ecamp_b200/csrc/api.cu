// C ABI (extern "C") of libecamp_b200.so — see include/ecamp_b200.h for the contract.
#include "../../include/ecamp_b200.h"

#include "common.cuh"
#include "gemm.cuh"

namespace ecamp {
const char* last_error_cstr();
}

using namespace ecamp;

extern "C" {

int ecamp_abi_version(void) { return ECAMP_ABI_VERSION; }
const char* ecamp_last_error(void) { return ecamp::last_error_cstr(); }

int ecamp_gemm_bf16(const void* A, int32_t lda, int32_t a_mn, const void* B, int32_t ldb, int32_t b_mn, int32_t M,
                    int32_t N, int32_t K, const ecamp_epilogue* ep, int32_t tile_n, void* stream) {
  ECAMP_REQUIRE(A && B && ep, "ecamp_gemm_bf16: null argument");
  GemmEpilogue e;
  e.bias = ep->bias;
  e.aux_in = static_cast<const bf16*>(ep->aux_in);
  e.aux_out = static_cast<bf16*>(ep->aux_out);
  e.ld_aux = ep->ld_aux;
  e.residual = ep->residual;
  e.ld_res = ep->ld_res;
  e.out_f32 = ep->out_f32;
  e.ld_f32 = ep->ld_f32;
  e.out_bf16 = static_cast<bf16*>(ep->out_bf16);
  e.ld_bf16 = ep->ld_bf16;
  e.flags = ep->flags;
  e.drop_p = ep->drop_p;
  e.seed = ep->seed;
  e.stream = ep->site;
  return gemm_bf16(static_cast<const bf16*>(A), lda, a_mn, static_cast<const bf16*>(B), ldb, b_mn, M, N, K, e, tile_n,
                   static_cast<cudaStream_t>(stream));
}

}  // extern "C"
