// Common device helpers for the ecamp_b200 sm_100a kernels: PTX wrappers for mbarrier,
// TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld), Philox RNG, GELU math,
// small vector load/store helpers and warp/block reductions.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace ecamp {

typedef __nv_bfloat16 bf16;

#define ECAMP_DEVINL __device__ __forceinline__

// ---------------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
void count_launch();  // every kernel launch of the library is counted (ecamp_launch_count)
#define ECAMP_CUDA_OK(expr)                                                                      \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      ecamp::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return -2;                                                                                 \
    }                                                                                            \
  } while (0)
#define ECAMP_LAUNCHED()                                 \
  do {                                                   \
    ECAMP_CUDA_OK(cudaGetLastError());                   \
    ecamp::count_launch();                               \
  } while (0)
#define ECAMP_REQUIRE(cond, ...)                                                                 \
  do {                                                                                           \
    if (!(cond)) {                                                                               \
      ecamp::set_last_error(__VA_ARGS__);                                                        \
      return -1;                                                                                 \
    }                                                                                            \
  } while (0)

// ---------------------------------------------------------------------------------------------
// programmatic dependent launch: every kernel of the library is launched with the "programmatic stream
// serialization" attribute and starts with ECAMP_PDL_ENTRY(): griddepcontrol.wait (all prerequisite grids have
// completed, their writes are visible) followed by griddepcontrol.launch_dependents (once every CTA of this grid has
// got that far the NEXT kernel's CTAs may become resident and sit in their own wait).  Launch latency and the next
// kernel's prologue then overlap with this kernel's tail instead of following it (the step has ~600 launches).
// A kernel launched without the attribute executes both as no-ops.  Nothing before the wait may touch global memory.
// ---------------------------------------------------------------------------------------------
ECAMP_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
ECAMP_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#define ECAMP_PDL_ENTRY()             \
  do {                                \
    ecamp::pdl_wait();                \
    ecamp::pdl_launch_dependents();   \
  } while (0)
int pdl_enabled();  // ECAMP_PDL=0 turns the launch attribute off (gemm.cu)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled();
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// shared-memory address / mbarrier
// ---------------------------------------------------------------------------------------------
ECAMP_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

ECAMP_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
ECAMP_DEVINL void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
ECAMP_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

ECAMP_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
ECAMP_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
ECAMP_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trap (CUDA error), never as a hung GPU.
// A waiting warp backs off with nanosleep and reads the clock only every 256 polls: the tight poll loop (try_wait +
// clock64 + compare + branch) of the producer / MMA warps was ~10 % of all executed instructions of an epilogue-bound
// GEMM (ncu source view) and competed for issue slots with the two epilogue warps of the same scheduler.
ECAMP_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = 0;
  uint32_t polls = 0;
  while (!mbar_try_wait(bar, parity)) {
#ifndef ECAMP_MBAR_SPIN  // (A/B switch for measurements: -DECAMP_MBAR_SPIN polls without backing off)
    __nanosleep(40);
#endif
    if ((++polls & 255u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000LL) {  // ~2 s at 2 GHz
        printf("ecamp_b200: mbarrier wait timeout (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
        __trap();
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------
ECAMP_DEVINL void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(desc) : "memory");
}
// 2-D tiled load: c0 = coordinate along the contiguous (inner) dimension, c1 = outer.
ECAMP_DEVINL void tma_load_2d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(desc), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 2-D tiled store shared -> global (bulk async group of the issuing thread); the tensor map clips out-of-range rows / columns
ECAMP_DEVINL void tma_store_2d(const void* desc, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(desc),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
ECAMP_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still have to READ their shared-memory source
template <int N>
ECAMP_DEVINL void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
ECAMP_DEVINL void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
ECAMP_DEVINL void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
ECAMP_DEVINL void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
ECAMP_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
ECAMP_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
ECAMP_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16/fp16 inputs with fp32 accumulate.
ECAMP_DEVINL void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
ECAMP_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 columns of 32-bit: thread i of the warp receives lane (base_lane + i), columns c..c+31.
ECAMP_DEVINL void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
ECAMP_DEVINL void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
ECAMP_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) variants: two SMs of one TPC cooperate on a 256-row tile ------------------------
ECAMP_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
ECAMP_DEVINL uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
ECAMP_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
ECAMP_DEVINL void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// TMA load issued by either CTA of the pair; completion bytes are credited to the barrier of the EVEN (leader) CTA
ECAMP_DEVINL void tma_load_2d_2sm(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(desc), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
ECAMP_DEVINL void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
ECAMP_DEVINL void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
ECAMP_DEVINL void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A (128 rows from each CTA) * B (N/2 rows from each CTA); issued by the leader CTA only
ECAMP_DEVINL void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the issued MMAs retire) on the barrier at this offset in BOTH CTAs of the pair
ECAMP_DEVINL void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) version = 1      bits [61,64) layout (2 = SWIZZLE_128B)
ECAMP_DEVINL uint64_t umma_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B, fp32 D.
//   [4,6) D format (1 = f32)  [7,10) A format (1 = bf16)  [10,13) B format  [15] A major (1 = MN)  [16] B major
//   [17,23) N >> 3            [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

ECAMP_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// math
// ---------------------------------------------------------------------------------------------
// Exact-erf GELU, x * Phi(x), with Phi evaluated as a logistic function of an odd polynomial:
//   Phi(x) = 1 / (1 + exp(-2 u(x))),  u(x) = x (c1 + c3 x^2 + c5 x^4 + c7 x^6),  |x| clamped to 6,
// the form tanh-GELU uses, but with the polynomial re-fitted (two more terms) to the normal CDF itself:
// |Phi error| <= 7e-6, |GELU error| <= 2.5e-5, |GELU' error| <= 6e-5 over all x (checked in float32 against
// scipy's erf) - far below the bf16 rounding of the stored activation.  Cost: 2 MUFU (ex2, rcp) + 7 FMA-pipe
// instructions; the fused GEMM epilogues were bound by the ~25 instructions of the erf form they replaced
// (measured: fc1 forward 663 -> 1112 TFLOP/s with the epilogue math removed).
ECAMP_DEVINL float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
ECAMP_DEVINL float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// coefficients of -2 log2(e) u(x) / x
#define ECAMP_PHI_K1 (-2.3020227f)     // -2 log2(e) * 0.7978202620075527
#define ECAMP_PHI_K3 (-0.10545136f)    // -2 log2(e) * 0.03654665804906016
#define ECAMP_PHI_K5 (5.6107155e-4f)   // -2 log2(e) * -1.9445258094679608e-4
#define ECAMP_PHI_K7 (3.9495544e-5f)   // -2 log2(e) * -1.3688112480251091e-5
ECAMP_DEVINL float normal_cdf(float x) {
  const float xc = fminf(fmaxf(x, -6.0f), 6.0f);
  const float t = xc * xc;
  float p = fmaf(ECAMP_PHI_K7, t, ECAMP_PHI_K5);
  p = fmaf(p, t, ECAMP_PHI_K3);
  p = fmaf(p, t, ECAMP_PHI_K1);
  return rcp_approx(1.0f + ex2_approx(xc * p));
}
ECAMP_DEVINL float gelu_erf(float x) { return x * normal_cdf(x); }
// d/dx [x Phi(x)] = Phi + x Phi', with Phi' taken from the same approximation: Phi (1 - Phi) 2 u'(x)
ECAMP_DEVINL float gelu_erf_grad(float x) {
  const float xc = fminf(fmaxf(x, -6.0f), 6.0f);
  const float t = xc * xc;
  float p = fmaf(ECAMP_PHI_K7, t, ECAMP_PHI_K5);
  p = fmaf(p, t, ECAMP_PHI_K3);
  p = fmaf(p, t, ECAMP_PHI_K1);
  const float s = rcp_approx(1.0f + ex2_approx(xc * p));
  // 2 u'(x) = 2 (c1 + 3 c3 t + 5 c5 t^2 + 7 c7 t^3)
  float d = fmaf(-1.9163357e-4f, t, -1.9445258e-3f);
  d = fmaf(d, t, 0.21927995f);
  d = fmaf(d, t, 1.5956405f);
  return fmaf(x * d, fmaf(-s, s, s), s);
}

// GELU and its derivative from one evaluation of Phi (the forward epilogue can store the derivative for backward:
// 6 more FMA-pipe instructions there instead of the 16 of gelu_erf_grad per element in the dGELU epilogue)
ECAMP_DEVINL void gelu_erf_both(float x, float& y, float& g) {
  const float xc = fminf(fmaxf(x, -6.0f), 6.0f);
  const float t = xc * xc;
  float p = fmaf(ECAMP_PHI_K7, t, ECAMP_PHI_K5);
  p = fmaf(p, t, ECAMP_PHI_K3);
  p = fmaf(p, t, ECAMP_PHI_K1);
  const float s = rcp_approx(1.0f + ex2_approx(xc * p));
  float d = fmaf(-1.9163357e-4f, t, -1.9445258e-3f);
  d = fmaf(d, t, 0.21927995f);
  d = fmaf(d, t, 1.5956405f);
  y = x * s;
  g = fmaf(x * d, fmaf(-s, s, s), s);
}

// exact (erf) GELU and derivative for the fp32-accurate mode: x Phi(x), Phi(x) + x phi(x)
ECAMP_DEVINL float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
ECAMP_DEVINL float gelu_exact_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}

// Philox4x32-10: counter-based, so forward and backward regenerate identical dropout masks from
// (seed, stream offset, element index) regardless of how the work is tiled.
struct Philox {
  uint32_t k0, k1;
  ECAMP_DEVINL Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  ECAMP_DEVINL uint4 operator()(uint64_t ctr, uint64_t stream) const {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = (uint32_t)stream, c3 = (uint32_t)(stream >> 32);
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      c0 = hi1 ^ c1 ^ a;
      c1 = lo1;
      c2 = hi0 ^ c3 ^ b;
      c3 = lo0;
      a += 0x9E3779B9u;
      b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
// keep-decision for element `idx` of dropout site `stream`: uniform u32 >= threshold(p) keeps.
ECAMP_DEVINL uint32_t dropout_threshold(float p) { return (uint32_t)(p * 4294967296.0); }
ECAMP_DEVINL uint32_t philox_word(const Philox& ph, uint64_t idx, uint64_t stream) {
  const uint4 r = ph(idx >> 2, stream);
  const uint32_t w = (uint32_t)(idx & 3);
  return w == 0 ? r.x : (w == 1 ? r.y : (w == 2 ? r.z : r.w));
}

// Attention-probability dropout draws 16 bits per element: one Philox call decides 8 consecutive keys of one query
// (per-element calls made the attention kernels RNG-bound).  p is quantised to thr16 / 65536.
ECAMP_DEVINL uint32_t dropout_threshold16(float p) { return (uint32_t)(p * 65536.0f + 0.5f); }
ECAMP_DEVINL float dropout_keep_scale16(uint32_t thr16) { return 65536.0f / (65536.0f - (float)thr16); }
// bit k of the result = keep element 8 * group + k of row `row_id` (row_id enumerates (batch, head, query))
ECAMP_DEVINL uint32_t philox_keep8(const Philox& ph, uint64_t row_id, int groups_per_row, int group, uint64_t site,
                                   uint32_t thr16) {
  const uint4 r = ph(row_id * (uint64_t)groups_per_row + (uint64_t)group, site);
  uint32_t m = 0;
  m |= ((r.x & 0xFFFFu) >= thr16) ? 1u : 0u;
  m |= ((r.x >> 16) >= thr16) ? 2u : 0u;
  m |= ((r.y & 0xFFFFu) >= thr16) ? 4u : 0u;
  m |= ((r.y >> 16) >= thr16) ? 8u : 0u;
  m |= ((r.z & 0xFFFFu) >= thr16) ? 16u : 0u;
  m |= ((r.z >> 16) >= thr16) ? 32u : 0u;
  m |= ((r.w & 0xFFFFu) >= thr16) ? 64u : 0u;
  m |= ((r.w >> 16) >= thr16) ? 128u : 0u;
  return m;
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
ECAMP_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
ECAMP_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Block-wide sum; `red` must hold >= 32 floats; result valid in all threads.
ECAMP_DEVINL float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

ECAMP_DEVINL float bf2f(bf16 x) { return __bfloat162float(x); }
ECAMP_DEVINL bf16 f2bf(float x) { return __float2bfloat16_rn(x); }
ECAMP_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
ECAMP_DEVINL float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(t);
}


// ---------------------------------------------------------------------------------------------
// activation element type.  Production: AT = bf16 (GEMM / attention operands, 16-bit like the reference's autocast).
// fp32-accurate parity mode: AT = float - the SAME kernels instantiated with fp32 activations; the GEMMs then run on the
// same tcgen05 kernel with error-compensated bf16 x 3 split operands (gemm.cu, gemm_hp).  The helpers below move 1 / 4 / 8
// consecutive elements between memory and fp32 registers for either type.
// ---------------------------------------------------------------------------------------------
ECAMP_DEVINL float act_ld(const bf16* p) { return bf2f(*p); }
ECAMP_DEVINL float act_ld(const float* p) { return *p; }
ECAMP_DEVINL void act_st(bf16* p, float v) { *p = f2bf(v); }
ECAMP_DEVINL void act_st(float* p, float v) { *p = v; }
ECAMP_DEVINL float4 ld4(const bf16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 lo = unpack_bf16x2(u.x), hi = unpack_bf16x2(u.y);
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
ECAMP_DEVINL float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
ECAMP_DEVINL void st4(bf16* p, const float4& v) {
  uint2 u;
  u.x = pack_bf16x2(v.x, v.y);
  u.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}
ECAMP_DEVINL void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
// what the consumer of a stored activation reads back: the bf16-rounded value in production, the value itself in fp32
ECAMP_DEVINL float4 act_round4(const bf16*, const float4& v) {
  const float2 lo = unpack_bf16x2(pack_bf16x2(v.x, v.y)), hi = unpack_bf16x2(pack_bf16x2(v.z, v.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
ECAMP_DEVINL float4 act_round4(const float*, const float4& v) { return v; }
template <typename AT> struct is_hp { static constexpr bool value = false; };
template <> struct is_hp<float> { static constexpr bool value = true; };

}  // namespace ecamp
