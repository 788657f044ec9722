// Fused multi-tensor SGD with momentum + global gradient-norm clipping for the fine-tune trainer — replaces
//   torch.nn.utils.clip_grad_norm_(model.parameters(), args.max_grad_norm)   (Fine-tuning/Classification/train.py:459-461)
//   torch.optim.SGD(model.parameters(), lr, momentum=0.9, weight_decay=wd).step()          (train.py:377-380,463)
// Two launches per step, no host synchronisation: (1) sum of squares of every gradient -> one fp32 scalar,
// (2) the update, which derives the clip coefficient from that scalar on the device.  HBM-bound: launch (1) reads
// g (4 B / element); launch (2) reads p, g, buf and writes p, buf (20 B / element).
//
// Update rule (torch/optim/sgd.py _single_tensor_sgd, dampening 0, no nesterov):
//   coef = min(1, max_norm / (||g||_2 + 1e-6))                 (torch.nn.utils.clip_grad_norm_)
//   d = coef * g + wd * p ;  buf = first ? d : momentum * buf + d ;  p -= lr * buf
#include "kernels.cuh"

#include <vector>

namespace ecamp {
namespace {

constexpr int kSgdChunk = 4096;  // elements per CTA: 256 threads x 4 float4

struct SgdChunk {
  int tensor;
  int pad;
  long long start;
};

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const SgdTensor* __restrict__ table,
                                                         const SgdChunk* __restrict__ chunks, float* __restrict__ out) {
  ECAMP_PDL_ENTRY();
  __shared__ float red[32];
  const SgdChunk ch = chunks[blockIdx.x];
  const SgdTensor t = table[ch.tensor];
  const bool aligned = (reinterpret_cast<uintptr_t>(t.g) & 15) == 0;
  float s = 0.f;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const long long i = ch.start + ((long long)it * 256 + threadIdx.x) * 4;
    if (i >= t.numel) break;
    if (aligned && i + 4 <= t.numel) {
      const float4 g = *reinterpret_cast<const float4*>(t.g + i);
      s += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
    } else {
      for (long long e = i; e < i + 4 && e < t.numel; ++e) s += t.g[e] * t.g[e];
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

__global__ void __launch_bounds__(256) sgd_momentum_kernel(const SgdTensor* __restrict__ table,
                                                           const SgdChunk* __restrict__ chunks, float lr, float momentum,
                                                           float wd, int first, float max_norm,
                                                           const float* __restrict__ sumsq, int write_grads) {
  ECAMP_PDL_ENTRY();
  const SgdChunk ch = chunks[blockIdx.x];
  const SgdTensor t = table[ch.tensor];
  float coef = 1.0f;
  if (max_norm > 0.f && sumsq) coef = fminf(1.0f, max_norm / (sqrtf(*sumsq) + 1e-6f));
  const bool aligned =
      ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) | reinterpret_cast<uintptr_t>(t.buf)) & 15) == 0;
#define ECAMP_SGD1(P, G, M)                      \
  {                                              \
    G = G * coef;                                \
    const float d = fmaf(wd, P, G);              \
    M = first ? d : __fadd_rn(__fmul_rn(momentum, M), d); \
    P = fmaf(-lr, M, P);                         \
  }
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const long long i = ch.start + ((long long)it * 256 + threadIdx.x) * 4;
    if (i >= t.numel) break;
    if (aligned && i + 4 <= t.numel) {
      float4 p = *reinterpret_cast<const float4*>(t.p + i);
      float4 g = *reinterpret_cast<const float4*>(t.g + i);
      float4 m = first ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(t.buf + i);
      ECAMP_SGD1(p.x, g.x, m.x)
      ECAMP_SGD1(p.y, g.y, m.y)
      ECAMP_SGD1(p.z, g.z, m.z)
      ECAMP_SGD1(p.w, g.w, m.w)
      *reinterpret_cast<float4*>(t.p + i) = p;
      *reinterpret_cast<float4*>(t.buf + i) = m;
      if (write_grads) *reinterpret_cast<float4*>(t.g + i) = g;
    } else {
      for (long long e = i; e < i + 4 && e < t.numel; ++e) {
        float p = t.p[e], g = t.g[e], m = first ? 0.f : t.buf[e];
        ECAMP_SGD1(p, g, m)
        t.p[e] = p;
        t.buf[e] = m;
        if (write_grads) t.g[e] = g;
      }
    }
  }
#undef ECAMP_SGD1
}

}  // namespace

size_t sgd_table_bytes(int n) { return (size_t)(n > 0 ? n : 0) * sizeof(SgdTensor); }
size_t sgd_chunk_bytes(const long long* numel, int n) {
  long long c = 0;
  for (int i = 0; i < n; ++i) c += (numel[i] + kSgdChunk - 1) / kSgdChunk;
  return (size_t)c * sizeof(SgdChunk);
}

int sgd_build_tables(const SgdTensor* host, int n, void* dev_table, void* dev_chunks, long long* n_chunks) {
  ECAMP_REQUIRE(host && dev_table && dev_chunks && n_chunks && n > 0, "sgd_build_tables: null argument");
  std::vector<SgdChunk> chunks;
  for (int i = 0; i < n; ++i) {
    ECAMP_REQUIRE(host[i].p && host[i].g && host[i].buf && host[i].numel > 0, "sgd_build_tables: tensor %d incomplete", i);
    for (long long s = 0; s < host[i].numel; s += kSgdChunk) chunks.push_back(SgdChunk{i, 0, s});
  }
  ECAMP_CUDA_OK(cudaMemcpy(dev_table, host, (size_t)n * sizeof(SgdTensor), cudaMemcpyHostToDevice));
  ECAMP_CUDA_OK(cudaMemcpy(dev_chunks, chunks.data(), chunks.size() * sizeof(SgdChunk), cudaMemcpyHostToDevice));
  *n_chunks = (long long)chunks.size();
  return 0;
}

int grad_sumsq(const void* dev_table, const void* dev_chunks, long long n_chunks, float* sumsq, cudaStream_t st) {
  ECAMP_REQUIRE(dev_table && dev_chunks && sumsq, "grad_sumsq: null argument");
  ECAMP_CUDA_OK(cudaMemsetAsync(sumsq, 0, sizeof(float), st));
  if (n_chunks <= 0) return 0;
  ECAMP_CUDA_OK(launch_pdl(grad_sumsq_kernel, (unsigned)n_chunks, 256, 0, st, static_cast<const SgdTensor*>(dev_table),
                           static_cast<const SgdChunk*>(dev_chunks), sumsq));
  ECAMP_LAUNCHED();
  return 0;
}

int sgd_momentum_step(const void* dev_table, const void* dev_chunks, long long n_chunks, float lr, float momentum,
                      float wd, int first, float max_norm, const float* sumsq, int write_grads, cudaStream_t st) {
  ECAMP_REQUIRE(dev_table && dev_chunks, "sgd_momentum_step: null argument");
  ECAMP_REQUIRE(max_norm <= 0.f || sumsq, "sgd_momentum_step: clipping needs the gradient sum of squares");
  if (n_chunks <= 0) return 0;
  ECAMP_CUDA_OK(launch_pdl(sgd_momentum_kernel, (unsigned)n_chunks, 256, 0, st, static_cast<const SgdTensor*>(dev_table),
                           static_cast<const SgdChunk*>(dev_chunks), lr, momentum, wd, first, max_norm, sumsq,
                           write_grads));
  ECAMP_LAUNCHED();
  return 0;
}

}  // namespace ecamp
