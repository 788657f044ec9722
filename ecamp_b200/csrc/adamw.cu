// Fused multi-tensor AdamW (+ bf16 shadow refresh) — replaces torch.optim.AdamW driven by
// timm add_weight_decay param groups (main_pretrain.py:253-254).  HBM-bound:
// per element it reads p, g, m, v (16 B) and writes p, m, v (12 B) + the bf16 GEMM copy (2 B).
//
// Update rule (torch/optim/adam.py _single_tensor_adam, decoupled decay):
//   p *= 1 - lr * wd ; m += (g - m) * (1 - b1) ; v = b2 * v + (1 - b2) * g * g
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "kernels.cuh"

#include <vector>

namespace ecamp {
namespace {

constexpr int kChunk = 4096;  // elements per CTA: 256 threads x 4 float4

struct Chunk {
  int tensor;
  int pad;
  long long start;
};

ECAMP_DEVINL void store_shadow(const AdamTensor& t, long long i, float4 p, bool vec) {
  if (t.shadow32) *reinterpret_cast<float4*>(t.shadow32 + i) = p;
  if (t.shadow_f) {  // fp32-accurate mode: the same copies (incl. the patch-embed K-order) in fp32
    if (t.shadow_kind == 0) {
      *reinterpret_cast<float4*>(t.shadow_f + i) = p;
    } else {
      const float vals[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long e = i + k;
        const int n = (int)(e / 768), kk = (int)(e % 768), c = kk / 256, pq = kk % 256;
        t.shadow_f[(size_t)n * 768 + pq * 3 + c] = vals[k];
      }
    }
  }
  if (!t.shadow) return;
  if (t.shadow_kind == 0) {
    if (vec) {
      uint2 u;
      u.x = pack_bf16x2(p.x, p.y);
      u.y = pack_bf16x2(p.z, p.w);
      *reinterpret_cast<uint2*>(t.shadow + i) = u;
    }
  } else {
    // patch-embed weight: canonical [n, c*256 + pq] -> GEMM K-order [n, pq*3 + c]
    const float vals[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long e = i + k;
      const int n = (int)(e / 768), kk = (int)(e % 768), c = kk / 256, pq = kk % 256;
      t.shadow[(size_t)n * 768 + pq * 3 + c] = f2bf(vals[k]);
    }
  }
}

// THREADS = 256: the stand-alone step (one chunk of 4096 elements per CTA, four float4 per thread).  THREADS = 128 with at
// most 48 registers per thread: the form used while backward is still running (adamw_step_range) - small enough to sit on an
// SM next to a resident GEMM CTA (11 warps x 168 registers leave 6400 registers and no shared memory is needed), so the
// HBM-bound update proceeds under the tensor-bound GEMMs instead of after them.
template <bool UPDATE, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 128 ? 10 : 1) adamw_kernel(const AdamTensor* __restrict__ table,
                                                    const Chunk* __restrict__ chunks, float lr_decay, float lr_nodecay,
                                                    float beta1, float beta2, float eps, float wd, float bc1,
                                                    float bc2_sqrt, float grad_scale) {
  ECAMP_PDL_ENTRY();
  const Chunk ch = chunks[blockIdx.x];
  const AdamTensor t = table[ch.tensor];
  const float lr = t.decay ? lr_decay : lr_nodecay;  // timm add_weight_decay: two parameter groups, each with its own lr
  const float decay = t.decay ? 1.0f - lr * wd : 1.0f;
  const float step_size = lr / bc1;
  const bool aligned = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) |
                         reinterpret_cast<uintptr_t>(t.m) | reinterpret_cast<uintptr_t>(t.v)) & 15) == 0 &&
                       (t.shadow == nullptr || (reinterpret_cast<uintptr_t>(t.shadow) & 7) == 0) &&
                       (t.shadow32 == nullptr || (reinterpret_cast<uintptr_t>(t.shadow32) & 15) == 0) &&
                       (t.shadow_f == nullptr || (reinterpret_cast<uintptr_t>(t.shadow_f) & 15) == 0);
#pragma unroll 4
  for (int it = 0; it < kChunk / (THREADS * 4); ++it) {
    const long long i = ch.start + ((long long)it * THREADS + threadIdx.x) * 4;
    if (i >= t.numel) break;
    if (aligned && i + 4 <= t.numel) {
      float4 p = *reinterpret_cast<const float4*>(t.p + i);
      if (UPDATE) {
        float4 g = *reinterpret_cast<const float4*>(t.g + i);
        float4 m = *reinterpret_cast<const float4*>(t.m + i);
        float4 v = *reinterpret_cast<const float4*>(t.v + i);
#define ECAMP_ADAM1(P, G, M, V)                                   \
  {                                                               \
    const float gg = G * grad_scale;                              \
    P *= decay;                                                   \
    M = M + (gg - M) * (1.0f - beta1);                            \
    V = beta2 * V + (1.0f - beta2) * gg * gg;                     \
    P = P - step_size * (M / (sqrtf(V) / bc2_sqrt + eps));        \
  }
        ECAMP_ADAM1(p.x, g.x, m.x, v.x)
        ECAMP_ADAM1(p.y, g.y, m.y, v.y)
        ECAMP_ADAM1(p.z, g.z, m.z, v.z)
        ECAMP_ADAM1(p.w, g.w, m.w, v.w)
        *reinterpret_cast<float4*>(t.p + i) = p;
        *reinterpret_cast<float4*>(t.m + i) = m;
        *reinterpret_cast<float4*>(t.v + i) = v;
      }
      store_shadow(t, i, p, true);
    } else {
      for (long long e = i; e < i + 4 && e < t.numel; ++e) {
        float p = t.p[e];
        if (UPDATE) {
          float m = t.m[e], v = t.v[e];
          ECAMP_ADAM1(p, t.g[e], m, v)
          t.p[e] = p;
          t.m[e] = m;
          t.v[e] = v;
        }
        if (t.shadow32) t.shadow32[e] = p;
        if (t.shadow_f) {
          if (t.shadow_kind == 0) {
            t.shadow_f[e] = p;
          } else {
            const int n = (int)(e / 768), kk = (int)(e % 768), c = kk / 256, pq = kk % 256;
            t.shadow_f[(size_t)n * 768 + pq * 3 + c] = p;
          }
        }
        if (t.shadow) {
          if (t.shadow_kind == 0) {
            t.shadow[e] = f2bf(p);
          } else {
            const int n = (int)(e / 768), kk = (int)(e % 768), c = kk / 256, pq = kk % 256;
            t.shadow[(size_t)n * 768 + pq * 3 + c] = f2bf(p);
          }
        }
      }
    }
  }
}

}  // namespace

size_t adamw_table_bytes(int n) { return (size_t)n * sizeof(AdamTensor); }
size_t adamw_chunk_bytes(const AdamTensor* host, int n) {
  long long c = 0;
  for (int i = 0; i < n; ++i) c += (host[i].numel + kChunk - 1) / kChunk;
  return (size_t)c * sizeof(Chunk);
}

int adamw_build_tables(const AdamTensor* host, int n, void* dev_table, void* dev_chunks, long long* n_chunks) {
  std::vector<Chunk> chunks;
  for (int i = 0; i < n; ++i)
    for (long long s = 0; s < host[i].numel; s += kChunk) chunks.push_back(Chunk{i, 0, s});
  ECAMP_CUDA_OK(cudaMemcpy(dev_table, host, (size_t)n * sizeof(AdamTensor), cudaMemcpyHostToDevice));
  ECAMP_CUDA_OK(cudaMemcpy(dev_chunks, chunks.data(), chunks.size() * sizeof(Chunk), cudaMemcpyHostToDevice));
  *n_chunks = (long long)chunks.size();
  return 0;
}

int adamw_step(const void* dev_table, const void* dev_chunks, long long n_chunks, float lr, float lr_nodecay, float beta1,
               float beta2, float eps, float wd, int step, float grad_scale, cudaStream_t st) {
  ECAMP_REQUIRE(step >= 1, "adamw: step counts from 1");
  if (n_chunks <= 0) return 0;
  const float bc1 = 1.0f - (float)pow((double)beta1, (double)step);
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  ECAMP_CUDA_OK(launch_pdl(adamw_kernel<true, 256>, (unsigned)n_chunks, 256, 0, st, static_cast<const AdamTensor*>(dev_table),
                                                         static_cast<const Chunk*>(dev_chunks), lr, lr_nodecay, beta1, beta2,
                                                         eps, wd, bc1, bc2_sqrt, grad_scale));
  ECAMP_LAUNCHED();
  return 0;
}

// the same update for chunks [chunk_begin, chunk_end) only (a contiguous run of tensors), in the small-footprint form
int adamw_step_range(const void* dev_table, const void* dev_chunks, long long chunk_begin, long long chunk_end, float lr,
                     float lr_nodecay, float beta1, float beta2, float eps, float wd, int step, float grad_scale,
                     cudaStream_t st) {
  ECAMP_REQUIRE(step >= 1, "adamw: step counts from 1");
  if (chunk_end <= chunk_begin) return 0;
  const float bc1 = 1.0f - (float)pow((double)beta1, (double)step);
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  ECAMP_CUDA_OK(launch_pdl(adamw_kernel<true, 128>, (unsigned)(chunk_end - chunk_begin), 128, 0, st,
                           static_cast<const AdamTensor*>(dev_table), static_cast<const Chunk*>(dev_chunks) + chunk_begin, lr,
                           lr_nodecay, beta1, beta2, eps, wd, bc1, bc2_sqrt, grad_scale));
  ECAMP_LAUNCHED();
  return 0;
}
long long adamw_chunks_of(long long numel) { return (numel + kChunk - 1) / kChunk; }

int refresh_shadows(const void* dev_table, const void* dev_chunks, long long n_chunks, cudaStream_t st) {
  if (n_chunks <= 0) return 0;
  ECAMP_CUDA_OK(launch_pdl(adamw_kernel<false, 256>, (unsigned)n_chunks, 256, 0, st, static_cast<const AdamTensor*>(dev_table),
                                                          static_cast<const Chunk*>(dev_chunks), 0.f, 0.f, 0.f, 0.f, 0.f,
                                                          0.f, 1.f, 1.f, 1.f));
  ECAMP_LAUNCHED();
  return 0;
}

}  // namespace ecamp
