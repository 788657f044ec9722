"""GPU diagnostic for the tcgen05 GEMM: every operand-major combination, tile width, tails and epilogue.
Each case runs in-process; run the whole script under `timeout` (mbarrier waits trap after ~2 s)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from ecamp_b200 import _lib as L

torch.manual_seed(0)
dev = "cuda"
MODE = int(sys.argv[1]) if len(sys.argv) > 1 else 0   # 0 auto, 1 single-CTA kernel, 2 CTA-pair kernel
L.lib().ecamp_gemm_set_cta_pair(MODE)
print(json.dumps(dict(cta_pair_mode=MODE)), flush=True)
results = []


def run_case(name, M, N, K, a_mn, b_mn, tile_n, **kw):
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    b = torch.randn(N, K, device=dev).to(torch.bfloat16)
    a_st = a.t().contiguous() if a_mn else a
    b_st = b.t().contiguous() if b_mn else b
    ref = a.float() @ b.float().t()
    out = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32)
    try:
        L.gemm(a_st, b_st, a_mn=a_mn, b_mn=b_mn, M=M, N=N, K=K, out_f32=out, tile_n=tile_n)
        torch.cuda.synchronize()
        err = (out - ref).abs().max().item()
        scale = ref.abs().max().item()
        nan = int(torch.isnan(out).sum().item())
        ok = nan == 0 and err <= 2e-3 * max(scale, 1.0)
        results.append(dict(name=name, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, bn=tile_n, err=err, scale=scale, nan=nan, ok=ok))
        if not ok:
            # error pattern: which rows / columns are wrong
            bad = ((out - ref).abs() > 2e-3 * max(scale, 1.0)) | torch.isnan(out)
            rows = bad.any(1).nonzero().flatten()[:8].tolist()
            cols = bad.any(0).nonzero().flatten()[:8].tolist()
            results[-1].update(bad_frac=bad.float().mean().item(), bad_rows=rows, bad_cols=cols,
                               sample=[out[0, :4].tolist(), ref[0, :4].tolist()])
    except Exception as e:  # noqa
        results.append(dict(name=name, error=str(e)[:300], ok=False))
    print(json.dumps(results[-1]), flush=True)


for bn in (128, 192, 256):
    if MODE == 2 and bn == 192:
        for (nm, a_, b_) in (("kk", False, False), ("mn_k", True, False)):
            run_case(nm + "_192_pair", 512, 384, 512, a_, b_, bn)
        continue
    run_case("kk_single_kblock", 128, bn, 64, False, False, bn)
    run_case("kk_multi_k", 256, 2 * bn, 512, False, False, bn)
    run_case("kk_tails", 200, 300, 136, False, False, bn)
    run_case("k_mn", 256, 2 * bn, 512, False, True, bn)
    run_case("mn_mn", 256, 2 * bn, 512, True, True, bn)
    run_case("mn_k", 256, 2 * bn, 512, True, False, bn)
    run_case("mn_mn_tails", 200, 304, 136, True, True, bn)
# split-K candidates (fp32-only outputs with few tiles and a long contraction)
for rep in range(3):
    run_case("splitk_dgrad_small_m", 100, 768, 3072, False, True, 0)
    run_case("splitk_dgrad_394", 394, 512, 2048, False, True, 0)
    run_case("splitk_wgrad_768", 768, 768, 32768, True, True, 0)
    run_case("splitk_wgrad_3072", 3072, 768, 12800, True, True, 0)
    run_case("splitk_wgrad_tail", 768, 1536, 12544 + 37, True, True, 0)
run_case("many_tiles", 12800, 768, 768, False, False, 0)
run_case("vocab_like", 1024, 30000, 768, False, False, 0)
run_case("dgrad_vocab", 512, 768, 30000, False, True, 0)

# epilogue checks (bias, gelu + aux, dgelu, residual, bf16 out, dropout statistics)
M, N, K = 384, 768, 256
a = torch.randn(M, K, device=dev).to(torch.bfloat16)
b = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
bias = torch.randn(N, device=dev)
res = torch.randn(M, N, device=dev)
ref_lin = a.float() @ b.float().t() + bias
pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
o16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
o32 = torch.empty(M, N, device=dev, dtype=torch.float32)
L.gemm(a, b, bias=bias, aux_out=pre, residual=res, out_f32=o32, out_bf16=o16, flags=L.GEMM_GELU)
torch.cuda.synchronize()
pre_ref = ref_lin.to(torch.bfloat16)
ref = torch.nn.functional.gelu(pre_ref.float()) + res
e1 = (pre.float() - pre_ref.float()).abs().max().item()
e2 = (o32 - ref).abs().max().item()
e3 = (o16.float() - ref).abs().max().item()
print(json.dumps(dict(name="epi_gelu", pre_err=e1, out_err=e2, bf16_err=e3, ok=e1 < 0.05 and e2 < 0.05 and e3 < 0.05)), flush=True)
results.append(dict(name="epi_gelu", ok=e1 < 0.05 and e2 < 0.05 and e3 < 0.05))

g = torch.randn(M, K, device=dev).to(torch.bfloat16)
L.gemm(g, b.t().contiguous(), b_mn=False, aux_in=pre, out_f32=o32, flags=L.GEMM_DGELU, M=M, N=N, K=K) if False else None
x = pre.float().requires_grad_(True)
torch.nn.functional.gelu(x).sum().backward()
L.gemm(a, b, aux_in=pre, out_f32=o32, flags=L.GEMM_DGELU)
torch.cuda.synchronize()
ref = (a.float() @ b.float().t()) * x.grad
e = (o32 - ref).abs().max().item()
print(json.dumps(dict(name="epi_dgelu", err=e, ok=e < 1e-2)), flush=True)
results.append(dict(name="epi_dgelu", ok=e < 1e-2))

L.gemm(a, b, out_f32=o32, flags=L.GEMM_DROPOUT, drop_p=0.1, seed=1234, site=7)
o32b = torch.empty_like(o32)
L.gemm(a, b, out_f32=o32b, flags=L.GEMM_DROPOUT, drop_p=0.1, seed=1234, site=7, tile_n=128)
torch.cuda.synchronize()
lin = a.float() @ b.float().t()
kept = o32 != 0
frac = 1 - kept.float().mean().item()
e = ((o32 - lin / 0.9) * kept).abs().max().item()
same = bool((o32 == o32b).all().item())
ok = abs(frac - 0.1) < 0.01 and e < 1e-2 and same
print(json.dumps(dict(name="epi_dropout", drop_frac=frac, err=e, tiling_invariant=same, ok=ok)), flush=True)
results.append(dict(name="epi_dropout", ok=ok))

# accumulate (residual aliases out)
acc = torch.randn(M, N, device=dev)
acc0 = acc.clone()
L.gemm(a, b, residual=acc, out_f32=acc)
torch.cuda.synchronize()
e = (acc - (acc0 + lin)).abs().max().item()
print(json.dumps(dict(name="epi_accumulate", err=e, ok=e < 1e-2)), flush=True)
results.append(dict(name="epi_accumulate", ok=e < 1e-2))

# quick timing of the big shapes
def bench(M, N, K, a_mn=False, b_mn=False, tile_n=0, iters=20):
    a = torch.randn((K, M) if a_mn else (M, K), device=dev).to(torch.bfloat16)
    b = torch.randn((K, N) if b_mn else (N, K), device=dev).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        L.gemm(a, b, a_mn=a_mn, b_mn=b_mn, M=M, N=N, K=K, out_bf16=out, tile_n=tile_n)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        L.gemm(a, b, a_mn=a_mn, b_mn=b_mn, M=M, N=N, K=K, out_bf16=out, tile_n=tile_n)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    # cuBLAS comparison
    a2 = a.t() if a_mn else a
    b2 = b if b_mn else b.t()
    for _ in range(3):
        torch.matmul(a2, b2)
    s.record()
    for _ in range(iters):
        torch.matmul(a2, b2)
    e.record()
    torch.cuda.synchronize()
    ms2 = s.elapsed_time(e) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    print(json.dumps(dict(name="bench", M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, bn=tile_n, ms=ms, tflops=tf,
                          cublas_ms=ms2, cublas_tflops=2.0 * M * N * K / ms2 / 1e9)), flush=True)


if all(r.get("ok") for r in results):
    for bn in (0, 128, 192, 256):
        bench(12800, 2304, 768, tile_n=bn)
        bench(12800, 3072, 768, tile_n=bn)
        bench(12800, 768, 3072, tile_n=bn)
    bench(8192, 8192, 8192)
    bench(50432, 2048, 512)
    bench(32768, 1536, 768)
    bench(12800, 768, 3072, b_mn=True)       # dgrad
    bench(3072, 768, 12800, a_mn=True, b_mn=True)  # wgrad
    bench(4096, 30000, 768)
print("ALL_OK" if all(r.get("ok") for r in results) else "SOME_FAILED")
