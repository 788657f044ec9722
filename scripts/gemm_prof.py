"""A handful of GEMM launches for `ncu --set full`: argv[1] = cta-pair mode (1 single, 2 pair)."""
import sys

import torch

sys.path.insert(0, ".")
from ecamp_b200 import _lib as L

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
L.lib().ecamp_gemm_set_cta_pair(mode)
dev = "cuda"
torch.manual_seed(0)


def run(M, N, K, a_mn=False, b_mn=False, tile_n=0, f32=False, res=False):
    a = torch.randn((K, M) if a_mn else (M, K), device=dev).to(torch.bfloat16)
    b = torch.randn((K, N) if b_mn else (N, K), device=dev).to(torch.bfloat16)
    kw = {}
    if f32:
        kw["out_f32"] = torch.empty(M, N, device=dev)
        if res:
            kw["residual"] = torch.randn(M, N, device=dev)
            kw["bias"] = torch.randn(N, device=dev)
    else:
        kw["out_bf16"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    L.gemm(a, b, a_mn=a_mn, b_mn=b_mn, M=M, N=N, K=K, tile_n=tile_n, **kw)
    torch.cuda.synchronize()


run(8192, 8192, 8192, tile_n=256)
run(12800, 2304, 768, tile_n=256)
run(12800, 768, 768, tile_n=192, f32=True, res=True)
run(12800, 768, 3072, b_mn=True, tile_n=192, f32=True)
run(3072, 768, 12800, a_mn=True, b_mn=True, f32=True)
