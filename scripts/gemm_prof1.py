"""One GEMM launch for an ncu source-level capture: argv = mode M N K kind."""
import sys

import torch

sys.path.insert(0, ".")
from ecamp_b200 import _lib as L

mode, M, N, K, kind = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
L.lib().ecamp_gemm_set_cta_pair(mode)
dev = "cuda"
a = torch.randn(M, K, device=dev).to(torch.bfloat16)
b = torch.randn(N, K, device=dev).to(torch.bfloat16)
kw = {}
if kind == "bf16":
    kw = dict(out_bf16=torch.empty(M, N, device=dev, dtype=torch.bfloat16), bias=torch.randn(N, device=dev))
elif kind == "gelu":
    kw = dict(out_bf16=torch.empty(M, N, device=dev, dtype=torch.bfloat16), aux_out=torch.empty(M, N, device=dev, dtype=torch.bfloat16),
              bias=torch.randn(N, device=dev), flags=L.GEMM_GELU)
elif kind == "res":
    kw = dict(out_f32=torch.empty(M, N, device=dev), residual=torch.randn(M, N, device=dev), bias=torch.randn(N, device=dev))
elif kind == "dgelu":   # fc2 dgrad: x gelu'(pre-activation) -> bf16, + column sums (fc1 bias gradient)
    kw = dict(out_bf16=torch.empty(M, N, device=dev, dtype=torch.bfloat16), aux_in=torch.randn(M, N, device=dev).to(torch.bfloat16),
              colsum_out=torch.zeros(N, device=dev), flags=L.GEMM_DGELU)
elif kind == "f32":
    kw = dict(out_f32=torch.empty(M, N, device=dev))
L.gemm(a, b, M=M, N=N, K=K, tile_n=256, **kw)
torch.cuda.synchronize()
