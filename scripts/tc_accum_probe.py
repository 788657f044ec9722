"""How accurate is a bf16x3 split-operand GEMM on the tcgen05 kernel (fp32 accumulation in TMEM)?  Emulates the term
loop by concatenating the six operand-plane pairs along K and calling the library's GEMM with an fp32 output; compares
with float64.  Also the variant that accumulates 1024-wide K chunks outside the tensor core (fp32 adds on CUDA cores)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from ecamp_b200 import _lib as L

dev = "cuda"
torch.manual_seed(0)


def split3(x):
    h = x.to(torch.bfloat16)
    r = x - h.float()
    m = r.to(torch.bfloat16)
    l = (r - m.float()).to(torch.bfloat16)
    return h, m, l


TERMS = [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)]
for dist_name in ("randn", "uniform01"):
    for K in (768, 3072, 12800, 32768):
        M, N = 512, 768
        gen = torch.randn if dist_name == "randn" else torch.rand
        A = gen(M, K, device=dev); B = gen(N, K, device=dev)
        ref = A.double() @ B.double().t()
        mag = A.double().abs() @ B.double().abs().t()
        a3, b3 = split3(A), split3(B)
        for nterms in (1, 3, 6):
            ac = torch.cat([a3[i] for i, _ in TERMS[:nterms]], 1).contiguous()
            bc = torch.cat([b3[j] for _, j in TERMS[:nterms]], 1).contiguous()
            out = torch.empty(M, N, device=dev)
            L.gemm(ac, bc, out_f32=out)
            # chunked: every 1024 columns of the concatenated contraction in its own launch, summed in fp32
            out2 = torch.zeros(M, N, device=dev)
            tmp = torch.empty(M, N, device=dev)
            for k0 in range(0, ac.shape[1], 1024):
                L.gemm(ac[:, k0:k0 + 1024], bc[:, k0:k0 + 1024], out_f32=tmp)
                out2 += tmp
            torch.cuda.synchronize()
            e1 = ((out.double() - ref).norm() / ref.norm()).item()
            e2 = ((out2.double() - ref).norm() / ref.norm()).item()
            c1 = ((out.double() - ref).abs() / mag).max().item()
            c2 = ((out2.double() - ref).abs() / mag).max().item()
            bias1 = ((out.double() - ref) / mag).mean().item()
            print(json.dumps(dict(dist=dist_name, K=K, terms=nterms, rel_l2=e1, rel_l2_chunked=e2, max_comp=c1, max_comp_chunked=c2,
                                  mean_signed_comp=bias1)), flush=True)
        f32 = (A @ B.t())
        print(json.dumps(dict(dist=dist_name, K=K, terms="torch_fp32", rel_l2=((f32.double() - ref).norm() / ref.norm()).item(),
                              max_comp=((f32.double() - ref).abs() / mag).max().item())), flush=True)
