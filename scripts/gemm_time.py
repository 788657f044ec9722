"""Times a list of GEMM shapes (CUDA events, L2 flushed between launches): argv = cta-pair mode."""
import json
import sys

import torch

sys.path.insert(0, ".")
from ecamp_b200 import _lib as L

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
L.lib().ecamp_gemm_set_cta_pair(mode)
dev = "cuda"
torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(name, M, N, K, a_mn=False, b_mn=False, tile_n=0, kind="bf16", iters=10):
    a = torch.randn((K, M) if a_mn else (M, K), device=dev).to(torch.bfloat16)
    b = torch.randn((K, N) if b_mn else (N, K), device=dev).to(torch.bfloat16)
    kw = {}
    if kind == "bf16":
        kw["out_bf16"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        kw["bias"] = torch.randn(N, device=dev)
    elif kind == "gelu":
        kw["out_bf16"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        kw["aux_out"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        kw["bias"] = torch.randn(N, device=dev)
        kw["flags"] = L.GEMM_GELU
    elif kind == "dgelu":
        kw["out_bf16"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        kw["aux_in"] = torch.randn(M, N, device=dev).to(torch.bfloat16)
        kw["flags"] = L.GEMM_DGELU
    elif kind == "f32":
        kw["out_f32"] = torch.empty(M, N, device=dev)
    elif kind == "res":
        kw["out_f32"] = torch.empty(M, N, device=dev)
        kw["residual"] = torch.randn(M, N, device=dev)
        kw["bias"] = torch.randn(N, device=dev)
    ts = []
    for it in range(iters + 2):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        L.gemm(a, b, a_mn=a_mn, b_mn=b_mn, M=M, N=N, K=K, tile_n=tile_n, **kw)
        e.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(s.elapsed_time(e))
    ms = sorted(ts)[len(ts) // 2]
    print(json.dumps(dict(name=name, M=M, N=N, K=K, bn=tile_n, kind=kind, us=round(ms * 1e3, 1),
                          tflops=round(2.0 * M * N * K / ms / 1e9))), flush=True)


shapes = [("enc_qkv", 12800, 2304, 768, False, False, "bf16"), ("enc_fc1", 12800, 3072, 768, False, False, "gelu"),
          ("enc_proj", 12800, 768, 768, False, False, "res"), ("enc_fc2", 12800, 768, 3072, False, False, "res"),
          ("enc_fc2_dgrad", 12800, 3072, 768, False, True, "dgelu"), ("enc_fc1_dgrad", 12800, 768, 3072, False, True, "f32"),
          ("dec_fc1", 50432, 2048, 512, False, False, "gelu"), ("dec_fc2", 50432, 512, 2048, False, False, "res"),
          ("bert_qkv", 32768, 2304, 768, False, False, "bf16"), ("bert_ao", 32768, 768, 768, False, False, "res"),
          ("enc_fc1_wgrad", 3072, 768, 12800, True, True, "f32"), ("big", 8192, 8192, 8192, False, False, "bf16")]
bns = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
for (nm, M, N, K, am, bm, kind) in shapes:
    for bn in bns:
        run(nm, M, N, K, am, bm, bn, kind)
