"""Times attention forward / backward for the shapes of the step (C ABI, CUDA events)."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from ecamp_b200 import _lib as L

lib = L.lib()
dev = "cuda"
torch.manual_seed(0)
tc = int(sys.argv[1]) if len(sys.argv) > 1 else 1
lib.ecamp_attention_set_tcgen05(tc)


def run(name, B, H, S, D, masked=False, iters=int(sys.argv[3]) if len(sys.argv) > 3 else 5):
    W = H * D
    qkv = torch.randn(B * S, 3 * W, device=dev).to(torch.bfloat16)
    o = torch.empty(B * S, W, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B * H * S, device=dev)
    d_o = torch.randn(B * S, W, device=dev).to(torch.bfloat16)
    dqkv = torch.empty(B * S, 3 * W, device=dev, dtype=torch.bfloat16)
    delta = torch.empty(B * H * S, device=dev)
    km = None
    if masked:
        ln = torch.randint(S // 4, S + 1, (B,), device=dev)
        km = (torch.arange(S, device=dev)[None, :] < ln[:, None]).to(torch.int64).contiguous()
    a = L.Attn()
    a.q, a.k, a.v = qkv.data_ptr(), qkv.data_ptr() + 2 * W, qkv.data_ptr() + 4 * W
    a.ldq = a.ldk = a.ldv = 3 * W
    a.o, a.ldo, a.lse = o.data_ptr(), W, lse.data_ptr()
    a.key_mask = km.data_ptr() if km is not None else None
    a.B, a.H, a.Sq, a.Sk, a.D = B, H, S, S, D
    a.scale = D ** -0.5
    a.drop_p, a.seed, a.site = 0.0, 0, 0
    a.d_o, a.ld_do, a.delta = d_o.data_ptr(), W, delta.data_ptr()
    a.dq, a.dk, a.dv = dqkv.data_ptr(), dqkv.data_ptr() + 2 * W, dqkv.data_ptr() + 4 * W
    a.lddq = a.lddk = a.lddv = 3 * W
    for nm, fn in (("fwd", lib.ecamp_attention_fwd), ("bwd", lib.ecamp_attention_bwd)):
        for _ in range(2):
            L.check(fn(ctypes.byref(a), L.cur_stream()), nm)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            L.check(fn(ctypes.byref(a), L.cur_stream()), nm)
        e.record()
        torch.cuda.synchronize()
        print(name, nm, round(s.elapsed_time(e) / iters * 1e3, 1), "us", flush=True)


run("dec  16x32  S197", 256, 16, 197, 32)
if len(sys.argv) > 2 and sys.argv[2] == "dec":
    sys.exit(0)
run("enc  12x64  S50 ", 256, 12, 50, 64)
run("bert 6x128  S128", 256, 6, 128, 128, masked=True)
