"""Times LayerNorm backward (with bf16 output + column sums, the variant the step uses) through the C ABI, for both kernel
forms (slab = 740 CTAs with per-warp shared-memory column sums; staged = one CTA per SM, cp.async.bulk row rings) over a
sweep of row counts: the intercept of the sweep is the fixed cost of a launch."""
import sys

import torch

sys.path.insert(0, ".")
from ecamp_b200 import _lib as L

lib = L.lib()
dev = "cuda"
torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for slab in (1, 0):
    lib.ecamp_layernorm_set_bwd_slab(slab)
    for name, M, D in (("tiny", 512, 768), ("small", 4096, 768), ("enc", 12800, 768), ("bert", 32768, 768), ("big", 65536, 768),
                       ("tiny", 512, 512), ("dec", 50432, 512)):
        dy = torch.randn(M, D, device=dev); x = torch.randn(M, D, device=dev); add = torch.randn(M, D, device=dev)
        mean = x.mean(1).contiguous(); rstd = (x.var(1, unbiased=False) + 1e-6).rsqrt().contiguous()
        g = torch.randn(D, device=dev); dx = torch.empty(M, D, device=dev); dxb = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
        dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev); cs = torch.zeros(D, device=dev)
        for with_add in (1, 0):
            ts = []
            for it in range(12):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                L.check(lib.ecamp_layernorm_bwd(L.ptr(dy), L.ptr(x), L.ptr(mean), L.ptr(rstd), L.ptr(g), M, D, L.ptr(add) if with_add else None,
                                                L.ptr(dx), L.ptr(dxb), L.ptr(dg), L.ptr(db), L.ptr(cs), 1, None, L.cur_stream()), "lnb")
                e.record()
                torch.cuda.synchronize()
                if it >= 2:
                    ts.append(s.elapsed_time(e))
            us = sorted(ts)[len(ts) // 2] * 1e3
            gb = M * D * (4 + 4 + 4 * with_add + 4 + 2) / 1e9
            print(f"{'slab  ' if slab else 'staged'} {name} M={M} D={D} addend={with_add}: {us:.1f} us  {gb / us * 1e6:.0f} GB/s", flush=True)
