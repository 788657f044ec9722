"""SR backward: cost of in-window / bordering / outside tiles, with and without the window-limited stages.
Windows are 12 x 12 tiles of 32 px at (column, row); tiles of a 14 x 14 grid that border the window only feed d_u through a
2-pixel halo."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from ecamp_b200 import _lib as L

lib = L.lib()
dev = "cuda"
B = 256
torch.manual_seed(0)
pred = torch.randn(B, 197, 768, device=dev) * 0.5
big = torch.randn(B, 3, 448, 448, device=dev)
w1 = torch.randn(3, 3, 3, 3, device=dev) * 0.3; b1 = torch.randn(3, device=dev) * 0.1
w2 = torch.randn(3, 3, 3, 3, device=dev) * 0.3; b2 = torch.randn(3, device=dev) * 0.1
ws = torch.empty(max(lib.ecamp_sr_ws_floats(B), B * 196), device=dev)
g = torch.tensor([0.7, 1.3, 1.0], device=dev)
d_u = torch.empty(B, 3, 448, 448, device=dev); d_conv = torch.zeros(168, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref = {}
for skip in (0, 1):
    lib.ecamp_sr_set_window_skip(skip)
    for name, c, r in (("window at (1,1): 144 in, 52 bordering", 1, 1), ("window at (0,0): 144 in, 25 bordering, 27 outside", 0, 0),
                       ("window at (8,1): 72 in, 38 bordering", 8, 1), ("window at (14,14): empty", 14, 14), ("random (0..2)", -1, -1)):
        if c < 0:
            gen = torch.Generator().manual_seed(1)
            column = torch.randint(0, 3, (B,), generator=gen).to(dev); row = torch.randint(0, 3, (B,), generator=gen).to(dev)
        else:
            column = torch.full((B,), c, dtype=torch.int64, device=dev); row = torch.full((B,), r, dtype=torch.int64, device=dev)
        ts = []
        for it in range(7):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            L.check(lib.ecamp_sr_loss_bwd(L.ptr(pred), L.ptr(big), L.ptr(column), L.ptr(row), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), B,
                                          ctypes.c_void_p(g.data_ptr() + 4), L.ptr(d_u), L.ptr(d_conv), 0, L.ptr(ws), L.cur_stream()), "srb")
            e.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(s.elapsed_time(e))
        key = (name,)
        if skip == 0:
            ref[key] = (d_u.clone(), d_conv.clone())
            same = ""
        else:
            same = f"  d_u identical: {bool((ref[key][0] == d_u).all().item())}  conv-grad rel diff: {((ref[key][1] - d_conv).norm() / ref[key][1].norm().clamp_min(1e-30)).item():.2e}"
        print(f"skip={skip} {name}: {sorted(ts)[len(ts) // 2]:.3f} ms{same}", flush=True)
