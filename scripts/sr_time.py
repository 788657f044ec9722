"""Times the image-loss kernels (SR forward / backward, pred_grad) at B = 256 through the C ABI."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from ecamp_b200 import _lib as L

lib = L.lib()
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
pred = torch.randn(B, 197, 768, device=dev) * 0.5
big = torch.randn(B, 3, 448, 448, device=dev)
tgt = torch.randn(B, 196, 768, device=dev)
mask = (torch.rand(B, 196, device=dev) < 0.75).float()
column = torch.randint(0, 3, (B,), device=dev); row = torch.randint(0, 3, (B,), device=dev)
w1 = torch.randn(3, 3, 3, 3, device=dev) * 0.3; b1 = torch.randn(3, device=dev) * 0.1
w2 = torch.randn(3, 3, 3, 3, device=dev) * 0.3; b2 = torch.randn(3, device=dev) * 0.1
loss = torch.zeros(2, device=dev)
ws = torch.empty(max(lib.ecamp_sr_ws_floats(B), B * 196), device=dev)
g = torch.tensor([0.7, 1.3, 1.0], device=dev)
d_u = torch.empty(B, 3, 448, 448, device=dev); d_conv = torch.zeros(168, device=dev)
d_pred = torch.empty(B, 197, 768, dtype=torch.bfloat16, device=dev)


def fwd():
    L.check(lib.ecamp_sr_loss_fwd(L.ptr(pred), L.ptr(big), L.ptr(column), L.ptr(row), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), B, ctypes.c_void_p(loss.data_ptr() + 4), L.ptr(ws), L.cur_stream()), "sr")


def bwd():
    L.check(lib.ecamp_sr_loss_bwd(L.ptr(pred), L.ptr(big), L.ptr(column), L.ptr(row), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), B, ctypes.c_void_p(g.data_ptr() + 4), L.ptr(d_u), L.ptr(d_conv), 0, L.ptr(ws), L.cur_stream()), "srb")


def pg():
    L.check(lib.ecamp_pred_grad(L.ptr(pred), L.ptr(tgt), L.ptr(mask), L.ptr(d_u), L.ptr(g), B, L.ptr(d_pred), L.cur_stream()), "pg")


for name, fn in (("sr_fwd", fwd), ("sr_bwd", bwd), ("pred_grad", pg)):
    for _ in range(2):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        fn()
    e.record()
    torch.cuda.synchronize()
    print(name, round(s.elapsed_time(e) / 5, 3), "ms")
