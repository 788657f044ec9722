"""BASELINE config 5: fine-tune ViT-B/16 14-class classification, batch 512 synthetic 224-px images, forward + BCE loss +
backward on one B200 through the public module (ecamp_b200.models_vit).  Prints one JSON line (images/s; 105.38 GFLOP per
image for forward + backward)."""
import argparse
import json
import sys

import torch

sys.path.insert(0, ".")
from ecamp_b200.models_vit import vit_base_patch16

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--sgd", action="store_true", help="add grad-norm clip + SGD-momentum (FusedSGD) to the step, as train.py:459-463")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = vit_base_patch16(num_classes=14, drop_path_rate=0.1, global_pool=True).to(dev).train()
xs = [torch.randn(a.batch, 3, 224, 224, device=dev) for _ in range(2)]
ys = [(torch.rand(a.batch, 14, device=dev) < 0.3).float() for _ in range(2)]
loss_fct = torch.nn.BCEWithLogitsLoss()
opt = None
if a.sgd:
    from ecamp_b200.optim import FusedSGD
    opt = FusedSGD(m.parameters(), lr=3e-3, momentum=0.9, weight_decay=0.0, max_grad_norm=1.0)


def step(i):
    m.zero_grad(set_to_none=True)
    loss = loss_fct(m(xs[i % 2]), ys[i % 2])
    loss.backward()
    if opt is not None:
        opt.step()
    return loss


for i in range(a.warmup):
    step(i)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(a.steps):
    loss = step(i)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / a.steps
ips = a.batch / (ms * 1e-3)
peaks = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else dict(bf16_tflops=1590.0)
tf = ips * 105.38 / 1e3
print(json.dumps(dict(metric="finetune_cls_images_per_sec", value=round(ips, 1), unit="images/s", ms_per_step=round(ms, 3), batch=a.batch,
                      steps=a.steps, warmup=a.warmup, dtype="bf16", data="synthetic",
                      config=dict(workload="BASELINE config 5: ViT-B/16 14-class fine-tune, 224 px, 197 tokens, DropPath 0.1, fwd + BCE + bwd" + (" + clip + SGD-momentum" if a.sgd else "")),
                      tflops=round(tf, 1), frac_of_bf16_burst_peak=round(tf / peaks["bf16_tflops"], 4), loss=float(loss.detach()))))
