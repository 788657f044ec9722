"""One pre-training step at BASELINE config 2 for ncu: two warm-up steps, then exactly one step between
cudaProfilerStart/Stop (run ncu with --profile-from-start off)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ecamp_b200.model_ecamp import ecamp
from ecamp_b200.optim import FusedAdamW
from ecamp_b200.parallel import DataParallelStep
from ecamp_b200.synthetic import make_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--seq", type=int, default=128)
ap.add_argument("--warm", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = ecamp().to(dev).train()
opt = FusedAdamW(m)
dp = DataParallelStep(m, opt)
b = {k: v.to(dev) for k, v in make_batch(a.batch, T=a.seq, big=True, seed=1).items()}
b.pop("noise")
for _ in range(a.warm):
    dp.step(b)
torch.cuda.synchronize()
torch.cuda.profiler.start()
losses = dp.step(b)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("losses", losses.tolist())
