"""One GPU: cost of driving backward stage by stage (what the all-reduce overlap needs) against one native call, and against
one call per all-reduce bucket (ecamp_backward_stages).  No communication: the callback is empty."""
import sys
import time

import torch

sys.path.insert(0, ".")
from ecamp_b200 import _lib as L
from ecamp_b200.model_ecamp import ecamp
from ecamp_b200.optim import FusedAdamW
from ecamp_b200.parallel import plan_buckets, stage_ranges
from ecamp_b200.synthetic import make_batch

dev = "cuda"
torch.manual_seed(0)
m = ecamp().to(dev).train()
opt = FusedAdamW(m, lr=1.5e-4, betas=(0.9, 0.95), weight_decay=0.05)
batches = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in make_batch(256, seed=s).items()} for s in (1, 2)]
buckets = plan_buckets(stage_ranges(L.lib()), 64 * (1 << 20) // 4, 96 * (1 << 20) // 4)
ends = [b[0] for b in buckets]
noop = lambda s, lo, hi: None
modes = (("one call", dict()), ("per stage", dict(stage_callback=noop)), ("per bucket", dict(stage_callback=noop, callback_stages=ends)))
for rep in range(2):
    for name, kw in modes:
        for i in range(3):
            m.forward_backward(batches[i & 1], **kw); opt.step(); opt.zero_grad(set_to_none=True)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s.record()
        n = 30
        for i in range(n):
            m.forward_backward(batches[i & 1], **kw); opt.step(); opt.zero_grad(set_to_none=True)
        e.record()
        host_ms = (time.perf_counter() - t0) * 1e3 / n
        torch.cuda.synchronize()
        print(f"{name}: {s.elapsed_time(e) / n:.3f} ms/step (host enqueue {host_ms:.2f} ms/step)", flush=True)
