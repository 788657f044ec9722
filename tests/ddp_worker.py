"""Worker of tests/test_ddp_gpu.py: one of WORLD_SIZE processes (gloo rendezvous on 127.0.0.1; every rank uses GPU
LOCAL_RANK % device_count, so the test also runs with both ranks on ONE GPU).  Checks, on hardware, what data
parallelism rests on (main_pretrain.py:247-250):
  1. stock DistributedDataParallel(find_unused_parameters=True) around the drop-in module synchronises gradients: after
     backward of HALF the batch per rank, every rank holds the gradient of the FULL batch (vs a single-process run);
  2. the same through ecamp_b200.parallel.DataParallelStep (bucketed all-reduce of flat-buffer slices overlapped with the
     staged backward): reduced flat buffer / world == full-batch gradient;
  3. replicas stay bit-identical over optimizer steps on different data (parameters and Adam moments).
Prints one JSON line per rank."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    backend = os.environ.get("ECAMP_TEST_BACKEND", "gloo")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)) % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world)
    from ecamp_b200.model_ecamp import ecamp
    from ecamp_b200.optim import FusedAdamW
    from ecamp_b200.parallel import DataParallelStep
    from oracle.ecamp_oracle import ecamp_oracle, seeded_state_dict, synthetic_batch
    w = seeded_state_dict(ecamp_oracle(), 0)
    m = ecamp().to(dev).eval()
    m.load_state_dict(w)
    per = 2
    full = synthetic_batch(per * world, T=32, seed=60, device=dev)
    mine = {k: v[per * rank:per * (rank + 1)].contiguous() for k, v in full.items()}
    out = dict(rank=rank, world=world, backend=backend)

    # single-process reference: gradient of the mean losses over the FULL batch
    m.zero_grad(set_to_none=True)
    l = m(full)
    (l[0] + l[1] + l[2]).backward()
    names = [k for k, p in m.named_parameters() if p.grad is not None]
    g_full = torch.cat([p.grad.flatten() for k, p in m.named_parameters() if p.grad is not None]).clone()
    flat_full = m.flat_grads().clone()

    # 1. stock DDP, as main_pretrain.py:249 wraps it
    ddp = torch.nn.parallel.DistributedDataParallel(m, device_ids=[dev.index], find_unused_parameters=True)
    for it in range(2):                      # second iteration: DDP has rebuilt its buckets in arrival order
        m.zero_grad(set_to_none=True)
        l = ddp(mine)
        (l[0] + l[1] + l[2]).backward()
    torch.cuda.synchronize()
    g_ddp = torch.cat([p.grad.flatten() for k, p in m.named_parameters() if p.grad is not None])
    out["ddp_names_match"] = names == [k for k, p in m.named_parameters() if p.grad is not None]
    out["ddp_grad_rel_vs_full_batch"] = rel(g_ddp, g_full)
    chk = torch.stack([g_ddp.double().sum(), g_ddp.double().abs().sum()])
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    out["ddp_replica_grads_identical"] = all(torch.equal(allc[0], c) for c in allc)
    # without synchronisation the half-batch gradient differs from the full-batch one by far more than the tolerance
    m.zero_grad(set_to_none=True)
    l = m(mine)
    (l[0] + l[1] + l[2]).backward()
    out["unsynchronised_rel"] = rel(torch.cat([p.grad.flatten() for p in m.parameters() if p.grad is not None]), g_full)
    del ddp

    # 2. DataParallelStep: reduced flat buffer / world == full-batch gradient (lr = 0: the update changes nothing)
    m.zero_grad(set_to_none=True)
    opt = FusedAdamW(m, lr=0.0, betas=(0.9, 0.95), weight_decay=0.05)
    dp = DataParallelStep(m, opt, bucket_mb=8)
    dp.step(mine)
    torch.cuda.synchronize()
    out["dp_grad_rel_vs_full_batch"] = rel(m.flat_grads() / world, flat_full)
    out["dp_buckets"] = len(dp.buckets)

    # 3. replicas stay identical over real steps on different data
    for g in opt.param_groups:
        g["lr"] = 1e-3
    m.train()
    for it in range(3):
        b = synthetic_batch(per, T=32, seed=70 + 10 * it + rank, device=dev)
        dp.step(b)
    torch.cuda.synchronize()
    rt = m._rt
    vec = torch.cat([torch.cat([p.detach().flatten() for p in rt["params"]]), rt["M1"], rt["M2"]])
    ref = vec.clone()
    dist.broadcast(ref, src=0)
    out["replicas_identical_after_steps"] = bool(torch.equal(ref, vec))
    out["max_abs_replica_diff"] = (ref - vec).abs().max().item()
    print("DDPWORKER " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
