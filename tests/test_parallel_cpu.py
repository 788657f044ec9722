"""Host-side logic of the data-parallel path on CPU: bucket planning over the library's backward stages and a
world_size-2 gloo run showing that the bucketed all-reduce of the flat gradient buffer equals the gradient of
the concatenated batch (mean of per-rank means)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ecamp_b200.parallel import allreduce_flat, plan_buckets, stage_ranges


def test_plan_buckets(built_lib):
    ranges = stage_ranges(built_lib.lib())
    total = built_lib.lib().ecamp_grad_floats()
    for mb in (1, 25, 64, 1024):
        buckets = plan_buckets(ranges, mb * (1 << 20) // 4)
        assert buckets[0][2] == total and buckets[-1][1] == 0
        assert all(b[1] == nb[2] for b, nb in zip(buckets, buckets[1:]))      # contiguous, back to front
        assert all(b[0] < nb[0] for b, nb in zip(buckets, buckets[1:]))       # in stage order
        assert all(hi - lo >= mb * (1 << 20) // 4 for _, lo, hi in buckets[:-1])
    assert len(plan_buckets(ranges, 1 << 40)) == 1
    # tapered tail: below 96 MB from the front of the buffer every stage is its own bucket, the last one the patch embedding
    taper = plan_buckets(ranges, 64 * (1 << 20) // 4, 96 * (1 << 20) // 4)
    assert taper[0][2] == total and taper[-1][1] == 0 and all(b[1] == nb[2] for b, nb in zip(taper, taper[1:]))
    assert (taper[-1][2] - taper[-1][1]) * 4 < 4 * (1 << 20)           # ~2.4 MB: what cannot overlap with anything
    assert len(taper) >= len(plan_buckets(ranges, 64 * (1 << 20) // 4)) + 1


def _worker(rank, world, port, n, ranges):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        flat = torch.randn(n, generator=g)
        ref = flat.clone()
        dist.all_reduce(ref)
        buckets = plan_buckets(ranges, n // 7)
        seen = 0
        for s in range(len(ranges)):                      # the overlapped schedule: reduce as stages finish
            for w in allreduce_flat(flat, buckets, after_stage=s):
                w.wait()
                seen += 1
        assert seen == len(buckets)
        assert torch.equal(flat, ref)
        # mean of rank means == global-batch mean for equal per-rank batches
        both = torch.stack([torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)])
        assert torch.allclose(flat / world, both.mean(0), atol=1e-6)
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n = 10007
    edges = [n, 9000, 7001, 7000, 4242, 1000, 17, 0]
    ranges = [(lo, hi) for hi, lo in zip(edges, edges[1:])]
    mp.spawn(_worker, args=(2, port, n, ranges), nprocs=2, join=True)


def test_numa_binding_is_a_noop_without_topology():
    """bind_host_to_gpu must never fail or shrink the affinity to nothing when the GPU / sysfs topology is unavailable."""
    import os
    from ecamp_b200.synthetic import bind_host_to_gpu
    before = os.sched_getaffinity(0)
    r = bind_host_to_gpu(0)
    after = os.sched_getaffinity(0)
    assert (r is None and after == before) or (set(r) == after and after <= before and len(after) > 0)
    os.sched_setaffinity(0, before)
