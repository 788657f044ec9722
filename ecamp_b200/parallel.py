"""Data-parallel step: one process per GPU, gradients averaged once per optimizer step.

The reference wraps the model in stock DistributedDataParallel(find_unused_parameters=True) and, with
accum_iter = 8 and no no_sync(), all-reduces on every micro-step (ECAMP/Pre-training/main_pretrain.py:135-153,249;
SURVEY D10).  Here the native backward runs in stages that finish CONTIGUOUS ranges of one flat fp32 gradient
buffer from the back (LM head first, patch-embed last), so a bucket is just a slice: when the stages of a
bucket are done an event is recorded and the slice is all-reduced on a side stream (NCCL over NVLink /
NVSwitch) while the remaining stages keep computing.  Averaging (1 / world) is folded into the fused AdamW.
The path shards by samples with exactly this one exchange step; there is no other collective.
"""
import ctypes
import os

import torch
import torch.distributed as dist


def stage_ranges(lib):
    lo, hi = ctypes.c_int64(), ctypes.c_int64()
    out = []
    for s in range(lib.ecamp_backward_stage_count()):
        rc = lib.ecamp_backward_stage_range(s, ctypes.byref(lo), ctypes.byref(hi))
        if rc != 0:
            raise RuntimeError("ecamp_backward_stage_range failed")
        out.append((lo.value, hi.value))
    return out


def plan_buckets(ranges, bucket_floats, tail_floats=0):
    """Merge consecutive backward stages (whose ranges tile the buffer back to front) into buckets of at least
    `bucket_floats`.  Returns [(last_stage, lo, hi)]: after `last_stage` has run, flat[lo:hi] is final.
    `tail_floats` > 0 tapers the end of backward: once at most that many floats remain below the current position,
    every stage becomes its own bucket - the all-reduce of the LAST bucket cannot overlap with anything, so it should be
    the smallest possible slice (patch embedding: 2.4 MB) instead of a full-size bucket that has waited for it."""
    buckets, cur_hi, cur_lo = [], None, None
    for s, (lo, hi) in enumerate(ranges):
        if cur_hi is None:
            cur_hi = hi
        elif hi != cur_lo:
            raise ValueError("backward stage ranges must be contiguous and descending")
        cur_lo = lo
        if cur_hi - cur_lo >= bucket_floats or s == len(ranges) - 1 or cur_hi <= tail_floats:
            buckets.append((s, cur_lo, cur_hi))
            cur_hi = None
    return buckets


def allreduce_flat(flat, buckets, group=None, after_stage=None):
    """All-reduce (sum) `flat` bucket by bucket.  With `after_stage` given, only the buckets that end at that
    stage are reduced (the overlapped path); otherwise all of them.  Returns the async work handles."""
    works = []
    for last, lo, hi in buckets:
        if after_stage is None or last == after_stage:
            works.append(dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True))
    return works


class DataParallelStep:
    def __init__(self, model, optimizer, bucket_mb=64, group=None, tail_mb=96):
        """bucket_mb: minimum bucket size; tail_mb: below this many MB from the front of the buffer (= the end of
        backward: encoder blocks 2..0 and the patch embedding) every backward stage is reduced on its own."""
        from . import _lib as L
        self.model, self.optimizer, self.group = model, optimizer, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buckets = plan_buckets(stage_ranges(L.lib()), int(bucket_mb * (1 << 20) // 4), int(tail_mb * (1 << 20) // 4))
        # priority of the communication stream (ECAMP_DP_COMM_PRIORITY: 0 = default, negative = higher): NCCL's CTAs and the
        # persistent GEMM grids compete for SMs at every kernel boundary
        self.comm = torch.cuda.Stream(priority=int(os.environ.get("ECAMP_DP_COMM_PRIORITY", "0"))) if self.world > 1 else None
        self.timeline = None   # set to [] to record per-bucket CUDA events of the next step (see bucket_timeline)
        self.group_stages = os.environ.get("ECAMP_DP_GROUP_STAGES", "1") != "0"   # 0: one native backward call per stage
        self.overlap_update = os.environ.get("ECAMP_DP_OVERLAP_UPDATE", "1") != "0"  # 0: one AdamW launch after backward
        self.upd = torch.cuda.Stream() if torch.cuda.is_available() else None

    def step(self, batch, loss_weights=(1.0, 1.0, 1.0), update=True):
        """forward + backward (+ overlapped gradient all-reduce) (+ fused AdamW).  Returns the 3 local losses.

        With `overlap_update` (default) the optimizer step is applied bucket by bucket on a side stream as soon as a bucket's
        gradients are final (and all-reduced): nothing in the rest of this backward pass reads those parameters again, and the
        HBM-bound update then runs under the tensor-bound GEMMs of the remaining stages instead of after them."""
        cur = torch.cuda.current_stream()
        overlap = update and self.overlap_update and hasattr(self.optimizer, "step_range")
        if self.world == 1 and not overlap:
            losses = self.model.forward_backward(batch, loss_weights)
        else:
            if self.comm is not None:
                self.comm.wait_stream(cur)
            ends = {b[0]: b for b in self.buckets}
            rt = self.optimizer.begin_step() if overlap else None
            if overlap:
                self.upd.wait_stream(cur)

            rec = self.timeline is not None and self.world > 1
            if rec:
                t0 = torch.cuda.Event(enable_timing=True)
                t0.record(cur)

            def on_stage(stage, lo, hi):
                b = ends.get(stage)
                if b is None or not update:
                    return
                ev = torch.cuda.Event(enable_timing=rec)
                ev.record(cur)
                done = ev
                if self.world > 1:
                    self.comm.wait_event(ev)
                    with torch.cuda.stream(self.comm):
                        if rec:
                            st = torch.cuda.Event(enable_timing=True); st.record(self.comm)
                        dist.all_reduce(self.model.flat_grads()[b[1]:b[2]], op=dist.ReduceOp.SUM, group=self.group)
                        if rec:
                            en = torch.cuda.Event(enable_timing=True); en.record(self.comm)
                            self.timeline.append((stage, b[1], b[2], ev, st, en))
                        if overlap:
                            done = torch.cuda.Event(); done.record(self.comm)
                if overlap:   # the update has its own stream: it must not delay the next bucket's all-reduce
                    self.upd.wait_event(done)
                    with torch.cuda.stream(self.upd):
                        self.optimizer.step_range(rt, b[1], b[2], grad_scale=1.0 / self.world)

            # one native call per bucket: finality is only needed where an all-reduce / an update starts
            losses = self.model.forward_backward(batch, loss_weights, stage_callback=on_stage,
                                                 callback_stages=[b[0] for b in self.buckets] if self.group_stages else None)
            if rec:
                bwd_end = torch.cuda.Event(enable_timing=True)
                bwd_end.record(cur)
            if self.comm is not None:
                cur.wait_stream(self.comm)
            if overlap:
                cur.wait_stream(self.upd)
            if rec:
                joined = torch.cuda.Event(enable_timing=True)
                joined.record(cur)
                self._marks = (t0, bwd_end, joined)
        if update:
            if not overlap:
                self.optimizer.step(grad_scale=1.0 / self.world)
            self.optimizer.zero_grad(set_to_none=True)
        return losses

    def bucket_timeline(self):
        """After a step recorded with `self.timeline = []`: per bucket (stage, MB, ready_ms, start_ms, end_ms) on the
        step's clock (0 = step start), plus when backward ended and when the compute stream had joined the last
        all-reduce; exposed = joined - backward_end is the communication that did not overlap."""
        torch.cuda.synchronize()
        t0, bwd_end, joined = self._marks
        rows = [dict(stage=s, mb=round((hi - lo) * 4 / 2 ** 20, 1), ready_ms=round(t0.elapsed_time(ev), 3),
                     start_ms=round(t0.elapsed_time(st), 3), end_ms=round(t0.elapsed_time(en), 3)) for s, lo, hi, ev, st, en in self.timeline]
        out = dict(buckets=rows, backward_end_ms=round(t0.elapsed_time(bwd_end), 3), joined_ms=round(t0.elapsed_time(joined), 3))
        out["exposed_ms"] = round(out["joined_ms"] - out["backward_end_ms"], 3)
        self.timeline = None
        return out
