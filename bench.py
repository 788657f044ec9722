#!/usr/bin/env python
"""Benchmark of the ECAMP pre-training step (BASELINE.json config 2 / 3): forward + backward + fused AdamW in bf16,
per-GPU batch 256 synthetic pairs (448-px radiograph -> in-model bicubic 224 px -> ViT-B/16 MAE with 75 % masking
+ SR branch; 128-token report through the 6-layer BERT with context fusion; three losses), data-parallel over N GPUs.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU path (oracle port), reported baseline

Prints ONE JSON line (rank 0).  `value` = pairs/s with inputs resident in HBM; `e2e` = the same through the public
module API with pinned-host inputs copied every step and the losses read back every step (the reference collate's fp32
image batch by default; `e2e_u8_input` = the same with the loader's 8-bit crops normalised on the GPU, `e2e_f32_input`
repeats the headline with its input format spelled out).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_PAIR_STEP = {128: 88.60, 256: 136.25}   # BASELINE.md §3 (2*M*N*K, backward = 2 x forward)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0), "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def to_device(batch, dev, stream=None):
    if stream is None:
        return {k: v.to(dev, non_blocking=True) for k, v in batch.items()}
    with torch.cuda.stream(stream):
        return {k: v.to(dev, non_blocking=True) for k, v in batch.items()}


def gemm_roofline(B, T, peak_tflops):
    """Time every GEMM of the step through the C ABI WITH THE EPILOGUE THE STEP USES for it (CUDA events on the launching
    stream, L2 flushed between launches) and return achieved TFLOP/s over all tcgen05 GEMM launches of one step.
    Per Linear: (rows, out, in, count, forward epilogue, dgrad epilogue or None); the weight gradient is always a plain
    fp32 store (split-K atomics where the library chooses them)."""
    from ecamp_b200 import _lib as L
    dev = "cuda"
    Me, Mi, Md, Mt = B * 50, B * 49, B * 197, B * T
    lin = [
        (Mi, 768, 768, 1, "f32", None),                       # patch embed (49 kept patches); no input gradient
        (Me, 2304, 768, 12, "bf16", "f32"), (Me, 768, 768, 12, "res", "bf16"),          # encoder qkv, proj
        (Me, 3072, 768, 12, "gelu", "f32"), (Me, 768, 3072, 12, "res", "dgelu"),        # encoder fc1, fc2
        (Me, 512, 768, 1, "bf16", "res"),                                               # decoder_embed
        (Md, 1536, 512, 4, "bf16", "f32"), (Md, 512, 512, 4, "res", "bf16"),            # decoder qkv, proj
        (Md, 2048, 512, 4, "gelu", "f32"), (Md, 512, 2048, 4, "res", "dgelu"),          # decoder fc1, fc2
        (Md, 768, 512, 1, "f32", "f32"),                                                # decoder_pred
        (Me, 768, 768, 1, "bf16", "f32"),                                               # bert_mlp
        (Mt, 2304, 768, 7, "bf16", "res"),                                              # BERT / fusion q|k|v
        (Mt, 768, 768, 8, "res_drop", "bf16"),                                          # attention.output.dense x7, out_layer.dense
        (Mt, 768, 768, 1, "bf16", "res"), (Mt, 768, 768, 1, "gelu", "f32"),             # cross query; LM-head transform
        (Mi, 1536, 768, 1, "bf16", "bf16"),                                             # cross key|value
        (Mt, 1536, 768, 7, "gelu", "res"), (Mt, 768, 1536, 7, "res_drop", "dgelu"),     # BERT intermediate, output
        (Mt, 30000, 768, 1, "bf16", "f32"),                                             # vocabulary projection (row chunks)
    ]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot_flops = tot_ms = 0.0
    launches = 0
    per_shape = []
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, reps=3):
        fn()
        ms = []
        for _ in range(reps):
            flush.zero_()
            s.record(); fn(); e.record()
            torch.cuda.synchronize()
            ms.append(s.elapsed_time(e))
        return statistics.median(ms)

    def epilogue(kind, rows, n):
        kw = {}
        if kind in ("bf16", "gelu", "dgelu"):
            kw["out_bf16"] = torch.empty(rows, n, dtype=torch.bfloat16, device=dev)
        else:
            kw["out_f32"] = torch.empty(rows, n, dtype=torch.float32, device=dev)
        if kind in ("bf16", "gelu", "res", "res_drop", "f32"):
            kw["bias"] = torch.zeros(n, device=dev)
        if kind == "gelu":
            kw["aux_out"] = torch.empty(rows, n, dtype=torch.bfloat16, device=dev); kw["flags"] = L.GEMM_GELU | L.GEMM_AUX_GRAD
        if kind == "dgelu":
            kw["aux_in"] = torch.randn(rows, n, device=dev).to(torch.bfloat16); kw["flags"] = L.GEMM_DGELU | L.GEMM_AUX_GRAD
            kw["colsum_out"] = torch.zeros(n, device=dev); kw.pop("bias", None)
        if kind in ("res", "res_drop"):
            kw["residual"] = torch.randn(rows, n, device=dev)
        if kind == "res_drop":
            kw["flags"] = L.GEMM_DROPOUT; kw["drop_p"] = 0.1; kw["seed"] = 1; kw["site"] = 1
        return kw

    for rows, n_out, n_in, count, k_f, k_d in lin:
        r = min(rows, 8192) if n_out == 30000 else rows          # the vocabulary projection runs in row chunks
        scale = rows / r
        x = torch.randn(r, n_in, device=dev).to(torch.bfloat16)
        w = torch.randn(n_out, n_in, device=dev).to(torch.bfloat16)
        dy = torch.randn(r, n_out, device=dev).to(torch.bfloat16)
        dw = torch.empty(n_out, n_in, dtype=torch.float32, device=dev)
        kf = epilogue(k_f, r, n_out)
        t_f = timed(lambda: L.gemm(x, w, **kf))
        del kf
        t_d = 0.0
        if k_d is not None:
            kd = epilogue(k_d, r, n_in)
            kd.pop("bias", None)
            t_d = timed(lambda: L.gemm(dy, w, b_mn=True, M=r, N=n_in, K=n_out, **kd))
            del kd
        t_w = timed(lambda: L.gemm(dy, x, a_mn=True, b_mn=True, M=n_out, N=n_in, K=r, out_f32=dw))
        fl = 2.0 * r * n_out * n_in
        n_g = 3 if k_d is not None else 2
        tot_flops += n_g * fl * scale * count
        tot_ms += (t_f + t_d + t_w) * scale * count
        launches += n_g * count * int(round(scale))
        per_shape.append(dict(rows=rows, out=n_out, inp=n_in, count=count, fwd=k_f, dgrad=k_d, fwd_tflops=round(fl / t_f / 1e9, 0),
                              dgrad_tflops=round(fl / t_d / 1e9, 0) if k_d else None, wgrad_tflops=round(fl / t_w / 1e9, 0)))
        del x, w, dy, dw
    achieved = tot_flops / (tot_ms * 1e-3) / 1e12
    traffic = None
    try:  # DRAM bytes per GEMM launch from the committed ncu capture of one step (profiles/, see r01c_ncu_summary.md)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r02g_ncu_gemm_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass
    return dict(bound="tensor", achieved=round(achieved, 1), peak=peak_tflops, unit="TFLOP/s", frac=round(achieved / peak_tflops, 4),
                traffic=traffic, kernel="gemm_tcgen05_2cta_kernel (every Linear forward / dgrad / wgrad launch of one step, with the step's fused epilogues)",
                flops_per_step=tot_flops, gemm_ms_per_step=round(tot_ms, 3), launches_per_step=launches, shapes=per_shape)



def _timed_steps(step, steps, warmup):
    """CUDA-event timing of `steps` calls of step(i) after `warmup` untimed calls; returns ms per step."""
    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        step(i)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps


def _frac(pairs_per_s, gflop_per_unit, pk):
    tf = pairs_per_s * gflop_per_unit / 1e3
    return dict(tflops=round(tf, 1), frac_of_bf16_burst_peak=round(tf / pk["bf16_tflops"], 4),
                frac_of_bf16_sustained_peak=round(tf / pk["bf16_tflops_sustained"], 4))


def e2e_gpu_loader(dp, host, B, dev, steps=10, frame=512):
    """End to end with the image half of the loader on the GPU (ecamp_b200.image_pipeline): per step the host draws the
    RandomResizedCrop / flip parameters for B decoded 8-bit frames (torch generator, torchvision's order), packs the crop boxes
    into pinned memory, ships them (H2D), the GPU resamples them to 448 x 448 with Pillow's arithmetic, and the step runs on
    the uint8 batch; the losses are read back every step."""
    from ecamp_b200.image_pipeline import GpuImageTransform
    t = GpuImageTransform(device=dev)
    g = torch.Generator().manual_seed(99)
    frames = [torch.randint(0, 256, (frame, frame), generator=g, dtype=torch.uint8) for _ in range(B)]
    text = [{k: v for k, v in hb.items() if k != "image"} for hb in host]
    h2d = [0]

    def step(i):
        params = [t.draw_params(frame, frame) for _ in range(B)]
        img = t(frames, params)
        b = to_device(text[i % 2], dev)
        b["image"] = img
        h2d[0] = sum(p[2] * p[3] for p in params) + sum(v.numel() * v.element_size() for v in text[0].values())
        return dp.step(b).tolist()

    ms = _timed_steps(step, steps, 3)
    return dict(value=round(B / (ms * 1e-3), 1), unit="pairs/s", ms_per_step=round(ms, 3), steps=steps, h2d_bytes_per_step=int(h2d[0]),
                d2h_bytes_per_step=12, frame=f"{frame} x {frame} uint8",
                what="decoded 8-bit frames on the host -> crop boxes H2D -> GPU RandomResizedCrop(448, bicubic) + flip -> step on the uint8 batch; "
                     "parameter draw and packing run in this (single) Python thread inside the timed region")


def dropin_path(model, batches, B, T, pk, steps=10):
    """The reference trainer's call sequence on the drop-in module (main_pretrain.py:139-153): model(batch) under
    autocast -> loss.backward() (staged autograd nodes, ordinary .grad tensors) -> torch.optim.AdamW.step() -> zero_grad.
    Device-resident inputs; the vocabulary head is evaluated in forward AND recomputed in backward on this path."""
    decay = [p for n, p in model.named_parameters() if p.requires_grad and not (p.dim() == 1 or n.endswith(".bias"))]
    no_decay = [p for n, p in model.named_parameters() if p.requires_grad and (p.dim() == 1 or n.endswith(".bias"))]
    opt = torch.optim.AdamW([dict(params=no_decay, weight_decay=0.0), dict(params=decay, weight_decay=0.05)], lr=1.5e-4, betas=(0.9, 0.95))
    model.zero_grad(set_to_none=True)

    def step(i):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            mim, res, mlm = model(batches[i % 2])
            (mim + res + mlm).backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    ms = _timed_steps(step, steps, 3)
    v = B / (ms * 1e-3)
    del opt
    model.zero_grad(set_to_none=True)
    return dict(value=round(v, 1), unit="pairs/s", ms_per_step=round(ms, 3), steps=steps,
                what="model(batch) + loss.backward() + torch.optim.AdamW.step() through the reference API", **_frac(v, GFLOP_PER_PAIR_STEP[T], pk))


def gpu_eager_baseline(batches, B, T, pk, steps=5):
    """The reference's own PyTorch path ON THIS GPU: the oracle restatement (plain torch modules,
    cuBLAS / torch kernels) under bf16 autocast + torch.optim.AdamW, same batch size and inputs.  A reported baseline."""
    from oracle.ecamp_oracle import ecamp_oracle
    torch.manual_seed(0)
    dev = batches[0]["image"].device
    m = ecamp_oracle(dropout=0.1).to(dev).train()
    decay = [p for n, p in m.named_parameters() if p.requires_grad and not (p.dim() == 1 or n.endswith(".bias"))]
    no_decay = [p for n, p in m.named_parameters() if p.requires_grad and (p.dim() == 1 or n.endswith(".bias"))]
    opt = torch.optim.AdamW([dict(params=no_decay, weight_decay=0.0), dict(params=decay, weight_decay=0.05)], lr=1.5e-4, betas=(0.9, 0.95))

    def step(i):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            mim, res, mlm = m(batches[i % 2])
            loss = mim + res + mlm
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    ms = _timed_steps(step, steps, 2)
    v = B / (ms * 1e-3)
    del m, opt
    torch.cuda.empty_cache()
    return dict(value=round(v, 1), unit="pairs/s", ms_per_step=round(ms, 3), steps=steps, kind="port",
                what="oracle (plain torch modules) under bf16 autocast + torch.optim.AdamW on the same B200, same batch", **_frac(v, GFLOP_PER_PAIR_STEP[T], pk))


def config4_quick(dp, model, B, pk, steps=10):
    """BASELINE config 4's per-GPU shape: 448-px input, 256-token reports (the reference default), SR branch."""
    from ecamp_b200.synthetic import make_batch
    dev = next(model.parameters()).device
    bs = [{k: v.to(dev) for k, v in make_batch(B, T=256, big=True, seed=4000 + i).items()} for i in range(2)]
    for b_ in bs:
        b_.pop("noise")
    ms = _timed_steps(lambda i: dp.step(bs[i % 2]), steps, 3)
    v = B / (ms * 1e-3)
    return dict(value=round(v, 1), unit="pairs/s", ms_per_step=round(ms, 3), steps=steps, per_gpu_batch=B, seq_len=256,
                what="config 4 shape on one GPU: fwd + bwd + fused AdamW, T = 256", **_frac(v, GFLOP_PER_PAIR_STEP[256], pk))


def config5_quick(pk, B=512, steps=10):
    """BASELINE config 5: fine-tune ViT-B/16, 14 classes, batch 512, 224 px: fwd + BCE + bwd + clip + SGD-momentum."""
    from ecamp_b200.models_vit import vit_base_patch16
    from ecamp_b200.optim import FusedSGD
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(0)
    m = vit_base_patch16(num_classes=14, drop_path_rate=0.1, global_pool=True).to(dev).train()
    xs = [torch.randn(B, 3, 224, 224, device=dev) for _ in range(2)]
    ys = [(torch.rand(B, 14, device=dev) < 0.3).float() for _ in range(2)]
    loss_fct = torch.nn.BCEWithLogitsLoss()
    opt = FusedSGD(m.parameters(), lr=3e-3, momentum=0.9, weight_decay=0.0, max_grad_norm=1.0)

    def step(i):
        m.zero_grad(set_to_none=True)
        loss_fct(m(xs[i % 2]), ys[i % 2]).backward()
        opt.step()

    ms = _timed_steps(step, steps, 3)
    v = B / (ms * 1e-3)
    del m, opt, xs
    torch.cuda.empty_cache()
    return dict(value=round(v, 1), unit="images/s", ms_per_step=round(ms, 3), steps=steps, batch=B,
                what="config 5: ViT-B/16 14-class fine-tune step through ecamp_b200.models_vit + FusedSGD", **_frac(v, 105.38, pk))


def multi_rank_checks(model, opt, dp, world, rank, dev, seq, per_gpu_batch=256):
    """N > 1, after the timed loops: (1) replicas bit-identical (parameters + Adam moments vs rank 0); (2) gradient of
    the data-parallel step on per-rank slices of one batch == gradient of the whole batch on one GPU."""
    import torch.distributed as dist
    from ecamp_b200.synthetic import make_batch
    rt = model._rt
    vec = torch.cat([torch.cat([p.detach().flatten() for p in rt["params"]]), rt["M1"], rt["M2"]])
    ref = vec.clone()
    dist.broadcast(ref, src=0)
    d = (ref - vec).abs().max().reshape(1)
    dist.all_reduce(d, op=dist.ReduceOp.MAX)
    out = dict(replicas_identical=bool(d.item() == 0.0), max_abs_replica_diff=float(d.item()))
    del vec, ref
    per = 8
    full = {k: v.to(dev) for k, v in make_batch(per * world, T=seq, big=True, seed=777).items()}
    mine = {k: v[per * rank:per * (rank + 1)].contiguous() for k, v in full.items()}
    model.eval()
    for g in opt.param_groups:
        g["lr"] = 0.0
    model.zero_grad(set_to_none=True)
    dp.step(mine)                      # bucketed all-reduce overlapped with the staged backward; lr = 0: no update
    g_dp = model.flat_grads() / world
    g_dp = g_dp.clone()
    model.zero_grad(set_to_none=True)
    model.forward_backward(full)       # every rank: the whole batch on its own GPU, no communication
    g_full = model.flat_grads()
    r = ((g_dp.double() - g_full.double()).norm() / g_full.double().norm()).reshape(1).float()
    dist.all_reduce(r, op=dist.ReduceOp.MAX)
    out["dp_grad_rel"] = float(r.item())
    out["dp_grad_what"] = f"{per} pairs per rank through DataParallelStep vs the same {per * world} pairs on one GPU, rel-L2 over all gradients, max over ranks"
    model.train()
    # (3) how fast is each GPU on its own while all of them are busy?  Every rank runs the same full-size fused step WITHOUT
    # the all-reduce, all ranks at the same time (lr = 0: parameters stay identical): the spread is the part of the N-GPU
    # step time that is the pace of the slowest GPU of the box, not communication.
    batch = {k: v.to(dev) for k, v in make_batch(per_gpu_batch, T=seq, big=True, seed=778).items()}
    for _ in range(3):
        model.forward_backward(batch); opt.step(grad_scale=1.0); opt.zero_grad(set_to_none=True)
    dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    n = 25
    for _ in range(n):
        model.forward_backward(batch); opt.step(grad_scale=1.0); opt.zero_grad(set_to_none=True)
    e.record()
    torch.cuda.synchronize()
    mine_ms = torch.tensor([s.elapsed_time(e) / n], device=dev)
    allms = [torch.zeros_like(mine_ms) for _ in range(world)]
    dist.all_gather(allms, mine_ms)
    out["independent_step_ms_per_rank"] = [round(float(t.item()), 3) for t in allms]
    out["independent_step_what"] = f"{n} fused steps per rank at batch {per_gpu_batch}, no all-reduce, all ranks concurrently (lr = 0)"
    return out


def cpu_baseline(T):
    """The reference's CPU path (oracle port) on config 1: batch 8, fp32, forward + three losses."""
    from oracle.ecamp_oracle import ecamp_oracle, synthetic_batch
    torch.set_num_threads(os.cpu_count() or 1)
    m = ecamp_oracle().eval()
    b = synthetic_batch(8, T=T, seed=1234)
    with torch.no_grad():
        m(b)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); m(b); ts.append(time.perf_counter() - t0)
    return dict(value=round(8 / statistics.median(ts), 3), unit="pairs/s", cores=torch.get_num_threads(), kind="port",
                sample=f"BASELINE config 1: batch 8, fp32, 448-px input, T={T}, forward + 3 losses, median of 3 (1 warm-up)")


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the step (oracle port: the reference cannot be
    installed or imported unmodified here), forward + backward + torch.optim.AdamW, on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.ecamp_oracle import ecamp_oracle, synthetic_batch
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    Bc = 2
    m = ecamp_oracle(dropout=0.1).train()
    decay = [p for n, p in m.named_parameters() if p.requires_grad and not (p.dim() == 1 or n.endswith(".bias"))]
    no_decay = [p for n, p in m.named_parameters() if p.requires_grad and (p.dim() == 1 or n.endswith(".bias"))]
    opt = torch.optim.AdamW([dict(params=no_decay, weight_decay=0.0), dict(params=decay, weight_decay=0.05)], lr=1.5e-4, betas=(0.9, 0.95))
    b = synthetic_batch(Bc, T=args.seq, seed=1234)

    def step():
        mim, res, mlm = m(b)
        (mim + res + mlm).backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = Bc * args.steps / dt
    sample = f"{Bc} pairs per step (448-px input, T={args.seq}), fp32 forward + backward + AdamW, {torch.get_num_threads()} threads"
    print(json.dumps(dict(impl="reference", metric="pretrain_pairs_per_sec", value=round(v, 4), unit="pairs/s", n_gpus=args.gpus,
                          steps=args.steps, warmup=args.warmup, ms_per_step=round(1e3 * dt / args.steps, 2), higher_is_better=True,
                          scaling="strong" if args.global_batch else "weak", vs_baseline=None, dtype="f32", data="synthetic",
                          config=workload_config(args, int(os.environ.get("WORLD_SIZE", "1"))),
                          cpu_baseline=dict(value=round(v, 4), unit="pairs/s", cores=torch.get_num_threads(), kind="port", sample=sample),
                          e2e=dict(value=round(v, 4), unit="pairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))), flush=True)


E2E_INPUT_TEXT = {
    "u8": ("loader's 8-bit grayscale crop [B,448,448]; Grayscale(3)+ToTensor+Normalize run on the GPU, bit-exact with the "
           "CPU transform (tests/parity_checks.py::check_image_u8)"),
    "f32": "reference collate format: normalised fp32 [B,3,448,448]",
}


def workload_config(args, world):
    """The `config` object of the JSON line: a function of the command line only, so that the reference arm prints the
    SAME object (the driver compares the two arms' configs); run-specific notes go to `run_notes`."""
    return dict(workload=workload_name(args), global_batch=world * args.batch, per_gpu_batch=args.batch,
                seq_len=args.seq, image_px="448 -> 224", mask_ratio=0.75, parallelism=f"dp{world}",
                optimizer="AdamW lr 1.5e-4 betas (0.9,0.95) wd 0.05 over timm add_weight_decay groups",
                weights="random init (reference initialize_weights)",
                l2="per-step working set (~16 GB of activations) is far larger than the 126 MB L2; two input batches alternate",
                e2e_input=E2E_INPUT_TEXT[args.e2e_input])


def workload_name(args):
    return (f"BASELINE config {'2' if args.gpus == 1 else '3'}: ECAMP pretrain fwd+bwd+AdamW bf16, batch {args.batch}/GPU, 448-px input -> "
            f"224-px ViT-B/16 (196 patches, mask 0.75) + SR branch, {args.seq}-token reports, dropout 0.1")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=80, help="timed steps (default: >= 3 s of timed region at ~40 ms per step)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="pairs per GPU (BASELINE config 2: 256)")
    ap.add_argument("--seq", type=int, default=128)
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling (BASELINE config 3: 2048): the per-GPU batch becomes global / n_gpus and `scaling` "
                         "is reported as strong; 0 = weak scaling with --batch pairs per GPU")
    ap.add_argument("--e2e-input", default="u8", choices=["f32", "u8"],
                    help="host image format of the headline e2e arm: u8 (default) = the loader's 8-bit grayscale crop [B,448,448] "
                         "(INTEGRATION.md section 1, data loader: Grayscale(3) + ToTensor + Normalize run on the GPU, bit-identical "
                         "batch, 52 MB per step); f32 = the reference collate's normalised [B,3,448,448] tensor (617 MB per step). "
                         "The other format is always measured too and reported next to it")
    ap.add_argument("--no-extras", action="store_true", help="skip the roofline micro-timing and the CPU baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.global_batch:
        w_ = int(os.environ.get("WORLD_SIZE", "1"))
        if args.global_batch % w_:
            raise SystemExit("--global-batch must be a multiple of the number of GPUs")
        args.batch = args.global_batch // w_
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from ecamp_b200 import _lib as L
    from ecamp_b200.model_ecamp import ecamp
    from ecamp_b200.optim import FusedAdamW
    from ecamp_b200.parallel import DataParallelStep
    from ecamp_b200.synthetic import bind_host_to_gpu, make_batch

    torch.manual_seed(0)                      # identical random-init replicas on every rank
    model = ecamp(norm_pix_loss=True).to(dev).train()
    opt = FusedAdamW(model, lr=1.5e-4, betas=(0.9, 0.95), weight_decay=0.05)
    dp = DataParallelStep(model, opt, bucket_mb=64)
    torch.manual_seed(1234 + rank)            # per-rank masking noise / dropout (main_pretrain.py:189)
    affinity = os.sched_getaffinity(0)
    numa_cpus = bind_host_to_gpu(local)        # first-touch the pinned staging buffers on the GPU's own NUMA node ...
    host = [make_batch(args.batch, T=args.seq, big=True, seed=1234 + 17 * rank + i, pin=True) for i in range(2)]
    os.sched_setaffinity(0, affinity)          # ... then give the process its CPUs back (the CPU baseline uses them all)
    for hb in host:
        hb.pop("noise")                       # the module draws torch.rand(N, L) itself, like the reference
    resident = [to_device(hb, dev) for hb in host]
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- arm 1: inputs resident in HBM ---------------------------------------------------------------------
    for i in range(args.warmup):
        dp.step(resident[i % 2])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.lib().ecamp_launch_count()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(args.steps):
        losses = dp.step(resident[i % 2])
    e.record()
    barrier()
    ms = max_over_ranks(s.elapsed_time(e))
    launches = L.lib().ecamp_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    last_losses = [float(x) for x in losses.tolist()]
    value = world * args.batch * args.steps / (ms * 1e-3)

    # ---- arm 2: end to end through the public API, pinned-host inputs copied every step, losses read every step ----
    del resident
    copy_stream = torch.cuda.Stream()

    def run_e2e(host):
        h2d = sum(v.numel() * v.element_size() for v in host[0].values())
        for i in range(2):
            dp.step(to_device(host[i % 2], dev))
        barrier()
        pinned_out = [torch.empty(3, dtype=torch.float32).pin_memory() for _ in range(2)]
        read_ev = [torch.cuda.Event() for _ in range(2)]
        host_losses = []
        nxt = to_device(host[0], dev, copy_stream)                     # pipeline fill (a loader's first prefetch)
        s.record()
        for i in range(args.steps):
            torch.cuda.current_stream().wait_stream(copy_stream)
            cur_batch = nxt
            for v in cur_batch.values():
                v.record_stream(torch.cuda.current_stream())
            nxt = to_device(host[(i + 1) % 2], dev, copy_stream)      # prefetch the next step's inputs: exactly one
                                                                       # H2D copy of a full batch per timed step
            losses = dp.step(cur_batch)
            pinned_out[i % 2].copy_(losses, non_blocking=True)
            read_ev[i % 2].record()
            if i > 0:                                                  # the host reads step i-1's result while step i runs
                read_ev[(i - 1) % 2].synchronize()
                host_losses.append(pinned_out[(i - 1) % 2].tolist())
        read_ev[(args.steps - 1) % 2].synchronize()
        host_losses.append(pinned_out[(args.steps - 1) % 2].tolist())
        assert len(host_losses) == args.steps
        e.record()
        barrier()
        ms_ = max_over_ranks(s.elapsed_time(e))
        return dict(value=round(world * args.batch * args.steps / (ms_ * 1e-3), 2), unit="pairs/s", h2d_bytes_per_step=int(h2d),
                    d2h_bytes_per_step=12, ms_per_step=round(ms_ / args.steps, 3))

    def u8_batches():
        os.sched_setaffinity(0, numa_cpus or affinity)
        hb = [make_batch(args.batch, T=args.seq, big=True, seed=1234 + 17 * rank + i, pin=True, u8=True) for i in range(2)]
        os.sched_setaffinity(0, affinity)
        for b_ in hb:
            b_.pop("noise")
        return hb

    # headline: the reference collate's format (normalised fp32 [B,3,448,448]) unless --e2e-input u8; the other format is
    # measured right after it and reported next to it
    e2e_main = run_e2e(host if args.e2e_input == "f32" else u8_batches())
    e2e_other = run_e2e(u8_batches() if args.e2e_input == "f32" else host)
    e2e_f32, e2e_u8 = (e2e_main, e2e_other) if args.e2e_input == "f32" else (e2e_other, e2e_main)
    e2e_u8["input"] = E2E_INPUT_TEXT["u8"]
    e2e_f32["input"] = E2E_INPUT_TEXT["f32"]

    tl = None
    if world > 1:   # one recorded step: per-bucket all-reduce timeline (CUDA events), exposed communication per rank
        try:
            resident2 = to_device(host[0], dev)
            dp.timeline = []
            dp.step(resident2)
            t = dp.bucket_timeline()
            ex = torch.tensor([t["exposed_ms"]], device=dev)
            dist.all_reduce(ex, op=dist.ReduceOp.MAX)
            mine = torch.tensor([t["backward_end_ms"], t["joined_ms"], t["exposed_ms"]], device=dev)
            allr = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            per_rank = [dict(rank=r, backward_end_ms=round(float(v[0]), 3), joined_ms=round(float(v[1]), 3), exposed_ms=round(float(v[2]), 3))
                        for r, v in enumerate(allr)]
            # the all-reduce of a bucket completes when the SLOWEST rank has delivered it: on the fast ranks "exposed" is time
            # spent waiting for the straggler (GPU-to-GPU clock differences under the power cap), on the slowest rank it is
            # the communication that really did not overlap
            slowest = max(per_rank, key=lambda r: r["backward_end_ms"])
            tl = dict(rank0=t, per_rank=per_rank, slowest_rank=slowest, backward_end_skew_ms=round(slowest["backward_end_ms"] -
                      min(r["backward_end_ms"] for r in per_rank), 3),
                      exposed_ms_max_over_ranks=round(float(ex.item()), 3), n_buckets=len(t["buckets"]),
                      nccl=dict(NCCL_ALGO=os.environ.get("NCCL_ALGO"), NCCL_MAX_CTAS=os.environ.get("NCCL_MAX_CTAS"),
                                NCCL_NVLS_ENABLE=os.environ.get("NCCL_NVLS_ENABLE")))
            del resident2
        except Exception as ex_:  # noqa
            tl = dict(error=str(ex_)[:300])
    mr = None
    if world > 1 and not args.no_extras:
        try:
            mr = multi_rank_checks(model, opt, dp, world, rank, dev, args.seq, args.batch)
        except Exception as ex:  # noqa
            mr = dict(error=str(ex)[:300])
    if rank == 0:
        pk, pk_kind = peaks()
        gf = GFLOP_PER_PAIR_STEP.get(args.seq)
        out = dict(metric="pretrain_pairs_per_sec", value=round(value, 2), unit="pairs/s", n_gpus=world, steps=args.steps,
                   warmup=args.warmup, ms_per_step=round(ms / args.steps, 3), higher_is_better=True,
                   scaling="strong" if args.global_batch else "weak",
                   vs_baseline=None, dtype="bf16", data="synthetic",
                   config=workload_config(args, world),
                   run_notes=dict(optimizer_impl="fused AdamW (ecamp_b200.optim.FusedAdamW)",
                                  e2e_pipeline="per timed step: one H2D copy of a full pinned batch on a copy stream (prefetching the "
                                               "next step's inputs while this step runs) and one D2H read of the 3 losses, which the "
                                               "host consumes one step late so that it never stalls the launch queue",
                                  host_buffers=("pinned, first-touched on the GPU's NUMA node (%d CPUs)" % len(numa_cpus)) if numa_cpus
                                  else "pinned (NUMA topology unknown or single node)"),
                   clocks=clocks, gpu_launches=int(launches),
                   e2e={k: v for k, v in e2e_main.items() if k != "input"},
                   e2e_u8_input=e2e_u8, e2e_f32_input=e2e_f32, losses_last_step=last_losses)
        if gf:
            tf = value / world * gf / 1e3
            out["step_tflops_per_gpu"] = round(tf, 1)
            out["step_frac_of_bf16_peak"] = dict(burst=round(tf / pk["bf16_tflops"], 4), sustained=round(tf / pk["bf16_tflops_sustained"], 4),
                                                 peaks=pk_kind, gflop_per_pair=gf)
        if mr is not None:
            out["multi_rank_checks"] = mr
        if tl is not None:
            out["allreduce_timeline"] = tl
        if not args.no_extras and world == 1 and args.seq in GFLOP_PER_PAIR_STEP:
            # outside the timed headline: the reference-API path, the eager-torch baseline on this GPU, configs 4 and 5
            res2 = [to_device(hb, dev) for hb in host]
            for name, fn in (("e2e_gpu_loader", lambda: e2e_gpu_loader(dp, host, args.batch, dev)),
                             ("dropin_path", lambda: dropin_path(model, res2, args.batch, args.seq, pk)),
                             ("gpu_eager_baseline", lambda: gpu_eager_baseline(res2, args.batch, args.seq, pk)),
                             ("config4", lambda: config4_quick(dp, model, args.batch, pk)),
                             ("config5", lambda: config5_quick(pk))):
                try:
                    out[name] = fn()
                except Exception as ex:  # noqa
                    out[name] = dict(error=str(ex)[:300])
                torch.cuda.empty_cache()
            del res2
        if not args.no_extras:
            del host
            torch.cuda.empty_cache()
            try:
                out["roofline"] = gemm_roofline(args.batch, args.seq, pk["bf16_tflops"])
                out["roofline"]["peak_kind"] = f"{pk_kind} burst bf16 (kernel timed alone)"
            except Exception as ex:  # noqa
                out["roofline"] = dict(error=str(ex)[:200])
            if world == 1:
                try:
                    out["cpu_baseline"] = cpu_baseline(args.seq)
                except Exception as ex:  # noqa
                    out["cpu_baseline"] = dict(error=str(ex)[:200])
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
