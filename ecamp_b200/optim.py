"""Fused AdamW over the ECAMP parameters (one multi-tensor kernel, bf16 GEMM copies refreshed in the same pass).

Host-side mirror of what the reference builds at ECAMP/Pre-training/main_pretrain.py:253-254:
    param_groups = optim_factory.add_weight_decay(model, wd); AdamW(param_groups, lr, betas=(0.9, 0.95))
`param_groups` has the same two groups in the same order (no-decay first), so the reference's
util/lr_sched.py:9-21 (which writes param_group["lr"]) and util/misc.py:295-338 (optimizer state_dict in a
checkpoint) keep working.  Parameters that never receive a gradient (bert.pooler.*) get no state and no
decay, exactly like torch.optim.AdamW skips `p.grad is None`.
"""
import ctypes

import torch

from . import _lib as L


class FusedAdamW:
    def __init__(self, model, lr=1.5e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05):
        self.model = model
        self.defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        no_decay, decay = [], []
        for name, p in model.named_parameters():
            if not p.requires_grad:
                continue
            (no_decay if (p.dim() == 1 or name.endswith(".bias")) else decay).append(p)  # timm add_weight_decay
        self.param_groups = [dict(params=no_decay, weight_decay=0.0, lr=lr, betas=betas, eps=eps),
                             dict(params=decay, weight_decay=weight_decay, lr=lr, betas=betas, eps=eps)]
        self.step_count = 0

    def zero_grad(self, set_to_none=True):
        for g in self.param_groups:
            for p in g["params"]:
                if set_to_none:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.zero_()

    def _rt(self):
        """The module's native runtime (flat gradient / Adam-state buffers).  Bound eagerly, so the optimizer state can be
        restored before the first forward - where the reference's misc.load_model does it (util/misc.py:331-335)."""
        p0 = self.param_groups[0]["params"][0]
        if p0.device.type != "cuda":
            raise RuntimeError("FusedAdamW: move the ECAMP module to the GPU first (there is no CPU path)")
        return self.model._runtime(p0.device)

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        rt = self._rt()
        if not self.model._sync_flat_grads(rt):   # .grad tensors left by autograd / DDP / GradScaler -> flat buffer
            return                                # no parameter has a gradient: torch.optim.AdamW would skip them all
        # util/lr_sched.py:9-21 has already folded "lr_scale" into param_group["lr"]: use it as is, one rate per group
        g0, g1 = self.param_groups
        if tuple(g0["betas"]) != tuple(g1["betas"]) or g0["eps"] != g1["eps"]:
            raise RuntimeError("FusedAdamW: the two parameter groups must share betas and eps")
        self.step_count += 1
        L.check(L.lib().ecamp_adamw_step_groups(rt["ctx"], ctypes.c_float(g1["lr"]), ctypes.c_float(g0["lr"]),
                                                ctypes.c_float(g1["betas"][0]), ctypes.c_float(g1["betas"][1]),
                                                ctypes.c_float(g1["eps"]), ctypes.c_float(g1["weight_decay"]),
                                                ctypes.c_int32(self.step_count), ctypes.c_float(grad_scale), L.cur_stream()),
                "ecamp_adamw_step_groups")

    # ---- one optimizer step in pieces, each as soon as its gradient slice is final (parallel.DataParallelStep) ----------
    def begin_step(self):
        """Start an optimizer step that will be applied slice by slice with `step_range`; every parameter must be covered
        by exactly one slice.  Reads the learning rates of the two groups now (the scheduler has run before the step)."""
        g0, g1 = self.param_groups
        if tuple(g0["betas"]) != tuple(g1["betas"]) or g0["eps"] != g1["eps"]:
            raise RuntimeError("FusedAdamW: the two parameter groups must share betas and eps")
        self.step_count += 1
        return self._rt()

    @torch.no_grad()
    def step_range(self, rt, lo, hi, grad_scale=1.0):
        """AdamW for the tensors whose gradients are flat[lo:hi] (a backward stage range or a union of consecutive ones), on
        the current stream.  The gradients are read from the flat buffer, where the fused backward leaves them."""
        g0, g1 = self.param_groups
        L.check(L.lib().ecamp_adamw_step_range(rt["ctx"], ctypes.c_float(g1["lr"]), ctypes.c_float(g0["lr"]),
                                               ctypes.c_float(g1["betas"][0]), ctypes.c_float(g1["betas"][1]),
                                               ctypes.c_float(g1["eps"]), ctypes.c_float(g1["weight_decay"]),
                                               ctypes.c_int32(self.step_count), ctypes.c_float(grad_scale),
                                               ctypes.c_int64(lo), ctypes.c_int64(hi), L.cur_stream()),
                "ecamp_adamw_step_range")

    # ---- torch.optim-compatible checkpoint layout -------------------------------------------------------
    def _index(self):
        idx, i = {}, 0
        for g in self.param_groups:
            for p in g["params"]:
                idx[id(p)] = i
                i += 1
        return idx

    def state_dict(self):
        rt = self._rt()
        idx = self._index()
        state = {}
        for p, off, n in zip(rt["params"], rt["goff"], rt["numel"]):
            if self.step_count == 0:
                break
            state[idx[id(p)]] = dict(step=torch.tensor(float(self.step_count)),
                                     exp_avg=rt["M1"][off:off + n].view(p.shape).clone(),
                                     exp_avg_sq=rt["M2"][off:off + n].view(p.shape).clone())
        groups, i = [], 0
        for g in self.param_groups:
            d = {k: v for k, v in g.items() if k != "params"}
            d["params"] = list(range(i, i + len(g["params"])))
            i += len(g["params"])
            groups.append(d)
        return dict(state=state, param_groups=groups)

    def load_state_dict(self, sd):
        rt = self._rt()
        idx = self._index()
        for g, sg in zip(self.param_groups, sd["param_groups"]):
            for k, v in sg.items():
                if k != "params":
                    g[k] = v
        steps = set()
        for p, off, n in zip(rt["params"], rt["goff"], rt["numel"]):
            st = sd["state"].get(idx[id(p)])
            if st is None:
                continue
            rt["M1"][off:off + n].copy_(st["exp_avg"].reshape(-1))
            rt["M2"][off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(st["step"]))
        if len(steps) > 1:
            raise RuntimeError("FusedAdamW keeps one step counter; the checkpoint has several")
        self.step_count = steps.pop() if steps else 0


class FusedSGD(torch.optim.Optimizer):
    """SGD with momentum + global gradient-norm clipping in two native launches (csrc/sgd.cu), no host sync.

    Mirrors the fine-tune trainer's optimizer (ECAMP/Fine-tuning/Classification/train.py:377-380:
    `torch.optim.SGD(model.parameters(), lr, momentum=0.9, weight_decay=wd)`) and folds in the
    `torch.nn.utils.clip_grad_norm_(model.parameters(), args.max_grad_norm)` the trainer calls right before
    `optimizer.step()` (train.py:459-463): pass `max_grad_norm` here and drop that call (or keep it: clipping twice
    is idempotent).  It is a torch.optim.Optimizer, so the reference's WarmupCosineSchedule / LambdaLR
    (train.py:388-392) drive `param_groups[i]["lr"]` unchanged and `state_dict()` has torch.optim.SGD's layout
    (`momentum_buffer` per parameter).  Dampening and Nesterov are not supported (the reference uses neither).
    Parameters must be contiguous fp32 CUDA tensors; there is no CPU path."""

    def __init__(self, params, lr, momentum=0.9, weight_decay=0.0, max_grad_norm=0.0, write_clipped_grads=False):
        if lr < 0 or momentum < 0 or weight_decay < 0:
            raise ValueError("FusedSGD: lr, momentum and weight_decay must be non-negative")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        self.max_grad_norm = float(max_grad_norm or 0.0)
        self.write_clipped_grads = bool(write_clipped_grads)
        self._tables = {}
        self._sumsq = None

    def _table(self, gi, group):
        ps = [p for p in group["params"] if p.grad is not None]
        for p in ps:
            if p.device.type != "cuda" or p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                raise RuntimeError("FusedSGD: parameters and gradients must be contiguous fp32 CUDA tensors (no CPU path)")
            st = self.state[p]
            if "momentum_buffer" not in st or st["momentum_buffer"] is None:
                st["momentum_buffer"] = torch.zeros_like(p)   # momentum * 0 + g == torch's first-step buf = g
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["momentum_buffer"].data_ptr()) for p in ps)
        t = self._tables.get(gi)
        if t is None or t["key"] != key:
            lib = L.lib()
            n = len(ps)
            host = (L.SgdTensor * max(n, 1))()
            numel = (ctypes.c_int64 * max(n, 1))()
            for i, p in enumerate(ps):
                host[i].p, host[i].g = p.data_ptr(), p.grad.data_ptr()
                host[i].buf, host[i].numel = self.state[p]["momentum_buffer"].data_ptr(), p.numel()
                numel[i] = p.numel()
            dev = ps[0].device if ps else torch.device("cuda")
            t = dict(key=key, n=n, chunks=ctypes.c_int64(0), params=ps)
            if n:
                t["tab"] = torch.empty(lib.ecamp_sgd_table_bytes(n), dtype=torch.uint8, device=dev)
                t["chk"] = torch.empty(lib.ecamp_sgd_chunk_bytes(numel, n), dtype=torch.uint8, device=dev)
                L.check(lib.ecamp_sgd_build_tables(host, n, L.ptr(t["tab"]), L.ptr(t["chk"]), ctypes.byref(t["chunks"])),
                        "ecamp_sgd_build_tables")
            self._tables[gi] = t
        return t

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.lib()
        tabs = [(g, self._table(gi, g)) for gi, g in enumerate(self.param_groups)]
        tabs = [(g, t) for g, t in tabs if t["n"]]
        if not tabs:
            return loss
        dev = tabs[0][1]["params"][0].device
        clip = self.max_grad_norm > 0
        if clip:
            if self._sumsq is None or self._sumsq.device != dev:
                self._sumsq = torch.zeros(len(self.param_groups) + 1, dtype=torch.float32, device=dev)
            parts = self._sumsq[1:1 + len(tabs)]
            for i, (g, t) in enumerate(tabs):      # one norm over ALL groups, as clip_grad_norm_(model.parameters())
                L.check(lib.ecamp_grad_sumsq(L.ptr(t["tab"]), L.ptr(t["chk"]), t["chunks"], L.ptr(parts[i:i + 1]), L.cur_stream()),
                        "ecamp_grad_sumsq")
            total = self._sumsq[0:1]
            torch.sum(parts, dim=0, keepdim=True, out=total) if len(tabs) > 1 else total.copy_(parts[0:1])
        for g, t in tabs:
            L.check(lib.ecamp_sgd_momentum_step(L.ptr(t["tab"]), L.ptr(t["chk"]), t["chunks"], ctypes.c_float(g["lr"]),
                                                ctypes.c_float(g["momentum"]), ctypes.c_float(g["weight_decay"]),
                                                ctypes.c_int32(0), ctypes.c_float(self.max_grad_norm),
                                                L.ptr(self._sumsq[0:1]) if clip else ctypes.c_void_p(0),
                                                ctypes.c_int32(1 if self.write_clipped_grads else 0), L.cur_stream()),
                    "ecamp_sgd_momentum_step")
            for p in t["params"]:   # updated in place by the kernel: bump ._version so modules re-derive their bf16 copies
                torch.autograd.graph.increment_version(p)
        return loss

    def grad_norm(self):
        """The global gradient norm of the last clipped step (device scalar), as clip_grad_norm_ returns it."""
        if self._sumsq is None:
            raise RuntimeError("FusedSGD.grad_norm(): no clipped step yet")
        return self._sumsq[0].sqrt()
