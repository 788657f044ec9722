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
        rt = self.model._rt
        if rt is None or rt.get("G") is None:
            raise RuntimeError("FusedAdamW.step() before any backward of the ECAMP module")
        return rt

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        rt = self._rt()
        lrs = {g["lr"] * g.get("lr_scale", 1.0) for g in self.param_groups}
        if len(lrs) != 1:
            raise RuntimeError("FusedAdamW applies one learning rate to both groups (as the reference schedule does)")
        g1 = self.param_groups[1]
        self.step_count += 1
        L.check(L.lib().ecamp_adamw_step(rt["ctx"], ctypes.c_float(lrs.pop()), ctypes.c_float(g1["betas"][0]),
                                         ctypes.c_float(g1["betas"][1]), ctypes.c_float(g1["eps"]),
                                         ctypes.c_float(g1["weight_decay"]), ctypes.c_int32(self.step_count),
                                         ctypes.c_float(grad_scale), L.cur_stream()), "ecamp_adamw_step")

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
