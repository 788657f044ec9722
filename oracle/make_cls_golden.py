"""TEST INFRASTRUCTURE.  Runs ONLY in the build container (needs /root/reference).

Pins oracle/vit_cls_oracle.py against the reference's own fine-tune classifier, executed FROM SOURCE:
`/root/reference/ECAMP/Fine-tuning/Classification/models_vit.py` is imported unmodified and
`vit_base_patch16(num_classes=14, drop_path_rate=0.1, global_pool=True)` (train.py:124-128) is constructed and run.
The class derives from timm 0.4.12's VisionTransformer, which is not installed here and not vendored in the reference
(environment.yml:128 pins the version): its constructor, Block and DropPath are restated below from the published
0.4.12 source (constructor: patch_embed, cls_token, pos_embed, pos_drop, blocks with drop_path = linspace(0, rate, depth),
norm, head; forward = head(forward_features(x)); DropPath: x / keep * floor(keep + U[0,1)) per sample, training only).
What runs from the reference is therefore its `__init__` override (fc_norm, `del self.norm`) and its `forward_features`
(cls prepend, + pos_embed, blocks, mean over the PATCH tokens, fc_norm) — the part the oracle header used to call unpinned.

Checks (asserted here) and fixtures (tests/golden/cls_cases.json, re-checked on CPU by tests/test_oracle_cpu.py and used
on the GPU by tests/parity_checks.py::check_finetune_cls):
  * state_dict keys / shapes of the reference class == the oracle's == ecamp_b200.models_vit's
  * eval mode: logits, BCE-with-logits loss (train.py:422-423,443) and every parameter gradient, oracle vs reference
  * train mode with DropPath: the same, with the reference's torch.rand draws injected into the oracle's scales

Usage:  python oracle/make_cls_golden.py
"""
import json
import os
import sys
import types

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import _ref_shims  # noqa: E402

REF_FT = "/root/reference/ECAMP/Fine-tuning/Classification"
TOL = 2e-5


def install_timm_vit():
    """timm 0.4.12 VisionTransformer / Block-with-DropPath / layers.to_2tuple, restated (see the module docstring)."""
    _ref_shims.install()
    vt = sys.modules["timm.models.vision_transformer"]
    base_block, patch_embed = vt.Block, vt.PatchEmbed

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0. or not self.training:
                return x
            keep = 1 - self.drop_prob
            shape = (x.shape[0],) + (1,) * (x.ndim - 1)
            r = keep + torch.rand(shape, dtype=x.dtype, device=x.device)
            r.floor_()
            return x.div(keep) * r

    class Block(base_block):
        def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, drop_path=0., norm_layer=nn.LayerNorm, **kw):
            super().__init__(dim, num_heads, mlp_ratio, qkv_bias, norm_layer)
            self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()

        def forward(self, x):
            x = x + self.drop_path(self.attn(self.norm1(x)))
            x = x + self.drop_path(self.mlp(self.norm2(x)))
            return x

    class VisionTransformer(nn.Module):
        def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                     mlp_ratio=4., qkv_bias=True, drop_rate=0., drop_path_rate=0., norm_layer=None, **kw):
            super().__init__()
            self.num_classes, self.embed_dim = num_classes, embed_dim
            norm_layer = norm_layer or nn.LayerNorm
            self.patch_embed = patch_embed(img_size, patch_size, in_chans, embed_dim)
            self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
            self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
            self.pos_drop = nn.Dropout(p=drop_rate)
            dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
            self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, mlp_ratio, qkv_bias, drop_path=dpr[i], norm_layer=norm_layer)
                                          for i in range(depth)])
            self.norm = norm_layer(embed_dim)
            self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()

        def forward(self, x):
            return self.head(self.forward_features(x))

    vt.VisionTransformer, vt.Block = VisionTransformer, Block
    layers = types.ModuleType("timm.models.layers")
    layers.to_2tuple = lambda v: v if isinstance(v, tuple) else (v, v)
    sys.modules["timm.models.layers"] = layers
    sys.modules["timm.models"].layers = layers
    ml = types.ModuleType("ml_collections")
    ml.ConfigDict = dict
    sys.modules.setdefault("ml_collections", ml)
    if REF_FT not in sys.path:
        sys.path.insert(0, REF_FT)


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def run(model, x, y, scales=None):
    model.zero_grad(set_to_none=True)
    logits = model(x) if scales is None else model(x, scales)
    loss = nn.BCEWithLogitsLoss()(logits, y)
    loss.backward()
    return logits.detach(), loss.detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}


def main():
    install_timm_vit()
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_models_vit", os.path.join(REF_FT, "models_vit.py"))
    ref_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_mod)
    from oracle.vit_cls_oracle import VitClsOracle, seeded_cls_state

    torch.manual_seed(0)
    ref = ref_mod.vit_base_patch16(num_classes=14, drop_path_rate=0.1, global_pool=True)   # train.py:124-128
    orc = VitClsOracle(num_classes=14, drop_path_rate=0.1)
    rkeys = {k: list(v.shape) for k, v in ref.state_dict().items()}
    okeys = {k: list(v.shape) for k, v in orc.state_dict().items()}
    assert rkeys == okeys, sorted(set(rkeys) ^ set(okeys))
    assert not hasattr(ref, "norm") and isinstance(ref.fc_norm, nn.LayerNorm) and ref.fc_norm.eps == 1e-6
    sd = seeded_cls_state(ref, 11)
    ref.load_state_dict(sd)
    orc.load_state_dict(sd)

    B = 2
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, 3, 224, 224, generator=g)
    y = (torch.rand(B, 14, generator=g) < 0.3).float()
    out = {"seed_weights": 11, "seed_inputs": 5, "B": B, "layout": rkeys, "cases": {}}

    # ---- eval ----
    ref.eval(); orc.eval()
    lr_, ls_r, gr = run(ref, x, y)
    lo_, ls_o, go = run(orc, x, y)
    worst = max(rel(go[k], gr[k]) for k in gr)
    print(f"eval : logits rel {rel(lo_, lr_):.2e}  loss {ls_o.item():.6f} vs {ls_r.item():.6f}  worst grad rel {worst:.2e}")
    assert rel(lo_, lr_) < TOL and abs(ls_o.item() - ls_r.item()) < TOL * abs(ls_r.item()) and worst < 5e-4
    out["cases"]["eval"] = dict(logits=lr_.tolist(), loss=ls_r.item(), grad_norms={k: v.norm().item() for k, v in gr.items()})

    # ---- train: the reference draws torch.rand((B,1,1)) per DropPath call, in execution order ----
    ref.train(); orc.train()
    draws = []
    orig = torch.rand

    def spy(*a, **k):
        r = orig(*a, **k)
        draws.append(r.reshape(-1).clone())
        return r
    torch.manual_seed(21)
    torch.rand = spy
    try:
        lr_, ls_r, gr = run(ref, x, y)
    finally:
        torch.rand = orig
    rates = [v.item() for v in torch.linspace(0, 0.1, 12)]
    assert len(draws) == 2 * sum(r > 0 for r in rates)
    scales = torch.ones(12, 2, B)
    it = iter(draws)
    for l, p in enumerate(rates):
        if p > 0:
            for br in range(2):
                scales[l, br] = torch.floor((1 - p) + next(it)) / (1 - p)
    lo_, ls_o, go = run(orc, x, y, scales)
    worst = max(rel(go[k], gr[k]) for k in gr)
    print(f"train: logits rel {rel(lo_, lr_):.2e}  loss {ls_o.item():.6f} vs {ls_r.item():.6f}  worst grad rel {worst:.2e}  "
          f"dropped branches {(scales == 0).sum().item()}")
    assert rel(lo_, lr_) < TOL and abs(ls_o.item() - ls_r.item()) < TOL * abs(ls_r.item()) and worst < 5e-4
    out["cases"]["train_droppath"] = dict(scales=scales.tolist(), logits=lr_.tolist(), loss=ls_r.item(),
                                          grad_norms={k: v.norm().item() for k, v in gr.items()})
    path = os.path.join(ROOT, "tests", "golden", "cls_cases.json")
    json.dump(out, open(path, "w"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
