"""CPU/GPU oracle of the fine-tune classification model — TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench cpu_baseline).

Plain-torch restatement of ECAMP/Fine-tuning/Classification/models_vit.py:60-128 (`vit_base_patch16(num_classes=14,
drop_path_rate=0.1, global_pool=True)`, constructed at train.py:124-128): timm 0.4.12 VisionTransformer with the
overridden forward_features (:78-98): patch_embed -> prepend cls -> + pos_embed (learnable) -> 12 Blocks with DropPath
-> mean over the PATCH tokens -> fc_norm -> head; loss = BCEWithLogitsLoss (train.py:422-423,443).
Pinned by oracle/make_cls_golden.py: the reference's models_vit.py is imported UNMODIFIED there (its timm 0.4.12 base class
restated, since timm is neither installed nor vendored) and its logits, BCE loss and every parameter gradient equal this
oracle's in eval mode and with injected DropPath draws (fixtures: tests/golden/cls_cases.json, re-checked on CPU by
tests/test_oracle_cpu.py).
DropPath follows timm.models.layers.drop_path: per-sample mask floor(keep + U[0,1)) / keep, rates linspace(0, rate, depth);
the masks can be injected so that the CUDA path and the oracle see the same draw.
"""
import torch
import torch.nn as nn

from .ecamp_oracle import Block, PatchEmbed


class VitClsOracle(nn.Module):
    def __init__(self, num_classes=14, drop_path_rate=0.1, depth=12, embed_dim=768, num_heads=12):
        super().__init__()
        self.patch_embed = PatchEmbed(224, 16, 3, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, 197, embed_dim))
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, 4, 1e-6) for _ in range(depth)])
        self.fc_norm = nn.LayerNorm(embed_dim, eps=1e-6)   # models_vit.py:70 (self.norm is deleted, :72)
        self.head = nn.Linear(embed_dim, num_classes)
        self.drop_rates = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.trunc_normal_(self.cls_token, std=.02)

    def draw_drop_path(self, B, device, generator=None):
        """[depth, 2, B] scales mask / keep_prob (what timm's DropPath multiplies the branch with)."""
        out = torch.ones(len(self.blocks), 2, B, device=device)
        for l, p in enumerate(self.drop_rates):
            if p > 0:
                keep = 1 - p
                out[l] = torch.floor(keep + torch.rand(2, B, device=device, generator=generator)) / keep
        return out

    def forward(self, x, drop_path_scales=None):
        B = x.shape[0]
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(B, -1, -1), x), dim=1) + self.pos_embed
        for l, blk in enumerate(self.blocks):
            if drop_path_scales is None:
                x = blk(x)
            else:
                x = x + drop_path_scales[l, 0].view(B, 1, 1) * blk.attn(blk.norm1(x))
                x = x + drop_path_scales[l, 1].view(B, 1, 1) * blk.mlp(blk.norm2(x))
        x = x[:, 1:, :].mean(dim=1)          # global pool without cls token (models_vit.py:92)
        return self.head(self.fc_norm(x))


def seeded_cls_state(model, seed):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in sorted(model.state_dict().items()):   # sorted: independent of the registration order of the class
        if k.endswith("norm1.weight") or k.endswith("norm2.weight") or k == "fc_norm.weight":
            sd[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        elif v.dim() == 1:
            sd[k] = 0.05 * torch.randn(v.shape, generator=g)
        else:
            sd[k] = 0.03 * torch.randn(v.shape, generator=g)
    return sd
