"""Test infrastructure only.  Shims that let the UNMODIFIED reference sources under
/root/reference/ECAMP/Pre-training be imported in this container (no timm / ipdb, torch 2.x,
numpy 2.x, transformers 5.x).  Used by oracle/make_golden.py to validate the oracle restatement
and to generate tests/golden/*.  Never imported by the product (ecamp_b200/) and never available
on the GPU box (/root/reference does not exist there).

The timm stub restates timm 0.4.12 `PatchEmbed` / `Block` (environment.yml:128 pins the version;
the source is not vendored in the reference): Block = x + attn(norm1(x)); x + mlp(norm2(x));
Attention = fused qkv Linear, reshape(B,N,3,H,C/H).permute(2,0,3,1,4), softmax(q k^T / sqrt(d)) v,
proj; Mlp = fc1 -> GELU(erf) -> fc2; every drop rate is 0 on this path.
"""
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF_PT = "/root/reference/ECAMP/Pre-training"


def _timm_stub():
    class PatchEmbed(nn.Module):
        def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
            super().__init__()
            self.img_size = (img_size, img_size)
            self.patch_size = (patch_size, patch_size)
            self.num_patches = (img_size // patch_size) ** 2
            self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

        def forward(self, x):
            return self.proj(x).flatten(2).transpose(1, 2)

    class Attention(nn.Module):
        def __init__(self, dim, num_heads=8, qkv_bias=False):
            super().__init__()
            self.num_heads = num_heads
            self.scale = (dim // num_heads) ** -0.5
            self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
            self.proj = nn.Linear(dim, dim)

        def forward(self, x):
            B, N, C = x.shape
            qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
            q, k, v = qkv[0], qkv[1], qkv[2]
            attn = (q @ k.transpose(-2, -1)) * self.scale
            attn = attn.softmax(dim=-1)
            x = (attn @ v).transpose(1, 2).reshape(B, N, C)
            return self.proj(x)

    class Mlp(nn.Module):
        def __init__(self, in_features, hidden_features):
            super().__init__()
            self.fc1 = nn.Linear(in_features, hidden_features)
            self.act = nn.GELU()
            self.fc2 = nn.Linear(hidden_features, in_features)

        def forward(self, x):
            return self.fc2(self.act(self.fc1(x)))

    class Block(nn.Module):
        def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, norm_layer=nn.LayerNorm, **kw):
            super().__init__()
            self.norm1 = norm_layer(dim)
            self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias)
            self.norm2 = norm_layer(dim)
            self.mlp = Mlp(dim, int(dim * mlp_ratio))

        def forward(self, x):
            x = x + self.attn(self.norm1(x))
            x = x + self.mlp(self.norm2(x))
            return x

    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    vt = types.ModuleType("timm.models.vision_transformer")
    vt.PatchEmbed, vt.Block = PatchEmbed, Block
    timm.models, models.vision_transformer = models, vt
    timm.__spec__ = None
    return {"timm": timm, "timm.models": models, "timm.models.vision_transformer": vt}


def install():
    """Make `from module.model_ecamp import ecamp` importable from the reference tree."""
    import transformers  # noqa: F401  (must be imported BEFORE the spec-less timm stub is registered)
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    from transformers.configuration_utils import PretrainedConfig
    if not hasattr(mu, "apply_chunking_to_forward"):
        mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    for k, v in dict(is_decoder=False, add_cross_attention=False, chunk_size_feed_forward=0).items():
        if not hasattr(PretrainedConfig, k):
            setattr(PretrainedConfig, k, v)
    sys.modules.setdefault("ipdb", types.ModuleType("ipdb"))
    for k, v in _timm_stub().items():
        sys.modules.setdefault(k, v)
    if not hasattr(np, "float"):
        np.float = float
    # the reference hard-codes .cuda() (model_ecamp.py:211-212,312-317); on this CPU box make it a no-op
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    if REF_PT not in sys.path:
        sys.path.insert(0, REF_PT)
