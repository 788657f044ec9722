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
    # model_ecamp.py:318 calls torchvision.transforms.Resize([224, 224], BICUBIC) without `antialias`: on tensors the pinned
    # torchvision 0.14.1 (environment.yml) does NOT antialias (default None), torchvision >= 0.17 does (default True).
    import torchvision
    _Resize = torchvision.transforms.Resize
    if not getattr(_Resize, "_ecamp_pinned_default", False):
        class Resize(_Resize):
            _ecamp_pinned_default = True

            def __init__(self, size, interpolation=torchvision.transforms.InterpolationMode.BILINEAR, max_size=None, antialias=False):
                super().__init__(size, interpolation=interpolation, max_size=max_size, antialias=antialias)
        torchvision.transforms.Resize = Resize


# ------------------------------------------------------------------------------------------------------------------
# transformers 4.42.4 call conventions for the leaf modules the reference's BERT code drives
# ------------------------------------------------------------------------------------------------------------------
def install_bert_adapters(ref_model):
    """Let the reference's OWN BERT orchestration run from source on transformers 5.x: `MultiModalBertEncoder.forward`
    (module/bert_encoder.py:19-22), `MultimodalBertMaskedLM.forward` (module/bert_modeling.py:165-228),
    `MultimodalBertModel.forward` (:15-156) and `ECAMPFusionLayer.forward` (module/context_fusion.py:21-72) are executed
    unmodified; only the transformers-4.42.4 entry points they call, whose signatures changed in 5.x, are given their
    4.42.4 calling convention back on THIS model instance (restating the published 4.42.4 algorithms on the module's own
    parameters): `get_extended_attention_mask(mask, shape, device)`, `get_head_mask`, `BertSelfAttention.forward` with
    its seven positional arguments (eager path: softmax(q k^T / sqrt(d) + mask) -> dropout -> @ v, keys / values and the
    mask taken from the encoder arguments when given), `BertAttention.forward(hidden, mask, head_mask=, output_attentions=,
    past_key_value=)` and `BertEncoder.forward(..., head_mask=, output_attentions=, output_hidden_states=, return_dict=)`."""
    import math
    import types as _types

    from transformers.modeling_outputs import BaseModelOutputWithPastAndCrossAttentions
    from transformers.models.bert.modeling_bert import BertAttention, BertSelfAttention

    bert = ref_model.bert_encoder.model.bert
    cfg = bert.config
    for k, v in dict(output_attentions=False, output_hidden_states=False, use_cache=False).items():
        if getattr(cfg, k, None) is None:
            setattr(cfg, k, v)
    if not hasattr(cfg, "use_return_dict"):
        type(cfg).use_return_dict = property(lambda self: True)

    def get_extended_attention_mask(self, attention_mask, input_shape, device=None, dtype=None):  # modeling_utils.py (4.42.4)
        dtype = dtype or next(self.parameters()).dtype
        assert attention_mask.dim() == 2 and not self.config.is_decoder
        ext = attention_mask[:, None, None, :].to(dtype=dtype)
        return (1.0 - ext) * torch.finfo(dtype).min

    def get_head_mask(self, head_mask, num_hidden_layers, is_attention_chunked=False):
        assert head_mask is None
        return [None] * num_hidden_layers

    bert.get_extended_attention_mask = _types.MethodType(get_extended_attention_mask, bert)
    bert.get_head_mask = _types.MethodType(get_head_mask, bert)

    def self_attention_forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                               encoder_attention_mask=None, past_key_value=None, output_attentions=False):
        H = self.num_attention_heads
        d = self.attention_head_size

        def split(x):
            return x.view(x.shape[0], x.shape[1], H, d).permute(0, 2, 1, 3)

        q = split(self.query(hidden_states))
        if encoder_hidden_states is not None:                       # cross-attention
            k, v = split(self.key(encoder_hidden_states)), split(self.value(encoder_hidden_states))
            attention_mask = encoder_attention_mask
        else:
            k, v = split(self.key(hidden_states)), split(self.value(hidden_states))
        scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d)
        if attention_mask is not None:
            scores = scores + attention_mask
        probs = self.dropout(torch.nn.functional.softmax(scores, dim=-1))
        if head_mask is not None:
            probs = probs * head_mask
        ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous()
        ctx = ctx.view(ctx.shape[0], ctx.shape[1], H * d)
        return (ctx, probs) if output_attentions else (ctx,)

    def attention_forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                          encoder_attention_mask=None, past_key_value=None, output_attentions=False):
        so = self.self(hidden_states, attention_mask, head_mask, encoder_hidden_states, encoder_attention_mask, past_key_value,
                       output_attentions)
        return (self.output(so[0], hidden_states),) + so[1:]

    def layer_forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                      encoder_attention_mask=None, past_key_value=None, output_attentions=False):
        a = self.attention(hidden_states, attention_mask, head_mask, output_attentions=output_attentions)[0]
        return (self.output(self.intermediate(a), a),)

    def encoder_forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                        encoder_attention_mask=None, past_key_values=None, use_cache=None, output_attentions=False,
                        output_hidden_states=False, return_dict=True):
        for i, layer in enumerate(self.layer):
            hidden_states = layer(hidden_states, attention_mask, head_mask[i] if head_mask is not None else None)[0]
        return BaseModelOutputWithPastAndCrossAttentions(last_hidden_state=hidden_states, past_key_values=None, hidden_states=None,
                                                         attentions=None, cross_attentions=None)

    for m in bert.modules():
        if isinstance(m, BertSelfAttention):
            if not hasattr(m, "dropout") or not isinstance(m.dropout, torch.nn.Module):
                m.dropout = torch.nn.Dropout(cfg.attention_probs_dropout_prob)
            m.forward = _types.MethodType(self_attention_forward, m)
        elif isinstance(m, BertAttention):
            m.forward = _types.MethodType(attention_forward, m)
    for layer in bert.encoder.layer:
        layer.forward = _types.MethodType(layer_forward, layer)
    bert.encoder.forward = _types.MethodType(encoder_forward, bert.encoder)
    return ref_model
