"""Drop-in for the reference fine-tune classifier (ECAMP/Fine-tuning/Classification/models_vit.py:60-128):
`vit_base_patch16(num_classes=14, drop_path_rate=0.1, global_pool=True)` -> nn.Module with timm's parameter names
(cls_token, pos_embed, patch_embed.proj.*, blocks.N.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}.*, fc_norm.*,
head.*), so `load_state_dict(checkpoint['model'], strict=False)` of a pre-training checkpoint works as at train.py:131-143.

`forward(x) -> logits [B, num_classes]`; every FLOP of forward and backward runs in libecamp_b200.so (the same encoder
kernels as the pre-training step, on the full 197-token sequence, plus DropPath, mean-pool, fc_norm and the head)
through the C ABI (`ecamp_cls_*` in include/ecamp_b200.h).  There is no CPU / PyTorch fallback.  The loss
(BCEWithLogitsLoss over [B, 14], train.py:443) stays in PyTorch: backward receives d(logits).
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib as L
from .model_ecamp import _Box, _vit_block

__all__ = ["VisionTransformer", "vit_base_patch16"]

PAD = 16  # the head is padded to 16 outputs (16-byte GEMM operand pitch)


class _ClsStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, handle, *params):
        ctx.model, ctx.handle = model, handle
        return handle["logits"].clone()

    @staticmethod
    def backward(ctx, g):
        # ordinary autograd semantics: the native backward overwrites the flat buffer, its views are returned as the
        # parameter gradients and AccumulateGrad (hence DDP hooks, gradient accumulation, clip_grad_norm_) does the rest
        views = ctx.model._backward(ctx.handle, g)
        return (None, None) + tuple(v if n else None for v, n in zip(views, ctx.needs_input_grad[2:]))


class VisionTransformer(nn.Module):
    def __init__(self, num_classes=14, drop_path_rate=0.1, global_pool=True, img_size=224, patch_size=16, in_chans=3,
                 embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, **_):
        super().__init__()
        if (img_size, patch_size, in_chans, embed_dim, depth, num_heads, int(mlp_ratio), bool(global_pool)) != \
                (224, 16, 3, 768, 12, 12, 4, True):
            raise ValueError("ecamp_b200.models_vit implements vit_base_patch16(global_pool=True) at 224 px only")
        if not 0 < num_classes <= PAD:
            raise ValueError(f"ecamp_b200.models_vit: num_classes must be in 1..{PAD}")
        self.num_classes, self.drop_path_rate, self.depth = num_classes, float(drop_path_rate), depth
        self.patch_embed = _Box()
        self.patch_embed.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, 197, embed_dim))
        self.blocks = nn.ModuleList([_vit_block(embed_dim, embed_dim * mlp_ratio, 1e-6) for _ in range(depth)])
        self.fc_norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.head = nn.Linear(embed_dim, num_classes)
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.trunc_normal_(self.cls_token, std=.02)
        for m in self.modules():                         # timm _init_vit_weights
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        self._rt = None

    # ---- native runtime ---------------------------------------------------------------------------------------
    def _runtime(self, device):
        lib = L.lib()
        rt = self._rt
        named = dict(self.named_parameters())
        if rt is None:
            rt = dict(ctx=ctypes.c_void_p(), ws=None, B=None, ptrs=None, versions=None, gen=0)
            L.check(lib.ecamp_ctx_create(ctypes.byref(rt["ctx"])), "ecamp_ctx_create")
            n = lib.ecamp_param_count()
            rt["names"] = [lib.ecamp_param_name(i).decode() for i in range(n)]
            rt["numel"] = [lib.ecamp_param_numel(i) for i in range(n)]
            rt["goff"] = [lib.ecamp_param_grad_offset(i) for i in range(n)]
            rt["mine"] = [k for k in rt["names"] if k in named]   # patch_embed.proj.*, cls_token, blocks.*
            self._rt = rt
        for p in named.values():
            if p.device.type != "cuda" or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("ecamp_b200: parameters must be contiguous fp32 CUDA tensors (call model.cuda())")
        ptrs = tuple(p.data_ptr() for p in named.values())
        if rt["ptrs"] != ptrs:
            gf = lib.ecamp_grad_floats()
            rt["G"] = torch.zeros(gf, dtype=torch.float32, device=device)
            rt["SH"] = torch.zeros(lib.ecamp_shadow_bytes(), dtype=torch.uint8, device=device)
            rt["AT"] = torch.zeros(lib.ecamp_adam_table_bytes(), dtype=torch.uint8, device=device)
            rt["AC"] = torch.zeros(lib.ecamp_adam_chunk_bytes(), dtype=torch.uint8, device=device)
            rt["dummy"] = torch.zeros(max(rt["numel"]), dtype=torch.float32, device=device)  # table entries this model lacks
            arr = (ctypes.c_void_p * len(rt["names"]))(*[(named[k].data_ptr() if k in named else rt["dummy"].data_ptr())
                                                         for k in rt["names"]])
            L.check(lib.ecamp_ctx_bind(rt["ctx"], arr, len(rt["names"]), L.ptr(rt["G"]), None, None, L.ptr(rt["SH"]),
                                       L.ptr(rt["dummy"]), L.ptr(rt["dummy"]), L.ptr(rt["AT"]), L.ptr(rt["AC"])), "ecamp_ctx_bind")
            rt["views"] = {k: rt["G"][o:o + nel].view(named[k].shape) for k, o, nel in zip(rt["names"], rt["goff"], rt["numel"])
                           if k in named}
            ex = dict(pos_embed=torch.zeros(197 * 768, device=device), fc_norm_w=torch.zeros(768, device=device),
                      fc_norm_b=torch.zeros(768, device=device), head_w=torch.zeros(PAD, 768, device=device),
                      head_b=torch.zeros(PAD, device=device))
            rt["extra"] = ex
            rt["views"].update({"pos_embed": ex["pos_embed"].view(1, 197, 768), "fc_norm.weight": ex["fc_norm_w"],
                                "fc_norm.bias": ex["fc_norm_b"], "head.weight": ex["head_w"][:self.num_classes],
                                "head.bias": ex["head_b"][:self.num_classes]})
            rt["ptrs"], rt["versions"] = ptrs, None
        versions = tuple(p._version for p in named.values())
        if rt["versions"] != versions:   # parameters changed (optimizer step, load_state_dict): refresh the bf16 GEMM copies
            L.check(lib.ecamp_refresh_shadows(rt["ctx"], L.cur_stream()), "ecamp_refresh_shadows")
            hw = torch.zeros(PAD, 768, dtype=torch.bfloat16, device=device)
            hw[:self.num_classes] = self.head.weight.detach().to(torch.bfloat16)
            hb = torch.zeros(PAD, dtype=torch.float32, device=device)
            hb[:self.num_classes] = self.head.bias.detach()
            rt["head_w16"], rt["head_b"], rt["versions"] = hw, hb, versions
        return rt

    def draw_drop_path(self, B, device):
        """[depth, 2, B] per-sample branch scales (timm DropPath: floor(keep + U) / keep, rates linspace(0, rate, depth))."""
        out = torch.ones(self.depth, 2, B, device=device)
        for l, p in enumerate(torch.linspace(0, self.drop_path_rate, self.depth).tolist()):
            if p > 0:
                out[l] = torch.floor((1 - p) + torch.rand(2, B, device=device)) / (1 - p)
        return out

    def _io(self, rt, handle):
        io = L.ClsIO()
        ex = rt["extra"]
        io.image, io.pos_embed = handle["x"].data_ptr(), self.pos_embed.data_ptr()
        io.fc_norm_w, io.fc_norm_b = self.fc_norm.weight.data_ptr(), self.fc_norm.bias.data_ptr()
        io.head_w16, io.head_b = rt["head_w16"].data_ptr(), rt["head_b"].data_ptr()
        io.dp_scale = handle["dp"].data_ptr() if handle["dp"] is not None else None
        io.logits = handle["logits"].data_ptr()
        io.d_logits = handle["d_logits"].data_ptr() if handle.get("d_logits") is not None else None
        io.g_pos_embed, io.g_fc_norm_w, io.g_fc_norm_b = ex["pos_embed"].data_ptr(), ex["fc_norm_w"].data_ptr(), ex["fc_norm_b"].data_ptr()
        io.g_head_w, io.g_head_b = ex["head_w"].data_ptr(), ex["head_b"].data_ptr()
        return io

    def forward(self, x, drop_path_scales=None):
        device = self.cls_token.device
        if device.type != "cuda":
            raise RuntimeError("ecamp_b200 runs on CUDA (sm_100a) only: move the module to the GPU; there is no CPU path")
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 224, 224):
            raise ValueError(f"ecamp_b200.models_vit: expected [B, 3, 224, 224], got {tuple(x.shape)}")
        lib = L.lib()
        B = x.shape[0]
        x = x.to(device=device, dtype=torch.float32).contiguous()
        rt = self._runtime(device)
        if rt["B"] != B:
            need = lib.ecamp_cls_workspace_bytes(B)
            if rt["ws"] is None or rt["ws"].numel() < need:
                rt["ws"] = None
                rt["ws"] = torch.empty(need, dtype=torch.uint8, device=device)
            L.check(lib.ecamp_cls_set_workspace(rt["ctx"], L.ptr(rt["ws"]), ctypes.c_int64(rt["ws"].numel()), B), "ecamp_cls_set_workspace")
            rt["B"] = B
        dp = drop_path_scales
        if dp is None and self.training and self.drop_path_rate > 0:
            dp = self.draw_drop_path(B, device)
        if dp is not None:
            dp = dp.to(device=device, dtype=torch.float32).contiguous()
            if tuple(dp.shape) != (self.depth, 2, B):
                raise ValueError("drop_path_scales must be [depth, 2, B]")
        rt["gen"] += 1
        handle = dict(x=x, dp=dp, logits=torch.empty(B, PAD, dtype=torch.float32, device=device), rt=rt, B=B, gen=rt["gen"])
        L.check(lib.ecamp_cls_forward(rt["ctx"], ctypes.byref(self._io(rt, handle)), L.cur_stream()), "ecamp_cls_forward")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            out = _ClsStep.apply(self, handle, *self.parameters())
        else:
            out = handle["logits"]
        return out[:, :self.num_classes]

    def _backward(self, handle, g):
        lib = L.lib()
        rt = handle["rt"]
        if rt["B"] != handle["B"] or rt["gen"] != handle["gen"]:
            raise RuntimeError("ecamp_b200: another forward() of this module ran between this forward() and its backward(); "
                               "the saved activations live in one shared workspace and have been overwritten")
        views = rt["views"]
        handle["d_logits"] = g.detach().to(torch.float32).contiguous()
        L.check(lib.ecamp_cls_backward(rt["ctx"], ctypes.byref(self._io(rt, handle)), ctypes.c_int32(0), L.cur_stream()),
                "ecamp_cls_backward")
        return [views[k] for k, _ in self.named_parameters()]

    def __del__(self):
        try:
            if self._rt is not None and self._rt.get("ctx"):
                L.lib().ecamp_ctx_destroy(self._rt["ctx"])
        except Exception:
            pass


def vit_base_patch16(**kwargs):
    """The reference factory (models_vit.py:122-126)."""
    return VisionTransformer(patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, **kwargs)
