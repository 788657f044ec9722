"""Drop-in replacement for the reference `module.model_ecamp` (ECAMP/Pre-training/module/model_ecamp.py).

`ecamp(**kwargs)` returns an nn.Module with the reference's parameter names / shapes (350 state_dict
entries, 349 distinct parameters: SURVEY.md §8b) and the reference's `forward(batch, mask_ratio=0.75)
-> (mim_loss, res_loss, mlm_loss)` contract (model_ecamp.py:303-325), plus the positional form
`forward(imgs, input_ids, attention_mask, labels, mask_ratio, *, type_ids, weights, big_imgs, column,
row, noise)`.  The sub-modules below are PARAMETER CONTAINERS only: no torch op of theirs ever runs.
Every FLOP of forward and backward is executed by libecamp_b200.so (hand-written sm_100a kernels)
through the C ABI in include/ecamp_b200.h; there is no CPU / PyTorch fallback — without the library
or without a CUDA device `forward` raises.

Gradient contract.  Autograd path (`loss.backward()`, what the reference trainer and DistributedDataParallel use):
the native backward runs as a chain of per-stage autograd nodes; each stage returns the slices of ONE flat fp32
gradient buffer it has just finished as the gradients of its parameters, so `.grad` are ordinary tensors filled by
AccumulateGrad stage by stage (gradient accumulation, GradScaler.unscale_, clip_grad_norm_, torch.optim.AdamW, DDP
hooks all work unchanged).  Fused path (`forward_backward`, used by ecamp_b200.parallel / FusedAdamW): `.grad` are
views of the flat buffer; a backward that finds the views still attached accumulates, otherwise it overwrites.
bert.pooler.* never receives a gradient (its output is dead code in the reference: bert_modeling.py:144).
"""
import ctypes
import os
import math
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L

__all__ = ["ECAMP", "ecamp"]

VOCAB, HID, MAXPOS = 30000, 768, 256


# --------------------------------------------------------------------------------------------------
# util/pos_embed.py:20-67 (init-time only, host numpy as in the reference)
# --------------------------------------------------------------------------------------------------
def _sincos_1d(embed_dim, pos):
    omega = np.arange(embed_dim // 2, dtype=np.float64)
    omega = omega / embed_dim / 2.  # sic, pos_embed.py:57
    omega = 1. / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def get_2d_sincos_pos_embed(embed_dim, grid_size, cls_token=False):
    grid_h = np.arange(grid_size, dtype=np.float32)
    grid_w = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(grid_w, grid_h), axis=0).reshape([2, 1, grid_size, grid_size])
    emb = np.concatenate([_sincos_1d(embed_dim // 2, grid[0]), _sincos_1d(embed_dim // 2, grid[1])], axis=1)
    if cls_token:
        emb = np.concatenate([np.zeros([1, embed_dim]), emb], axis=0)
    return emb


# --------------------------------------------------------------------------------------------------
# parameter containers mirroring timm 0.4.12 / transformers 4.42.4 attribute names
# --------------------------------------------------------------------------------------------------
class _Box(nn.Module):
    """A module that only holds children / parameters."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("ecamp_b200 sub-modules are parameter containers; call the ECAMP module itself")


def _vit_block(dim, hidden, eps):
    b = _Box()
    b.norm1 = nn.LayerNorm(dim, eps=eps)
    b.attn = _Box()
    b.attn.qkv = nn.Linear(dim, dim * 3, bias=True)
    b.attn.proj = nn.Linear(dim, dim)
    b.norm2 = nn.LayerNorm(dim, eps=eps)
    b.mlp = _Box()
    b.mlp.fc1 = nn.Linear(dim, hidden)
    b.mlp.fc2 = nn.Linear(hidden, dim)
    return b


def _bert_self_attention():
    s = _Box()
    s.query, s.key, s.value = nn.Linear(HID, HID), nn.Linear(HID, HID), nn.Linear(HID, HID)
    return s


def _bert_self_output():
    o = _Box()
    o.dense = nn.Linear(HID, HID)
    o.LayerNorm = nn.LayerNorm(HID, eps=1e-12)
    return o


def _bert_attention():
    a = _Box()
    a.self = _bert_self_attention()
    a.output = _bert_self_output()
    return a


def _bert_intermediate(ffn):
    i = _Box()
    i.dense = nn.Linear(HID, ffn)
    return i


def _bert_output(ffn):
    o = _Box()
    o.dense = nn.Linear(ffn, HID)
    o.LayerNorm = nn.LayerNorm(HID, eps=1e-12)
    return o


def _bert_model(layers=6, ffn=1536):
    bert = _Box()
    bert.embeddings = _Box()
    bert.embeddings.word_embeddings = nn.Embedding(VOCAB, HID, padding_idx=0)
    bert.embeddings.position_embeddings = nn.Embedding(MAXPOS, HID)
    bert.embeddings.token_type_embeddings = nn.Embedding(2, HID)
    bert.embeddings.LayerNorm = nn.LayerNorm(HID, eps=1e-12)
    bert.encoder = _Box()
    bert.encoder.layer = nn.ModuleList()
    for _ in range(layers):
        l = _Box()
        l.attention = _bert_attention()
        l.intermediate = _bert_intermediate(ffn)
        l.output = _bert_output(ffn)
        bert.encoder.layer.append(l)
    bert.pooler = _Box()                       # created although unused (bert_modeling.py:11-13)
    bert.pooler.dense = nn.Linear(HID, HID)
    f = _Box()                                 # context_fusion.py:7-19
    f.attention = _bert_attention()
    f.cross_self_attention = _bert_self_attention()
    f.intermediate = _bert_intermediate(ffn)
    f.output = _bert_output(ffn)
    f.gap_mlp = nn.Linear(HID, HID)
    f.out_layer = _bert_self_output()
    bert.context_fusion_layer = f
    return bert


def _bert_masked_lm():
    m = _Box()
    m.bert = _bert_model()
    m.cls = _Box()
    p = _Box()
    p.transform = _Box()
    p.transform.dense = nn.Linear(HID, HID)
    p.transform.LayerNorm = nn.LayerNorm(HID, eps=1e-12)
    p.decoder = nn.Linear(HID, VOCAB, bias=False)
    p.bias = nn.Parameter(torch.zeros(VOCAB))
    p.decoder.bias = p.bias                    # transformers 4.42.4: one Parameter, two state_dict keys
    m.cls.predictions = p
    return m


# --------------------------------------------------------------------------------------------------
class _StageNode(torch.autograd.Function):
    """ONE backward stage of the native schedule as an autograd node (ecamp_backward_stage_range: LM head, BERT layers
    5..0, fusion + embeddings, decoder head, decoder blocks, ..., encoder blocks 11..0, patch embed).  The nodes are
    chained through a dummy token in REVERSE stage order, so autograd runs stage 0 first; every node returns the slices of
    the flat gradient buffer that its stage has just made final as the gradients of its parameters.  AccumulateGrad
    (and with it the hooks of a stock DistributedDataParallel wrapper, main_pretrain.py:249) therefore fires stage by
    stage while later stages still run: gradient accumulation, GradScaler.unscale_, clip_grad_norm_, torch.optim.AdamW
    and DDP's bucketed all-reduce all see ordinary `.grad` tensors (hard part 2)."""

    @staticmethod
    def forward(ctx, model, handle, stage, token, *params):
        ctx.model, ctx.handle, ctx.stage = model, handle, stage
        return torch.empty((), dtype=torch.float32, device=token.device)   # value never read: ordering only

    @staticmethod
    def backward(ctx, gtoken):
        views = ctx.model._backward_stage(ctx.handle, ctx.stage)
        need = ctx.needs_input_grad[4:]
        return (None, None, None, gtoken) + tuple(v if n else None for v, n in zip(views, need))


class _LossNode(torch.autograd.Function):
    """Head of the chain: hands out the three losses; its backward records the upstream gradients (3 floats)."""

    @staticmethod
    def forward(ctx, handle, token):
        ctx.handle = handle
        return handle["losses"].clone()

    @staticmethod
    def backward(ctx, g):
        h = ctx.handle
        if h.get("consumed"):
            raise RuntimeError("ecamp_b200: backward() called twice on the same forward (retain_graph is not supported)")
        h["consumed"] = True
        h["g"] = g.detach().to(torch.float32).contiguous()
        h["acc"] = 0   # the flat buffer is overwritten; accumulation over micro-steps happens in AccumulateGrad
        return None, torch.empty((), dtype=torch.float32, device=g.device)


class ECAMP(nn.Module):
    """Masked Autoencoder with VisionTransformer backbone + multimodal BERT report decoder (model_ecamp.py:49)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12,
                 decoder_embed_dim=768, decoder_depth=4, decoder_num_heads=6, mlp_ratio=4., norm_layer=nn.LayerNorm,
                 norm_pix_loss=False, dropout=0.1):
        super().__init__()
        if (img_size, patch_size, in_chans, embed_dim, depth, num_heads, decoder_embed_dim, decoder_depth,
                decoder_num_heads, int(mlp_ratio)) != (224, 16, 3, 768, 12, 12, 512, 4, 16, 4):
            raise ValueError("ecamp_b200 implements the configuration of the reference factory ecamp() only "
                             "(model_ecamp.py:328-333): ViT-B/16 encoder, 4x512 decoder with 16 heads")
        eps = 1e-6
        self.patch_embed = _Box()
        self.patch_embed.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.patch_embed.num_patches = (img_size // patch_size) ** 2
        self.patch_embed.patch_size = (patch_size, patch_size)
        n = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim), requires_grad=False)
        self.blocks = nn.ModuleList([_vit_block(embed_dim, int(embed_dim * mlp_ratio), eps) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=eps)
        self.decoder_embed = nn.Linear(embed_dim, decoder_embed_dim, bias=True)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, decoder_embed_dim))
        self.decoder_pos_embed = nn.Parameter(torch.zeros(1, n + 1, decoder_embed_dim), requires_grad=False)
        self.decoder_blocks = nn.ModuleList(
            [_vit_block(decoder_embed_dim, int(decoder_embed_dim * mlp_ratio), eps) for _ in range(decoder_depth)])
        self.decoder_norm = nn.LayerNorm(decoder_embed_dim, eps=eps)
        self.decoder_pred = nn.Linear(decoder_embed_dim, patch_size ** 2 * in_chans, bias=True)
        self.super_res = _Box()
        self.super_res.conv1 = nn.Conv2d(3, 3, 3, 1, 1)
        self.super_res.conv2 = nn.Conv2d(3, 3, 3, 1, 1)
        self.bert_encoder = _Box()
        self.bert_encoder.model = _bert_masked_lm()
        self.bert_mlp = nn.Linear(embed_dim, 768, bias=True)
        self.norm_pix_loss = norm_pix_loss  # accepted and ignored, like the reference
        self.dropout = float(dropout)       # bert_config.py:71-72 (hidden and attention-prob dropout)
        self.initialize_weights()
        # runtime state (not part of the state_dict)
        self.image_mean, self.image_std = 0.4721, 0.3037   # transforms.Normalize of pretrain_datasets.py:52 (uint8 inputs)
        self._rt = None
        self._precision = 0                 # 0 = production (bf16 operands), 1 = fp32-accurate parity mode (set_precision)
        self._dropout_step = 0
        self.ce_rows = int(os.environ.get("ECAMP_CE_ROWS", "4096"))   # rows per vocabulary-head chunk (tuning knob)

    # ---- model_ecamp.py:105-137 -------------------------------------------------------------------
    def initialize_weights(self):
        g = int(self.patch_embed.num_patches ** .5)
        self.pos_embed.data.copy_(torch.from_numpy(get_2d_sincos_pos_embed(self.pos_embed.shape[-1], g, True)).float().unsqueeze(0))
        self.decoder_pos_embed.data.copy_(
            torch.from_numpy(get_2d_sincos_pos_embed(self.decoder_pos_embed.shape[-1], g, True)).float().unsqueeze(0))
        w = self.patch_embed.proj.weight.data
        torch.nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        torch.nn.init.normal_(self.cls_token, std=.02)
        torch.nn.init.normal_(self.mask_token, std=.02)
        for m in self.modules():
            if isinstance(m, nn.Embedding):  # HF _init_weights: normal(0, initializer_range), zero pad row
                nn.init.normal_(m.weight, std=0.02)
                if m.padding_idx is not None:
                    m.weight.data[m.padding_idx].zero_()
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            torch.nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # ---- host-side helpers of the reference that stay host-side -------------------------------------
    def patchify(self, imgs):  # model_ecamp.py:140-151 (p = 2 * patch, as written there)
        p = self.patch_embed.patch_size[0] * 2
        assert imgs.shape[2] == imgs.shape[3] and imgs.shape[2] % p == 0
        h = w = imgs.shape[2] // p
        x = imgs.reshape(shape=(imgs.shape[0], 3, h, p, w, p))
        x = torch.einsum('nchpwq->nhwpqc', x)
        return x.reshape(shape=(imgs.shape[0], h * w, p ** 2 * 3))

    def unpatchify(self, x):  # model_ecamp.py:153-165
        p = self.patch_embed.patch_size[0]
        h = w = int(x.shape[1] ** .5)
        assert h * w == x.shape[1]
        x = x.reshape(shape=(x.shape[0], h, w, p, p, 3))
        x = torch.einsum('nhwpqc->nchpwq', x)
        return x.reshape(shape=(x.shape[0], 3, h * p, h * p))

    # ---- runtime plumbing ------------------------------------------------------------------------------
    def _runtime(self, device):
        """Create / re-validate the native context: flat gradient + shadow buffers, parameter binding."""
        lib = L.lib()
        # one process per GPU: the native library launches on the CURRENT device's streams (and keeps its side stream there)
        if device.type == "cuda" and device.index is not None and device.index != torch.cuda.current_device():
            raise RuntimeError(f"ecamp_b200: the module lives on cuda:{device.index} but the current device is "
                               f"cuda:{torch.cuda.current_device()}; call torch.cuda.set_device({device.index}) first")
        rt = self._rt
        named = dict(self.named_parameters())
        if rt is None:
            rt = dict(ctx=ctypes.c_void_p(), names=[], ws=None, shape=None, versions=None, ptrs=None, precision=0)
            L.check(lib.ecamp_ctx_create(ctypes.byref(rt["ctx"])), "ecamp_ctx_create")
            n = lib.ecamp_param_count()
            for i in range(n):
                rt["names"].append(lib.ecamp_param_name(i).decode())
            rt["numel"] = [lib.ecamp_param_numel(i) for i in range(n)]
            rt["goff"] = [lib.ecamp_param_grad_offset(i) for i in range(n)]
            missing = [k for k in rt["names"] if k not in named]
            if missing:
                raise RuntimeError(f"ecamp_b200: module is missing parameters {missing[:4]}...")
            for k, nel in zip(rt["names"], rt["numel"]):
                if named[k].numel() != nel:
                    raise RuntimeError(f"ecamp_b200: parameter {k} has {named[k].numel()} elements, expected {nel}")
            ranges = []
            lo, hi = ctypes.c_int64(), ctypes.c_int64()
            for st in range(lib.ecamp_backward_stage_count()):
                L.check(lib.ecamp_backward_stage_range(st, ctypes.byref(lo), ctypes.byref(hi)), "ecamp_backward_stage_range")
                ranges.append((lo.value, hi.value))
            rt["stage_params"] = [[i for i, o in enumerate(rt["goff"]) if a <= o < b] for a, b in ranges]
            assert sorted(i for sp in rt["stage_params"] for i in sp) == list(range(n))
            rt["gen"] = 0
            self._rt = rt
        params = [named[k] for k in rt["names"]]
        rt["params"] = params
        for p in params + [self.pos_embed, self.decoder_pos_embed]:
            if p.device.type != "cuda" or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("ecamp_b200: parameters must be contiguous fp32 CUDA tensors (call model.cuda())")
        ptrs = tuple(p.data_ptr() for p in params) + (self.pos_embed.data_ptr(), self.decoder_pos_embed.data_ptr())
        if rt["precision"] != self._precision:   # buffers are laid out per precision: re-bind, re-plan the workspace
            L.check(lib.ecamp_ctx_set_precision(rt["ctx"], ctypes.c_int32(self._precision)), "ecamp_ctx_set_precision")
            rt["precision"], rt["ptrs"], rt["shape"], rt["ws"] = self._precision, None, None, None
        if rt["ptrs"] != ptrs:
            if rt.get("G") is None or rt["G"].device != device:
                gf = lib.ecamp_grad_floats()
                rt["G"] = torch.zeros(gf, dtype=torch.float32, device=device)
                rt["M1"] = torch.zeros(gf, dtype=torch.float32, device=device)
                rt["M2"] = torch.zeros(gf, dtype=torch.float32, device=device)
                rt["SH"] = torch.zeros(lib.ecamp_shadow_bytes(), dtype=torch.uint8, device=device)
                rt["AT"] = torch.zeros(lib.ecamp_adam_table_bytes(), dtype=torch.uint8, device=device)
                rt["AC"] = torch.zeros(lib.ecamp_adam_chunk_bytes(), dtype=torch.uint8, device=device)
                rt["grad_views"] = [rt["G"][o:o + nel].view(p.shape) for o, nel, p in zip(rt["goff"], rt["numel"], params)]
                rt["ws"], rt["shape"] = None, None
            arr = (ctypes.c_void_p * len(params))(*[p.data_ptr() for p in params])
            L.check(lib.ecamp_ctx_bind(rt["ctx"], arr, len(params), L.ptr(rt["G"]), L.ptr(rt["M1"]), L.ptr(rt["M2"]),
                                       L.ptr(rt["SH"]), L.ptr(self.pos_embed), L.ptr(self.decoder_pos_embed),
                                       L.ptr(rt["AT"]), L.ptr(rt["AC"])), "ecamp_ctx_bind")
            rt["ptrs"] = ptrs
            rt["versions"] = None
        versions = tuple(p._version for p in params)
        if rt["versions"] != versions:  # parameters were modified outside the fused optimizer: refresh GEMM copies
            L.check(lib.ecamp_refresh_shadows(rt["ctx"], L.cur_stream()), "ecamp_refresh_shadows")
            rt["versions"] = versions
        return rt

    def _workspace(self, rt, B, T, keep, has_big, device):
        shape = (B, T, keep, int(has_big), int(self.ce_rows))
        if rt["shape"] != shape:
            lib = L.lib()
            s = L.Shape(*shape)
            need = lib.ecamp_ctx_workspace_bytes(rt["ctx"], ctypes.byref(s))
            if rt["ws"] is None or rt["ws"].numel() < need:
                rt["ws"] = None
                rt["ws"] = torch.empty(need, dtype=torch.uint8, device=device)
            L.check(lib.ecamp_ctx_set_workspace(rt["ctx"], L.ptr(rt["ws"]), ctypes.c_int64(rt["ws"].numel()),
                                                ctypes.byref(s)), "ecamp_ctx_set_workspace")
            rt["shape"] = shape
        return rt

    @staticmethod
    def _dev(t, dtype, device):
        if t.device != device or t.dtype != dtype or not t.is_contiguous():
            t = t.to(device=device, dtype=dtype, non_blocking=True).contiguous()
        return t

    def _prepare(self, image, ids, labels, attention_mask, type_ids, weights, column, row, noise, mask_ratio, has_big):
        device = self.cls_token.device
        if device.type != "cuda":
            raise RuntimeError("ecamp_b200 runs on CUDA (sm_100a) only: move the module to the GPU; there is no CPU path")
        B, T = ids.shape
        keep = int(196 * (1 - mask_ratio))  # model_ecamp.py:175, python double arithmetic
        side = 448 if has_big else 224
        f32, i64 = torch.float32, torch.int64
        if image.dtype == torch.uint8:
            # the loader's 8-bit grayscale crop, [B, side, side] or [B, 1, side, side]: Grayscale(3) + ToTensor + Normalize
            # (pretrain_datasets.py:49-52) run on the GPU, bit-exact with the CPU transform - 1 byte per pixel over PCIe
            if tuple(image.shape) not in ((B, side, side), (B, 1, side, side)):
                raise ValueError(f"ecamp_b200: expected a uint8 image of shape {(B, side, side)} or {(B, 1, side, side)}, "
                                 f"got {tuple(image.shape)}")
            gray = self._dev(image, torch.uint8, device)
            img = torch.empty(B, 3, side, side, dtype=f32, device=device)
            L.check(L.lib().ecamp_image_u8_normalize(L.ptr(gray), ctypes.c_int64(B), ctypes.c_int64(side * side),
                                                     ctypes.c_float(self.image_mean), ctypes.c_float(self.image_std),
                                                     L.ptr(img), L.cur_stream()), "ecamp_image_u8_normalize")
            image = img
        if tuple(image.shape) != (B, 3, side, side):
            raise ValueError(f"ecamp_b200: expected image of shape {(B, 3, side, side)}, got {tuple(image.shape)}")
        if not (0 < keep <= 196) or T > MAXPOS:
            raise ValueError(f"ecamp_b200: unsupported mask_ratio {mask_ratio} / sequence length {T}")
        t = dict(image=self._dev(image, f32, device), ids=self._dev(ids, i64, device), labels=self._dev(labels, i64, device),
                 attention_mask=self._dev(attention_mask, i64, device))
        t["type_ids"] = self._dev(type_ids, i64, device) if type_ids is not None else torch.zeros_like(t["ids"])
        t["weights"] = (self._dev(weights, f32, device) if weights is not None
                        else torch.ones(B, T, dtype=f32, device=device))
        if has_big:
            t["column"] = self._dev(torch.as_tensor(column), i64, device)
            t["row"] = self._dev(torch.as_tensor(row), i64, device)
        t["noise"] = (self._dev(noise, f32, device) if noise is not None
                      else torch.rand(B, 196, device=device))  # model_ecamp.py:177
        return t, B, T, keep

    def _launch_forward(self, t, B, T, keep, has_big, defer_mlm):
        device = self.cls_token.device
        lib = L.lib()
        rt = self._workspace(self._runtime(device), B, T, keep, has_big, device)
        losses = torch.zeros(3, dtype=torch.float32, device=device)
        mask = torch.empty(B, 196, dtype=torch.float32, device=device)
        ids_restore = torch.empty(B, 196, dtype=torch.int64, device=device)
        ids_keep = torch.empty(B, keep, dtype=torch.int64, device=device)
        bt = L.Batch()
        bt.image, bt.ids, bt.labels = t["image"].data_ptr(), t["ids"].data_ptr(), t["labels"].data_ptr()
        bt.attention_mask, bt.type_ids, bt.weights = (t["attention_mask"].data_ptr(), t["type_ids"].data_ptr(),
                                                      t["weights"].data_ptr())
        bt.column = t["column"].data_ptr() if has_big else None
        bt.row = t["row"].data_ptr() if has_big else None
        bt.noise = t["noise"].data_ptr()
        train = self.training and torch.is_grad_enabled()
        flags = (1 if self.training else 0) | (2 if defer_mlm else 0)
        self._dropout_step += 1
        seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + self._dropout_step) & 0xFFFFFFFFFFFFFFFF
        L.check(lib.ecamp_forward(rt["ctx"], ctypes.byref(bt), ctypes.c_int32(flags),
                                  ctypes.c_float(self.dropout if self.training else 0.0), ctypes.c_uint64(seed),
                                  L.ptr(losses), L.ptr(mask), L.ptr(ids_restore), L.ptr(ids_keep), L.cur_stream()),
                "ecamp_forward")
        rt["gen"] += 1   # the activations of this forward now own the (single) workspace
        handle = dict(losses=losses, tensors=t, rt=rt, shape=rt["shape"], train=train, gen=rt["gen"])
        self.last = dict(mask=mask, ids_restore=ids_restore, ids_keep=ids_keep)
        return handle

    @staticmethod
    def _check_handle(handle):
        rt = handle["rt"]
        if rt["shape"] != handle["shape"] or rt["gen"] != handle["gen"]:
            raise RuntimeError("ecamp_b200: another forward() of this module ran between this forward() and its backward(); "
                               "the saved activations live in one shared workspace and have been overwritten")

    def _backward_stage(self, handle, stage):
        """Autograd path: run native backward stage `stage` into the flat buffer (overwriting) and return the gradient
        views of the parameters that the stage finishes."""
        lib = L.lib()
        rt = handle["rt"]
        self._check_handle(handle)
        L.check(lib.ecamp_backward(rt["ctx"], L.ptr(handle["g"]), ctypes.c_int32(0), ctypes.c_int32(stage), L.cur_stream()),
                "ecamp_backward")
        return [rt["grad_views"][i] for i in rt["stage_params"][stage]]

    def _sync_flat_grads(self, rt):
        """Make the flat gradient buffer agree with the `.grad` tensors (autograd / DDP / GradScaler leave ordinary
        tensors there, the fused path leaves views of the buffer): called by the fused optimizer before it reads it."""
        src, dst, zero = [], [], []
        for p, v in zip(rt["params"], rt["grad_views"]):
            if p.grad is None:
                zero.append(v)
            elif p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad); dst.append(v)
        if dst:
            torch._foreach_copy_(dst, src)
        if zero and len(zero) != len(rt["params"]):
            torch._foreach_zero_(zero)
        return len(zero) != len(rt["params"])

    def _autograd_losses(self, handle):
        rt = handle["rt"]
        token = torch.empty((), dtype=torch.float32, device=handle["losses"].device)
        for st in reversed(range(len(rt["stage_params"]))):
            token = _StageNode.apply(self, handle, st, token, *[rt["params"][i] for i in rt["stage_params"][st]])
        return _LossNode.apply(handle, token)

    def _backward(self, handle, g, stage=-1, stage_end=None):
        """Run the native backward for upstream gradients g (3 floats); attaches p.grad views.  stage = -1: all stages;
        stage_end given: stages [stage, stage_end) in one native call."""
        lib = L.lib()
        rt = handle["rt"]
        self._check_handle(handle)
        params, views = rt["params"], rt["grad_views"]
        if stage <= 0:
            attached = [p.grad is v for p, v in zip(params, views)]
            if all(attached):
                handle["acc"] = 1
            else:
                for p, v, a in zip(params, views, attached):
                    if not a:
                        if p.grad is not None:
                            raise RuntimeError("ecamp_b200: a parameter has a foreign .grad tensor; call zero_grad() first")
                        if any(attached):
                            v.zero_()
                handle["acc"] = 1 if any(attached) else 0
            handle["g"] = g.detach().to(torch.float32).contiguous()
        if stage_end is not None and stage >= 0:
            L.check(lib.ecamp_backward_stages(rt["ctx"], L.ptr(handle["g"]), ctypes.c_int32(handle["acc"]), ctypes.c_int32(stage),
                                              ctypes.c_int32(stage_end), L.cur_stream()), "ecamp_backward_stages")
        else:
            L.check(lib.ecamp_backward(rt["ctx"], L.ptr(handle["g"]), ctypes.c_int32(handle["acc"]), ctypes.c_int32(stage),
                                       L.cur_stream()), "ecamp_backward")
        last = stage_end - 1 if (stage_end is not None and stage >= 0) else stage
        if stage < 0 or last == lib.ecamp_backward_stage_count() - 1:
            for p, v in zip(params, views):
                p.grad = v

    # ---- public API -----------------------------------------------------------------------------------------
    def forward(self, batch, input_ids=None, attention_mask=None, labels=None, mask_ratio=0.75, *, type_ids=None,
                weights=None, big_imgs=None, column=None, row=None, noise=None):
        if isinstance(batch, dict):  # reference form: forward(batch, mask_ratio=0.75)   (model_ecamp.py:303)
            if input_ids is not None:
                mask_ratio = input_ids
            b = batch
            image, has_big = b["image"], True
            input_ids, labels, attention_mask = b["ids"], b["labels"], b["attention_mask"]
            type_ids, weights, column, row = b["type_ids"], b["weights"], b["column"], b["row"]
            noise = b.get("noise", noise)
        else:
            has_big = big_imgs is not None
            image = big_imgs if has_big else batch
        t, B, T, keep = self._prepare(image, input_ids, labels, attention_mask, type_ids, weights, column, row, noise,
                                      mask_ratio, has_big)
        handle = self._launch_forward(t, B, T, keep, has_big, defer_mlm=False)
        if torch.is_grad_enabled() and any(p.requires_grad for p in handle["rt"]["params"]):
            losses = self._autograd_losses(handle)
        else:
            losses = handle["losses"]
        return losses[0], losses[1], losses[2]

    def forward_backward(self, batch, loss_weights=(1.0, 1.0, 1.0), mask_ratio=0.75, stage_callback=None, callback_stages=None):
        """Fused training step without autograd: forward with the vocabulary head deferred, then the staged
        backward (the head is evaluated once, fused with its gradient).  Returns the 3 losses (device tensor).
        `stage_callback(stage, lo, hi)` is invoked after each backward stage with the range of the flat gradient
        buffer that became final (used by ecamp_b200.parallel for bucketed all-reduce overlap).  `callback_stages` (optional,
        increasing stage numbers, the last stage included): the callback is only needed after these stages - the stages in
        between run in one native call each group, which lets the library overlap across their boundaries; the callback then
        gets the union of the group's slices."""
        b = batch
        t, B, T, keep = self._prepare(b["image"], b["ids"], b["labels"], b["attention_mask"], b.get("type_ids"),
                                      b.get("weights"), b.get("column"), b.get("row"), b.get("noise"), mask_ratio,
                                      b["image"].shape[-1] == 448)
        has_big = b["image"].shape[-1] == 448
        handle = self._launch_forward(t, B, T, keep, has_big, defer_mlm=True)
        g = torch.tensor(loss_weights, dtype=torch.float32, device=handle["losses"].device) \
            if not torch.is_tensor(loss_weights) else loss_weights
        lib = L.lib()
        if stage_callback is None:
            self._backward(handle, g, -1)
        elif callback_stages is None:
            lo, hi = ctypes.c_int64(), ctypes.c_int64()
            for s in range(lib.ecamp_backward_stage_count()):
                self._backward(handle, g, s)
                lib.ecamp_backward_stage_range(s, ctypes.byref(lo), ctypes.byref(hi))
                stage_callback(s, lo.value, hi.value)
        else:
            n = lib.ecamp_backward_stage_count()
            ends = sorted(set(int(s) for s in callback_stages))
            if not ends or ends[-1] != n - 1 or ends[0] < 0:
                raise ValueError("callback_stages must be stage numbers in [0, n) and include the last stage")
            lo, hi = ctypes.c_int64(), ctypes.c_int64()
            first = 0
            for e in ends:
                self._backward(handle, g, first, e + 1)
                glo, ghi = None, None
                for s in range(first, e + 1):   # slices of consecutive stages are adjacent, walking down the buffer
                    lib.ecamp_backward_stage_range(s, ctypes.byref(lo), ctypes.byref(hi))
                    glo = lo.value if glo is None else min(glo, lo.value)
                    ghi = hi.value if ghi is None else max(ghi, hi.value)
                stage_callback(e, glo, ghi)
                first = e + 1
        return handle["losses"]

    _OLD_FUSION_KEY = "bert_encoder.model.bert.cross_attn_layer"
    _FUSION_KEY = "bert_encoder.model.bert.context_fusion_layer"

    def load_state_dict(self, state_dict, strict=True, assign=False):
        """Accepts the released checkpoints' spelling `cross_attn_layer` for `context_fusion_layer`, the rename the
        reference applies by hand at Visualization/main_visualization.py:88-93; otherwise nn.Module.load_state_dict."""
        if any(k.startswith(self._OLD_FUSION_KEY) for k in state_dict):
            state_dict = type(state_dict)((k.replace(self._OLD_FUSION_KEY, self._FUSION_KEY, 1)
                                           if k.startswith(self._OLD_FUSION_KEY) else k, v)
                                          for k, v in state_dict.items())
        return super().load_state_dict(state_dict, strict=strict, assign=assign)

    @torch.no_grad()
    def cross_attention_map(self, imgs, text_ids, attention_mask, type_ids=None, mask_ratio=0.0, noise=None,
                            restore_order=False):
        """Probabilities of the fusion layer's text -> image cross-attention, `[B, 6, T, keep]` fp32: what the reference's
        heat-map model returns (Visualization/module/model_ecamp.py:308-319 -> context_fusion.py:45-57,
        `cross_self_outputs[1]`; used at main_visualization.py:153-154 as `attention[0, :, 4]`).  `imgs` is 224 px.
        As in the reference the image tokens - hence the columns - are in `ids_keep` order (random_masking shuffles even
        at mask_ratio = 0); `restore_order=True` scatters them back to raster order (only at mask_ratio = 0).  Call
        `model.eval()` first, as the tool does: the probabilities are re-derived without dropout."""
        if self.training:
            raise RuntimeError("cross_attention_map: call model.eval() first (main_visualization.py:151)")
        zeros = torch.zeros_like(text_ids)
        t, B, T, keep = self._prepare(imgs, text_ids, zeros, attention_mask, type_ids, None, None, None, noise, mask_ratio, False)
        handle = self._launch_forward(t, B, T, keep, False, defer_mlm=True)   # the vocabulary head is not needed
        probs = torch.empty(B, 6, T, keep, dtype=torch.float32, device=handle["losses"].device)
        L.check(L.lib().ecamp_cross_attention_probs(handle["rt"]["ctx"], L.ptr(probs), L.cur_stream()),
                "ecamp_cross_attention_probs")
        if restore_order:
            if keep != 196:
                raise ValueError("restore_order needs mask_ratio = 0 (all 196 patches kept)")
            idx = self.last["ids_restore"][:, None, None, :].expand(B, 6, T, 196)
            probs = torch.gather(probs, 3, idx)
        return probs

    def set_precision(self, precision):
        """"bf16" (default): production - GEMM / attention operands in bf16 with fp32 accumulation, the reference's
        autocast.  "fp32": the fp32-accurate parity mode - the same native schedule with fp32 activations, every GEMM on
        the same tcgen05 kernel with error-compensated bf16 x 3 split operands, attention in fp32; exists to pin the
        algebra of the step against the fp32 oracle at 1e-5 and is roughly an order of magnitude slower."""
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self._precision = 1 if precision == "fp32" else 0
        return self

    def flat_grads(self):
        return self._rt["G"] if self._rt else None

    def debug_buffer(self, name, shape, dtype):
        """View of a named native intermediate buffer (parity tests)."""
        lib = L.lib()
        lib.ecamp_debug_buffer.restype = ctypes.c_void_p
        p = lib.ecamp_debug_buffer(self._rt["ctx"], name.encode())
        if not p:
            raise KeyError(name)
        n = int(np.prod(shape))
        esz = torch.empty(0, dtype=dtype).element_size()
        off = p - self._rt["ws"].data_ptr()
        return self._rt["ws"][off:off + n * esz].view(dtype).view(shape)

    def __del__(self):
        try:
            if self._rt is not None and self._rt.get("ctx"):
                L.lib().ecamp_ctx_destroy(self._rt["ctx"])
        except Exception:
            pass


class ECAMPVis(ECAMP):
    """Drop-in for the heat-map tool's model (Visualization/module/model_ecamp.py): same parameters and state_dict,
    `forward(imgs, text_ids, attention_mask, type_ids, mask_ratio=0)` returns the cross-attention probabilities."""

    def forward(self, imgs, text_ids, attention_mask, type_ids, mask_ratio=0):
        return self.cross_attention_map(imgs, text_ids, attention_mask, type_ids, mask_ratio=mask_ratio)


def ecamp_vis(**kwargs):
    """`Visualization/module/model_ecamp.py:322-327` (`ecamp(**kwargs)` there)."""
    return ECAMPVis(patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12, decoder_embed_dim=512,
                    decoder_depth=4, decoder_num_heads=16, mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def ecamp(**kwargs):
    """The reference factory (model_ecamp.py:328-333)."""
    return ECAMP(patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12, decoder_embed_dim=512,
                 decoder_depth=4, decoder_num_heads=16, mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
