"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.

A plain-PyTorch restatement of the ECAMP pre-training step (forward + the three losses; backward is
autograd) that runs on CPU or GPU with no timm / transformers dependency, so that it travels to the
GPU box.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it, and only as the checker / CPU baseline — never as the thing shipped.  ecamp_b200/ must not
import this file.

Parity pinning: the reference has NO golden vectors, tests or fixtures for this path (SURVEY.md §8c:
"parity unpinned" by the reference's own tests).  This oracle is instead pinned against outputs of the
reference's own sources run in the build container (oracle/make_golden.py imports
/root/reference/ECAMP/Pre-training/module/model_ecamp.py through oracle/_ref_shims.py and the installed
Hugging Face BERT modules) — those outputs are committed under tests/golden/ and checked by
tests/test_oracle.py.

Each function cites the reference file:line it follows (paths relative to
/root/reference/ECAMP/Pre-training/).  Third-party arithmetic not vendored in the reference:
  * timm 0.4.12 (environment.yml:128) — PatchEmbed, Block, Attention, Mlp;
  * transformers 4.42.4 (environment.yml:138) — BertEmbeddings, BertSelfAttention (eager path),
    BertSelfOutput, BertIntermediate, BertOutput, BertLayer, BertPooler, BertLMPredictionHead,
    get_extended_attention_mask; their published algorithms are restated here.
"""
import math
from functools import partial

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# util/pos_embed.py:20-67 — frozen 2-D sin-cos position embedding (float64 numpy, cast to fp32)
# ----------------------------------------------------------------------------------------------


def _sincos_1d(embed_dim, pos):
    omega = np.arange(embed_dim // 2, dtype=np.float64)  # pos_embed.py:56 (np.float == float64)
    omega = omega / embed_dim / 2.  # pos_embed.py:57 as written: (omega / embed_dim) / 2, NOT MAE's omega / (embed_dim / 2)
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_2d(embed_dim, grid_size, cls_token=True):
    gh = np.arange(grid_size, dtype=np.float32)
    gw = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape([2, 1, grid_size, grid_size])  # pos_embed.py:28-31, w first
    emb = np.concatenate([_sincos_1d(embed_dim // 2, grid[0]), _sincos_1d(embed_dim // 2, grid[1])], axis=1)
    if cls_token:
        emb = np.concatenate([np.zeros([1, embed_dim]), emb], axis=0)
    return emb


# ----------------------------------------------------------------------------------------------
# timm 0.4.12 vision_transformer: PatchEmbed / Attention / Mlp / Block (used at model_ecamp.py:60,66-68,80-82)
# ----------------------------------------------------------------------------------------------


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
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = ((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1)
        return self.proj((attn @ v).transpose(1, 2).reshape(B, N, C))


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))  # exact (erf) GELU


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, eps):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=eps)
        self.attn = Attention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


# ----------------------------------------------------------------------------------------------
# module/model_ecamp.py:28-46 — super-resolution head
# ----------------------------------------------------------------------------------------------


class InterpolateConvSuperResolution(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 3, 3, 1, 1)
        self.conv2 = nn.Conv2d(3, 3, 3, 1, 1)

    def forward(self, x):
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
        return F.relu(self.conv2(F.relu(self.conv1(x))) + x)


# ----------------------------------------------------------------------------------------------
# transformers 4.42.4 modeling_bert pieces, with the reference's config (module/bert_config.py:63-80):
# 6 layers, 6 heads x 128, hidden 768, FFN 1536, vocab 30000, max_pos 256, LN eps 1e-12, dropout 0.1
# ----------------------------------------------------------------------------------------------
VOCAB, HID, LAYERS, HEADS, FFN, MAXPOS, TYPES, BERT_EPS = 30000, 768, 6, 6, 1536, 256, 2, 1e-12


class BertEmbeddings(nn.Module):
    def __init__(self, p_drop):
        super().__init__()
        self.word_embeddings = nn.Embedding(VOCAB, HID, padding_idx=0)
        self.position_embeddings = nn.Embedding(MAXPOS, HID)
        self.token_type_embeddings = nn.Embedding(TYPES, HID)
        self.LayerNorm = nn.LayerNorm(HID, eps=BERT_EPS)
        self.dropout = nn.Dropout(p_drop)

    def forward(self, input_ids, token_type_ids):
        T = input_ids.shape[1]
        pos = torch.arange(T, device=input_ids.device).unsqueeze(0)
        e = self.word_embeddings(input_ids) + self.token_type_embeddings(token_type_ids) + self.position_embeddings(pos)
        return self.dropout(self.LayerNorm(e))


class BertSelfAttention(nn.Module):
    """Eager path (torch 1.13 has no SDPA): softmax(QK^T / sqrt(d) + mask) -> dropout -> @V."""

    def __init__(self, p_drop):
        super().__init__()
        self.query = nn.Linear(HID, HID)
        self.key = nn.Linear(HID, HID)
        self.value = nn.Linear(HID, HID)
        self.dropout = nn.Dropout(p_drop)

    def _split(self, x):
        B, S, _ = x.shape
        return x.view(B, S, HEADS, HID // HEADS).permute(0, 2, 1, 3)

    def forward(self, hidden, ext_mask=None, encoder_hidden=None, encoder_ext_mask=None, return_probs=False):
        q = self._split(self.query(hidden))
        if encoder_hidden is not None:  # cross-attention: K/V from the image tokens, mask = encoder mask
            k, v, ext_mask = self._split(self.key(encoder_hidden)), self._split(self.value(encoder_hidden)), encoder_ext_mask
        else:
            k, v = self._split(self.key(hidden)), self._split(self.value(hidden))
        scores = q @ k.transpose(-1, -2) / math.sqrt(HID // HEADS)
        if ext_mask is not None:
            scores = scores + ext_mask
        probs = self.dropout(scores.softmax(dim=-1))
        ctx = (probs @ v).permute(0, 2, 1, 3).contiguous()
        ctx = ctx.view(ctx.shape[0], ctx.shape[1], HID)
        return (ctx, probs) if return_probs else ctx


class BertSelfOutput(nn.Module):
    def __init__(self, p_drop):
        super().__init__()
        self.dense = nn.Linear(HID, HID)
        self.LayerNorm = nn.LayerNorm(HID, eps=BERT_EPS)
        self.dropout = nn.Dropout(p_drop)

    def forward(self, hidden, input_tensor):
        return self.LayerNorm(self.dropout(self.dense(hidden)) + input_tensor)


class BertAttention(nn.Module):
    def __init__(self, p_drop):
        super().__init__()
        self.self = BertSelfAttention(p_drop)
        self.output = BertSelfOutput(p_drop)

    def forward(self, hidden, ext_mask):
        return self.output(self.self(hidden, ext_mask), hidden)


class BertIntermediate(nn.Module):
    def __init__(self):
        super().__init__()
        self.dense = nn.Linear(HID, FFN)

    def forward(self, x):
        return F.gelu(self.dense(x))


class BertOutput(nn.Module):
    def __init__(self, p_drop):
        super().__init__()
        self.dense = nn.Linear(FFN, HID)
        self.LayerNorm = nn.LayerNorm(HID, eps=BERT_EPS)
        self.dropout = nn.Dropout(p_drop)

    def forward(self, hidden, input_tensor):
        return self.LayerNorm(self.dropout(self.dense(hidden)) + input_tensor)


class BertLayer(nn.Module):
    def __init__(self, p_drop):
        super().__init__()
        self.attention = BertAttention(p_drop)
        self.intermediate = BertIntermediate()
        self.output = BertOutput(p_drop)

    def forward(self, hidden, ext_mask):
        a = self.attention(hidden, ext_mask)
        return self.output(self.intermediate(a), a)


class BertEncoder(nn.Module):
    def __init__(self, p_drop):
        super().__init__()
        self.layer = nn.ModuleList([BertLayer(p_drop) for _ in range(LAYERS)])

    def forward(self, hidden, ext_mask):
        for l in self.layer:
            hidden = l(hidden, ext_mask)
        return hidden


class BertPooler(nn.Module):
    def __init__(self):
        super().__init__()
        self.dense = nn.Linear(HID, HID)

    def forward(self, hidden):
        return torch.tanh(self.dense(hidden[:, 0]))


class ECAMPFusionLayer(nn.Module):
    """module/context_fusion.py:7-72."""

    def __init__(self, p_drop):
        super().__init__()
        self.attention = BertAttention(p_drop)
        self.cross_self_attention = BertSelfAttention(p_drop)
        self.intermediate = BertIntermediate()
        self.output = BertOutput(p_drop)
        self.gap_mlp = nn.Linear(HID, HID)
        self.out_layer = BertSelfOutput(p_drop)

    def forward(self, hidden, latent, gap_token, ext_text_mask, ext_img_mask, return_probs=False):
        attention_output = self.attention(hidden, ext_text_mask)                      # context_fusion.py:32-39
        cross = self.cross_self_attention(attention_output, ext_text_mask, latent, ext_img_mask,
                                          return_probs=return_probs)                  # :45-53
        probs = None
        if return_probs:
            cross, probs = cross
        cross = cross + self.gap_mlp(gap_token)                                       # :54-55
        attention_output = self.out_layer(cross, attention_output)                    # :56
        out = self.output(self.intermediate(attention_output), attention_output)      # :62-72
        return (out, probs) if return_probs else out


class MultimodalBertModel(nn.Module):
    """module/bert_modeling.py:10-156."""

    def __init__(self, p_drop):
        super().__init__()
        self.embeddings = BertEmbeddings(p_drop)
        self.encoder = BertEncoder(p_drop)
        self.pooler = BertPooler()  # created and executed although unused (bert_modeling.py:11-13,144)
        self.context_fusion_layer = ECAMPFusionLayer(p_drop)

    def forward(self, latent, gap_token, input_ids, attention_mask, token_type_ids):
        dtype = latent.dtype
        # get_extended_attention_mask: (1 - mask) * finfo(dtype).min, broadcast [B,1,1,S]  (bert_modeling.py:92-93)
        ext_text = (1.0 - attention_mask[:, None, None, :].to(dtype)) * torch.finfo(dtype).min
        ext_img = torch.zeros(latent.shape[0], 1, 1, latent.shape[1], dtype=dtype, device=latent.device)
        emb = self.embeddings(input_ids, token_type_ids)                              # :113-119
        fused = self.context_fusion_layer(emb, latent, gap_token, ext_text, ext_img)  # :121-129
        seq = self.encoder(fused, ext_text)                                           # :131-142
        _ = self.pooler(seq)                                                          # :144 (dead)
        return seq


class BertPredictionHeadTransform(nn.Module):
    def __init__(self):
        super().__init__()
        self.dense = nn.Linear(HID, HID)
        self.LayerNorm = nn.LayerNorm(HID, eps=BERT_EPS)

    def forward(self, x):
        return self.LayerNorm(F.gelu(self.dense(x)))


class BertLMPredictionHead(nn.Module):
    """transformers 4.42.4: decoder = Linear(bias=False); self.bias = Parameter; decoder.bias = self.bias
    -> ONE parameter under two state_dict keys.  The decoder weight is NOT tied to the live word
    embeddings because bert_modeling.py:161-163 replaces self.bert after post_init()."""

    def __init__(self):
        super().__init__()
        self.transform = BertPredictionHeadTransform()
        self.decoder = nn.Linear(HID, VOCAB, bias=False)
        self.bias = nn.Parameter(torch.zeros(VOCAB))
        self.decoder.bias = self.bias

    def forward(self, x):
        return self.decoder(self.transform(x))


class BertOnlyMLMHead(nn.Module):
    def __init__(self):
        super().__init__()
        self.predictions = BertLMPredictionHead()

    def forward(self, x):
        return self.predictions(x)


class MultimodalBertMaskedLM(nn.Module):
    """module/bert_modeling.py:160-228."""

    def __init__(self, p_drop):
        super().__init__()
        self.bert = MultimodalBertModel(p_drop)
        self.cls = BertOnlyMLMHead()

    def forward(self, latent, gap_token, input_ids, attention_mask, token_type_ids, weights, labels):
        seq = self.bert(latent, gap_token, input_ids, attention_mask, token_type_ids)
        logits = self.cls(seq)                                                        # :208-209
        ce = F.cross_entropy(logits.view(-1, VOCAB).float(), labels.view(-1), reduction="none")  # :211-213
        return (ce * weights.view(-1)).mean()                                         # :214-217


class MultiModalBertEncoder(nn.Module):
    """module/bert_encoder.py:12-22."""

    def __init__(self, p_drop):
        super().__init__()
        self.model = MultimodalBertMaskedLM(p_drop)

    def forward(self, latent, gap_token, ids, labels, attn_mask, token_type, weights):
        return self.model(latent, gap_token, ids, attn_mask, token_type, weights, labels)


# ----------------------------------------------------------------------------------------------
# module/model_ecamp.py:49-333 — the model
# ----------------------------------------------------------------------------------------------


def bicubic_downsample_2x(big):
    """torchvision 0.14.1 Resize([224,224], BICUBIC) on a tensor == F.interpolate(bicubic, align_corners=False,
    antialias=False) (model_ecamp.py:318)."""
    return F.interpolate(big, size=(big.shape[2] // 2, big.shape[3] // 2), mode="bicubic", align_corners=False,
                         antialias=False)


def len_keep_of(L, mask_ratio):
    return int(L * (1 - mask_ratio))  # model_ecamp.py:175 — python double arithmetic


def random_masking_ids(noise, len_keep):
    """model_ecamp.py:168-193 on a given noise tensor.  Contract (SURVEY §7 hard part 1): stable ascending
    argsort; ids_restore is the inverse permutation."""
    ids_shuffle = torch.argsort(noise, dim=1, stable=True)
    ids_restore = torch.argsort(ids_shuffle, dim=1, stable=True)
    ids_keep = ids_shuffle[:, :len_keep]
    mask = torch.ones(noise.shape, device=noise.device)
    mask[:, :len_keep] = 0
    mask = torch.gather(mask, 1, ids_restore)
    return ids_restore, ids_keep, mask


class EcampOracle(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12,
                 decoder_embed_dim=512, decoder_depth=4, decoder_num_heads=16, mlp_ratio=4, norm_pix_loss=False,
                 dropout=0.0):
        super().__init__()
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        n = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim), requires_grad=False)
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio, 1e-6) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.decoder_embed = nn.Linear(embed_dim, decoder_embed_dim)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, decoder_embed_dim))
        self.decoder_pos_embed = nn.Parameter(torch.zeros(1, n + 1, decoder_embed_dim), requires_grad=False)
        self.decoder_blocks = nn.ModuleList(
            [Block(decoder_embed_dim, decoder_num_heads, mlp_ratio, 1e-6) for _ in range(decoder_depth)])
        self.decoder_norm = nn.LayerNorm(decoder_embed_dim, eps=1e-6)
        self.decoder_pred = nn.Linear(decoder_embed_dim, patch_size ** 2 * in_chans)
        self.super_res = InterpolateConvSuperResolution()
        self.bert_encoder = MultiModalBertEncoder(dropout)
        self.bert_mlp = nn.Linear(embed_dim, 768)
        self.norm_pix_loss = norm_pix_loss  # stored, never read (model_ecamp.py:100)
        self.patch = patch_size
        self.initialize_weights()

    # model_ecamp.py:105-137 (+ HF BertPreTrainedModel._init_weights for the embeddings: normal(0, 0.02), pad row 0)
    def initialize_weights(self):
        g = int(self.patch_embed.num_patches ** .5)
        self.pos_embed.data.copy_(torch.from_numpy(sincos_2d(self.pos_embed.shape[-1], g)).float().unsqueeze(0))
        self.decoder_pos_embed.data.copy_(
            torch.from_numpy(sincos_2d(self.decoder_pos_embed.shape[-1], g)).float().unsqueeze(0))
        w = self.patch_embed.proj.weight.data
        nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        nn.init.normal_(self.cls_token, std=.02)
        nn.init.normal_(self.mask_token, std=.02)
        for m in self.modules():
            if isinstance(m, nn.Embedding):
                nn.init.normal_(m.weight, std=0.02)
                if m.padding_idx is not None:
                    m.weight.data[m.padding_idx].zero_()
            elif isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

    def unpatchify(self, x):  # model_ecamp.py:153-165
        p = self.patch
        h = w = int(x.shape[1] ** .5)
        x = x.reshape(x.shape[0], h, w, p, p, 3)
        return torch.einsum("nhwpqc->nchpwq", x).reshape(x.shape[0], 3, h * p, h * p)

    def mask_2_pixel(self, mask, column, row):  # model_ecamp.py:196-215
        p = self.patch
        g = int(mask.shape[1] ** .5)
        m = mask.reshape(mask.shape[0], g, g)
        sm = torch.zeros_like(m)
        for i in range(m.shape[0]):
            sm[i, int(column[i]):int(column[i]) + 12, int(row[i]):int(row[i]) + 12] = 1
        pm = torch.kron(m, torch.ones(p, p, device=m.device))
        spm = torch.kron(sm, torch.ones(2 * p, 2 * p, device=m.device))
        return pm.unsqueeze(1).repeat(1, 3, 1, 1), spm.unsqueeze(1).repeat(1, 3, 1, 1)

    def image_encoder(self, x, mask_ratio, noise=None):  # model_ecamp.py:218-237
        x = self.patch_embed(x) + self.pos_embed[:, 1:, :]
        N, L, D = x.shape
        if noise is None:
            noise = torch.rand(N, L, device=x.device)
        ids_restore, ids_keep, mask = random_masking_ids(noise, len_keep_of(L, mask_ratio))
        x = torch.gather(x, 1, ids_keep.unsqueeze(-1).repeat(1, 1, D))
        cls = (self.cls_token + self.pos_embed[:, :1, :]).expand(N, -1, -1)
        x = torch.cat((cls, x), dim=1)
        for blk in self.blocks:
            x = blk(x)
        return self.norm(x), mask, ids_restore, ids_keep

    def image_decoder(self, x, ids_restore):  # model_ecamp.py:240-264
        x = self.decoder_embed(x)
        mask_tokens = self.mask_token.repeat(x.shape[0], ids_restore.shape[1] + 1 - x.shape[1], 1)
        x_ = torch.cat([x[:, 1:, :], mask_tokens], dim=1)
        x_ = torch.gather(x_, 1, ids_restore.unsqueeze(-1).repeat(1, 1, x.shape[2]))
        x = torch.cat([x[:, :1, :], x_], dim=1) + self.decoder_pos_embed
        for blk in self.decoder_blocks:
            x = blk(x)
        return self.decoder_pred(self.decoder_norm(x))[:, 1:, :]

    def forward_loss(self, imgs, big_imgs, pred, mask, column, row):  # model_ecamp.py:276-300
        pred_img = self.unpatchify(pred.float())
        if big_imgs is None:  # positional 224-px form: no SR branch
            pm = torch.kron(mask.reshape(mask.shape[0], 14, 14), torch.ones(16, 16, device=mask.device))
            pm = pm.unsqueeze(1).repeat(1, 3, 1, 1)
            return F.mse_loss(pred_img * pm, imgs * pm), torch.zeros((), device=imgs.device)
        pixel_mask, super_mask = self.mask_2_pixel(mask, column, row)
        sr = self.super_res(pred_img)
        mim = F.mse_loss(pred_img * pixel_mask, imgs * pixel_mask, reduction="mean")
        res = F.mse_loss(sr * super_mask, big_imgs * super_mask, reduction="mean")
        return mim, res

    def forward_report_decoder(self, latent, ids, labels, attention_mask, type_ids, weights):  # :267-273
        latent = self.bert_mlp(latent)
        gap = latent[:, 1:, :].mean(dim=1).unsqueeze(1)
        return self.bert_encoder(latent[:, 1:, :], gap, ids, labels, attention_mask, type_ids, weights)

    def cross_attention_map(self, imgs, text_ids, attention_mask, type_ids, mask_ratio=0.0, noise=None):
        """Visualization/module/model_ecamp.py:272-278,308-319 + Visualization/module/context_fusion.py:26-57: image encoder
        at mask_ratio (0 by default), bert_mlp, GAP token, embeddings, the fusion layer's self-attention, then the
        probabilities `[B, 6, T, keep]` of its cross-attention (columns in ids_keep order, as in the reference)."""
        latent, mask, ids_restore, ids_keep = self.image_encoder(imgs, mask_ratio, noise)
        latent = self.bert_mlp(latent)
        gap = latent[:, 1:, :].mean(dim=1).unsqueeze(1)
        bert = self.bert_encoder.model.bert
        lat = latent[:, 1:, :]
        ext_text = (1.0 - attention_mask[:, None, None, :].to(lat.dtype)) * torch.finfo(lat.dtype).min
        ext_img = torch.zeros(lat.shape[0], 1, 1, lat.shape[1], dtype=lat.dtype, device=lat.device)
        emb = bert.embeddings(text_ids, type_ids)
        _, probs = bert.context_fusion_layer(emb, lat, gap, ext_text, ext_img, return_probs=True)
        self.last = dict(mask=mask, ids_restore=ids_restore, ids_keep=ids_keep, latent=latent)
        return probs

    def forward(self, batch, input_ids=None, attention_mask=None, labels=None, mask_ratio=0.75, *, type_ids=None,
                weights=None, big_imgs=None, column=None, row=None, noise=None):
        """Dict form = model_ecamp.py:303-325; positional form = SURVEY §8(b)(ii)."""
        if isinstance(batch, dict):
            if input_ids is not None:
                mask_ratio = input_ids  # forward(batch, mask_ratio)
            b = batch
            big_imgs, input_ids, labels = b["image"], b["ids"], b["labels"]
            attention_mask, type_ids, weights = b["attention_mask"], b["type_ids"], b["weights"]
            column, row = b["column"], b["row"]
            noise = b.get("noise", noise)
            imgs = bicubic_downsample_2x(big_imgs)
        else:
            imgs = batch
            if big_imgs is not None:
                imgs = bicubic_downsample_2x(big_imgs)
        if type_ids is None:
            type_ids = torch.zeros_like(input_ids)
        if weights is None:
            weights = torch.ones(input_ids.shape, dtype=torch.float32, device=input_ids.device)
        latent, mask, ids_restore, ids_keep = self.image_encoder(imgs, mask_ratio, noise)
        pred = self.image_decoder(latent, ids_restore)
        mim, res = self.forward_loss(imgs, big_imgs, pred, mask, column, row)
        mlm = self.forward_report_decoder(latent, input_ids, labels, attention_mask, type_ids, weights)
        self.last = dict(mask=mask, ids_restore=ids_restore, ids_keep=ids_keep, pred=pred, latent=latent)
        return mim, res, mlm


def ecamp_oracle(**kw):  # model_ecamp.py:328-333
    return EcampOracle(patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12, decoder_embed_dim=512,
                       decoder_depth=4, decoder_num_heads=16, mlp_ratio=4, **kw)


# ----------------------------------------------------------------------------------------------
# deterministic weights and synthetic batches shared by the golden generator, the tests and bench.py
# ----------------------------------------------------------------------------------------------


def seeded_state_dict(model, seed=0, std_scale=1.0):
    """Deterministic, construction-order-independent weights: every tensor is drawn from its own generator
    seeded by (seed, crc32(key)); LayerNorm weights ~ 1 + N(0, .1), biases ~ N(0, .02), matrices xavier-like
    normal.  Frozen sin-cos tables are left as built.  Gives every parameter (incl. biases, LN affine, conv
    kernels) a non-trivial value so that gradient parity is exercised everywhere."""
    import zlib
    out = {}
    sd = model.state_dict()
    for k, v in sd.items():
        if k in ("pos_embed", "decoder_pos_embed"):
            out[k] = v.clone()
            continue
        if k.endswith("cls.predictions.decoder.bias"):
            continue  # alias of cls.predictions.bias
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(k.encode())) % (2 ** 63))
        if "LayerNorm.weight" in k or k.endswith("norm.weight") or k.endswith("norm1.weight") or k.endswith("norm2.weight"):
            t = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        elif v.dim() == 1:
            t = 0.02 * torch.randn(v.shape, generator=g)
        elif "embeddings" in k and "LayerNorm" not in k:
            t = 0.02 * torch.randn(v.shape, generator=g)
            if "word_embeddings" in k:
                t[0].zero_()
        elif k.startswith("super_res"):
            t = 0.2 * torch.randn(v.shape, generator=g)
        else:
            fan_out = v.shape[0]
            fan_in = v[0].numel()
            t = math.sqrt(2.0 / (fan_in + fan_out)) * torch.randn(v.shape, generator=g)
        out[k] = (t * std_scale).to(v.dtype) if ("norm" not in k.lower()) else t.to(v.dtype)
    if "bert_encoder.model.cls.predictions.bias" in out:
        out["bert_encoder.model.cls.predictions.decoder.bias"] = out["bert_encoder.model.cls.predictions.bias"]
    return out


def synthetic_batch(B, T=128, big=True, seed=1234, device="cpu"):
    """SURVEY §8(d) synthetic inputs (shape and value ranges of pretrain_datasets.py collate output)."""
    g = torch.Generator().manual_seed(seed)
    side = 448 if big else 224
    img = torch.randn(B, 1, side, side, generator=g).expand(B, 3, side, side).contiguous()
    length = torch.randint(T // 4, T + 1, (B,), generator=g)
    pos = torch.arange(T).unsqueeze(0)
    attn = (pos < length.unsqueeze(1)).long()
    labels = torch.randint(5, VOCAB, (B, T), generator=g) * attn
    labels[:, 0] = 2
    ids = labels.clone()
    mask_here = (torch.rand(B, T, generator=g) < 0.45) & (attn == 1) & (pos > 0)
    ids[mask_here] = 3
    weights = torch.ones(B, T)
    rows = (torch.rand(B, generator=g) < 0.05).nonzero().flatten().tolist()
    if B <= 8 and not rows:
        rows = [0]
    for r in rows:
        s = int(torch.randint(1, max(2, T - 6), (1,), generator=g))
        weights[r] = T / (T - 0.95 * 5)
        weights[r, s:s + 5] = 0.05
    col = torch.randint(0, 3, (B,), generator=g)
    row = torch.randint(0, 3, (B,), generator=g)
    noise = torch.rand(B, 196, generator=g)
    batch = dict(image=img, ids=ids, labels=labels, attention_mask=attn, type_ids=torch.zeros(B, T, dtype=torch.long),
                 weights=weights, column=col, row=row, noise=noise)
    return {k: v.to(device) for k, v in batch.items()}
