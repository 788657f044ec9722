"""TEST INFRASTRUCTURE.  Runs ONLY in the build container (needs /root/reference).

Pins oracle/ecamp_oracle.py against outputs of the reference's own sources and writes the small
fixtures under tests/golden/ that tests/test_oracle.py (CPU) and tests/test_parity_gpu.py (GPU) check:

  * state_dict layout of the reference class                       -> state_dict_layout.json
  * the WHOLE step executed FROM THE REFERENCE SOURCE: `ECAMP.forward(batch)` (module/model_ecamp.py:303-325) unmodified,
    i.e. bicubic resize, random_masking, image_encoder, image_decoder, mask_2_pixel, unpatchify, super_res, forward_loss,
    forward_report_decoder, MultiModalBertEncoder.forward (bert_encoder.py), MultimodalBertMaskedLM.forward and
    MultimodalBertModel.forward (bert_modeling.py:15-228), ECAMPFusionLayer.forward (context_fusion.py:21-72), with
    injected noise; losses and every parameter gradient                 -> golden_cases.json
    (timm 0.4.12 PatchEmbed / Block are stubbed and the transformers-4.42.4 leaf entry points the reference calls -
    BertSelfAttention / BertAttention / BertEncoder.forward signatures, get_extended_attention_mask, get_head_mask -
    get their 4.42.4 calling convention back on transformers 5.x, both per oracle/_ref_shims.py; torchvision's Resize
    keeps the pinned 0.14.1 default antialias=False)
  * tie-breaking of argsort on deliberately tied noise, from the reference's random_masking.

Usage:  python oracle/make_golden.py          (re-generates tests/golden/*.json; asserts oracle == reference)
"""
import json
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import _ref_shims  # noqa: E402

_ref_shims.install()
from module.model_ecamp import ecamp  # noqa: E402  (the reference, from /root/reference)
from oracle.ecamp_oracle import ecamp_oracle, seeded_state_dict, synthetic_batch  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)
TOL = 2e-5


def with_noise(noise, fn):
    orig = torch.rand
    torch.rand = lambda *a, **k: noise.clone()
    try:
        return fn()
    finally:
        torch.rand = orig


def main():
    torch.manual_seed(0)
    ref = ecamp(norm_pix_loss=True)
    orc = ecamp_oracle(norm_pix_loss=True)
    sd_ref = ref.state_dict()
    layout = {k: list(v.shape) for k, v in sd_ref.items()}
    assert list(sd_ref.keys()) == list(orc.state_dict().keys())
    assert all(tuple(orc.state_dict()[k].shape) == tuple(v.shape) for k, v in sd_ref.items())
    assert torch.equal(ref.pos_embed, orc.pos_embed) and torch.equal(ref.decoder_pos_embed, orc.decoder_pos_embed)
    frozen = [k for k, p in ref.named_parameters() if not p.requires_grad]
    json.dump(dict(keys=layout, frozen=frozen,
                   note="transformers 4.42.4 aliases cls.predictions.decoder.bias to cls.predictions.bias "
                        "(349 distinct parameters); transformers 5.x, installed here, does not."),
              open(os.path.join(GOLD, "state_dict_layout.json"), "w"), indent=0)

    w = seeded_state_dict(orc, 0)
    orc.load_state_dict(w)
    ref.load_state_dict(w)
    # transformers 5.x un-aliases decoder.bias: make the reference model use the 4.42.4 semantics
    ref.bert_encoder.model.cls.predictions.decoder.bias = ref.bert_encoder.model.cls.predictions.bias
    ref.eval()
    orc.eval()
    _ref_shims.install_bert_adapters(ref)

    cases = []
    for (B, T, seed) in [(2, 32, 1), (3, 128, 2), (2, 256, 3)]:
        b = synthetic_batch(B, T=T, seed=seed)
        big = b["image"]
        import torchvision
        from torchvision.transforms.functional import InterpolationMode
        imgs = torchvision.transforms.Resize([224, 224], interpolation=InterpolationMode.BICUBIC, antialias=False)(big)

        def ref_pieces():   # intermediates for the fixtures (the reference's forward does not return them)
            with torch.no_grad():
                lat, mask, idr, idk = ref.image_encoder(imgs, 0.75)
                pred = ref.image_decoder(lat, idr)
            return lat, mask, idr, idk, pred

        for p in ref.parameters():
            p.grad = None
        lat, mask, idr, idk, pred = with_noise(b["noise"], ref_pieces)
        mim, res, mlm = with_noise(b["noise"], lambda: ref(b))   # ECAMP.forward from the reference source, end to end
        (mim + res + mlm).backward()
        gref = {k: p.grad.clone() for k, p in ref.named_parameters() if p.grad is not None}

        for p in orc.parameters():
            p.grad = None
        mim_o, res_o, mlm_o = orc(b)
        (mim_o + res_o + mlm_o).backward()
        gorc = {k: p.grad for k, p in orc.named_parameters() if p.grad is not None}

        def rel(a, c):
            return abs(a - c) / max(abs(c), 1e-12)

        assert torch.equal(mask, orc.last["mask"]) and torch.equal(idr, orc.last["ids_restore"])
        assert torch.equal(idk, orc.last["ids_keep"])
        assert rel(mim_o.item(), mim.item()) < TOL and rel(res_o.item(), res.item()) < TOL, (mim_o, mim, res_o, res)
        assert rel(mlm_o.item(), mlm.item()) < TOL, (mlm_o.item(), mlm.item())
        assert set(gref) == set(gorc), set(gref) ^ set(gorc)
        worst = 0.0
        for k in gref:
            # key biases have an analytically zero gradient (softmax shift invariance): floor the denominator
            e = (gref[k] - gorc[k]).norm().item() / max(gref[k].norm().item(), 1e-5)
            worst = max(worst, e)
            assert e < 5e-4, (k, e)
        print(f"case B={B} T={T}: mim {mim.item():.6f} res {res.item():.6f} mlm {mlm.item():.6f} "
              f"oracle==reference; worst grad rel-L2 {worst:.2e}")
        picks = ["cls_token", "mask_token", "patch_embed.proj.weight", "blocks.0.attn.qkv.weight", "blocks.11.mlp.fc2.bias",
                 "norm.weight", "decoder_embed.weight", "decoder_blocks.3.attn.proj.weight", "decoder_pred.bias",
                 "super_res.conv1.weight", "super_res.conv2.bias", "bert_mlp.weight",
                 "bert_encoder.model.bert.embeddings.word_embeddings.weight",
                 "bert_encoder.model.bert.embeddings.position_embeddings.weight",
                 "bert_encoder.model.bert.context_fusion_layer.gap_mlp.weight",
                 "bert_encoder.model.bert.context_fusion_layer.cross_self_attention.key.weight",
                 "bert_encoder.model.bert.encoder.layer.5.output.LayerNorm.weight",
                 "bert_encoder.model.cls.predictions.decoder.weight", "bert_encoder.model.cls.predictions.bias"]
        cases.append(dict(
            B=B, T=T, seed=seed, weight_seed=0, mask_ratio=0.75,
            input_checksums=dict(image_sum=big.double().sum().item(), ids_sum=int(b["ids"].sum()),
                                 noise_sum=b["noise"].double().sum().item(), weights_sum=b["weights"].double().sum().item()),
            mim_loss=mim.item(), res_loss=res.item(), mlm_loss=mlm.item(),
            ids_restore=idr.tolist(), ids_keep=idk.tolist(), mask_sum=mask.sum(1).tolist(),
            latent_abs_mean=lat.abs().mean().item(), pred_abs_mean=pred.abs().mean().item(),
            no_grad_params=sorted(k for k, p in ref.named_parameters() if p.grad is None and p.requires_grad),
            grad_norms={k: gref[k].norm().item() for k in picks},
            word_emb_grad_row0_abs_max=gref["bert_encoder.model.bert.embeddings.word_embeddings.weight"][0].abs().max().item(),
        ))

    # argsort tie-breaking, from the reference's random_masking (model_ecamp.py:168-193)
    g = torch.Generator().manual_seed(7)
    noise = torch.rand(4, 196, generator=g)
    noise[0, 10] = noise[0, 150]
    noise[0, 3] = noise[0, 150]
    noise[1, :] = 0.5                       # everything tied
    noise[2, 100:] = noise[2, :96]          # every value appears twice
    noise[3, ::2] = 0.25
    x = torch.zeros(4, 196, 8)
    _, mask, idr, idk = with_noise(noise, lambda: ref.random_masking(x, 0.75))
    from oracle.ecamp_oracle import random_masking_ids, len_keep_of
    idr_o, idk_o, mask_o = random_masking_ids(noise, len_keep_of(196, 0.75))
    # The reference calls torch.argsort(noise) with stable=False: on tied keys its order is implementation-defined
    # (this container's CPU sort and torch 1.13's CUDA bitonic sort both reorder ties).  The reference result must be
    # a VALID ascending sort with the same kept count; the pinned contract is the stable order (lower index first).
    ref_shuffle = torch.argsort(idr, dim=1)
    assert bool((torch.gather(noise, 1, ref_shuffle).diff(dim=1) >= 0).all())
    assert torch.equal(mask.sum(1), mask_o.sum(1))
    assert torch.equal(torch.gather(noise, 1, idk).sort(1).values, torch.gather(noise, 1, idk_o).sort(1).values)
    untied = torch.rand(2, 196, generator=g)
    _, m_u, idr_u, idk_u = with_noise(untied, lambda: ref.random_masking(torch.zeros(2, 196, 8), 0.75))
    idr_uo, idk_uo, m_uo = random_masking_ids(untied, 49)
    assert torch.equal(idr_u, idr_uo) and torch.equal(idk_u, idk_uo) and torch.equal(m_u, m_uo)
    ties = dict(noise_bits=noise.view(torch.int32).tolist(), ids_restore=idr_o.tolist(), ids_keep=idk_o.tolist(),
                mask=mask_o.int().tolist(), contract="stable ascending argsort (lower index first on ties)",
                len_keep={str(r): int(196 * (1 - r)) for r in (0.75, 0.9, 0.7, 0.5, 0.6)})
    json.dump(dict(cases=cases, ties=ties, tolerance=TOL,
                   generator="oracle/make_golden.py: ECAMP.forward run from the /root/reference sources end to end; transformers "
                             + __import__("transformers").__version__),
              open(os.path.join(GOLD, "golden_cases.json"), "w"))
    print("wrote", GOLD)


if __name__ == "__main__":
    main()
