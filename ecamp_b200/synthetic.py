"""Synthetic pre-training batches with the shapes / value ranges of the reference's collate output
(ECAMP/Pre-training/module/pretrain_datasets.py:202-239; recipe in SURVEY.md §8d).  Used by bench.py and smoke()."""
import torch

VOCAB = 30000


def make_batch(B, T=128, big=True, seed=1234, pin=False):
    g = torch.Generator().manual_seed(seed)
    side = 448 if big else 224
    img = torch.randn(B, 1, side, side, generator=g).expand(B, 3, side, side).contiguous()  # Grayscale(3) + Normalize
    length = torch.randint(T // 4, T + 1, (B,), generator=g)
    pos = torch.arange(T).unsqueeze(0)
    attn = (pos < length.unsqueeze(1)).long()
    labels = torch.randint(5, VOCAB, (B, T), generator=g) * attn
    labels[:, 0] = 2                                                   # [CLS]
    ids = labels.clone()
    ids[(torch.rand(B, T, generator=g) < 0.45) & (attn == 1) & (pos > 0)] = 3   # [MASK]
    weights = torch.ones(B, T)
    for r in (torch.rand(B, generator=g) < 0.05).nonzero().flatten().tolist():  # re-balanced rows (:141-184)
        s = int(torch.randint(1, max(2, T - 6), (1,), generator=g))
        weights[r] = T / (T - 0.95 * 5)
        weights[r, s:s + 5] = 0.05
    batch = dict(image=img, ids=ids, labels=labels, attention_mask=attn, type_ids=torch.zeros(B, T, dtype=torch.long),
                 weights=weights, column=torch.randint(0, 3, (B,), generator=g), row=torch.randint(0, 3, (B,), generator=g),
                 noise=torch.rand(B, 196, generator=g))
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    return batch
