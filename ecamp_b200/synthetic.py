"""Synthetic pre-training batches with the shapes / value ranges of the reference's collate output
(ECAMP/Pre-training/module/pretrain_datasets.py:202-239; recipe in SURVEY.md §8d).  Used by bench.py and smoke()."""
import os

import torch

VOCAB = 30000


def bind_host_to_gpu(device_index=0):
    """Restrict this process to the CPUs of the NUMA node the GPU hangs off (sysfs `local_cpulist` of its PCI function),
    so that pinned staging buffers allocated afterwards are first-touched on that node: a 617 MB batch copied from the
    remote socket can take longer than the whole step.  Returns the CPU list used, or None if the topology is unknown
    (single node, container without sysfs, ...) - then nothing is changed."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        text = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
        cpus = []
        for part in text.split(","):
            if part:
                a, _, b = part.partition("-")
                cpus += list(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus = sorted(c for c in cpus if c in allowed)
        if not cpus or len(cpus) == len(allowed):
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def make_batch(B, T=128, big=True, seed=1234, pin=False, u8=False):
    """u8=True: `image` is the loader's 8-bit grayscale crop [B, side, side] (the module applies Grayscale(3) + ToTensor +
    Normalize on the GPU); otherwise the normalised fp32 [B, 3, side, side] tensor the reference's collate produces."""
    g = torch.Generator().manual_seed(seed)
    side = 448 if big else 224
    if u8:
        img = torch.randint(0, 256, (B, side, side), generator=g, dtype=torch.uint8)
    else:
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
