"""Generates tests/golden/image_pipeline.json by running the UNMODIFIED reference image transform
(/root/reference/ECAMP/Pre-training/module/pretrain_datasets.py: `pil_loader` + `ContextBertDataset.transform` =
RandomResizedCrop(448, scale=(0.2, 1.0), BICUBIC) -> RandomHorizontalFlip -> Grayscale(3) -> ToTensor -> Normalize) on
procedurally generated grayscale frames (saved losslessly, loaded with the reference's `pil_loader`) with a seeded torch
generator.  Test infrastructure only; runs in the build container (needs /root/reference).  Per case the fixture stores the
frame recipe, the torch seed, the parameters torchvision drew, a draw from the generator AFTER the sample (stream position),
the SHA-256 of the 448 x 448 output bytes and of the normalised fp32 tensor, and 64 sampled output pixels."""
import hashlib
import json
import os
import sys
import tempfile
import types

import numpy as np
import pandas as pd
import torch
from PIL import Image

REF_PT = "/root/reference/ECAMP/Pre-training"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "image_pipeline.json")
sys.path.insert(0, ROOT)
from tests.image_frames import make_frame  # noqa: E402  (the procedural frames, shared with the tests)

CASES = [(300, 260, 0), (512, 512, 1), (700, 1000, 2), (1200, 900, 3), (448, 448, 4), (2048, 2500, 5), (200, 640, 6)]


def main():
    sys.modules.setdefault("ipdb", types.ModuleType("ipdb"))
    sys.path.insert(0, REF_PT)
    from module.pretrain_datasets import ContextBertDataset, pil_loader      # the reference, unmodified
    import torchvision.transforms as T
    with tempfile.TemporaryDirectory() as root:
        os.symlink(os.path.join(REF_PT, "dataset", "mimic_wordpiece.json"), os.path.join(root, "mimic_wordpiece.json"))
        paths = []
        for (H, W, fs) in CASES:
            p = os.path.join(root, f"f{fs}.png")
            Image.fromarray(make_frame(H, W, fs)).save(p)
            paths.append(p)
        pd.DataFrame(dict(img_path=paths, report=["no change."] * len(paths), llm_output=["unremarkable study."] * len(paths))) \
            .to_csv(os.path.join(root, "mimic-cxr-2.0.0-entity-llm.csv"), index=False)
        pd.DataFrame(dict(label_i=[0] * len(paths), label_j=[1] * len(paths))).to_csv(os.path.join(root, "mimic-cxr-2.0.0-attn-label.csv"), index=False)
        ds = ContextBertDataset(root, max_caption_length=64)
        rrc = ds.transform.transforms[0]
        assert isinstance(rrc, T.RandomResizedCrop) and isinstance(ds.transform.transforms[1], T.RandomHorizontalFlip)
        cases = []
        for (H, W, fs), path in zip(CASES, paths):
            for seed in (11, 12, 13):
                ts = 100 * fs + seed
                torch.manual_seed(ts)
                out = ds.transform(pil_loader(path))                              # the reference pipeline, fp32 [3, 448, 448]
                after = float(torch.rand(1))
                torch.manual_seed(ts)                                             # what torchvision drew for it
                i, j, h, w = rrc.get_params(pil_loader(path), rrc.scale, rrc.ratio)
                flip = bool(torch.rand(1) < 0.5)
                assert torch.equal(out[0], out[1]) and torch.equal(out[0], out[2])
                mean, std = torch.tensor(0.4721), torch.tensor(0.3037)
                u8 = torch.round((out[0] * std + mean) * 255).to(torch.uint8)
                assert torch.equal(u8.float().div(255).sub(mean).div(std), out[0])  # bytes <-> normalised tensor, exactly
                samp = [(int(y), int(x)) for y, x in np.random.RandomState(ts).randint(0, 448, (64, 2))]
                cases.append(dict(H=H, W=W, frame_seed=fs, torch_seed=ts, i=i, j=j, h=h, w=w, flip=flip, next_rand=after,
                                  sha256_u8=hashlib.sha256(u8.numpy().tobytes()).hexdigest(),
                                  sha256_f32=hashlib.sha256(out.numpy().tobytes()).hexdigest(),
                                  samples=[[y, x, int(u8[y, x])] for y, x in samp]))
        json.dump(dict(generator="oracle/make_image_golden.py", reference="ECAMP/Pre-training/module/pretrain_datasets.py:27-31,47-52",
                       pillow=Image.__version__, torchvision=__import__("torchvision").__version__, cases=cases), open(OUT, "w"))
        print("wrote", OUT, len(cases), "cases;", sum(c["flip"] for c in cases), "flipped")


if __name__ == "__main__":
    main()
