"""Generates tests/golden/text_masking.json by running the UNMODIFIED reference dataset
(/root/reference/ECAMP/Pre-training/module/pretrain_datasets.py: ContextBertDataset.__getitem__ / _context_mask) on a
throw-away data root (two CSV files, one JPEG, the reference's tokenizer file) with seeded `random`.  Test
infrastructure only; runs in the build container (needs /root/reference).  The fixture stores, per case, the RNG seed,
the report / LLM strings, the unmasked ids, the masked ids, the fp32 weights and - so that the CPU test needs neither
the tokenizer file nor the reference - which of the ids that occur are '##' pieces / entity words."""
import json
import os
import random
import sys
import tempfile
import types

import numpy as np
import pandas as pd
from PIL import Image

REF_PT = "/root/reference/ECAMP/Pre-training"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "text_masking.json")

REPORTS = [
    ("there is no evidence of pneumothorax. the cardiomediastinal silhouette is normal. mild bibasilar atelectasis is present.",
     "small left pleural effusion with adjacent atelectasis."),
    ("the lungs are clear. there is no focal consolidation. heart size is normal. no acute osseous abnormality.",
     "no acute cardiopulmonary process."),
    ("there is no pleural effusion or pneumothorax. there is no evidence of pulmonary edema. stable cardiomegaly.",
     "cardiomegaly without edema."),
    ("lines and tubes are in standard position. lung volumes are low. the patient is rotated.",
     "low lung volumes."),
    ("right lower lobe opacity is concerning for pneumonia. there is no evidence of a large effusion. "
     "the hilar contours are unremarkable. degenerative changes of the spine are noted. recommend follow up.",
     "right lower lobe pneumonia."),
    ("no change.", "unremarkable study."),
]


def main():
    sys.modules.setdefault("ipdb", types.ModuleType("ipdb"))
    sys.path.insert(0, REF_PT)
    from module.pretrain_datasets import ContextBertDataset          # the reference, unmodified
    with tempfile.TemporaryDirectory() as root:
        os.symlink(os.path.join(REF_PT, "dataset", "mimic_wordpiece.json"), os.path.join(root, "mimic_wordpiece.json"))
        img = os.path.join(root, "x.jpg")
        Image.fromarray((np.random.RandomState(0).rand(64, 64) * 255).astype(np.uint8)).save(img)
        pd.DataFrame(dict(img_path=[img] * len(REPORTS), report=[r for r, _ in REPORTS], llm_output=[l for _, l in REPORTS])) \
            .to_csv(os.path.join(root, "mimic-cxr-2.0.0-entity-llm.csv"), index=False)
        pd.DataFrame(dict(label_i=[0] * len(REPORTS), label_j=[1] * len(REPORTS))).to_csv(os.path.join(root, "mimic-cxr-2.0.0-attn-label.csv"), index=False)
        cases, sub_ids, ent_ids = [], set(), set()
        for T in (256, 64):
            ds = ContextBertDataset(root, max_caption_length=T)
            ents = __import__("module.pretrain_datasets", fromlist=["entities"]).entities
            for idx in range(len(REPORTS)):
                for seed in (0, 1, 2):
                    random.seed(1000 * T + 10 * idx + seed)
                    image, ids, attention_mask, type_ids, masked_ids, weights, column, row = ds[idx]
                    after = random.random()                          # position of the stream after the sample
                    ids_l = ids[0].tolist()
                    for v in set(ids_l):
                        w = ds.idxtoword[v]
                        if w[0:2] == "##":
                            sub_ids.add(v)
                        if w in ents:
                            ent_ids.add(v)
                    cases.append(dict(T=T, idx=idx, seed=1000 * T + 10 * idx + seed, report=ds.report_list[idx],
                                      llm_output=ds.llm_out_list[idx], ids=ids_l, attention_mask=attention_mask[0].tolist(),
                                      masked_ids=masked_ids[0].tolist(), weights=[float(x) for x in weights[0].tolist()],
                                      next_random=after))
        json.dump(dict(generator="oracle/make_text_golden.py", reference="ECAMP/Pre-training/module/pretrain_datasets.py:60-191",
                       sub_ids=sorted(sub_ids), entity_ids=sorted(ent_ids), cases=cases), open(OUT, "w"))
        print("wrote", OUT, len(cases), "cases;", sum(1 for c in cases if 3 in c["masked_ids"]), "with masks")


if __name__ == "__main__":
    main()
