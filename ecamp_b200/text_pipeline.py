"""Host side of the report branch: sentence splice, entity-centred context masking and loss re-weighting — the part of
the reference's `ContextBertDataset.__getitem__` (ECAMP/Pre-training/module/pretrain_datasets.py:113-191) and `_context_mask`
(:60-110) that decides WHICH tokens the MLM loss sees and with which weight (north_star: "MLM label
selection bit-exact").  Same decisions, same order of `random.random()` / `random.randint()` draws, therefore the same
masked ids / mask positions / fp32 weights for the same RNG state - but on plain Python ints and two vocabulary
look-up tables instead of one tensor `.item()` + dict look-up per token test (the reference spends ~5 ms per report
there; this is ~50x less, which the 5.7 k pairs/s step needs from its 16 loader workers).

The outputs feed the model exactly as the reference's collate_fn does (:202-239): `labels` = unmasked ids, `ids` = masked
ids, `weights` = per-token loss weights.
"""
import random as _random

import numpy as np

PAD, UNK, CLS, MASK, SEP, PERIOD = 0, 1, 2, 3, 4, 16           # mimic_wordpiece.json ids (pretrain_datasets.py:72-90)
TEMPLATE1 = (219, 149, 152, 422, 158)                          # "there is no evidence of" (pretrain_datasets.py:23)
TEMPLATE2 = (219, 149, 152)                                    # "there is no"             (:24)
ENTITIES = ('abnormality', 'abscess', 'aerate', 'aorta', 'atelectasis', 'bronchiectasis', 'calcification', 'cardiomediastinal',
            'cardiomegaly', 'catheter', 'chf', 'collapse', 'congestion', 'consolidation', 'contour', 'COPD',
            'deformity', 'dilation', 'distention', 'edema', 'effusion', 'embolism', 'emphysema', 'engorgement',
            'fibrosis', 'fracture', 'granuloma', 'hernia', 'hilar', 'hyperinflate', 'hemidiaphragm', 'infiltrate',
            'mass', 'nodule', 'obscure', 'opacity', 'perihilar', 'pneumonia', 'pneumothorax', 'sarcoidosis',
            'silhouette', 'thickening', 'tuberculosis', 'vasculature')   # pretrain_datasets.py:17-22


class VocabTables:
    """is_sub[v]: the word piece of id v starts with '##'; is_entity[v]: it is one of the 44 entity words."""

    def __init__(self, is_sub, is_entity):
        self.is_sub, self.is_entity = is_sub, is_entity
        # byte tables for the native path (converted once, not per report)
        self.sub_u8 = np.ascontiguousarray(is_sub, dtype=np.uint8)
        self.entity_u8 = np.ascontiguousarray(is_entity, dtype=np.uint8)

    @classmethod
    def from_vocab(cls, vocab):
        """vocab: {word: id} as returned by tokenizers.Tokenizer.get_vocab()."""
        n = max(vocab.values()) + 1
        is_sub, is_entity = np.zeros(n, dtype=bool), np.zeros(n, dtype=bool)
        ent = set(ENTITIES)
        for w, v in vocab.items():
            is_sub[v] = w[0:2] == '##'
            is_entity[v] = w in ent
        return cls(is_sub, is_entity)

    @classmethod
    def from_sparse(cls, sub_ids, entity_ids, n=30000):
        is_sub, is_entity = np.zeros(n, dtype=bool), np.zeros(n, dtype=bool)
        is_sub[list(sub_ids)] = True
        is_entity[list(entity_ids)] = True
        return cls(is_sub, is_entity)


def splice_report(report, llm_output, rng=_random):
    """pretrain_datasets.py:116-133: with probability 0.8 the LLM summary is inserted before sentence `location`."""
    parts = report.split('.')
    n = len(parts)
    sent = ""
    add_prob = rng.random()
    if add_prob < 0.8:
        location = rng.randint(0, n)
        for i in range(0, location):
            sent += parts[i]
            sent += "."
        sent += llm_output
        for i in range(location, n):
            sent += parts[i]
            sent += "."
    else:
        sent = report
    sent = sent.replace("..", ".")
    return '[CLS] ' + sent


def context_mask(ids, tables, rng=_random):
    """pretrain_datasets.py:60-110 on a list of ints (one padded report).  Returns (masked ids, mask_pos).
    One rng.random() per visited non-continuation position of the second loop, then one per entity position."""
    tokens = [int(t) for t in ids]
    masked = list(tokens)
    T = len(tokens)
    is_sub, is_entity = tables.is_sub, tables.is_entity
    entity_pos, mask_pos = [], []
    entity_exist = False
    for i in range(1, T - 1):
        if is_entity[masked[i]]:
            entity_exist = True
            break
    for i in range(1, T - 1):
        cur = masked[i]
        if cur == PAD:
            break
        if is_sub[cur]:
            if masked[i - 1] == MASK:
                masked[i] = MASK      # a continuation piece follows its (masked) head
            continue
        if is_entity[cur]:
            entity_pos.append(i)
            for j in range(1, 3):
                if i - j <= 0:
                    break
                if tokens[i - j] != PERIOD:
                    if i - j not in mask_pos:
                        mask_pos.append(i - j)
                    # (the reference then tests `word(i) not in entities`, which is false in this branch: no-op)
        prob = rng.random()
        if not entity_exist:
            if prob < 0.75:
                masked[i] = MASK
        elif prob < 0.7 and i not in entity_pos and i not in mask_pos:
            masked[i] = MASK
    for i in range(1, T - 1):   # mask entity on 75% prob
        if i in entity_pos:
            if rng.random() < 0.75:
                masked[i] = MASK
    return masked, mask_pos


def template_weights(ids, mask_pos, max_len):
    """pretrain_datasets.py:141-184: 0.05 on the "there is no (evidence of)" templates, the removed weight handed to the
    entity-context positions (or spread over the report).  fp32 arithmetic like the reference's torch tensor."""
    w = np.ones(max_len, dtype=np.float32)
    t = [int(x) for x in ids]
    diminish_pos, diminish_cnt = [], 0
    i, n = 0, len(t)
    while i < n - 4:
        if tuple(t[i:i + 5]) == TEMPLATE1:
            w[i:i + 5] = 0.05
            diminish_pos.extend(range(i, i + 5))
            diminish_cnt += 5
            i += 5
        elif tuple(t[i:i + 3]) == TEMPLATE2:
            w[i:i + 3] = 0.05
            diminish_pos.extend(range(i, i + 3))
            diminish_cnt += 3
            i += 3
        else:
            i += 1
    dp = set(diminish_pos)
    len_dm = sum(1 for x in mask_pos if x in dp)
    mask_cnt = len(mask_pos)
    if mask_cnt > 0 and diminish_cnt > 0:
        expand = (0.95 * (diminish_cnt - len_dm) + mask_cnt) / (mask_cnt - 0.95 * len_dm)
        for p in mask_pos:
            w[p] = w[p] * np.float32(expand)
    elif diminish_cnt > 0:
        expand = max_len / (max_len - 0.95 * diminish_cnt)
        w = w * np.float32(expand)
    return w


# ---- native (C, no CUDA) versions of the two per-token loops: csrc/text_host.cu through the C ABI -----------------------
def _native():
    import ctypes
    from . import _lib as L
    return ctypes, L


def context_mask_native(ids, tables, rng=_random):
    """`context_mask` in native code.  The draws are taken from `rng` here - exactly as many as the reference consumes
    (a function of the tokens alone) - and handed to the library, so the generator ends at the same position."""
    ctypes, L = _native()
    lib = L.lib()
    a = np.ascontiguousarray(ids, dtype=np.int64)
    T = int(a.shape[0])
    sub, ent = tables.sub_u8, tables.entity_u8
    vocab = int(min(sub.shape[0], ent.shape[0]))
    pp = lambda x, t: x.ctypes.data_as(ctypes.POINTER(t))
    n = lib.ecamp_text_mask_draw_count(pp(a, ctypes.c_int64), T, pp(sub, ctypes.c_uint8), pp(ent, ctypes.c_uint8), vocab)
    if n < 0:
        raise ValueError("context_mask_native: bad arguments")
    draws = np.array([rng.random() for _ in range(n)], dtype=np.float64)
    masked = np.empty(T, dtype=np.int64)
    mask_pos = np.empty(max(T, 1), dtype=np.int32)
    cnt = ctypes.c_int32(0)
    L.check(lib.ecamp_text_context_mask(pp(a, ctypes.c_int64), T, pp(sub, ctypes.c_uint8), pp(ent, ctypes.c_uint8), vocab,
                                        pp(draws, ctypes.c_double), n, pp(masked, ctypes.c_int64), pp(mask_pos, ctypes.c_int32),
                                        ctypes.byref(cnt)), "ecamp_text_context_mask")
    return masked.tolist(), mask_pos[:cnt.value].tolist()


def template_weights_native(ids, mask_pos, max_len):
    ctypes, L = _native()
    a = np.ascontiguousarray(ids, dtype=np.int64)
    mp = np.ascontiguousarray(mask_pos, dtype=np.int32)
    w = np.empty(max_len, dtype=np.float32)
    pp = lambda x, t: x.ctypes.data_as(ctypes.POINTER(t))
    L.check(L.lib().ecamp_text_template_weights(pp(a, ctypes.c_int64), int(a.shape[0]), pp(mp, ctypes.c_int32), int(mp.shape[0]),
                                                int(max_len), pp(w, ctypes.c_float)), "ecamp_text_template_weights")
    return w


class ReportTextPipeline:
    """tokenizer (the reference's mimic_wordpiece.json, loaded with `tokenizers`) + the three steps above."""

    def __init__(self, tokenizer_json, max_caption_length=256):
        import tokenizers
        self.tokenizer = tokenizers.Tokenizer.from_file(tokenizer_json)
        self.tables = VocabTables.from_vocab(self.tokenizer.get_vocab())
        self.max_len = max_caption_length
        self.tokenizer.enable_truncation(max_length=max_caption_length)
        self.tokenizer.enable_padding(length=max_caption_length)

    def __call__(self, report, llm_output, rng=_random):
        enc = self.tokenizer.encode(splice_report(report, llm_output, rng))
        ids = list(enc.ids)
        masked, mask_pos = context_mask(ids, self.tables, rng)
        return dict(labels=np.asarray(ids, dtype=np.int64), ids=np.asarray(masked, dtype=np.int64),
                    attention_mask=np.asarray(enc.attention_mask, dtype=np.int64), type_ids=np.asarray(enc.type_ids, dtype=np.int64),
                    weights=template_weights(ids, mask_pos, self.max_len), mask_pos=mask_pos)


class NativeTextMasker:
    """`context_mask` + `template_weights` of one padded report in ONE native call (table pointers and output buffers are
    set up once; a report costs one ids conversion, the pre-drawn random numbers and two ctypes calls)."""

    def __init__(self, tables, max_len):
        ctypes, L = _native()
        self._ct, self._L, self._lib = ctypes, L, L.lib()
        self.tables, self.T = tables, int(max_len)
        self._vocab = int(min(tables.sub_u8.shape[0], tables.entity_u8.shape[0]))
        vp, i32 = ctypes.c_void_p, ctypes.c_int32
        self._lib.ecamp_text_mask_draw_count.argtypes = [vp, i32, vp, vp, i32]
        self._lib.ecamp_text_mask_and_weights.argtypes = [vp, i32, vp, vp, i32, vp, i32, vp, vp, vp, vp]
        self._sub, self._ent = tables.sub_u8.ctypes.data, tables.entity_u8.ctypes.data
        self._mask_pos = np.empty(self.T, dtype=np.int32)
        self._cnt = ctypes.c_int32(0)
        self._cnt_p = ctypes.addressof(self._cnt)

    def __call__(self, ids, rng=_random):
        a = np.ascontiguousarray(ids, dtype=np.int64)
        if a.shape[0] != self.T:
            raise ValueError(f"NativeTextMasker: expected {self.T} padded ids, got {a.shape[0]}")
        n = self._lib.ecamp_text_mask_draw_count(a.ctypes.data, self.T, self._sub, self._ent, self._vocab)
        if n < 0:
            raise ValueError("NativeTextMasker: bad arguments")
        draws = np.array([rng.random() for _ in range(n)], dtype=np.float64)
        masked = np.empty(self.T, dtype=np.int64)
        weights = np.empty(self.T, dtype=np.float32)
        self._L.check(self._lib.ecamp_text_mask_and_weights(a.ctypes.data, self.T, self._sub, self._ent, self._vocab, draws.ctypes.data, n,
                                                            masked.ctypes.data, weights.ctypes.data, self._mask_pos.ctypes.data,
                                                            self._cnt_p), "ecamp_text_mask_and_weights")
        return a, masked, weights, self._mask_pos[:self._cnt.value].tolist()


class NativeReportTextPipeline(ReportTextPipeline):
    """Same outputs, same `random` stream position; the masking and re-weighting loops run in the native library."""

    def __init__(self, tokenizer_json, max_caption_length=256):
        super().__init__(tokenizer_json, max_caption_length)
        self.masker = NativeTextMasker(self.tables, max_caption_length)

    def __call__(self, report, llm_output, rng=_random):
        enc = self.tokenizer.encode(splice_report(report, llm_output, rng))
        labels, masked, weights, mask_pos = self.masker(enc.ids, rng)
        return dict(labels=labels, ids=masked, attention_mask=np.asarray(enc.attention_mask, dtype=np.int64),
                    type_ids=np.asarray(enc.type_ids, dtype=np.int64), weights=weights, mask_pos=mask_pos)
