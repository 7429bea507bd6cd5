"""`CollatorGRAM` -- the producer of the hot path's input tensors, text in, `[B, N, L]` ids / masks out.

Drop-in for reference `src/processor/Collator.py:150-222` (`__call__`) and `:342-450` (`encode_texts_split`), `:286-340`
(`encode_target_split`) for `item_id_type == "split"`; same constructor (`tokenizer, args, mode`), same output dict.  Works
with ANY tokenizer that offers `batch_encode_plus(texts, max_length=..., truncation=True)` returning `{"input_ids": ...}` --
the T5 SentencePiece tokenizer the reference uses (when a `spiece.model` is available) or the surrogate tokenizer of
`gram_b200/data.py`.  `tests/test_collator.py` feeds the reference's own class and this one the same texts and asserts equal
tensors.

Contract restated (the reference does it passage by passage with torch ops, here once per batch on arrays):
  * every passage is tokenised (truncation at 999 tokens, EOS kept), the two separator ids 1820 ('|') and 9175 ('▁|') are
    dropped, the rest is cut to `item_prompt_max_len`; a passage that lost its EOS to the cut gets EOS(1) as its last token
  * ids are 0 and masks False behind the last token
  * every user is padded with all-masked passages to `min(max passages in the batch, max_his) + 1` -- the reference's
    off-by-one: a batch whose longest user is below the history cap carries one extra, empty passage
  * L is trimmed to the longest valid passage of the batch
  * targets: same separator / EOS treatment, cut to `target_max_len`, trimmed, pads replaced by -100
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

SEPARATOR_IDS = (1820, 9175)
EOS, PAD = 1, 0


def _encode(tokenizer, texts: Sequence[str], max_length: int) -> List[List[int]]:
    """token ids of every text, unpadded (pads of a padded encoding are stripped through its attention mask)"""
    if not texts:
        return []
    enc = tokenizer.batch_encode_plus(list(texts), max_length=max_length, truncation=True)
    ids = enc["input_ids"]
    if hasattr(ids, "tolist"):
        am = enc["attention_mask"]
        ids, am = ids.tolist(), am.tolist()
        return [[t for t, m in zip(r, a) if m] for r, a in zip(ids, am)]
    return [list(r) for r in ids]


def _filter_cut(row: Sequence[int], max_len: int, padded_len: int):
    """(ids[max_len], n_valid): separators dropped, cut to max_len, EOS forced when the cut removed it.  `padded_len` is the
    length the reference pads the tokenizer output to before filtering (999 / the batch's longest target): the reference
    cuts the PADDED, filtered row, so a row shorter than max_len after padding stays shorter there -- callers pass a
    padded_len that makes this impossible for passages and handle targets themselves."""
    kept = [t for t in row if t not in SEPARATOR_IDS]
    n_sep = len(row) - len(kept)
    width = min(max_len, padded_len - n_sep)           # length of the reference's filtered + cut row
    out = np.zeros(max_len, dtype=np.int64)
    n = min(len(kept), width)
    out[:n] = kept[:n]
    if EOS not in out[:width]:
        out[width - 1] = EOS                           # reference: tmp_input_ids[-1] = 1
    return out, n, width


class CollatorGRAM:
    def __init__(self, tokenizer, args=None, mode: str = "train"):
        self.tokenizer = tokenizer
        self.args = args
        self.mode = mode
        self.item_prompt_max_len = int(getattr(args, "item_prompt_max_len", 128))
        self.target_max_len = int(getattr(args, "target_max_len", 32))
        self.max_item_num = int(getattr(args, "max_his", 20))
        self.item_id_type = getattr(args, "item_id_type", "split")
        self.hierarchical_id_type = getattr(args, "hierarchical_id_type", None)

    def __call__(self, batch):
        if self.item_id_type != "split":
            raise NotImplementedError("gram_b200.CollatorGRAM implements item_id_type == 'split' (the shipped GRAM configuration)")
        input_texts = [b["input"] for b in batch]
        output_texts = [b["output"] for b in batch]
        target = self.encode_target_split(output_texts)
        target_masks = target["attention_mask"].bool()
        target_ids = target["input_ids"].masked_fill(~target_masks, -100)
        item_text_ids, item_text_masks = self.encode_texts_split(input_texts, self.tokenizer)
        return {"target_ids": target_ids, "target_masks": target_masks, "item_text_ids": item_text_ids,
                "item_text_masks": item_text_masks, "neg_item_ids": None, "neg_item_masks": None,
                "user_ids": [b["user_id"] for b in batch]}

    # reference Collator.py:342-450
    def encode_texts_split(self, batch_item_texts, tokenizer):
        L = self.item_prompt_max_len
        max_item_batch = max(len(t) for t in batch_item_texts)
        N = min(max_item_batch, self.max_item_num) + 1            # one for the coarse-grained user prompt
        B = len(batch_item_texts)
        ids = np.zeros((B, N, L), dtype=np.int64)
        mask = np.zeros((B, N, L), dtype=bool)
        flat = [p for passages in batch_item_texts for p in passages]
        rows = _encode(tokenizer, flat, 999)
        k = 0
        for b, passages in enumerate(batch_item_texts):
            if len(passages) > N:
                raise RuntimeError(f"user {b} has {len(passages)} passages, more than max_his + 1 = {N} (the reference fails in torch.cat)")
            for p in range(len(passages)):
                out, n, width = _filter_cut(rows[k], L, 999)
                k += 1
                if width < L:
                    raise RuntimeError("a passage holds more than 871 separator tokens (the reference fails in torch.stack)")
                ids[b, p] = out
                mask[b, p, :n] = True
        longest = int(mask.sum(-1).max())
        return torch.from_numpy(ids[:, :, :longest].copy()), torch.from_numpy(mask[:, :, :longest].copy())

    # reference Collator.py:286-340
    def encode_target_split(self, batch_output_texts):
        rows = _encode(self.tokenizer, batch_output_texts, 99)
        padded = max(len(r) for r in rows)                          # padding="longest"
        max_len = self.target_max_len
        outs, lens = [], []
        for r in rows:
            out, n, width = _filter_cut(r, max_len, padded)
            if width < max_len and EOS not in out[:width]:          # unreachable with a tokenizer that appends EOS
                raise RuntimeError("target without EOS shorter than target_max_len (the reference fails in torch.stack)")
            outs.append(out)
            lens.append(n)
        ids = np.stack(outs)
        am = np.arange(max_len)[None, :] < np.asarray(lens)[:, None]
        w = int(am.sum(-1).max())
        return {"input_ids": torch.from_numpy(ids[:, :w].copy()), "attention_mask": torch.from_numpy(am[:, :w].astype(np.int64))}
