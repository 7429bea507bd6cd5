"""The input contract of the hot path: `gram_b200.collator.CollatorGRAM` against the reference's own
`CollatorGRAM` (src/processor/Collator.py:150-450, loaded from /root/reference) on the same texts through the same
tokenizer object, and the pre-tokenised fast path of `GramTestData.collate` against the text path through the collator."""
import importlib.util
import os
import zlib
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from helpers import ROOT  # noqa: F401
from gram_b200.collator import CollatorGRAM
from gram_b200.data import GramTestData

REF = os.environ.get("GRAM_REFERENCE_ROOT", "/root/reference")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "src", "processor", "Collator.py")),
                               reason="reference tree not mounted")


class ToyTokenizer:
    """Deterministic stand-in with the T5 tokenizer's calling convention: words -> ids by CRC, '|' -> 1820 inside a word and
    9175 at the start of one (the two separator ids the collator drops), EOS(1) appended, truncation keeps the EOS."""

    def _ids(self, text, max_length):
        out = []
        for word in text.split():
            parts = word.split("|")
            for i, p in enumerate(parts):
                if i > 0:
                    out.append(9175 if i == 1 and parts[0] == "" else 1820)
                if p:
                    t = 2 + zlib.crc32(p.encode()) % 31000
                    out.append(t + 1 if t in (1820, 9175) else t)
        out = out[:max_length - 1] + [1]
        return out

    def batch_encode_plus(self, texts, max_length=None, pad_to_max_length=False, padding=None, return_tensors=None,
                          truncation=False):
        rows = [self._ids(t, max_length or 10 ** 9) for t in texts]
        if pad_to_max_length:
            width = max_length
        elif padding == "longest":
            width = max(len(r) for r in rows)
        else:
            width = None
        if width is None:
            return {"input_ids": rows, "attention_mask": [[1] * len(r) for r in rows]}
        ids = [r + [0] * (width - len(r)) for r in rows]
        am = [[1] * len(r) + [0] * (width - len(r)) for r in rows]
        if return_tensors == "pt":
            return {"input_ids": torch.tensor(ids), "attention_mask": torch.tensor(am)}
        return {"input_ids": ids, "attention_mask": am}


def _reference_collator(tok, args):
    spec = importlib.util.spec_from_file_location("gram_ref_collator", os.path.join(REF, "src", "processor", "Collator.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.CollatorGRAM(tok, args, mode="test")


def _texts(rng, n_users, max_hist, long_every=3):
    words = ["item:", "similar", "items:", ";", ",", "?", "What", "would", "user", "purchase", "after", "nail", "red", "soap",
             "|▁butter|▁mango|generation", "|loss|▁brake", "a|b|c", "|", "x|", "serum"]
    batch = []
    for u in range(n_users):
        h = int(rng.integers(0, max_hist + 1))
        passages = ["What would user purchase after " + " ; ".join(rng.choice(words, 3).tolist()) + " ?"]
        for j in range(h):
            n = 400 if (u + j) % long_every == 0 else int(rng.integers(1, 60))      # some passages exceed 128 tokens
            passages.append("item: " + " ".join(rng.choice(words, n).tolist()))
        batch.append({"input": passages, "output": " ".join(rng.choice(words, int(rng.integers(1, 6))).tolist()),
                      "user_id": f"U{u}"})
    return batch


@needs_ref
@pytest.mark.parametrize("max_his,item_len,target_len,seed", [(20, 128, 32, 0), (5, 64, 8, 1), (3, 16, 4, 2)])
def test_collator_equals_reference_collator(max_his, item_len, target_len, seed):
    args = SimpleNamespace(item_prompt_max_len=item_len, target_max_len=target_len, max_his=max_his, item_id_type="split",
                           hierarchical_id_type="hierarchy_v1")
    tok = ToyTokenizer()
    ref = _reference_collator(tok, args)
    ours = CollatorGRAM(tok, args, mode="test")
    rng = np.random.default_rng(seed)
    for n_users, hist_cap in ((1, 0), (4, max_his), (7, max(1, max_his // 2)), (3, 1)):
        batch = _texts(rng, n_users, hist_cap)
        want, got = ref(batch), ours(batch)
        for k in ("item_text_ids", "item_text_masks", "target_ids", "target_masks"):
            assert want[k].dtype == got[k].dtype and torch.equal(want[k], got[k]), k
        assert got["user_ids"] == want["user_ids"] and got["neg_item_ids"] is None
        # the contract the engine relies on: valid tokens form a prefix ending in EOS, ids are 0 behind it
        ids, mask = got["item_text_ids"], got["item_text_masks"]
        lens = mask.sum(-1)
        assert torch.equal(mask, torch.arange(mask.shape[-1])[None, None, :] < lens[..., None])
        assert bool(((lens == 0) | (ids.gather(-1, (lens - 1).clamp_min(0)[..., None])[..., 0] == 1)).all())
        assert not ids[~mask].any()


def test_pretokenised_collate_equals_text_path_through_the_collator():
    """`GramTestData.collate` (item passages tokenised once, the eval loop's fast path) == the same users rendered as TEXT
    passages (`GramTestData.text_samples`) and pushed through `CollatorGRAM` with the dataset's tokenizer."""
    data = GramTestData("Beauty")
    args = SimpleNamespace(item_prompt_max_len=data.L, target_max_len=32, max_his=data.max_his, item_id_type="split",
                           hierarchical_id_type="hierarchy_v1")
    coll = CollatorGRAM(data.tokenizer, args, mode="test")
    for users in ([0], [5, 17, 4000, 22362], list(range(100, 116))):
        fast = data.collate(users)
        slow = coll(data.text_samples(users))
        assert np.array_equal(fast["item_text_ids"], slow["item_text_ids"].numpy())
        assert np.array_equal(fast["item_text_masks"], slow["item_text_masks"].numpy())
        tgt = slow["target_ids"].numpy()
        for i, t in enumerate(fast["target_ids"]):
            # the candidate encoding carries the decoder start token in front (single_runner_gram.py:594-617)
            assert t[1:] == [x for x in tgt[i].tolist() if x >= 0]
        assert slow["user_ids"] == fast["user_ids"]


@needs_ref
def test_collator_with_the_real_t5_tokenizer():
    """With a SentencePiece model on the machine the same equality holds through the reference's actual tokenizer."""
    from gram_b200.data import find_t5_tokenizer
    tok = find_t5_tokenizer()
    if tok is None:
        pytest.skip("no spiece.model on this machine (offline image): the surrogate tokenizer is used and reports say so")
    args = SimpleNamespace(item_prompt_max_len=128, target_max_len=32, max_his=20, item_id_type="split",
                           hierarchical_id_type="hierarchy_v1")
    batch = [{"input": ["What would user purchase after |▁butter|▁mango|generation ; |▁loss|▁brake ?",
                        "item: |▁butter|▁mango|generation; similar items: |▁loss|▁brake; title: mango butter; brand: x " * 9,
                        "item: |▁loss|▁brake; similar items: ; title: brake pads"],
              "output": "|▁lend|▁obtained|▁said|▁kernel", "user_id": "A2CG5Y82ZZNY6W"},
             {"input": ["What would user purchase after |▁loss|▁brake ?", "item: |▁loss|▁brake; similar items: ; title: brake pads"],
              "output": "|▁butter|▁mango", "user_id": "A1YJEY40YUW4SE"}]
    want, got = _reference_collator(tok, args)(batch), CollatorGRAM(tok, args, mode="test")(batch)
    for k in ("item_text_ids", "item_text_masks", "target_ids", "target_masks"):
        assert torch.equal(want[k], got[k]), k
