"""Host-side input construction for the eval loop: surrogate tokenizer, leave-one-out test split and
the GRAM collator contract, over the compact fixtures in `gram_b200/assets/` (made from the shipped
`rec_datasets` splits by `scripts/make_dataset_fixture.py`).

What it mirrors in the reference (this is the *producer* of the hot path's input tensors, SURVEY.md
section 8(a) row 1 -- host Python there too, outside the `generate` timing):

  leave-one-out test split     src/data/test_dataset_gram.py:94-125 (target = last item, history =
                               previous <= max_his items, most recent first)
  validation split             src/data/test_dataset_gram.py:144-172 (target = items[-2])
  passages per user            src/data/test_dataset_gram.py:199-212: one user prompt
                               "What would user purchase after <lexids joined by ' ; '> ?" followed by
                               one passage per history item
  item passage text            src/utils/indexing.py:209-211,315-320:
                               "item: <lexid>; similar items: <k lexids joined by ', '>; <metadata>"
  collation                    src/processor/Collator.py:342-450: drop separator ids, truncate to
                               item_prompt_max_len, force EOS, zero-pad; passage count
                               min(max_in_batch, max_his) + 1; trim L to the longest valid passage
  candidate encoding           src/runner/single_runner_gram.py:594-617: [0] + pieces + [1]

What is NOT available offline (SURVEY.md "facts"): the SentencePiece model and `item_plain_text.txt`.
Hence the SURROGATE TOKENIZER (piece -> id by first appearance in the ID file, ids from 2, skipping
the two separator ids 1820/9175 the collator filters) and SYNTHETIC METADATA (hash-seeded filler
tokens up to the passage length, so item passages are full length as real ones are).  Every report
built on this module must say "surrogate tokenizer, synthetic metadata".
"""
from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np

from .synth import hash_u64

ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")
SEPARATOR_IDS = (1820, 9175)          # '|' and '▁|' in the T5 vocabulary (Collator.py:366)

# per-dataset launch parameters (reference command/train_gram_<dataset>.sh)
DATASET_PARAMS = {
    "Beauty": dict(top_k_similar=10, max_his=20, item_prompt_max_len=128),
    "Toys": dict(top_k_similar=5, max_his=20, item_prompt_max_len=128),
    "Sports": dict(top_k_similar=10, max_his=20, item_prompt_max_len=128),
    "Yelp": dict(top_k_similar=10, max_his=20, item_prompt_max_len=128),
}

_WORDS = ["item:", "similar items:", ";", ",", "?", "What", "would", "user", "purchase", "after"]


class SurrogateTokenizer:
    """piece <-> id by first appearance; 0 = pad, 1 = EOS; ids never collide with the separators."""

    pad_token_id = 0
    eos_token_id = 1

    def __init__(self, pieces: Sequence[str], vocab_size: int = 32128):
        self.vocab_size = vocab_size
        self.pieces = list(pieces)
        ids, nxt = [], 2
        for _ in range(len(self.pieces) + len(_WORDS)):
            while nxt in SEPARATOR_IDS:
                nxt += 1
            ids.append(nxt)
            nxt += 1
        if nxt >= vocab_size - 28:
            raise ValueError("surrogate vocabulary does not fit below the sentinel ids")
        self.piece_id = np.asarray(ids[:len(self.pieces)], dtype=np.int64)
        self.word_id = {w: ids[len(self.pieces) + i] for i, w in enumerate(_WORDS)}
        self.first_free = nxt
        self._id2piece = {int(i): p for i, p in zip(self.piece_id, self.pieces)}
        self._piece2id = None
        self._decode_cache = {}

    def encode(self, text: str) -> List[int]:
        """Lexical-id string ('|piece|piece...' or pieces separated by '|') -> ids + EOS, with the
        separator id between pieces exactly where the T5 tokenizer would emit it (callers filter it,
        as the reference does)."""
        if self._piece2id is None:
            self._piece2id = {p: int(i) for p, i in zip(self.pieces, self.piece_id)}
        out: List[int] = []
        for p in text.split("|"):
            if p == "":
                continue
            out.append(SEPARATOR_IDS[0])
            out.append(self._piece2id[p])
        out.append(self.eos_token_id)
        return out

    def encode_text(self, text: str) -> List[int]:
        """A whole passage as the collator sees it: known words, lexical ids ('|piece|piece', separator ids emitted as by
        `encode`), and raw ids written as '<123>' (the synthetic metadata); no EOS."""
        if self._piece2id is None:
            self._piece2id = {p: int(i) for p, i in zip(self.pieces, self.piece_id)}
        out: List[int] = []
        for word in text.replace("similar items:", "similar\x00items:").split():
            word = word.replace("\x00", " ")
            if word in self.word_id:
                out.append(self.word_id[word])
            elif word[0] == "|":
                for p in word.split("|"):
                    if p:
                        out.append(SEPARATOR_IDS[0])
                        out.append(self._piece2id[p])
            elif word[0] == "<" and word[-1] == ">":
                out.append(int(word[1:-1]))
            else:
                raise KeyError(f"surrogate tokenizer: unknown word {word!r}")
        return out

    def batch_encode_plus(self, texts, max_length=None, truncation=False, padding=None, pad_to_max_length=False,
                          return_tensors=None, **kw):
        """The slice of the HF tokenizer API the collators use (reference src/processor/Collator.py:354-360,287-293): EOS
        appended, truncation keeps it, optional padding with 0."""
        rows = []
        for t in texts:
            r = self.encode_text(t)
            if truncation and max_length is not None:
                r = r[:max_length - 1]
            rows.append(r + [self.eos_token_id])
        width = max_length if pad_to_max_length else (max(len(r) for r in rows) if padding == "longest" else None)
        if width is None:
            return {"input_ids": rows, "attention_mask": [[1] * len(r) for r in rows]}
        ids = [r + [0] * (width - len(r)) for r in rows]
        am = [[1] * len(r) + [0] * (width - len(r)) for r in rows]
        if return_tensors == "pt":
            import torch
            return {"input_ids": torch.tensor(ids), "attention_mask": torch.tensor(am)}
        return {"input_ids": ids, "attention_mask": am}

    def decode(self, ids, skip_special_tokens: bool = True) -> str:
        toks = []
        for i in ids:
            i = int(i)
            if i < 0:
                continue
            if skip_special_tokens and i in (0, 1):
                continue
            toks.append(self._id2piece.get(i, f"<{i}>"))
        return "".join(toks).replace("▁", " ").strip()

    def batch_decode(self, batch, skip_special_tokens: bool = True) -> List[str]:
        """Rows are decoded once and memoised: the eval loop decodes K predictions per user, all drawn from the
        same few thousand item ids."""
        if hasattr(batch, "tolist"):
            batch = batch.tolist()
        cache = self._decode_cache
        out = []
        for row in batch:
            key = (tuple(row), skip_special_tokens)
            text = cache.get(key)
            if text is None:
                text = self.decode(row, skip_special_tokens)
                if len(cache) < 1_000_000:
                    cache[key] = text
            out.append(text)
        return out


def find_t5_tokenizer(path: str = None):
    """The reference's tokenizer (`T5Tokenizer.from_pretrained(backbone)`, src/main_generative_gram.py) if a SentencePiece
    model exists on this machine: `path` / $GRAM_T5_TOKENIZER (a directory or a spiece.model file), the Hugging Face cache of
    t5-small / t5-base, or a copy under the reference tree.  Returns None when there is none (the offline image ships none):
    callers then use the surrogate tokenizer and say so.  With a real tokenizer the text path is
    `CollatorGRAM(tokenizer, args)(samples)` (gram_b200/collator.py), exactly as in the reference."""
    import glob
    cands = [path, os.environ.get("GRAM_T5_TOKENIZER")]
    home = os.path.expanduser("~/.cache/huggingface/hub")
    for name in ("t5-small", "t5-base", "google-t5--t5-small", "google-t5--t5-base"):
        cands += sorted(glob.glob(os.path.join(home, f"models--{name}", "snapshots", "*", "spiece.model")))
    cands += sorted(glob.glob("/root/reference/**/spiece.model", recursive=True))
    for c in cands:
        if not c:
            continue
        f = os.path.join(c, "spiece.model") if os.path.isdir(c) else c
        if os.path.isfile(f):
            from transformers import T5Tokenizer
            return T5Tokenizer(vocab_file=f)
    return None


class GramTestData:
    """Leave-one-out evaluation data of one dataset, tokenised with the surrogate tokenizer."""

    def __init__(self, dataset: str = "Beauty", mode: str = "test", max_his: int = None,
                 item_prompt_max_len: int = None, top_k_similar: int = None, vocab_size: int = 32128,
                 synthetic_users: int = 0, synthetic_history: int = 10, seed: int = 2023, packed_path: str = None):
        # packed_path: a cache written by gram_b200.formats.save_packed(pack_dataset(...)) from the reference's text files
        path = packed_path or os.path.join(ASSET_DIR, f"{dataset}.npz")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: run scripts/make_dataset_fixture.py in the build container")
        z = np.load(path)
        p = DATASET_PARAMS[dataset]
        self.dataset = dataset
        self.task = "sequential"
        self.mode = mode
        self.max_his = max_his if max_his is not None else p["max_his"]
        self.L = item_prompt_max_len or p["item_prompt_max_len"]
        self.top_k = top_k_similar if top_k_similar is not None else p["top_k_similar"]
        self.vocab_size = vocab_size
        self.tokenizer = SurrogateTokenizer(z["pieces"].tolist(), vocab_size)
        self.item_asin = z["item_asin"]
        self.item_lex = z["item_lex"]
        self.n_items = len(self.item_lex)
        self.similar = z["similar"] if "similar" in z.files else None
        # token ids of every lexical id (no separators, no EOS)
        self.item_tok = [self.tokenizer.piece_id[r[r >= 0]] for r in self.item_lex]
        self.all_items = ["|" + "|".join(self.tokenizer.pieces[j] for j in r[r >= 0]) for r in self.item_lex]
        if "user_off" in z.files and synthetic_users == 0:
            self.user_off = z["user_off"].astype(np.int64)
            self.user_items = z["user_items"].astype(np.int64)
            self.synthetic = False
        else:
            # no user_sequence.txt -> synthetic users with a fixed history length,
            # items drawn Zipf(1.0) over the real item ids
            n = synthetic_users or 30000
            h = synthetic_history + (2 if mode == "validation" else 1)
            ranks = np.arange(1, self.n_items + 1, dtype=np.float64)
            cdf = np.cumsum(1.0 / ranks)
            cdf /= cdf[-1]
            u = (hash_u64(n * h, seed) >> np.uint64(11)).astype(np.float64) / float(1 << 53)
            self.user_items = np.searchsorted(cdf, u).astype(np.int64).clip(0, self.n_items - 1)
            self.user_off = np.arange(0, n * h + 1, h, dtype=np.int64)
            self.synthetic = True
        self.n_users = len(self.user_off) - 1
        self._passages = None

    def __len__(self):
        return self.n_users

    # ---- candidates / trie input (single_runner_gram.py:594-617, item_id_type == "split") -------------
    def encoded_candidates(self) -> List[List[int]]:
        out = []
        for cand in self.all_items:
            enc = [0]
            for tok in self.tokenizer.encode(cand):
                if tok in SEPARATOR_IDS:
                    continue
                enc.append(tok)
            out.append(enc)
        return out

    # ---- per-item passages (built once: 12k distinct item passages vs ~160k instances on Beauty) ------
    def item_passages(self):
        if self._passages is not None:
            return self._passages
        L, tk = self.L, self.tokenizer
        w = tk.word_id
        tab = np.zeros((self.n_items, L), dtype=np.int64)
        lens = np.zeros(self.n_items, dtype=np.int32)
        lo, hi = tk.first_free, self.vocab_size - 28
        filler = lo + (hash_u64(self.n_items * L, 0xF111E7) % np.uint64(hi - lo)).astype(np.int64).reshape(self.n_items, L)
        for i in range(self.n_items):
            row = [w["item:"]] + self.item_tok[i].tolist() + [w[";"], w["similar items:"]]
            if self.similar is not None and self.top_k > 0:
                sims = [s for s in self.similar[i, :self.top_k] if s >= 0]
                for j, s in enumerate(sims):
                    if j:
                        row.append(w[","])
                    row.extend(self.item_tok[s].tolist())
            row.append(w[";"])
            n = min(len(row), L - 1)
            tab[i, :n] = row[:n]
            tab[i, n:L - 1] = filler[i, n:L - 1]       # synthetic metadata up to full length
            tab[i, L - 1] = 1                           # truncated passage: last token forced to EOS
            lens[i] = L
        self._passages = (tab, lens)
        return self._passages

    def split(self, u: int):
        """(history item indices, most recent first; target item index)."""
        items = self.user_items[self.user_off[u]:self.user_off[u + 1]]
        if self.mode == "validation":
            target, hist = items[-2], items[:-2]
        else:
            target, hist = items[-1], items[:-1]
        if self.max_his > 0:
            hist = hist[-self.max_his:]
        return hist[::-1], int(target)

    def user_prompt(self, hist) -> np.ndarray:
        w = self.tokenizer.word_id
        row = [w["What"], w["would"], w["user"], w["purchase"], w["after"]]
        for j, it in enumerate(hist):
            if j:
                row.append(w[";"])
            row.extend(self.item_tok[it].tolist())
        row.append(w["?"])
        row = row[:self.L - 1] + [1]
        return np.asarray(row, dtype=np.int64)

    def collate(self, users: Sequence[int]):
        """-> dict(item_text_ids int64 [B,N,L], item_text_masks bool [B,N,L], target_ids (list of id
        lists [0]+pieces+[1]), target_items, user_ids)."""
        tab, lens = self.item_passages()
        hists, targets = zip(*(self.split(u) for u in users))
        max_in_batch = max(len(h) for h in hists) + 1
        N = min(max_in_batch, self.max_his) + 1
        B, L = len(users), self.L
        ids = np.zeros((B, N, L), dtype=np.int64)
        mask = np.zeros((B, N, L), dtype=bool)
        for b, h in enumerate(hists):
            up = self.user_prompt(h)
            ids[b, 0, :len(up)] = up
            mask[b, 0, :len(up)] = True
            k = min(len(h), N - 1)
            if k:
                ids[b, 1:1 + k] = tab[h[:k]]
                mask[b, 1:1 + k] = np.arange(L)[None, :] < lens[h[:k]][:, None]
        longest = int(mask.sum(-1).max())
        ids, mask = ids[:, :, :longest], mask[:, :, :longest]
        tgt = [[0] + self.item_tok[t].tolist() + [1] for t in targets]
        return dict(item_text_ids=ids, item_text_masks=mask, target_ids=tgt, target_items=list(targets),
                    user_ids=[f"u{u}" for u in users])

    # ---- the same users as TEXT, for the collator (reference TestDatasetGRAM.construct_sentence / get_item) -----------
    def item_text(self, i: int) -> str:
        """'item: <lexid> ; similar items: <lexids joined by ' , '> ; <metadata>' (src/utils/indexing.py:209-211,315-320);
        the synthetic metadata tokens are written as '<id>' words."""
        tab, _ = self.item_passages()
        parts = ["item:", self.all_items[i], ";", "similar items:"]
        n_row = 1 + len(self.item_tok[i]) + 2
        if self.similar is not None and self.top_k > 0:
            sims = [s for s in self.similar[i, :self.top_k] if s >= 0]
            for j, sidx in enumerate(sims):
                if j:
                    parts.append(",")
                    n_row += 1
                parts.append(self.all_items[sidx])
                n_row += len(self.item_tok[sidx])
        parts.append(";")
        n_row += 1
        parts += [f"<{int(t)}>" for t in tab[i, n_row:self.L - 1]]
        return " ".join(parts)

    def text_samples(self, users: Sequence[int]):
        """[{'input': [user sentence, item passage, ...], 'output': target lexid, 'user_id': ...}] -- what the reference's
        `TestDatasetGRAM.__getitem__` yields (src/data/test_dataset_gram.py:179-231) and `CollatorGRAM` consumes."""
        out = []
        for u in users:
            hist, target = self.split(u)
            sent = "What would user purchase after " + " ; ".join(self.all_items[h] for h in hist) + " ?"
            out.append({"input": [sent] + [self.item_text(h) for h in hist], "output": self.all_items[target],
                        "user_id": f"u{u}"})
        return out

    # ---- cached-item path: the item passages once, users as (prompt, item indices) ----------
    def item_table(self):
        """(ids int64 [n_items, L], mask bool [n_items, L]): every item passage, for `GRAM.cache_items`."""
        tab, lens = self.item_passages()
        return tab, np.arange(self.L)[None, :] < lens[:, None]

    def collate_cached(self, users: Sequence[int]):
        """The same batch as `collate`, factored for `GRAM.generate_cached`: prompt_ids / prompt_masks [B, L]
        (passage 0, padded to the table's L) and item_index int32 [B, N-1] (passage 1+j = item, -1 = the empty
        passages `collate` pads with)."""
        hists, targets = zip(*(self.split(u) for u in users))
        max_in_batch = max(len(h) for h in hists) + 1
        N = min(max_in_batch, self.max_his) + 1
        B, L = len(users), self.L
        ids = np.zeros((B, L), dtype=np.int64)
        mask = np.zeros((B, L), dtype=bool)
        items = np.full((B, N - 1), -1, dtype=np.int32)
        for b, h in enumerate(hists):
            up = self.user_prompt(h)
            ids[b, :len(up)] = up
            mask[b, :len(up)] = True
            k = min(len(h), N - 1)
            items[b, :k] = h[:k]
        tgt = [[0] + self.item_tok[t].tolist() + [1] for t in targets]
        return dict(prompt_ids=ids, prompt_masks=mask, item_index=items, target_ids=tgt, target_items=list(targets),
                    user_ids=[f"u{u}" for u in users])

    def valid_tokens(self, users: Sequence[int]) -> int:
        """Packed (valid) encoder tokens of these users: the S_valid of SURVEY 8(d)."""
        total = 0
        for u in users:
            h, _ = self.split(u)
            total += len(self.user_prompt(h)) + min(len(h), self.max_his) * self.L
        return total
