"""Eval loop of the hot path -- host-side mirror of the reference runners' `test_dataset_task`.

Reference: `src/runner/single_runner_gram.py:570-719` (single GPU) and
`src/runner/distributed_runner_gram.py:685-874` (one process per GPU, `DistributedSampler`, metric
all-reduce, per-rank TSV merge through the file system).  Kept from the reference:

  * the candidate trie is built once from all item ids: `[0] + tokenizer.encode(lexid)` minus the two
    separator ids (`single:594-617`), `max_length = max(len(candidate))` (`single:633-636`)
  * `model_rec.generate(input_ids=..., attention_mask=..., max_length=..., prefix_allowed_tokens_fn=...,
    num_beams=G, num_return_sequences=G, output_scores=True, return_dict_in_generate=True,
    length_penalty=...)` with `G = max(max k in metrics, beam_size)` (`single:38-39,641-651`)
  * predictions and gold are DECODED TO STRINGS and compared as strings (`single:657-666`), metrics by
    `evaluate.rel_results` / `get_metrics_results`, summed then divided by the user count
    (`single:664-673,699-702`); optional TSV with the reference's columns (`single:580-588,675-694`)

Changed on purpose: the reference evaluates ONE user per `generate` call
(`eval_batch_size` 1) -- here a whole batch of users goes through one call; multi-GPU shards users
contiguously and exactly (the reference's `DistributedSampler` pads with duplicated users and shuffles),
gathers the ranked lists with ONE collective after the loop, and reduces integer hit ranks, so metrics
are identical at every world size.  No collective sits inside the data path.
"""
from __future__ import annotations

import logging
import os
import queue
import threading
import time
from types import SimpleNamespace
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import evaluate, formats
from . import generation_trie as gt
from .data import SEPARATOR_IDS


def shard_range(n: int, rank: int, world: int):
    """Contiguous, exact shard [lo, hi) of n users (no padding, no duplicates)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GramEvalLoader:
    """Minimal stand-in for the reference's test DataLoader: iterates collated batches of one rank's
    users and exposes `.dataset` (with `.all_items`, `.dataset`, `.task`) like the reference's."""

    def __init__(self, data, batch_size: int, rank: int = 0, world: int = 1, users: Optional[Sequence[int]] = None,
                 sort_by_length: bool = True, item_cache: bool = False):
        self.dataset = data
        self.batch_size = batch_size
        # item_cache: batches are (prompt, history item indices) for GRAM.generate_cached -- every item passage is
        # encoded once per eval instead of once per occurrence; same predictions bit for bit
        self.item_cache = item_cache
        all_users = list(users) if users is not None else list(range(data.n_users))
        lo, hi = shard_range(len(all_users), rank, world)
        mine = all_users[lo:hi]
        if sort_by_length:
            # users with similar history lengths share a batch: fewer all-masked passages, balanced CTAs
            mine = sorted(mine, key=lambda u: (len(data.split(u)[0]), u))
        self.users = mine

    def __len__(self):
        return (len(self.users) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        for i in range(0, len(self.users), self.batch_size):
            idx = self.users[i:i + self.batch_size]
            if self.item_cache:
                b = self.dataset.collate_cached(idx)
                for k in ("prompt_ids", "prompt_masks", "item_index"):
                    b[k] = torch.from_numpy(b[k])
            else:
                b = self.dataset.collate(idx)
                b["item_text_ids"] = torch.from_numpy(b["item_text_ids"])
                b["item_text_masks"] = torch.from_numpy(b["item_text_masks"])
            b["user_index"] = idx
            yield b


class _ItemStrings:
    """String comparison of predictions and gold items (reference single_runner_gram.py:657-666: both sides are DECODED and
    compared as strings) without decoding 20 rows per user in Python: every candidate id row is decoded once, rows are
    looked up by their bytes, and two rows are "equal" iff their decoded strings are (items whose strings coincide share a
    class).  Rows that are not candidates (cannot happen under the trie constraint, except -inf fillers) fall back to the
    tokenizer."""

    def __init__(self, tokenizer, encoded, width):
        self.tokenizer, self.width = tokenizer, width
        tab = np.zeros((len(encoded), width), dtype=np.int32)
        for i, c in enumerate(encoded):
            tab[i, :len(c)] = c
        strings = tokenizer.batch_decode(tab, skip_special_tokens=True)
        cls, self.class_string = {}, []
        item_class = np.empty(len(encoded), dtype=np.int64)
        for i, st in enumerate(strings):
            k = cls.get(st)
            if k is None:
                k = cls[st] = len(self.class_string)
                self.class_string.append(st)
            item_class[i] = k
        self._by_string = cls
        keys = np.ascontiguousarray(tab).view(np.dtype((np.void, width * 4))).ravel()
        order = np.argsort(keys)
        self._keys, self._class = keys[order], item_class[order]

    def classes(self, rows):
        """class id of every id row [n, <= width] (int), or a fresh negative id per distinct unknown string"""
        rows = np.asarray(rows)
        pad = np.zeros((rows.shape[0], self.width), dtype=np.int32)
        pad[:, :rows.shape[1]] = rows[:, :self.width]
        keys = np.ascontiguousarray(pad).view(np.dtype((np.void, self.width * 4))).ravel()
        pos = np.searchsorted(self._keys, keys).clip(0, len(self._keys) - 1)
        hit = self._keys[pos] == keys
        out = np.where(hit, self._class[pos], -1)
        if not hit.all():
            miss = np.nonzero(~hit)[0]
            for i, st in zip(miss, self.tokenizer.batch_decode(pad[miss], skip_special_tokens=True)):
                k = self._by_string.get(st)
                if k is None:
                    k = self._by_string[st] = -2 - len(self._by_string)      # a string no item has
                out[i] = k
        return out


def rel_rows_fast(pred_class, gold_class, scores, k):
    """`evaluate.rel_results` on class ids: per user the k predictions re-ordered by score (descending, stable, NaN last) and
    compared with the gold item -> uint8 [n, k]"""
    n = len(gold_class)
    sc = np.asarray(scores, dtype=np.float64).reshape(n, k)
    order = np.argsort(-sc, axis=1, kind="stable")
    pc = np.take_along_axis(np.asarray(pred_class).reshape(n, k), order, axis=1)
    return (pc == np.asarray(gold_class)[:, None]).astype(np.uint8)


class GramRunner:
    """`test_dataset_task` / `test` / `validate` of the reference runners for the inference path."""

    def __init__(self, model_rec, tokenizer, device, args=None, rank: int = 0, world_size: int = 1):
        self.model_rec = model_rec
        self.tokenizer = tokenizer
        self.device = device
        self.args = args or SimpleNamespace()
        self.rank, self.world_size = rank, world_size
        metrics = getattr(self.args, "metrics", "hit@5,hit@10,ndcg@5,ndcg@10")
        self.metrics = metrics.split(",") if isinstance(metrics, str) else list(metrics)
        beam = int(getattr(self.args, "beam_size", 20))
        self.generate_num = max(max(int(m.split("@")[1]) for m in self.metrics), beam)     # single:38-39
        self.length_penalty = float(getattr(self.args, "length_penalty", 1.0))
        self.item_id_type = getattr(self.args, "item_id_type", "split")
        self.save_predictions = bool(getattr(self.args, "save_predictions", False))
        self.pred_dir = getattr(self.args, "pred_dir", "../preds")

    # ---- trie input (single:594-617) ------------------------------------------------------------------
    def encode_candidates(self, candidates: Sequence[str]) -> List[List[int]]:
        enc = []
        for cand in candidates:
            row = [0]
            for tok in self.tokenizer.encode(cand):
                if self.item_id_type == "split" and tok in SEPARATOR_IDS:
                    continue
                row.append(tok)
            enc.append(row)
        return enc

    def _generate(self, batch, max_length, prefix_fn):
        model = getattr(self.model_rec, "module", self.model_rec)
        on_gpu = self.device is not None and torch.device(self.device).type == "cuda"
        if "prompt_ids" in batch:
            args = [batch[k].to(self.device, non_blocking=True) if on_gpu else batch[k]
                    for k in ("prompt_ids", "prompt_masks", "item_index")]
            return model.generate_cached(*args, max_length=max_length, prefix_allowed_tokens_fn=prefix_fn,
                                         num_beams=self.generate_num, num_return_sequences=self.generate_num,
                                         return_dict_in_generate=True, length_penalty=self.length_penalty)
        ids = batch["item_text_ids"]
        mask = batch["item_text_masks"]
        if self.device is not None and torch.device(self.device).type == "cuda":
            ids, mask = ids.to(self.device, non_blocking=True), mask.to(self.device, non_blocking=True)
        return model.generate(input_ids=ids, attention_mask=mask, max_length=max_length,
                              prefix_allowed_tokens_fn=prefix_fn, num_beams=self.generate_num,
                              num_return_sequences=self.generate_num, output_scores=True,
                              return_dict_in_generate=True, length_penalty=self.length_penalty)

    def test_dataset_task(self, testloader, mode: str = "test", pipeline: bool = True):
        """Eval loop.  With `pipeline`, batch collation runs one or two batches ahead in a background thread and
        decoding / metric bookkeeping of batch i runs while the GPU works on batch i+1 (the C-ABI call releases the
        GIL), so the loop is bound by `generate`, not by host Python."""
        data = testloader.dataset
        logging.info(f"[{mode}] testing {data.dataset} dataset on {data.task} task")
        G = self.generate_num
        # candidate encoding, trie and the item-string table depend on the item list only: built once per item list and
        # kept (the reference rebuilds them at every call, single:594-619; repeated validation / test calls reuse them, and
        # the engine keeps the uploaded CSR arrays as long as the trie object is the same)
        key = (id(data.all_items), len(data.all_items), self.item_id_type)
        cached = getattr(self, "_candidates", None)
        if cached is None or cached[0] != key:
            encoded = self.encode_candidates(data.all_items)
            candidate_trie = gt.Trie(encoded)
            max_length = max(len(c) for c in encoded)
            cached = (key, encoded, candidate_trie, gt.prefix_allowed_tokens_fn(candidate_trie), max_length,
                      _ItemStrings(self.tokenizer, encoded, max_length), data.all_items)
            self._candidates = cached
        _, encoded, candidate_trie, prefix_fn, max_length, strings, _ = cached
        if getattr(testloader, "item_cache", False):
            model = getattr(self.model_rec, "module", self.model_rec)
            if getattr(model, "_item_table_owner", None) is not data:
                model.cache_items(*data.item_table())
                model._item_table_owner = data
        # per-batch results as arrays (what the multi-GPU gather moves): user index, gold ids, ranked ids, scores and the
        # 0/1 relevance row of evaluate.rel_results (string comparison, as the reference)
        parts = []                               # (users [n], gold [n, ML], seqs [n, G, ML], scores [n, G], rel [n, G])

        def post(batch, seqs, scores):
            n = len(batch["user_index"])
            g_ids = np.zeros((n, max_length), dtype=np.int32)
            for i, t in enumerate(batch["target_ids"]):
                t = [int(x) for x in t if int(x) >= 0]
                if t and t[0] != 0 and len(t) < max_length:
                    t = [0] + t                           # collator targets carry no decoder start token; candidates do
                t = t[:max_length]
                g_ids[i, :len(t)] = t
            s_ids = np.zeros((n, G, max_length), dtype=np.int32)
            sq = seqs.numpy() if hasattr(seqs, "numpy") else np.asarray(seqs)
            s_ids[:, :, :sq.shape[1]] = sq.reshape(n, G, -1)[:, :, :max_length]
            sc = (scores.numpy() if hasattr(scores, "numpy") else np.asarray(scores)).astype(np.float32).reshape(n, G)
            # decoded-string equality through the item table (one decode per item, not per prediction)
            rel = rel_rows_fast(strings.classes(s_ids.reshape(n * G, max_length)), strings.classes(g_ids), sc, G)
            parts.append((np.asarray(batch["user_index"], dtype=np.int64), g_ids, s_ids, sc, rel))

        def batches():
            if not pipeline:
                yield from testloader
                return
            q = queue.Queue(maxsize=2)
            stop = object()

            def produce():
                try:
                    for b in testloader:
                        q.put(b)
                    q.put(stop)
                except BaseException as e:      # surface loader errors in the consumer
                    q.put(e)

            threading.Thread(target=produce, daemon=True).start()
            while True:
                b = q.get()
                if b is stop:
                    return
                if isinstance(b, BaseException):
                    raise b
                yield b

        total_time = 0.0
        pending = None                           # post-processing thread of the previous batch
        with torch.no_grad():
            for batch in batches():
                t0 = time.time()
                pred = self._generate(batch, max_length, prefix_fn)
                seqs = pred["sequences"].cpu()
                scores = pred["sequences_scores"].cpu()
                total_time += time.time() - t0
                if pending is not None:
                    pending.join()
                if pipeline:
                    pending = threading.Thread(target=post, args=(batch, seqs, scores))
                    pending.start()
                else:
                    post(batch, seqs, scores)
        if pending is not None:
            pending.join()
        # ---- after the loop: ONE all-gather of the ranked lists + an integer all-reduce of the hit counts
        #      (reference: all_reduce(metrics_res), all_reduce(test_total), TSV merge through the file system;
        #      distributed_runner_gram.py:832-838,854-874) ----
        t0 = time.time()
        local_rel = np.concatenate([p[4] for p in parts]) if parts else np.zeros((0, G), np.uint8)
        hist = self._reduce_hit_histogram(self._first_hit(local_rel), G)          # from each rank's OWN users
        users, gold_ids, seq_ids, scores, rel = self._gather(parts, G, max_length)
        gather_seconds = time.time() - t0
        order = np.argsort(users, kind="stable")
        users, gold_ids, seq_ids, scores, rel = users[order], gold_ids[order], seq_ids[order], scores[order], rel[order]
        test_total = len(users)
        # metrics from the relevance ROWS (a row may hold more than one 1 when two token paths decode to the same
        # string; the reference's ndcg_at_k sums every match), in user order on every rank: identical at any world size
        sums = evaluate.get_metrics_results(rel.tolist(), self.metrics) if test_total else np.zeros(len(self.metrics))
        metrics_res = sums / max(test_total, 1)
        first = self._first_hit(rel)
        if int(hist[-1]) != test_total or not np.array_equal(hist[:G + 1], np.bincount(first + 1, minlength=G + 1)):
            raise RuntimeError("multi-GPU eval: the all-reduced hit-rank histogram disagrees with the gathered rankings")
        result = dict(metrics={m: float(v) for m, v in zip(self.metrics, metrics_res)}, test_total=test_total,
                      generate_seconds=total_time, gather_seconds=gather_seconds, hit_ranks=first, rel_rows=rel,
                      hit_rank_histogram=hist[:G + 1], users=users, sequences=seq_ids, sequences_scores=scores, rows=None)
        if self.rank == 0:
            gold = self.tokenizer.batch_decode(gold_ids, skip_special_tokens=True)
            sents = self.tokenizer.batch_decode(seq_ids.reshape(-1, max_length), skip_special_tokens=True)
            result["rows"] = [(int(u), gold[i], sents[i * G:(i + 1) * G], scores[i].tolist(), int(first[i]))
                              for i, u in enumerate(users)]
            for m, v in result["metrics"].items():
                logging.info(f"{mode} {m}: {v}")
            logging.info(f"Total inference time: {total_time:.2f}s for {len(testloader)} batches")
            if self.save_predictions:
                result["pred_file"] = self._write_tsv(result["rows"], data, mode, result["metrics"], rel)
        return result

    # ---- the path's only collectives -----------------------------------------------------------------------------
    def _comm_device(self):
        import torch.distributed as dist
        if dist.get_backend() == "nccl":
            return torch.device(self.device) if self.device is not None else torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    def _gather(self, parts, G, max_length):
        """(users int64 [n], gold int32 [n, ML], seqs int32 [n, G, ML], scores fp32 [n, G], rel uint8 [n, G]) of ALL ranks.
        world > 1: every rank packs its users into one int32 matrix (ids | scores bit-cast | relevance | user index),
        pads it to the largest shard and ONE `all_gather_into_tensor` (ncclAllGather under NCCL) moves
        it; no pickling, no file system."""
        ML = max_length
        if parts:
            users = np.concatenate([p[0] for p in parts])
            gold = np.concatenate([p[1] for p in parts])
            seqs = np.concatenate([p[2] for p in parts])
            scores = np.concatenate([p[3] for p in parts])
            rel = np.concatenate([p[4] for p in parts])
        else:
            users, gold = np.zeros(0, np.int64), np.zeros((0, ML), np.int32)
            seqs, scores, rel = np.zeros((0, G, ML), np.int32), np.zeros((0, G), np.float32), np.zeros((0, G), np.uint8)
        if self.world_size <= 1:
            return users, gold, seqs, scores, rel
        import torch.distributed as dist
        dev = self._comm_device()
        n = len(users)
        width = ML + G * ML + G + G + 2
        mine = np.zeros((n, width), dtype=np.int32)
        o = 0
        mine[:, o:o + ML] = gold; o += ML
        mine[:, o:o + G * ML] = seqs.reshape(n, G * ML); o += G * ML
        mine[:, o:o + G] = scores.view(np.int32); o += G
        mine[:, o:o + G] = rel; o += G
        mine[:, o] = (users & 0x7fffffff).astype(np.int32)
        mine[:, o + 1] = (users >> 31).astype(np.int32)
        counts = torch.zeros(self.world_size, dtype=torch.int64, device=dev)
        counts[self.rank] = n
        dist.all_reduce(counts)                                   # shard sizes (a loader may carry a custom user list)
        counts = counts.cpu().tolist()
        n_pad = max(max(counts), 1)
        send = torch.zeros((n_pad, width), dtype=torch.int32, device=dev)
        send[:n] = torch.from_numpy(mine).to(dev)
        recv = torch.empty((self.world_size * n_pad, width), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(recv, send)
        recv = recv.cpu().numpy().reshape(self.world_size, n_pad, width)
        allr = np.concatenate([recv[r, :counts[r]] for r in range(self.world_size)])
        o = 0
        gold = allr[:, o:o + ML].copy(); o += ML
        seqs = allr[:, o:o + G * ML].reshape(-1, G, ML).copy(); o += G * ML
        scores = np.ascontiguousarray(allr[:, o:o + G]).view(np.float32); o += G
        rel = allr[:, o:o + G].astype(np.uint8); o += G
        users = allr[:, o].astype(np.int64) | (allr[:, o + 1].astype(np.int64) << 31)
        return users, gold, seqs, scores, rel

    @staticmethod
    def _first_hit(rel):
        """rank of the first 1 of every relevance row, -1 when the gold item is absent"""
        if len(rel) == 0:
            return np.zeros(0, np.int64)
        return np.where(rel.any(axis=1), rel.argmax(axis=1), -1).astype(np.int64)

    def _reduce_hit_histogram(self, first_hit, G):
        """int64 [G + 2]: users per first-hit rank (bin 0 = no hit, bin r + 1 = gold at rank r) and the user count, summed
        over the ranks from each rank's OWN users -- the integer counterpart of the reference's all_reduce(metrics_res) /
        all_reduce(test_total); it cross-checks the gathered rankings."""
        h = np.concatenate([np.bincount(first_hit + 1, minlength=G + 1), [len(first_hit)]]).astype(np.int64)
        if self.world_size <= 1:
            return h
        import torch.distributed as dist
        t = torch.from_numpy(h).to(self._comm_device())
        dist.all_reduce(t)
        return t.cpu().numpy()

    def _write_tsv(self, rows, data, mode, metrics, rel):
        os.makedirs(self.pred_dir, exist_ok=True)
        stamp = time.strftime("%Y%m%d_%H%M%S")
        path = os.path.join(self.pred_dir, f"{stamp}_{data.dataset}_{data.task}_pred_{mode}.tsv")
        per_user = [evaluate.get_metrics_results([r.tolist()], self.metrics) for r in rel]
        formats.write_predictions_tsv(path, rows, metrics, per_user, self.metrics)
        return path

    # reference entry points (single:370-408): checkpoints are state dicts
    def test(self, path: Optional[str] = None, loader=None):
        if path:
            self.model_rec.load_state_dict(torch.load(path, map_location="cpu"))
        return self.test_dataset_task(loader, mode="test")

    def test_from_model(self, loader):
        return self.test_dataset_task(loader, mode="test")

    def validate(self, loader):
        return self.test_dataset_task(loader, mode="validation")


def get_runner(model_rec, tokenizer, device, args=None, rank: int = 0, world_size: int = 1):
    """reference `src/runner/__init__.py:12-58` (inference side)."""
    return GramRunner(model_rec, tokenizer, device, args, rank, world_size)
