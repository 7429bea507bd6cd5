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
        encoded = self.encode_candidates(data.all_items)
        candidate_trie = gt.Trie(encoded)
        prefix_fn = gt.prefix_allowed_tokens_fn(candidate_trie)
        max_length = max(len(c) for c in encoded)
        if getattr(testloader, "item_cache", False):
            model = getattr(self.model_rec, "module", self.model_rec)
            if getattr(model, "_item_table_owner", None) is not data:
                model.cache_items(*data.item_table())
                model._item_table_owner = data
        rows = []                                # (user index, gold string, predictions, scores, hit rank)

        def post(batch, seqs, scores):
            gold = self.tokenizer.batch_decode(batch["target_ids"], skip_special_tokens=True)
            sents = self.tokenizer.batch_decode(seqs, skip_special_tokens=True)
            rel = evaluate.rel_results(sents, gold, scores, G)
            for i, u in enumerate(batch["user_index"]):
                r = rel[i]
                rows.append((u, gold[i], sents[i * G:(i + 1) * G], scores[i * G:(i + 1) * G].tolist(),
                             r.index(1) if 1 in r else -1))

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
        # ---- one gather after the loop (reference: all_reduce(metrics), all_reduce(total), TSV merge) ----
        rows = self._gather(rows)
        rows.sort(key=lambda r: r[0])
        ranks = np.array([r[4] for r in rows], dtype=np.int64)
        test_total = len(rows)
        sums = evaluate.metric_sums_from_ranks(ranks, self.metrics)
        metrics_res = sums / max(test_total, 1)
        result = dict(metrics={m: float(v) for m, v in zip(self.metrics, metrics_res)}, test_total=test_total,
                      generate_seconds=total_time, hit_ranks=ranks, rows=rows if self.rank == 0 else None)
        if self.rank == 0:
            for m, v in result["metrics"].items():
                logging.info(f"{mode} {m}: {v}")
            logging.info(f"Total inference time: {total_time:.2f}s for {len(testloader)} batches")
            if self.save_predictions:
                result["pred_file"] = self._write_tsv(rows, data, mode, result["metrics"])
        return result

    def _gather(self, rows):
        if self.world_size <= 1:
            return rows
        import torch.distributed as dist
        gathered = [None] * self.world_size
        dist.all_gather_object(gathered, rows)
        out = []
        for part in gathered:
            out.extend(part)
        return out

    def _write_tsv(self, rows, data, mode, metrics):
        os.makedirs(self.pred_dir, exist_ok=True)
        stamp = time.strftime("%Y%m%d_%H%M%S")
        path = os.path.join(self.pred_dir, f"{stamp}_{data.dataset}_{data.task}_pred_{mode}.tsv")
        per_user = [evaluate.metric_sums_from_ranks(np.array([r[4]]), self.metrics) for r in rows]
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
