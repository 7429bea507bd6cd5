"""Ranking metrics of the eval loop.

API-compatible with reference `src/utils/evaluate.py:5-58` (`rel_results`, `get_metrics_results`,
`hit_at_k`, `ndcg_at_k`) so `test_dataset_task` can call it unchanged, but computed on arrays:

  rel_results(preds, golds, scores, k)   per user, the k predictions are re-ordered by score
                                          (descending, stable for ties, like Python's sorted with
                                          reverse=True) and compared with the gold item -> 0/1 rows
  hit@k                                  number of users with a 1 among the first k entries
  ndcg@k                                 sum over users of sum_i rel_i / log2(i + 2)  (one relevant
                                          item per user under leave-one-out, so IDCG = 1)

Both metrics are SUMS over the users passed in; the runner divides by the user count
(reference `src/runner/single_runner_gram.py:699-702`).
"""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np


def _as_float_list(scores) -> List[float]:
    if hasattr(scores, "detach"):
        scores = scores.detach().cpu().tolist()
    return [float(s) for s in scores]


def rel_results(predictions: Sequence, targets: Sequence, scores, k: int) -> List[List[int]]:
    scores = _as_float_list(scores)
    rows = []
    for u, gold in enumerate(targets):
        lo = u * k
        # stable descending order: ties keep their original relative order
        order = sorted(range(lo, lo + k), key=lambda i: -scores[i] if scores[i] == scores[i] else math.inf)
        rows.append([int(predictions[i] == gold) for i in order])
    return rows


def _discounts(k: int) -> np.ndarray:
    return np.array([1.0 / math.log(i + 2, 2) for i in range(k)], dtype=np.float64)


def hit_at_k(relevance: Sequence[Sequence[int]], k: int) -> float:
    return float(sum(1 for row in relevance if any(row[:k])))


def ndcg_at_k(relevance: Sequence[Sequence[int]], k: int) -> float:
    total = 0.0
    for row in relevance:
        head = row[:k]
        # accumulate in rank order so the floating-point sum matches a left-to-right loop
        acc = 0.0
        for i, r in enumerate(head):
            acc += r / math.log(i + 2, 2)
        total += acc
    return total


def get_metrics_results(rel_results: Sequence[Sequence[int]], metrics: Sequence[str]) -> np.ndarray:
    out = []
    for name in metrics:
        kind, _, cut = name.partition("@")
        kind = kind.lower()
        if kind.startswith("hit"):
            out.append(hit_at_k(rel_results, int(cut)))
        elif kind.startswith("ndcg"):
            out.append(ndcg_at_k(rel_results, int(cut)))
    return np.array(out)


def metric_sums_from_ranks(ranks: np.ndarray, metrics: Sequence[str]) -> np.ndarray:
    """Same sums from integer hit ranks (rank of the gold item among a user's predictions, -1 when
    absent).  This is the form the multi-GPU path all-reduces: integer hit counts are exact and the
    ndcg sum is formed in fp64 in user order on every rank count."""
    ranks = np.asarray(ranks)
    out = []
    for name in metrics:
        kind, _, cut = name.partition("@")
        k = int(cut)
        hit = (ranks >= 0) & (ranks < k)
        if kind.lower().startswith("hit"):
            out.append(float(hit.sum()))
        else:
            total = 0.0
            for r in ranks[hit]:
                total += 1.0 / math.log(int(r) + 2, 2)
            out.append(total)
    return np.array(out)
