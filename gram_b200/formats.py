"""On-disk formats either side of the hot path: readers for the reference's
preprocessed text files, the prediction TSV, and binary caches so that start-up is a file map instead of a parse.

Text formats (reference `rec_datasets/README.md:21-72`; loaders `src/utils/indexing.py:150-176,236-246`):

  user_sequence.txt                       `user item1 item2 ... itemn`      one user per line, chronological
  item_generative_indexing_<type>.txt     `ASIN |piece|piece|...`           lexical id of every item, `|`-separated
  similar_item_<cf>.txt                   header line `anchor top1 ...`, then `anchor top1 ... top20`
  item_plain_text.txt                     `ASIN title: ...; brand: ...; ...` (absent from the shipped tree)

Prediction TSV (reference `src/runner/single_runner_gram.py:580-588,675-694`): header
`idx  H@5  H@10  NDCG@5  NDCG@10  gold  pred  scores`, one row per user with `||`-joined predictions and scores,
then one `metric: value` line per metric.

Binary caches written here: the packed dataset (`.npz`: piece table, item -> piece indices, user CSR, similar
items; what `gram_b200/assets/*.npz` are) and the CSR item-ID trie (`.npz`), loadable as a `CsrTrie` that the
device path uploads without ever building the nested dict.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .generation_trie import csr_children, csr_walk

TSV_METRIC_NAMES = {"hit@5": "H@5", "hit@10": "H@10", "ndcg@5": "NDCG@5", "ndcg@10": "NDCG@10"}


# ---- text readers ---------------------------------------------------------------------------------------
def read_item_index(path: str) -> Tuple[List[str], List[List[str]]]:
    """`ASIN |p1|p2|...` -> (asins, pieces per item).  Empty fields between bars are dropped, like
    `str.split('|')` followed by the reference's filtering of the separator ids."""
    asins, pieces = [], []
    with open(path, encoding="utf-8") as f:
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            asin, _, rest = line.partition(" ")
            asins.append(asin)
            pieces.append([p for p in rest.split("|") if p != ""])
    return asins, pieces


def read_user_sequence(path: str) -> List[Tuple[str, List[str]]]:
    """`user item1 ... itemn` -> [(user, [items])]; lines with no item are skipped (indexing.py:150-176)."""
    out = []
    with open(path, encoding="utf-8") as f:
        for line in f:
            parts = line.split()
            if len(parts) >= 2:
                out.append((parts[0], parts[1:]))
    return out


def read_similar_items(path: str, top_k: Optional[int] = None) -> Dict[str, List[str]]:
    """Header line skipped (indexing.py:240-241); anchor -> its neighbours, truncated to `top_k`."""
    out = {}
    with open(path, encoding="utf-8") as f:
        first = True
        for line in f:
            parts = line.split()
            if first:
                first = False
                if parts and parts[0] == "anchor":
                    continue
            if parts:
                out[parts[0]] = parts[1:] if top_k is None else parts[1:1 + top_k]
    return out


def read_item_plain_text(path: str) -> Dict[str, str]:
    """`ASIN text...` -> {ASIN: text} (rec_datasets/README.md:29-43)."""
    out = {}
    with open(path, encoding="utf-8") as f:
        for line in f:
            line = line.rstrip("\n")
            if line:
                asin, _, text = line.partition(" ")
                out[asin] = text
    return out


# ---- packed dataset cache -----------------------------------------------------------------------------------
def pack_dataset(id_file: str, user_sequence: Optional[str] = None, similar_file: Optional[str] = None,
                 top_k: int = 10) -> Dict[str, np.ndarray]:
    """Parse the text files once into the arrays `GramTestData` consumes: `pieces` (first-appearance order),
    `item_asin`, `item_lex` int32 [n_items, max_pieces] (-1 padded piece indices), `user_off` / `user_items`
    (CSR of item indices per user), `similar` int32 [n_items, top_k] (-1 padded)."""
    asins, item_pieces = read_item_index(id_file)
    table: Dict[str, int] = {}
    rows = []
    for ps in item_pieces:
        rows.append([table.setdefault(p, len(table)) for p in ps])
    width = max((len(r) for r in rows), default=0)
    item_lex = np.full((len(rows), width), -1, dtype=np.int32)
    for i, r in enumerate(rows):
        item_lex[i, :len(r)] = r
    out = dict(pieces=np.array(list(table.keys())), item_asin=np.array(asins), item_lex=item_lex)
    index = {a: i for i, a in enumerate(asins)}
    if user_sequence and os.path.exists(user_sequence):
        off, items = [0], []
        for _, seq in read_user_sequence(user_sequence):
            items.extend(index[a] for a in seq)
            off.append(len(items))
        out["user_off"] = np.asarray(off, dtype=np.int32)
        out["user_items"] = np.asarray(items, dtype=np.int32)
    if similar_file and os.path.exists(similar_file):
        sim = np.full((len(asins), top_k), -1, dtype=np.int32)
        for anchor, nbrs in read_similar_items(similar_file).items():
            if anchor in index:
                row = [index[a] for a in nbrs[:top_k] if a in index]
                sim[index[anchor], :len(row)] = row
        out["similar"] = sim
    return out


def save_packed(path: str, packed: Dict[str, np.ndarray]) -> None:
    np.savez_compressed(path, **packed)


def load_packed(path: str) -> Dict[str, np.ndarray]:
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


# ---- CSR trie file --------------------------------------------------------------------------------------------
class CsrTrie:
    """A trie that exists only in its flat form: `get(prefix)` answers from the CSR arrays (same SET as
    `Trie.get`, token-ascending), `to_csr()` hands the arrays to `GRAM.generate` unchanged.  Built from a `Trie`
    once (`save_trie_csr`), then loaded in O(file read) on every later start-up."""

    def __init__(self, csr: dict, n_sequences: int, start_token: int = 0):
        self._csr = csr
        self.len = int(n_sequences)
        self.start_token = int(start_token)
        self._version = 0
        self.append_trie = None

    def get(self, prefix_sequence: Sequence[int]) -> List[int]:
        return csr_children(self._csr, csr_walk(self._csr, [int(t) for t in prefix_sequence]))

    def __getitem__(self, prefix):
        return self.get(prefix)

    def __len__(self):
        return self.len

    def __iter__(self):
        """Every stored sequence, depth-first in token order."""
        off, tok, nxt = self._csr["child_offsets"], self._csr["child_tokens"], self._csr["child_nodes"]
        stack = [(0, [])]
        while stack:
            node, path = stack.pop()
            a, b = int(off[node]), int(off[node + 1])
            if a == b:
                if path:
                    yield path
                continue
            for e in range(b - 1, a - 1, -1):
                stack.append((int(nxt[e]), path + [int(tok[e])]))

    def to_csr(self, start_token: int = 0):
        if start_token != self.start_token:
            raise ValueError(f"this CSR trie was flattened for start token {self.start_token}, not {start_token}")
        return self._csr


def save_trie_csr(path: str, trie, start_token: int = 0) -> None:
    csr = trie.to_csr(start_token)
    np.savez(path, child_offsets=csr["child_offsets"], child_tokens=csr["child_tokens"], child_nodes=csr["child_nodes"],
             meta=np.asarray([csr["n_nodes"], csr["n_edges"], csr["root_node"], csr["max_fanout"], len(trie), start_token],
                             dtype=np.int64))


def load_trie_csr(path: str) -> CsrTrie:
    with np.load(path) as z:
        n_nodes, n_edges, root, fanout, n_seq, start = (int(v) for v in z["meta"])
        csr = dict(child_offsets=z["child_offsets"].astype(np.int32), child_tokens=z["child_tokens"].astype(np.int32),
                   child_nodes=z["child_nodes"].astype(np.int32), n_nodes=n_nodes, n_edges=n_edges, root_node=root,
                   max_fanout=fanout)
    if len(csr["child_offsets"]) != n_nodes + 1 or len(csr["child_tokens"]) != n_edges:
        raise ValueError(f"{path}: inconsistent CSR trie file")
    return CsrTrie(csr, n_seq, start)


# ---- prediction TSV -----------------------------------------------------------------------------------------------
def write_predictions_tsv(path: str, rows: Iterable[tuple], metrics: Dict[str, float],
                          per_user_metrics: Iterable[Sequence[float]], metric_names: Sequence[str]) -> None:
    """rows: (user index, gold string, [predictions], [scores], hit rank); one TSV row per user, then the totals."""
    names = [TSV_METRIC_NAMES.get(m, m) for m in metric_names]
    with open(path, "w", encoding="utf-8") as f:
        f.write("idx\t" + "\t".join(names) + "\tgold\tpred\tscores\n")
        for (u, gold, preds, scores, _), per_user in zip(rows, per_user_metrics):
            f.write("\t".join([f"u{u}", "\t".join(str(x) for x in per_user), gold, "||".join(preds),
                               "||".join(str(s) for s in scores)]) + "\n")
        for m, v in metrics.items():
            f.write(f"{m}: {v}\n")


def read_predictions_tsv(path: str):
    """-> (rows, metrics): rows = dict(idx, per_user {column: float}, gold, pred [str], scores [float])."""
    rows, metrics = [], {}
    with open(path, encoding="utf-8") as f:
        header = f.readline().rstrip("\n").split("\t")
        n_metric = len(header) - 4
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            cols = line.split("\t")
            if len(cols) == len(header):
                rows.append(dict(idx=cols[0], per_user={h: float(v) for h, v in zip(header[1:1 + n_metric], cols[1:1 + n_metric])},
                                 gold=cols[-3], pred=cols[-2].split("||") if cols[-2] else [],
                                 scores=[float(s) for s in cols[-1].split("||")] if cols[-1] else []))
            else:
                name, _, val = line.partition(": ")
                metrics[name] = float(val)
    return rows, metrics
