"""Item-ID prefix trie: the reference's host API plus a CSR flattening for the device.

Drop-in for reference `src/utils/generation_trie.py:5-95` (`Trie`, `prefix_allowed_tokens_fn`,
`exact_match`).  The nested-dict API (`trie_dict`, `add`, `get`, `__len__`, `__iter__`,
`__getitem__`, `append`, `load_from_dict`) is kept so callers such as
`src/runner/single_runner_gram.py:617-619` work unchanged; `to_csr()` produces the flat arrays
`gram_set_trie` uploads (include/gram_b200.h), and the closure returned by
`prefix_allowed_tokens_fn` exposes its trie as `.candidate_trie` so `GRAM.generate` can find it.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np


class Trie(object):
    def __init__(self, sequences: List[List[int]] = []):
        self.trie_dict: Dict[int, dict] = {}
        self.len = 0
        self._version = 0
        self._csr = None
        if sequences:
            for sequence in sequences:
                Trie._add_to_trie(sequence, self.trie_dict)
                self.len += 1
        self.append_trie = None
        self.bos_token_id = None

    def append(self, trie, bos_token_id):
        self.append_trie = trie
        self.bos_token_id = bos_token_id
        self._version += 1

    def add(self, sequence: List[int]):
        Trie._add_to_trie(sequence, self.trie_dict)
        self.len += 1
        self._version += 1

    def get(self, prefix_sequence: List[int]):
        return Trie._get_from_trie(prefix_sequence, self.trie_dict, self.append_trie, self.bos_token_id)

    @staticmethod
    def load_from_dict(trie_dict):
        trie = Trie()
        trie.trie_dict = trie_dict
        trie.len = sum(1 for _ in trie)
        return trie

    @staticmethod
    def _add_to_trie(sequence: List[int], trie_dict: Dict):
        # iterative (the reference recurses once per token; same resulting dict, insertion-ordered)
        node = trie_dict
        for tok in sequence:
            tok = int(tok)
            nxt = node.get(tok)
            if nxt is None:
                nxt = {}
                node[tok] = nxt
            node = nxt

    @staticmethod
    def _get_from_trie(prefix_sequence, trie_dict, append_trie=None, bos_token_id=None):
        node = trie_dict
        for i, tok in enumerate(prefix_sequence):
            if tok in node:
                node = node[tok]
            else:
                if append_trie:
                    return append_trie.get(list(prefix_sequence[i:]))
                return []
        output = list(node.keys())
        if append_trie and bos_token_id in output:
            output.remove(bos_token_id)
            output += list(append_trie.trie_dict.keys())
        return output

    def __iter__(self):
        """Yield every stored sequence (root-to-leaf path), depth-first in insertion order."""
        stack = [([], iter(self.trie_dict.items()), bool(self.trie_dict))]
        if not self.trie_dict:
            yield []
            return
        while stack:
            prefix, it, _ = stack[-1]
            step = next(it, None)
            if step is None:
                stack.pop()
                continue
            tok, child = step
            path = prefix + [tok]
            if child:
                stack.append((path, iter(child.items()), True))
            else:
                yield path

    def __len__(self):
        return self.len

    def __getitem__(self, value):
        return self.get(value)

    # ---- device form ---------------------------------------------------------------------------
    def to_csr(self, start_token: int = 0):
        """Flatten to CSR.  Node 0 is the empty prefix; nodes are numbered breadth-first; the children of a
        node are stored in ascending token order (the SET equals `get(prefix)`; the device enumerates candidates
        beam-major then token-ascending, which is the order of the flat index beam * V + token that HF's
        `topk` ranks, so a candidate's enumeration index doubles as its tie-break key).

        Returns dict(child_offsets int32[n_nodes+1], child_tokens int32[n_edges],
        child_nodes int32[n_edges], n_nodes, n_edges, root_node, max_fanout) where `root_node` is the
        node reached by `start_token` (-1 when the trie has no such sequence)."""
        if self.append_trie is not None:
            raise ValueError("Trie.append() chains are not supported by the device trie")
        if self._csr is not None and self._csr[0] == (self._version, self.len, start_token):
            return self._csr[1]
        offsets = [0]
        tokens: List[int] = []
        child_nodes: List[int] = []
        queue = [self.trie_dict]
        head = 0
        while head < len(queue):
            node = queue[head]
            head += 1
            for tok, child in sorted(node.items()):
                tokens.append(tok)
                child_nodes.append(len(queue))
                queue.append(child)
            offsets.append(len(tokens))
        root = -1
        for e in range(offsets[0], offsets[1]):
            if tokens[e] == start_token:
                root = child_nodes[e]
        off = np.asarray(offsets, dtype=np.int32)
        csr = dict(child_offsets=off, child_tokens=np.asarray(tokens, dtype=np.int32),
                   child_nodes=np.asarray(child_nodes, dtype=np.int32), n_nodes=len(queue),
                   n_edges=len(tokens), root_node=root,
                   max_fanout=int(np.diff(off).max()) if len(queue) else 0)
        self._csr = ((self._version, self.len, start_token), csr)
        return csr


def csr_children(csr, node: int) -> List[int]:
    """Allowed next tokens at a CSR node (host mirror of what the beam kernel gathers)."""
    if node < 0:
        return []
    a, b = csr["child_offsets"][node], csr["child_offsets"][node + 1]
    return csr["child_tokens"][a:b].tolist()


def csr_walk(csr, prefix: List[int]) -> int:
    """Node reached by `prefix` from the empty prefix, or -1."""
    node = 0
    for tok in prefix:
        a, b = csr["child_offsets"][node], csr["child_offsets"][node + 1]
        nxt = -1
        for e in range(a, b):
            if csr["child_tokens"][e] == tok:
                nxt = int(csr["child_nodes"][e])
                break
        if nxt < 0:
            return -1
        node = nxt
    return node


def prefix_allowed_tokens_fn(candidate_trie):
    def prefix_allowed_tokens(batch_id, sentence):
        sentence = sentence.tolist()
        trie_out = candidate_trie.get(sentence)
        return trie_out

    prefix_allowed_tokens.candidate_trie = candidate_trie
    return prefix_allowed_tokens


def exact_match(predictions, targets, k):
    """Number of users whose gold item is among their k predictions (reference helper, :98-108)."""
    return sum(1 for u, gold in enumerate(targets) if gold in predictions[u * k:(u + 1) * k])
