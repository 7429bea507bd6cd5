"""Shared builders for the parity tests: deterministic cases (config + weights + inputs + trie)."""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gram_b200 import synth  # noqa: E402
from gram_b200.config import GramConfig  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


@dataclass
class Case:
    name: str
    cfg: GramConfig
    weight_seed: int
    n_users: int
    n_passages: object          # int or (lo, hi)
    seq_len: int
    input_seed: int
    n_items: int
    branching: tuple
    trie_seed: int
    num_beams: int
    length_penalty: float = 1.0
    full: bool = False

    def build(self):
        sd = synth.make_state_dict(self.cfg, seed=self.weight_seed)
        ids, mask = synth.make_user_batch(self.cfg, self.n_users, self.n_passages, self.seq_len,
                                          seed=self.input_seed, full=self.full)
        seqs = synth.make_item_sequences(self.n_items, list(self.branching), self.cfg.vocab_size,
                                         seed=self.trie_seed, variable_tail=True)
        max_length = max(len(s) for s in seqs)
        return sd, torch.from_numpy(ids), torch.from_numpy(mask), seqs, max_length


# Cases shared by oracle/make_golden.py, the CPU tests and the GPU tests.
CASES = {
    # every code path at toy size: ragged passages, all-masked passages, 2 id lengths
    "tiny": Case("tiny", GramConfig.tiny(), weight_seed=3, n_users=3, n_passages=(2, 4), seq_len=16,
                 input_seed=5, n_items=60, branching=(6, 4, 3), trie_seed=11, num_beams=4),
    # wider beam than some trie levels, length penalty != 1
    "tiny_lp": Case("tiny_lp", GramConfig.tiny(), weight_seed=4, n_users=2, n_passages=(1, 3), seq_len=12,
                    input_seed=9, n_items=200, branching=(10, 5, 4, 2), trie_seed=12, num_beams=8,
                    length_penalty=0.6),
    # the real architecture (T5-small shapes), small batch: the reference's beam-20 configuration
    "small": Case("small", GramConfig.t5_small(max_seq_len=32, max_item_num=4), weight_seed=1, n_users=2,
                  n_passages=(2, 4), seq_len=32, input_seed=7, n_items=3000, branching=(40, 15, 5, 2),
                  trie_seed=13, num_beams=20),
}


def oracle_for(case: Case, sd):
    from oracle.gram_oracle import OracleGRAM
    return OracleGRAM(case.cfg, sd)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| relative to the magnitude of the reference tensor b."""
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def has_gpu() -> bool:
    return torch.cuda.is_available()
