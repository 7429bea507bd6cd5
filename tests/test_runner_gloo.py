"""Host-side eval loop: sharded (world_size 2, gloo, CPU) results equal the single-process results.

The model is a cheap deterministic stand-in with the `generate` signature the runner calls, so only
the runner's sharding / gather / metric logic is under test here (the CUDA path is covered by the
`-m gpu` tests)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from helpers import ROOT
from gram_b200.data import GramTestData
from gram_b200.runner import GramEvalLoader, GramRunner, shard_range

N_USERS = 37
G = 10


class FakeModel:
    """Returns, per user, G candidate sequences chosen by a hash of the user's input ids (the test
    wrapper below plants the gold item of every third user at a user-dependent rank)."""

    def __init__(self, data):
        self.data = data
        self.cands = data.encoded_candidates()
        self.max_len = max(len(c) for c in self.cands)

    def generate(self, input_ids, attention_mask, max_length, prefix_allowed_tokens_fn, num_beams, num_return_sequences,
                 **kw):
        assert prefix_allowed_tokens_fn.candidate_trie is not None and num_beams == num_return_sequences
        B = input_ids.shape[0]
        seqs = torch.zeros((B * num_beams, max_length), dtype=torch.long)
        scores = torch.zeros(B * num_beams)
        for b in range(B):
            h = int(input_ids[b].sum().item())
            for r in range(num_beams):
                c = self.cands[(h * 31 + r * 7919) % len(self.cands)]
                seqs[b * num_beams + r, :len(c)] = torch.tensor(c)
                scores[b * num_beams + r] = -float(r) - (h % 5) * 0.01
        return dict(sequences=seqs, sequences_scores=scores)


def _run(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    data = GramTestData("Beauty")
    users = list(range(N_USERS))
    model = FakeModel(data)
    # plant golds deterministically: user u with u % 3 == 0 gets its gold at rank u % G
    orig = model.generate

    def generate(input_ids, attention_mask, max_length, prefix_allowed_tokens_fn, num_beams, num_return_sequences, **kw):
        out = orig(input_ids, attention_mask, max_length, prefix_allowed_tokens_fn, num_beams, num_return_sequences, **kw)
        for b, u in enumerate(generate.current_users):
            if u % 3 == 0:
                gold = data.collate([u])["target_ids"][0]
                row = b * num_beams + (u % num_beams)
                out["sequences"][row] = 0
                out["sequences"][row, :len(gold)] = torch.tensor(gold)
                if u % 6 == 0:
                    # the same item twice in one ranking (two token paths that decode to one string do this in the
                    # reference): ndcg must count both matches, as evaluate.ndcg_at_k does
                    row2 = b * num_beams + ((u + 3) % num_beams)
                    out["sequences"][row2] = out["sequences"][row]
        return out

    model.generate = generate
    loader = GramEvalLoader(data, batch_size=4, rank=rank, world=world, users=users)

    class Args:
        metrics = "hit@5,hit@10,ndcg@5,ndcg@10"
        beam_size = G
        length_penalty = 1.0
        item_id_type = "split"

    runner = GramRunner(model, data.tokenizer, None, Args(), rank, world)
    # thread the user indices of the current batch to the fake model
    real_generate = runner._generate

    def _generate(batch, max_length, prefix_fn):
        generate.current_users = batch["user_index"]
        return real_generate(batch, max_length, prefix_fn)

    runner._generate = _generate
    res = runner.test_dataset_task(loader, "test")
    np.save(os.path.join(out_dir, f"ranks_w{world}_r{rank}.npy"), res["hit_ranks"])
    np.save(os.path.join(out_dir, f"rel_w{world}_r{rank}.npy"), res["rel_rows"])
    np.save(os.path.join(out_dir, f"seq_w{world}_r{rank}.npy"), res["sequences"])
    np.save(os.path.join(out_dir, f"sc_w{world}_r{rank}.npy"), res["sequences_scores"])
    assert res["gather_seconds"] >= 0 and int(res["hit_rank_histogram"].sum()) == N_USERS
    assert (res["rows"] is not None) == (rank == 0)
    np.save(os.path.join(out_dir, f"metrics_w{world}_r{rank}.npy"), np.array([res["metrics"][m] for m in runner.metrics]))
    assert res["test_total"] == N_USERS
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_is_exact():
    for n in (0, 1, 7, 37, 22363):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(600)
def test_sharded_eval_equals_single_process(tmp_path):
    out = str(tmp_path)
    _run(0, 1, 0, out)
    port = _free_port()
    mp.spawn(_run, args=(2, port, out), nprocs=2, join=True)
    single = np.load(os.path.join(out, "metrics_w1_r0.npy"))
    ranks1 = np.load(os.path.join(out, "ranks_w1_r0.npy"))
    for r in range(2):
        assert np.array_equal(np.load(os.path.join(out, f"ranks_w2_r{r}.npy")), ranks1)
        assert np.array_equal(np.load(os.path.join(out, f"metrics_w2_r{r}.npy")), single)     # bit-for-bit
        # the gathered ranked lists themselves (one all_gather_into_tensor of int32 ids + bit-cast fp32 scores)
        for name in ("rel", "seq", "sc"):
            assert np.array_equal(np.load(os.path.join(out, f"{name}_w2_r{r}.npy")), np.load(os.path.join(out, f"{name}_w1_r0.npy")))
    assert (ranks1 >= 0).sum() == len([u for u in range(N_USERS) if u % 3 == 0])
    assert single[1] > 0
    # metrics are the reference's functions applied to the 0/1 relevance rows, multi-hit rows included
    from gram_b200 import evaluate
    rel = np.load(os.path.join(out, "rel_w1_r0.npy"))
    assert (rel.sum(1) > 1).any()
    want = evaluate.get_metrics_results(rel.tolist(), ["hit@5", "hit@10", "ndcg@5", "ndcg@10"]) / N_USERS
    assert np.array_equal(want, single)
