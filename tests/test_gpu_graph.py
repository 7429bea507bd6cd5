"""GRAM_FLAG_CUDA_GRAPH: gram_generate replays one captured graph per call shape (first call eager, second captures, later
calls replay).  Same kernels, so the outputs must equal the eager engine's bit for bit on every call, with inputs that
change between calls, host or device tensors, and through the cached-item (decode-only graph) path."""
import numpy as np
import pytest
import torch

from helpers import CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,dtype", [("tiny", "fp32"), ("tiny_lp", "bf16"), ("small", "bf16")])
def test_graph_replay_equals_eager(name, dtype):
    from gram_b200 import GRAM, Trie, _cabi, prefix_allowed_tokens_fn
    case = CASES[name]
    sd, ids, mask, seqs, max_length = case.build()
    K = case.num_beams
    fn = prefix_allowed_tokens_fn(Trie(seqs))
    eager = GRAM(case.cfg, dtype=dtype, device="cuda:0")
    graph = GRAM(case.cfg, dtype=dtype, device="cuda:0", flags=_cabi.GRAM_FLAG_CUDA_GRAPH)
    for m in (eager, graph):
        m.load_state_dict(sd)
    B = ids.shape[0]
    for call in range(5):
        perm = torch.roll(torch.arange(B), call)                       # different users in each slot, same shape
        i, mk = ids[perm], mask[perm]
        if call != 3:
            i, mk = i.cuda(), mk.cuda()                                # call 3: host tensors through the same graph
        outs = []
        for m in (eager, graph):
            o = m.generate(input_ids=i, attention_mask=mk, max_length=max_length, prefix_allowed_tokens_fn=fn, num_beams=K,
                           num_return_sequences=K, return_dict_in_generate=True, length_penalty=case.length_penalty)
            outs.append((o["sequences"].cpu().numpy(), o["sequences_scores"].cpu().numpy()))
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]), f"call {call}"
    assert graph.stats()["launches"] == eager.stats()["launches"]
    # a different shape gets its own graph; a new trie drops the captured graphs (they hold the CSR pointers)
    o1 = graph.generate(input_ids=ids[:1].cuda(), attention_mask=mask[:1].cuda(), max_length=max_length, prefix_allowed_tokens_fn=fn,
                        num_beams=K, num_return_sequences=K, return_dict_in_generate=True, length_penalty=case.length_penalty)
    fn2 = prefix_allowed_tokens_fn(Trie(seqs[: len(seqs) // 2]))
    for m in (eager, graph):
        for _ in range(3):
            o = m.generate(input_ids=ids.cuda(), attention_mask=mask.cuda(), max_length=max_length, prefix_allowed_tokens_fn=fn2,
                           num_beams=K, num_return_sequences=K, return_dict_in_generate=True, length_penalty=case.length_penalty)
        outs.append(o["sequences"].cpu().numpy())
    assert np.array_equal(outs[-1], outs[-2]) and o1["sequences"].shape[0] == K


def test_graph_replay_cached_item_path():
    """generate_cached = gram_encode_cached (eager) + gram_generate(ids = NULL): the decode-only graph"""
    from gram_b200 import GRAM, GramConfig, Trie, _cabi, prefix_allowed_tokens_fn, synth
    from gram_b200.data import GramTestData
    data = GramTestData("Beauty")
    cfg = GramConfig.t5_small(max_seq_len=data.L, max_item_num=data.max_his)
    sd = synth.make_state_dict(cfg, seed=0)
    cands = data.encoded_candidates()
    fn = prefix_allowed_tokens_fn(Trie(cands))
    ml = max(len(c) for c in cands)
    tab, tmask = data.item_table()
    outs = []
    for flags in (0, _cabi.GRAM_FLAG_CUDA_GRAPH):
        m = GRAM(cfg, dtype="bf16", device="cuda:0", flags=flags, max_users=24)
        m.load_state_dict(sd)
        m.cache_items(torch.from_numpy(tab), torch.from_numpy(tmask))
        res = []
        for call in range(4):
            b = data.collate_cached([(call * 24 + i) % data.n_users for i in range(24)])
            # keep the shape fixed: pad the item-index matrix to the history cap
            items = np.full((24, data.max_his), -1, dtype=np.int32)
            items[:, :b["item_index"].shape[1]] = b["item_index"]
            o = m.generate_cached(torch.from_numpy(b["prompt_ids"]).cuda(), torch.from_numpy(b["prompt_masks"]).cuda(),
                                  torch.from_numpy(items).cuda(), ml, prefix_allowed_tokens_fn=fn, num_beams=20,
                                  num_return_sequences=20, return_dict_in_generate=True)
            res.append((o["sequences"].cpu().numpy(), o["sequences_scores"].cpu().numpy()))
        outs.append(res)
        del m
    for a, b in zip(*outs):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
