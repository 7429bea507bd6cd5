"""GPU tests of the per-item encoder-state cache.

The cached path (`cache_items` once, then `generate_cached(prompt, item indices)`) must reproduce the ordinary
passage-batched path bit for bit: item passages are encoded independently of the user, by the same kernels, and
the position row is added with the same separately-rounded fp32 add.
"""
import numpy as np
import pytest
import torch

from helpers import CASES
from gram_b200 import GRAM, GramConfig, Trie, prefix_allowed_tokens_fn, synth
from gram_b200.data import GramTestData

pytestmark = pytest.mark.gpu


def _synthetic(cfg, n_items, B, NI, L, seed):
    """Item table with ragged passage lengths, users with ragged histories (-1 padded, one user with no items)."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(1, L + 1, size=n_items)
    lens[0] = L
    item_ids = rng.integers(2, cfg.vocab_size, size=(n_items, L)).astype(np.int64)
    item_mask = np.arange(L)[None, :] < lens[:, None]
    item_ids[~item_mask] = 0
    plen = rng.integers(1, L + 1, size=B)
    prompt_ids = rng.integers(2, cfg.vocab_size, size=(B, L)).astype(np.int64)
    prompt_mask = np.arange(L)[None, :] < plen[:, None]
    prompt_ids[~prompt_mask] = 0
    items = np.full((B, NI), -1, dtype=np.int32)
    for b in range(B):
        k = 0 if b == 1 else int(rng.integers(1, NI + 1))
        items[b, :k] = rng.integers(0, n_items, size=k)
    ids = np.zeros((B, NI + 1, L), dtype=np.int64)
    mask = np.zeros((B, NI + 1, L), dtype=bool)
    ids[:, 0], mask[:, 0] = prompt_ids, prompt_mask
    for b in range(B):
        for j in range(NI):
            if items[b, j] >= 0:
                ids[b, 1 + j], mask[b, 1 + j] = item_ids[items[b, j]], item_mask[items[b, j]]
    return item_ids, item_mask, prompt_ids, prompt_mask, items, ids, mask


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("case", ["tiny", "small"])
def test_cached_path_is_bit_identical(case, dtype):
    c = CASES[case]
    cfg = c.cfg
    sd, _, _, seqs, ml = c.build()
    L, NI, B = c.seq_len, cfg.max_item_num, 5
    item_ids, item_mask, p_ids, p_mask, items, ids, mask = _synthetic(cfg, 37, B, NI, L, seed=21)
    fn = prefix_allowed_tokens_fn(Trie(seqs))
    K = c.num_beams
    m = GRAM(cfg, dtype=dtype, device="cuda:0")
    m.load_state_dict(sd)
    t = lambda a: torch.from_numpy(a).cuda()
    mem = m.encode(t(ids), t(mask))
    ref = m.generate(t(ids), t(mask), ml, prefix_allowed_tokens_fn=fn, num_beams=K, num_return_sequences=K,
                     return_dict_in_generate=True)
    m.cache_items(t(item_ids), t(item_mask))
    mem_c = m.encode_cached(t(p_ids), t(p_mask), t(items))
    assert torch.equal(mem, mem_c)
    for host in (False, True):                                   # device and host inputs
        cv = (lambda a: torch.from_numpy(a)) if host else t
        out = m.generate_cached(cv(p_ids), cv(p_mask), cv(items), ml, prefix_allowed_tokens_fn=fn, num_beams=K,
                                num_return_sequences=K, return_dict_in_generate=True)
        assert torch.equal(out["sequences"].cpu(), ref["sequences"].cpu())
        assert torch.equal(out["sequences_scores"].cpu(), ref["sequences_scores"].cpu())
    # the table survives an engine re-creation (capacity growth) and chunked calls (user_limit < B)
    m.user_limit = 2
    m.configure(max_users=2)
    out = m.generate_cached(t(p_ids), t(p_mask), t(items), ml, prefix_allowed_tokens_fn=fn, num_beams=K,
                            num_return_sequences=K, return_dict_in_generate=True)
    assert torch.equal(out["sequences"].cpu(), ref["sequences"].cpu())


def test_cached_path_errors():
    c = CASES["tiny"]
    sd, _, _, seqs, ml = c.build()
    L, NI = c.seq_len, c.cfg.max_item_num
    item_ids, item_mask, p_ids, p_mask, items, _, _ = _synthetic(c.cfg, 9, 3, NI, L, seed=4)
    fn = prefix_allowed_tokens_fn(Trie(seqs))
    m = GRAM(c.cfg, dtype="fp32", device="cuda:0")
    m.load_state_dict(sd)
    t = lambda a: torch.from_numpy(a).cuda()
    with pytest.raises(RuntimeError):                            # no table yet
        m.generate_cached(t(p_ids), t(p_mask), t(items), ml, prefix_allowed_tokens_fn=fn, num_beams=2)
    m.cache_items(t(item_ids), t(item_mask))
    bad = items.copy()
    bad[0, 0] = 9                                                # one past the table
    with pytest.raises(ValueError):
        m.generate_cached(p_ids, p_mask, bad, ml, prefix_allowed_tokens_fn=fn, num_beams=2)
    with pytest.raises(ValueError):                              # L differs from the table's L
        m.generate_cached(t(p_ids[:, :L - 1].copy()), t(p_mask[:, :L - 1].copy()), t(items), ml,
                          prefix_allowed_tokens_fn=fn, num_beams=2)
    out = m.generate_cached(t(p_ids), t(p_mask), t(items), ml, prefix_allowed_tokens_fn=fn, num_beams=2)
    assert out.shape[0] == 3 * 1


def test_cached_path_beauty_batch():
    """Headline configuration: 96 real Beauty users, T5-small bf16, beam 20 -- identical rankings and scores, and
    the loader's two collations describe the same batch."""
    data = GramTestData("Beauty")
    cfg = GramConfig.t5_small(max_seq_len=data.L, max_item_num=data.max_his)
    cands = data.encoded_candidates()
    fn = prefix_allowed_tokens_fn(Trie(cands))
    ml = max(len(x) for x in cands)
    users = [(i * 211) % data.n_users for i in range(96)]
    a, b = data.collate(users), data.collate_cached(users)
    assert a["item_text_ids"].shape[2] == data.L and a["target_ids"] == b["target_ids"]
    assert np.array_equal(a["item_text_ids"][:, 0], b["prompt_ids"])
    tab, tmask = data.item_table()
    for u in range(len(users)):
        for j in range(b["item_index"].shape[1]):
            it = b["item_index"][u, j]
            if it < 0:
                assert not a["item_text_masks"][u, 1 + j].any()
            else:
                assert np.array_equal(a["item_text_ids"][u, 1 + j], tab[it])
    m = GRAM(cfg, dtype="bf16", device="cuda:0")
    m.load_state_dict(synth.make_state_dict(cfg, seed=0))
    t = lambda x: torch.from_numpy(x).cuda()
    ref = m.generate(t(a["item_text_ids"]), t(a["item_text_masks"]), ml, prefix_allowed_tokens_fn=fn, num_beams=20,
                     num_return_sequences=20, return_dict_in_generate=True)
    m.cache_items(tab, tmask)                                     # host table, uploaded chunk by chunk
    out = m.generate_cached(t(b["prompt_ids"]), t(b["prompt_masks"]), t(b["item_index"]), ml, prefix_allowed_tokens_fn=fn,
                            num_beams=20, num_return_sequences=20, return_dict_in_generate=True)
    assert torch.equal(out["sequences"], ref["sequences"])
    assert torch.equal(out["sequences_scores"], ref["sequences_scores"])
