"""CPU tests of the oracle: against the committed golden fixtures (made by oracle/make_golden.py from
the REAL reference modules), against the reference itself when it is mounted, and against the
reference's known answers for the trie and the metrics."""
import math
import os

import numpy as np
import pytest
import torch

from helpers import CASES, GOLDEN_DIR, oracle_for
from oracle import ref_shim
from oracle.gram_oracle import (OracleGRAM, OracleTrie, get_metrics_results, hf426_beam_search, rel_results,
                                relative_position_bucket)

needs_ref = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def built():
    out = {}
    for name, case in CASES.items():
        sd, ids, mask, seqs, ml = case.build()
        out[name] = dict(case=case, sd=sd, ids=ids, mask=mask, seqs=seqs, ml=ml,
                         gold=np.load(os.path.join(GOLDEN_DIR, f"{name}.npz")))
    return out


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(built, name):
    b = built[name]
    case, gold = b["case"], b["gold"]
    ora = oracle_for(case, b["sd"])
    B = b["ids"].shape[0]
    mem = ora.encode(b["ids"], b["mask"])
    rows = gold["memory_rows"]
    assert torch.allclose(mem[rows[:, 0], rows[:, 1]], torch.from_numpy(gold["memory"]), atol=2e-5, rtol=1e-5)
    logits = ora.forward(b["ids"], b["mask"], torch.from_numpy(gold["dec_ids"]))
    vs = torch.from_numpy(gold["vocab_idx"]).long()
    assert torch.allclose(logits[:, :, vs], torch.from_numpy(gold["logits"]), atol=5e-5, rtol=1e-5)
    assert int(gold["max_length"]) == b["ml"]
    out = ora.generate(b["ids"], b["mask"], b["ml"], OracleTrie(b["seqs"]), case.num_beams, case.num_beams,
                       case.length_penalty)
    assert np.array_equal(out["sequences"].numpy(), gold["sequences"])
    assert np.allclose(out["sequences_scores"].numpy(), gold["sequences_scores"], atol=2e-5)
    assert out["n_steps"] == int(gold["n_steps"])
    assert out["sequences"].shape[0] == B * case.num_beams


@needs_ref
@pytest.mark.parametrize("name", ["tiny", "tiny_lp"])
def test_oracle_bit_exact_with_reference_modules(built, name):
    """The restated model math is the SAME sequence of torch ops as the reference modules: equal bits."""
    import sys
    b = built[name]
    cfg = b["case"].cfg
    hf = ref_shim.make_reference_config(
        vocab_size=cfg.vocab_size, d_model=cfg.d_model, d_kv=cfg.d_kv, d_ff=cfg.d_ff, num_layers=cfg.num_layers,
        num_decoder_layers=cfg.num_decoder_layers, num_heads=cfg.num_heads, max_seq_len=cfg.max_seq_len,
        max_item_num=cfg.max_item_num)
    ref = ref_shim.build_reference_model(hf, b["sd"])
    ora = oracle_for(b["case"], b["sd"])
    ids, mask = b["ids"], b["mask"]
    B, N, L = ids.shape
    ref.encoder.n_passages = N
    mem_ref = ref.encoder(input_ids=ids.view(B, -1), attention_mask=mask.view(B, -1), return_dict=True)[0]
    mem = ora.encode(ids, mask)
    assert torch.equal(mem_ref, mem)
    dec = torch.from_numpy(b["gold"]["dec_ids"])
    assert torch.equal(ref(input_ids=ids, attention_mask=mask, decoder_input_ids=dec, return_dict=True).logits,
                       ora.forward(ids, mask, dec))
    BO = sys.modules["gram_ref_model.gram_t5_outputs"].BaseModelOutputWithPastAndCrossAttentions
    past = opast = None
    for t in range(dec.shape[1]):
        di = dec[:, t:t + 1]
        o = ref(decoder_input_ids=di, past_key_values=past, encoder_outputs=BO(last_hidden_state=mem_ref),
                attention_mask=mask.view(B, -1), use_cache=True, return_dict=True)
        past = o.past_key_values
        lg, opast = ora.decode(di, mem, mask.view(B, -1), opast)
        assert torch.equal(o.logits, lg)
        for lr, lo in zip(past, opast):
            for a, c in zip(lr, lo):
                assert torch.equal(a, c)
    # the reference's _reorder_cache vs the restated one
    idx = torch.tensor([B - 1 - i for i in range(B)])
    for lr, lo in zip(ref._reorder_cache(past, idx), ora.reorder_cache(opast, idx)):
        for a, c in zip(lr, lo):
            assert torch.equal(a, c)


@needs_ref
def test_trie_and_metrics_match_reference_functions():
    ref = ref_shim.load_reference()
    seqs = CASES["tiny_lp"].build()[3]
    a, r = OracleTrie(seqs), ref.generation_trie.Trie(seqs)
    assert a.trie_dict == r.trie_dict and len(a) == len(r)
    for s in seqs[:50]:
        for j in range(len(s) + 1):
            assert a.get(s[:j]) == r.get(s[:j])
    assert a.get([0, 999999]) == r.get([0, 999999]) == []
    preds = ["a", "b", "c", "d", "b", "x", "y", "z"]
    golds = ["c", "q"]
    scores = [0.1, 0.9, 0.5, 0.5, 0.3, 0.2, 0.8, 0.1]
    mine, theirs = rel_results(preds, golds, scores, 4), ref.evaluate.rel_results(preds, golds, scores, 4)
    assert mine == theirs
    ms = ["hit@1", "hit@3", "ndcg@3", "ndcg@4"]
    assert np.array_equal(get_metrics_results(mine, ms), ref.evaluate.get_metrics_results(theirs, ms))


def test_trie_known_answer_from_reference_comment():
    """reference src/runner/single_runner_gram.py:591-593: candidate 'rene furterer complexe 5'."""
    seq = [0] + [3, 1536, 15, 4223, 449, 49, 1561, 15, 305, 1]
    t = OracleTrie([seq])
    expect = {0: {3: {1536: {15: {4223: {449: {49: {1561: {15: {305: {1: {}}}}}}}}}}}}
    assert t.trie_dict == expect
    assert t.get([0]) == [3] and t.get(seq[:-1]) == [1] and t.get(seq) == [] and t.get([5]) == []


def test_metrics_known_values():
    rel = [[0, 1, 0, 0], [0, 0, 0, 0], [1, 0, 0, 0]]
    got = get_metrics_results(rel, ["hit@1", "hit@2", "ndcg@2", "ndcg@4"])
    assert got[0] == 1.0 and got[1] == 2.0
    assert math.isclose(got[2], 1.0 / math.log(3, 2) + 1.0)
    assert math.isclose(got[3], 1.0 / math.log(3, 2) + 1.0)
    # score ties keep the original order (stable sort, as Python's sorted(reverse=True))
    assert rel_results(["a", "b"], ["b"], [0.5, 0.5], 2) == [[0, 1]]


def test_relative_position_buckets_edges():
    rel = torch.arange(-130, 131)
    b = relative_position_bucket(rel, True, 32, 128)
    assert int(b.min()) == 0 and int(b.max()) == 31
    assert b[130].item() == 0 and b[131].item() == 17 and b[129].item() == 1        # rel = 0, +1, -1
    assert b[130 + 128].item() == 31 and b[130 - 128].item() == 15
    d = relative_position_bucket(-torch.arange(0, 40), False, 32, 128)
    assert d[:16].tolist() == list(range(16))                                        # exact below 16


def test_beam_search_properties():
    """Scorer semantics on a hand-made 'model': logits depend only on the last token."""
    V, K = 12, 3
    seqs = [[0, 2, 3, 1], [0, 2, 4, 1], [0, 5, 1], [0, 6, 7, 1], [0, 6, 8, 1]]
    table = torch.linspace(-1, 1, V * V).view(V, V)

    def decode_fn(dec_in, mem, mm, past):
        return table[dec_in[:, -1]].unsqueeze(1), (past or 0)

    out = hf426_beam_search(decode_fn, lambda p, i: p, OracleTrie(seqs), torch.zeros(1, 2, 4), torch.ones(1, 2, dtype=torch.bool),
                            4, K, K, 1.0, V, 1, 0, 0)
    got = [tuple(r) for r in out["sequences"].tolist()]
    assert len(set(got)) == K
    allowed = {tuple((s + [0] * 4)[:4]) for s in seqs}
    assert set(got) <= allowed
    sc = out["sequences_scores"]
    assert torch.all(sc[:-1] >= sc[1:])
    # every returned score is sum of log-probs / hypothesis length (start token counted, EOS not)
    lsm = torch.log_softmax(table, -1)
    for row, s in zip(out["sequences"].tolist(), sc.tolist()):
        toks = [t for t in row if t != 0 or row.index(t) == 0]
        toks = row[:row.index(1) + 1]
        total = sum(lsm[toks[i], toks[i + 1]].item() for i in range(len(toks) - 1))
        assert math.isclose(s, total / (len(toks) - 1), rel_tol=1e-5, abs_tol=1e-5)


@pytest.mark.parametrize("lp", [1.0, 0.6])
def test_wide_beam_equals_exhaustive_teacher_forced_scoring(lp):
    """An anchor for the beam loop that does not go through the beam loop: when the beam is at least as wide as the set
    of live prefixes at every depth, nothing is ever pruned, so `generate` must return ALL items ranked by
    sum_t log p(s_t | s_<t) / len ** length_penalty with len = tokens before EOS, start token included
    (BeamHypotheses.add of transformers 4.26) -- computed here by teacher-forcing every item through the model
    (`forward`, bit-identical to the reference modules) with the full-vocabulary log-softmax."""
    from gram_b200 import synth
    case = CASES["tiny"]
    sd, ids, mask, _, _ = case.build()
    seqs = synth.make_item_sequences(11, [3, 2, 2], case.cfg.vocab_size, seed=5, variable_tail=True)
    assert len({len(s) for s in seqs}) == 2                      # two id lengths, as in the shipped ID files
    ml = max(len(s) for s in seqs)
    K = 16                                                       # >= number of prefixes at any depth (<= 11)
    ora = oracle_for(case, sd)
    out = ora.generate(ids, mask, ml, OracleTrie(seqs), K, K, lp)
    B = ids.shape[0]
    for u in range(B):
        want = []
        for s in seqs:
            dec = torch.tensor([s[:-1]], dtype=torch.long)
            logp = torch.log_softmax(ora.forward(ids[u:u + 1], mask[u:u + 1], dec)[0].float(), -1)
            total = sum(logp[i, s[i + 1]].item() for i in range(len(s) - 1))
            want.append((total / (len(s) - 1) ** lp, s))
        want.sort(key=lambda t: -t[0])
        got_seq = out["sequences"][u * K:(u + 1) * K].tolist()
        got_sc = out["sequences_scores"][u * K:(u + 1) * K].tolist()
        for rank, (sc, s) in enumerate(want):
            row = got_seq[rank]
            assert row[:len(s)] == s and all(t == 0 for t in row[len(s):]), (u, rank, row, s)
            assert math.isclose(got_sc[rank], sc, rel_tol=2e-5, abs_tol=2e-5)
        assert all(x == float("-inf") or x < want[-1][0] for x in got_sc[len(want):])


# ---------------------------------------------------------------------------------------------
# pins of the third-party (transformers 4.26.0) pieces that CAN still be pinned offline
# ---------------------------------------------------------------------------------------------
def test_prefix_mask_equals_installed_transformers_processor():
    """`PrefixConstrainedLogitsProcessor` still exists in the installed transformers (5.x) with the 4.26 mask-then-add
    body (plus a new empty-list ValueError): the oracle's mask step must equal it bit for bit, driven by the reference's
    own `Trie` / `prefix_allowed_tokens_fn` when the reference is mounted (ours otherwise -- tests/test_host.py shows
    they agree on every prefix of every shipped id file)."""
    from transformers.generation.logits_process import PrefixConstrainedLogitsProcessor
    from oracle.gram_oracle import prefix_constrained_scores
    from gram_b200.data import GramTestData
    if ref_shim.reference_available():
        gt = ref_shim.load_reference().generation_trie
    else:
        from gram_b200 import generation_trie as gt
    cands = GramTestData("Beauty").encoded_candidates()
    trie = gt.Trie(cands)
    fn = gt.prefix_allowed_tokens_fn(trie)
    K, V, users = 5, 32128, 4
    g = torch.Generator().manual_seed(3)
    for depth in (1, 2, 4, 6):
        pick = torch.randint(0, len(cands), (users * K,), generator=g).tolist()
        ids = torch.tensor([cands[i][:depth] for i in pick], dtype=torch.long)
        scores = torch.log_softmax(torch.randn(users * K, V, generator=g), -1)
        want = PrefixConstrainedLogitsProcessor(fn, K)(ids, scores)
        got = prefix_constrained_scores(trie, ids, scores)
        assert torch.equal(got, want)
        assert int(torch.isfinite(got).sum()) == sum(len(trie.get(r.tolist())) for r in ids)
    # a prefix outside the trie: 4.26 (and the oracle, and the device: a dead beam) give an all -inf row; 5.x raises
    bad = torch.tensor([[0, 31999]], dtype=torch.long)
    assert trie.get(bad[0].tolist()) == []
    assert not torch.isfinite(prefix_constrained_scores(trie, bad, torch.zeros(1, V))).any()
    with pytest.raises(ValueError):
        PrefixConstrainedLogitsProcessor(fn, 1)(bad, torch.zeros(1, V))


def test_beam_hypotheses_hand_computed_cases():
    """`BeamHypotheses.add / is_done` (transformers 4.26.0 generation/beam_search.py, early_stopping=False) on cases worked
    out by hand from the published source: length normalisation counts the start token and not EOS, the worst hypothesis
    is replaced (earliest among equal scores), `worst_score` follows, `is_done` compares with best_sum / cur_len**lp and is
    not monotone in the step."""
    from oracle.gram_oracle import _BeamHypotheses
    h = _BeamHypotheses(2, 1.0)
    assert not h.is_done(0.0, 1)                                 # fewer than num_beams hypotheses: never done
    h.add(torch.tensor([0, 7, 9]), -6.0)                         # [start, a, b]: len 3 -> -2.0
    assert h.beams[0][0] == -2.0 and h.worst_score == -2.0
    h.add(torch.tensor([0, 7]), -6.0)                            # len 2 -> -3.0
    assert h.worst_score == -3.0 and len(h) == 2
    h.add(torch.tensor([0, 5, 5, 5]), -10.0)                     # -2.5 > worst: joins, -3.0 leaves, worst becomes -2.5
    assert sorted(s for s, _ in h.beams) == [-2.5, -2.0] and h.worst_score == -2.5
    h.add(torch.tensor([0, 1]), -8.0)                            # -4.0 <= worst and full: ignored
    assert len(h) == 2 and h.worst_score == -2.5
    # is_done: worst_score >= best_sum_logprobs / cur_len ** lp
    assert not h.is_done(-2.0, 1)                                # -2.0 / 1 = -2.0 > -2.5
    assert h.is_done(-2.5, 1) and h.is_done(-6.0, 2)             # equality counts; -3.0 < -2.5
    assert not h.is_done(-6.0, 3)                                # the SAME candidate score later: -2.0 again -> not done
    # equal scores: sorted((score, idx)) removes the EARLIEST of the equal worst ones
    e = _BeamHypotheses(2, 1.0)
    e.add(torch.tensor([0, 2]), -4.0)                            # idx 0: -2.0
    e.add(torch.tensor([0, 3]), -4.0)                            # idx 1: -2.0
    e.add(torch.tensor([0, 4]), -2.0)                            # -1.0: idx 0 leaves
    assert [hyp.tolist() for _, hyp in e.beams] == [[0, 3], [0, 4]] and e.worst_score == -2.0
    # length penalty: score = sum / len ** lp
    p = _BeamHypotheses(1, 0.6)
    p.add(torch.tensor([0, 1, 2, 3]), -3.0)
    assert p.beams[0][0] == -3.0 / (4 ** 0.6)
    assert p.is_done(-3.0, 4) and not p.is_done(-3.0 * 0.999, 4)
