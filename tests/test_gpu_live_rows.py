"""Live-row compaction (engine.cu: live_compact before every decode step t >= 1) against GRAM_FLAG_ALL_ROWS, which decodes
every beam row at every step as the reference does (HF beam_search runs dead -inf beams and finished batches through
the decoder and discards them).  Skipping them must not change a single output bit."""
import numpy as np
import pytest
import torch

from helpers import CASES

pytestmark = pytest.mark.gpu


def _run(case, sd, ids, mask, seqs, max_length, dtype, flags, K=None, lp=None):
    from gram_b200 import GRAM, Trie, prefix_allowed_tokens_fn
    K = K or case.num_beams
    m = GRAM(case.cfg, dtype=dtype, device="cuda:0", flags=flags)
    m.load_state_dict(sd)
    fn = prefix_allowed_tokens_fn(Trie(seqs))
    out = m.generate(input_ids=ids.cuda(), attention_mask=mask.cuda(), max_length=max_length, prefix_allowed_tokens_fn=fn,
                     num_beams=K, num_return_sequences=K, output_scores=True, return_dict_in_generate=True,
                     length_penalty=case.length_penalty if lp is None else lp)
    torch.cuda.synchronize()
    return out["sequences"].cpu().numpy(), out["sequences_scores"].cpu().numpy(), m.stats()


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("name", list(CASES))
def test_live_rows_bit_identical_to_all_rows(name, dtype):
    from gram_b200 import _cabi
    case = CASES[name]
    built = case.build()
    seq_a, sc_a, st_a = _run(case, *built, dtype, _cabi.GRAM_FLAG_ALL_ROWS)
    seq_l, sc_l, st_l = _run(case, *built, dtype, 0)
    assert seq_a.shape == seq_l.shape and np.array_equal(seq_a, seq_l)
    assert np.array_equal(sc_a.view(np.uint32), sc_l.view(np.uint32)), "sequence scores differ bitwise"
    # executed-work counters: the reference-like mode runs one row per user at step 0 and every beam row afterwards, and
    # reads every user's K/V at every step; the default runs at most that
    B, K, T = case.n_users, case.num_beams, built[4] - 1
    assert st_a["decoded_rows"] == B + (T - 1) * B * K
    assert st_a["kv_tokens_read"] == T * st_a["packed_tokens"]
    assert B <= st_l["decoded_rows"] <= st_a["decoded_rows"]
    assert st_l["kv_tokens_read"] <= st_a["kv_tokens_read"]
    print(f"[live rows] case={name} {dtype}: decoder rows {st_l['decoded_rows']} / {st_a['decoded_rows']}, "
          f"K/V tokens {st_l['kv_tokens_read']} / {st_a['kv_tokens_read']}")


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_live_rows_with_many_dead_beams(dtype):
    """A trie with FEWER items than beams under some prefixes and two id lengths: most beams of the late steps are dead,
    some users finish early, and a wide beam (K > root fan-out) starts with dead beams at step 1."""
    from gram_b200 import _cabi, synth
    case = CASES["small"]
    sd, ids, mask, _, _ = case.build()
    seqs = synth.make_item_sequences(40, [5, 3, 2, 2], case.cfg.vocab_size, seed=21, variable_tail=True)
    max_length = max(len(s) for s in seqs)
    for K in (4, 20):
        a = _run(case, sd, ids, mask, seqs, max_length, dtype, _cabi.GRAM_FLAG_ALL_ROWS, K=K)
        b = _run(case, sd, ids, mask, seqs, max_length, dtype, 0, K=K)
        assert np.array_equal(a[0], b[0])
        assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


@pytest.mark.parametrize("lp", [1.0, 0.6])
def test_wide_beam_equals_exhaustive_scoring_on_the_gpu(lp):
    """The CUDA path (fp32) against a ranking that never went through a beam loop: with a beam wider than the live
    prefix set, `generate` must return every item ordered by sum log p / len**lp, recomputed by teacher-forcing each
    item through the CPU oracle's model (tests/test_oracle.py pins the same property for the oracle's own loop).  Most
    beams are dead from the first steps on, so this also drives the live-row compaction hard."""
    import math
    from helpers import oracle_for
    from gram_b200 import synth
    case = CASES["tiny"]
    sd, ids, mask, _, _ = case.build()
    seqs = synth.make_item_sequences(11, [3, 2, 2], case.cfg.vocab_size, seed=5, variable_tail=True)
    ml = max(len(s) for s in seqs)
    K = 16
    got_seq, got_sc, _ = _run(case, sd, ids, mask, seqs, ml, "fp32", 0, K=K, lp=lp)
    ora = oracle_for(case, sd)
    for u in range(ids.shape[0]):
        want = []
        for s in seqs:
            dec = torch.tensor([s[:-1]], dtype=torch.long)
            logp = torch.log_softmax(ora.forward(ids[u:u + 1], mask[u:u + 1], dec)[0].float(), -1)
            total = sum(logp[i, s[i + 1]].item() for i in range(len(s) - 1))
            want.append((total / (len(s) - 1) ** lp, s))
        want.sort(key=lambda t: -t[0])
        for rank, (sc, s) in enumerate(want):
            row = got_seq[u * K + rank].tolist()
            row = row + [0] * (len(s) - len(row))
            assert row[:len(s)] == s and all(t == 0 for t in row[len(s):]), (u, rank, row, s)
            assert math.isclose(float(got_sc[u * K + rank]), sc, rel_tol=1e-4, abs_tol=1e-4)
