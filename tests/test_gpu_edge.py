"""GPU edge cases and the scale-config shapes (BASELINE.json configs[4]: T5-base, beam 50, 256-token passages):
parity with the oracle computed live (the oracle itself is pinned bit-for-bit to the reference modules)."""
import numpy as np
import pytest
import torch

from helpers import CASES, ROOT  # noqa: F401
from gram_b200 import GRAM, GramConfig, Trie, prefix_allowed_tokens_fn, synth

pytestmark = pytest.mark.gpu


def _oracle(cfg, sd, ids, mask, seqs, K, max_length, lp=1.0):
    from oracle.gram_oracle import OracleGRAM, OracleTrie
    return OracleGRAM(cfg, sd).generate(ids, mask, max_length, OracleTrie(seqs), K, K, lp)


def _ours(cfg, sd, ids, mask, seqs, K, max_length, dtype, lp=1.0):
    m = GRAM(cfg, dtype=dtype, device="cuda:0")
    m.load_state_dict(sd)
    fn = prefix_allowed_tokens_fn(Trie(seqs))
    out = m.generate(input_ids=ids.cuda(), attention_mask=mask.cuda(), max_length=max_length, prefix_allowed_tokens_fn=fn,
                     num_beams=K, num_return_sequences=K, return_dict_in_generate=True, length_penalty=lp)
    return out["sequences"].cpu().numpy(), out["sequences_scores"].cpu().numpy()


def test_t5_base_beam50_long_passages():
    """T5-base shapes (d=768, 12 heads, 12+12 layers, d_ff=3072), L=256, K=50 (> 32 beams: the two-beam-half
    cross-attention kernel), root fan-out >= K as SURVEY 8(c) requires for a tie-free first step."""
    cfg = GramConfig.t5_base(max_seq_len=256, max_item_num=2)
    sd = synth.make_state_dict(cfg, seed=5)
    ids, mask = synth.make_user_batch(cfg, 2, (1, 2), 256, seed=21, min_len=200)
    ids, mask = torch.from_numpy(ids), torch.from_numpy(mask)
    seqs = synth.make_item_sequences(4000, [64, 8, 4, 2], cfg.vocab_size, seed=3, variable_tail=True)
    K, ml = 50, max(len(s) for s in seqs)
    ref = _oracle(cfg, sd, ids, mask, seqs, K, ml)
    want, wsc = ref["sequences"].numpy(), ref["sequences_scores"].numpy()
    got, gsc = _ours(cfg, sd, ids, mask, seqs, K, ml, "fp32")
    gap = np.abs(np.diff(wsc.reshape(2, K), axis=1)).min()
    print(f"[t5-base K=50] oracle min rank gap {gap:.3e}")
    assert got.shape == want.shape and np.array_equal(got, want)
    assert np.abs(gsc - wsc).max() < 2e-4
    got16, _ = _ours(cfg, sd, ids, mask, seqs, K, ml, "bf16")
    for u in range(2):
        a = {tuple(r) for r in want[u * K:u * K + 10].tolist()}
        b = {tuple(r) for r in got16[u * K:u * K + 10, :want.shape[1]].tolist()}
        print(f"[t5-base K=50 bf16] user {u} top-10 overlap {len(a & b) / 10:.1f}")
        assert len(a & b) >= 5


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_ragged_and_minimal_inputs(dtype):
    """Single user / single passage / one-token passages / interior masked tokens / beam 1."""
    cfg = GramConfig.t5_small(max_seq_len=32, max_item_num=4)
    sd = synth.make_state_dict(cfg, seed=2)
    seqs = synth.make_item_sequences(300, [12, 5, 5], cfg.vocab_size, seed=9)
    ml = max(len(s) for s in seqs)
    # user 0: one passage holding a single token (EOS); user 1: interior mask holes; user 2: full passages
    ids = torch.zeros((3, 3, 32), dtype=torch.long)
    mask = torch.zeros((3, 3, 32), dtype=torch.bool)
    ids[0, 0, 0] = 1
    mask[0, 0, 0] = True
    g = torch.Generator().manual_seed(4)
    ids[1, :2, :20] = torch.randint(2, 30000, (2, 20), generator=g)
    mask[1, :2, :20] = True
    mask[1, 0, 5:9] = False                      # holes inside the valid range (general masks are honoured)
    ids[1, 0, 5:9] = 0
    ids[2] = torch.randint(2, 30000, (3, 32), generator=g)
    mask[2] = True
    for K in (1, 7):
        ref = _oracle(cfg, sd, ids, mask, seqs, K, ml)
        got, gsc = _ours(cfg, sd, ids, mask, seqs, K, ml, dtype)
        want = ref["sequences"].numpy()
        if dtype == "fp32":
            assert np.array_equal(got, want)
            assert np.abs(gsc - ref["sequences_scores"].numpy()).max() < 2e-4
        else:
            assert got.shape == want.shape
            assert np.abs(gsc.reshape(3, K)[:, 0] - ref["sequences_scores"].numpy().reshape(3, K)[:, 0]).max() < 0.2
    # B = 1, N = 1
    got, _ = _ours(cfg, sd, ids[2:3, :1], mask[2:3, :1], seqs, 5, ml, dtype)
    ref = _oracle(cfg, sd, ids[2:3, :1], mask[2:3, :1], seqs, 5, ml)
    if dtype == "fp32":
        assert np.array_equal(got, ref["sequences"].numpy())


def test_capacity_and_argument_errors():
    cfg = GramConfig.tiny()
    sd = synth.make_state_dict(cfg, seed=1)
    m = GRAM(cfg, dtype="fp32", device="cuda:0")
    m.load_state_dict(sd)
    fn = prefix_allowed_tokens_fn(Trie([[0, 5, 1], [0, 6, 1]]))
    ids = torch.ones((1, 2, 8), dtype=torch.long).cuda()
    mask = torch.ones((1, 2, 8), dtype=torch.bool).cuda()
    # fewer reachable items (2) than beams (4): documented filler zone -- must not crash, finite hypotheses come first
    out = m.generate(ids, mask, 3, prefix_allowed_tokens_fn=fn, num_beams=4, num_return_sequences=4, return_dict_in_generate=True)
    sc = out["sequences_scores"].cpu().numpy()
    assert np.isfinite(sc[:2]).all() and out["sequences"].shape[0] == 4
    assert {tuple(r) for r in out["sequences"][:2].cpu().tolist()} == {(0, 5, 1), (0, 6, 1)}
    with pytest.raises(ValueError):
        m.generate(ids, mask[:, :1], 3, prefix_allowed_tokens_fn=fn, num_beams=2)            # shape mismatch
    with pytest.raises(ValueError):
        bad = prefix_allowed_tokens_fn(Trie([[0, cfg.vocab_size + 5, 1]]))                   # token outside the vocabulary
        m.generate(ids, mask, 3, prefix_allowed_tokens_fn=bad, num_beams=2)
    with pytest.raises(ValueError):                                                          # input id outside the vocabulary
        bad_ids = ids.clone()
        bad_ids[0, 0, 3] = cfg.vocab_size + 9
        m.generate(bad_ids.cpu(), mask.cpu(), 3, prefix_allowed_tokens_fn=fn, num_beams=2)
    # a start token that is not in the trie: every beam would be dead and the result all padding (HF raises in that
    # situation) -- the trie is refused
    with pytest.raises(ValueError):
        dead = prefix_allowed_tokens_fn(Trie([[7, 5, 1]]))
        m.generate(ids, mask, 3, prefix_allowed_tokens_fn=dead, num_beams=2, num_return_sequences=2, return_dict_in_generate=True)
    # re-uploading a grown trie replaces (and frees) the old arrays: same handle, new candidates reachable
    t2 = Trie([[0, 5, 1], [0, 6, 1]])
    fn2 = prefix_allowed_tokens_fn(t2)
    o1 = m.generate(ids, mask, 3, prefix_allowed_tokens_fn=fn2, num_beams=2, num_return_sequences=2, return_dict_in_generate=True)
    t2.add([0, 7, 1])
    o2 = m.generate(ids, mask, 3, prefix_allowed_tokens_fn=fn2, num_beams=3, num_return_sequences=3, return_dict_in_generate=True)
    assert set(o1["sequences"][:2, 1].tolist()) == {5, 6} and set(o2["sequences"][:3, 1].tolist()) == {5, 6, 7}


def test_large_synthetic_trie_beam50():
    """BASELINE configs[4] trie shape at 1/5 scale: 200,000 items, per-level branching [64, 25, 25, 5, 1],
    beam 50 (root fan-out >= K), T5-small weights: every returned id is an item, rankings are sorted and unique."""
    cfg = GramConfig.t5_small(max_seq_len=64, max_item_num=3)
    sd = synth.make_state_dict(cfg, seed=8)
    seqs = synth.make_item_sequences(200000, [64, 25, 25, 5, 1], cfg.vocab_size, seed=7)
    assert len(seqs) == 200000
    trie = Trie(seqs)
    csr = trie.to_csr()
    assert csr["max_fanout"] == 64 and len(trie.get([0])) == 64
    ids, mask = synth.make_user_batch(cfg, 6, (2, 3), 64, seed=33)
    m = GRAM(cfg, dtype="bf16", device="cuda:0")
    m.load_state_dict(sd)
    K, ml = 50, max(len(s) for s in seqs)
    out = m.generate(input_ids=torch.from_numpy(ids).cuda(), attention_mask=torch.from_numpy(mask).cuda(), max_length=ml,
                     prefix_allowed_tokens_fn=prefix_allowed_tokens_fn(trie), num_beams=K, num_return_sequences=K,
                     return_dict_in_generate=True)
    seq = out["sequences"].cpu().numpy()
    sc = out["sequences_scores"].cpu().numpy().reshape(6, K)
    items = {tuple(s) for s in seqs}
    for u in range(6):
        rows = [tuple(r[:list(r).index(1) + 1]) for r in seq[u * K:(u + 1) * K]]
        assert all(r in items for r in rows) and len(set(rows)) == K
    assert np.all(sc[:, :-1] >= sc[:, 1:]) and np.isfinite(sc).all()


def test_generate_from_a_csr_trie_file(tmp_path):
    """A trie loaded from its CSR file (gram_b200.formats) drives `generate` exactly like the nested-dict Trie."""
    from gram_b200 import formats
    case = CASES["tiny_lp"]
    sd, ids, mask, seqs, ml = case.build()
    m = GRAM(case.cfg, dtype="fp32", device="cuda:0")
    m.load_state_dict(sd)
    t = Trie(seqs)
    path = str(tmp_path / "trie.npz")
    formats.save_trie_csr(path, t)
    K = case.num_beams
    a = m.generate(ids.cuda(), mask.cuda(), ml, prefix_allowed_tokens_fn=prefix_allowed_tokens_fn(t), num_beams=K,
                   num_return_sequences=K, return_dict_in_generate=True, length_penalty=case.length_penalty)
    b = m.generate(ids.cuda(), mask.cuda(), ml, prefix_allowed_tokens_fn=prefix_allowed_tokens_fn(formats.load_trie_csr(path)),
                   num_beams=K, num_return_sequences=K, return_dict_in_generate=True, length_penalty=case.length_penalty)
    assert torch.equal(a["sequences"], b["sequences"]) and torch.equal(a["sequences_scores"], b["sequences_scores"])


def test_max_tokens_capacity_is_checked_on_the_device():
    """max_tokens sizes the workspace by VALID tokens: a batch below it runs (identical results to the uncapped
    engine), a batch above it is reported, not silently truncated."""
    case = CASES["tiny"]
    sd, ids, mask, seqs, ml = case.build()
    fn = prefix_allowed_tokens_fn(Trie(seqs))
    K = case.num_beams
    valid = int(mask.sum())
    ref = GRAM(case.cfg, dtype="fp32", device="cuda:0")
    ref.load_state_dict(sd)
    want = ref.generate(ids.cuda(), mask.cuda(), ml, prefix_allowed_tokens_fn=fn, num_beams=K, num_return_sequences=K)
    m = GRAM(case.cfg, dtype="fp32", device="cuda:0")
    m.load_state_dict(sd)
    m.configure(max_users=ids.shape[0], max_passages=ids.shape[1], max_seq_len=ids.shape[2], max_tokens=valid + 1)
    assert valid + 1 < ids.numel() // 1                       # the padded size would not fit
    got = m.generate(ids.cuda(), mask.cuda(), ml, prefix_allowed_tokens_fn=fn, num_beams=K, num_return_sequences=K)
    assert torch.equal(got, want)
    m.configure(max_tokens=valid - 1)
    with pytest.raises(ValueError, match="max_tokens"):
        m.generate(ids.cuda(), mask.cuda(), ml, prefix_allowed_tokens_fn=fn, num_beams=K, num_return_sequences=K)
    with pytest.raises(ValueError, match="max_tokens"):      # host tensors: reported by gram_generate itself
        m.generate(ids, mask, ml, prefix_allowed_tokens_fn=fn, num_beams=K, num_return_sequences=K)
    m.configure(max_tokens=0)
    assert torch.equal(m.generate(ids.cuda(), mask.cuda(), ml, prefix_allowed_tokens_fn=fn, num_beams=K, num_return_sequences=K), want)
