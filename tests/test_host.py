"""CPU tests of the host logic and of the C-ABI library surface (no compute without a GPU)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import CASES, ROOT
from gram_b200 import GRAM, GramConfig, Trie, _cabi, prefix_allowed_tokens_fn, synth
from gram_b200.data import SEPARATOR_IDS, GramTestData
from gram_b200.generation_trie import csr_children, csr_walk, exact_match
from gram_b200.weights import canonical_name, canonicalize, relative_position_buckets
from oracle.gram_oracle import OracleTrie, relative_position_bucket


# ---- C ABI ------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    lib = _cabi.load_library()
    header = open(os.path.join(ROOT, "include", "gram_b200.h")).read()
    declared = set(re.findall(r"\b(gram_[a-z_0-9]+)\s*\(", header))
    declared -= {"gram_b200"}
    assert declared == set(_cabi.EXPORTED_SYMBOLS), declared ^ set(_cabi.EXPORTED_SYMBOLS)
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert b"sm_100a" in lib.gram_version()
    out = subprocess.run(["nm", "-D", _cabi.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (gram_[a-z_0-9]+)", out))
    assert declared <= exported


def test_config_struct_layout_matches_header():
    header = open(os.path.join(ROOT, "include", "gram_b200.h")).read()
    body = header[header.index("typedef struct gram_config {"):header.index("} gram_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)      # comments may span lines
    fields = []
    for line in body.splitlines()[1:]:
        line = line.strip().rstrip(";")
        if not line:
            continue
        typ, names = line.split(None, 1)
        for n in names.split(","):
            fields.append((n.strip(), typ))
    assert [f[0] for f in fields] == [f[0] for f in _cabi.GramConfigC._fields_]
    ctype = {"int32_t": C.c_int32, "int64_t": C.c_int64, "float": C.c_float}
    assert [ctype[f[1]] for f in fields] == [f[1] for f in _cabi.GramConfigC._fields_]


def test_stats_struct_layout_matches_header():
    header = open(os.path.join(ROOT, "include", "gram_b200.h")).read()
    body = header[header.index("typedef struct gram_stats {"):header.index("} gram_stats;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [line.strip().rstrip(";").split()[-1] for line in body.splitlines()[1:] if line.strip()]
    assert names == [f[0] for f in _cabi.GramStatsC._fields_]
    assert all(f[1] is C.c_int64 for f in _cabi.GramStatsC._fields_)


def test_flag_values_match_header():
    header = open(os.path.join(ROOT, "include", "gram_b200.h")).read()
    flags = dict(re.findall(r"\b(GRAM_FLAG_[A-Z0-9_]+)\s*=\s*(\d+)", header))
    assert flags, "no flags parsed"
    for name, value in flags.items():
        assert getattr(_cabi, name) == int(value), name


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly (never route through the oracle)."""
    case = CASES["tiny"]
    sd, ids, mask, seqs, ml = case.build()
    m = GRAM(case.cfg, dtype="fp32")
    m.load_state_dict(sd)
    fn = prefix_allowed_tokens_fn(Trie(seqs))
    with pytest.raises(_cabi.GramLibraryError):
        m.generate(ids, mask, ml, prefix_allowed_tokens_fn=fn, num_beams=2, num_return_sequences=2)
    lib = _cabi.load_library()
    cc = _cabi.GramConfigC(vocab_size=384, d_model=64, d_kv=16, d_ff=128, num_layers=1, num_decoder_layers=1, num_heads=4,
                           rel_buckets=32, rel_max_distance=128, ln_eps=1e-6, pad_id=0, eos_id=1, start_id=0,
                           tie_word_embeddings=1, n_positions=5, dtype=0, device=0, max_users=1, max_passages=2,
                           max_seq_len=8, max_beams=2, max_length=4, max_tokens=0, flags=0)
    hp = C.c_void_p()
    assert lib.gram_create(C.byref(cc), C.byref(hp)) != 0
    assert b"no CUDA device" in lib.gram_last_error(None)
    src = "".join(open(os.path.join(ROOT, "gram_b200", f)).read() for f in os.listdir(os.path.join(ROOT, "gram_b200"))
                  if f.endswith(".py"))
    assert "import oracle" not in src and "from oracle" not in src


# ---- weights -----------------------------------------------------------------------------------------
def test_state_dict_names_cover_the_reference_keys():
    cfg = GramConfig.tiny()
    sd = synth.make_state_dict(cfg, seed=1)
    canon = canonicalize(sd)
    need = {"shared", "lm_head", "pos_emb", "enc.final_ln", "dec.final_ln", "enc.0.rel_bias", "dec.0.rel_bias"}
    for i in range(cfg.num_layers):
        need |= {f"enc.{i}.{w}" for w in ("q", "k", "v", "o", "wi", "wo", "ln0", "ln1")}
    for i in range(cfg.num_decoder_layers):
        need |= {f"dec.{i}.{w}" for w in ("q", "k", "v", "o", "cq", "ck", "cv", "co", "wi", "wo", "ln0", "ln1", "ln2")}
    assert set(canon) == need
    # wrapped (FiD) and plain-T5 spellings map to the same names; aliases are ignored
    assert canonical_name("encoder.encoder.block.3.module.layer.0.SelfAttention.q.weight") == "enc.3.q"
    assert canonical_name("encoder.block.3.layer.0.SelfAttention.q.weight") == "enc.3.q"
    assert canonical_name("decoder.block.1.layer.1.EncDecAttention.k.weight") == "dec.1.ck"
    assert canonical_name("decoder.block.1.layer.2.DenseReluDense.wo.weight") == "dec.1.wo"
    assert canonical_name("decoder.embed_tokens.weight") is None
    # deterministic, platform-independent generator
    again = synth.make_state_dict(cfg, seed=1)
    assert all(torch.equal(sd[k], again[k]) for k in sd)
    assert abs(float(sd["shared.weight"].std()) - 1.0) < 0.05


def test_relative_position_bucket_tables():
    cfg = GramConfig.t5_small()
    enc, dec = relative_position_buckets(cfg, 128, 12)
    L = 128
    ctx = torch.arange(L)[:, None]
    mem = torch.arange(L)[None, :]
    full = relative_position_bucket(mem - ctx, True, 32, 128)
    lut = torch.from_numpy(enc.astype(np.int64))
    assert torch.equal(lut[(mem - ctx) + L - 1], full)
    causal = relative_position_bucket(mem[:, :12] - ctx[:12], False, 32, 128)
    for i in range(12):
        for j in range(i + 1):
            assert causal[i, j].item() == dec[i - j]


# ---- trie ----------------------------------------------------------------------------------------------
def test_trie_api_and_csr_equal_the_dict_walk():
    seqs = CASES["tiny_lp"].build()[3]
    t = Trie(seqs)
    o = OracleTrie(seqs)
    assert t.trie_dict == o.trie_dict and len(t) == len(seqs)
    assert sorted(map(tuple, t)) == sorted(map(tuple, seqs))
    csr = t.to_csr()
    assert csr["n_edges"] == csr["n_nodes"] - 1 and csr["root_node"] > 0
    for s in seqs:
        for j in range(len(s) + 1):
            assert csr_children(csr, csr_walk(csr, s[:j])) == sorted(t.get(s[:j])) and t.get(s[:j]) == o.get(s[:j])
    assert csr_children(csr, csr_walk(csr, [0, 31999])) == [] == t.get([0, 31999])
    t.add([0, 5, 6, 1])
    assert t.to_csr()["n_nodes"] > csr["n_nodes"]          # cache invalidated by add()
    assert t[[0, 5]] == [6]
    fn = prefix_allowed_tokens_fn(t)
    assert fn.candidate_trie is t and fn(0, torch.tensor([0, 5, 6])) == [1]
    assert exact_match(["a", "b", "c", "d"], ["b", "x"], 2) == 1
    assert Trie.load_from_dict(t.trie_dict).len == len(t)


@pytest.mark.parametrize("dataset", ["Beauty", "Toys", "Sports", "Yelp"])
def test_csr_masks_bit_exact_on_every_shipped_prefix(dataset):
    """'Trie masks bit-exact': for EVERY prefix of every shipped item id the CSR children are exactly
    `Trie.get(prefix)` (SURVEY.md section 8(c) pin (i))."""
    d = GramTestData(dataset, synthetic_users=4 if dataset == "Yelp" else 0)
    cands = d.encoded_candidates()
    t = Trie(cands)
    csr = t.to_csr()
    stats = {"Beauty": (12101, 75892, 108, 255), "Toys": (11924, None, 30, 186), "Sports": (18357, None, 28, 175),
             "Yelp": (20033, None, 21, 263)}[dataset]
    assert len(cands) == stats[0]
    if stats[1]:
        assert csr["n_nodes"] == stats[1]
    assert len(t.get([0])) == stats[2] and csr["max_fanout"] == stats[3]
    seen = set()
    for s in cands:
        node = 0
        for j in range(len(s) + 1):
            key = tuple(s[:j])
            if key in seen:
                node = csr_walk(csr, s[:j]) if j else 0
                continue
            seen.add(key)
            node = csr_walk(csr, s[:j])
            assert csr_children(csr, node) == sorted(t.get(s[:j]))      # same allowed SET, token-ascending


# ---- data ------------------------------------------------------------------------------------------------
def test_collator_contract():
    d = GramTestData("Beauty")
    assert d.n_users == 22363 and d.n_items == 12101
    users = [0, 1, 2, 3, 17, 100]
    b = d.collate(users)
    ids, mask = b["item_text_ids"], b["item_text_masks"]
    hist = [d.split(u)[0] for u in users]
    N = min(max(len(h) for h in hist) + 1, d.max_his) + 1
    assert ids.shape == mask.shape == (len(users), N, mask.sum(-1).max())
    assert ids.dtype == np.int64 and mask.dtype == bool
    for r, h in enumerate(hist):
        n_real = 1 + min(len(h), N - 1)
        assert not mask[r, n_real:].any() and (ids[r][~mask[r]] == 0).all()
        for p in range(n_real):
            ln = int(mask[r, p].sum())
            assert mask[r, p, :ln].all() and ids[r, p, ln - 1] == 1          # prefix mask, EOS last
    assert not np.isin(ids, SEPARATOR_IDS).any()
    # leave-one-out, most recent first, capped at max_his
    items = d.user_items[d.user_off[17]:d.user_off[18]]
    h, tgt = d.split(17)
    assert tgt == items[-1] and list(h) == list(items[:-1][-d.max_his:][::-1])
    assert b["target_ids"][4] == [0] + d.item_tok[tgt].tolist() + [1]
    cands = d.encoded_candidates()
    assert b["target_ids"][4] in cands and {len(c) for c in cands} == {9, 10}
    assert d.tokenizer.batch_decode([b["target_ids"][4]])[0] == d.tokenizer.decode(cands[tgt])
    v = GramTestData("Beauty", mode="validation")
    assert v.split(17)[1] == items[-2]


def test_gram_is_a_module_that_ddp_can_wrap(tmp_path):
    """The reference wraps the model in DistributedDataParallel and calls `self.model_rec.module.generate`
    (src/runner/distributed_runner_gram.py:47-52,775): GRAM must be an nn.Module DDP accepts, wrapped or not."""
    import torch
    import torch.distributed as dist
    from gram_b200 import GRAM, GramConfig
    m = GRAM(GramConfig.tiny())
    assert isinstance(m, torch.nn.Module) and not m.training
    assert m.module is m and m.eval() is m
    with pytest.raises(NotImplementedError):
        m.train()
    dist.init_process_group("gloo", init_method=f"file://{tmp_path}/pg", rank=0, world_size=1)
    try:
        ddp = torch.nn.parallel.DistributedDataParallel(m)
        assert ddp.module is m and ddp.module.generate.__func__ is GRAM.generate
        assert ddp.module.module is m                      # `.module` also on the unwrapped model
    finally:
        dist.destroy_process_group()


def test_fast_relevance_rows_equal_the_string_path():
    """runner.rel_rows_fast + _ItemStrings (row bytes -> item -> decoded-string class) == evaluate.rel_results on decoded
    strings, including two token paths that decode to ONE string, unknown rows, -inf fillers, ties and NaN scores."""
    from gram_b200 import evaluate
    from gram_b200.runner import _ItemStrings, rel_rows_fast

    class Tok:
        def batch_decode(self, rows, skip_special_tokens=True):
            names = {2: "a", 3: "b", 4: "ab", 5: "c", 6: "bc"}
            return ["".join(names.get(int(t), f"<{int(t)}>") for t in r if int(t) > 1) for r in rows]

    tok = Tok()
    cands = [[0, 2, 6, 1], [0, 4, 5, 1], [0, 2, 3, 1], [0, 5, 1], [0, 3, 5, 5, 1]]       # "abc", "abc", "ab", "c", "bcc"
    W, G = 5, 4
    st = _ItemStrings(tok, cands, W)
    rng = np.random.default_rng(0)
    n = 64
    rows = np.zeros((n * G, W), dtype=np.int64)
    for i in range(n * G):
        c = cands[rng.integers(0, len(cands))] if rng.random() < 0.85 else [0, 7, 1]       # sometimes a non-item row
        rows[i, :len(c)] = c
    rows[5] = 0                                                                            # an all-pad filler row
    gold = [cands[rng.integers(0, len(cands))] for _ in range(n)]
    scores = rng.integers(-5, 0, size=n * G).astype(np.float32)                            # many ties
    scores[7] = np.nan
    scores[5] = -np.inf
    want = evaluate.rel_results(tok.batch_decode(rows), tok.batch_decode(gold), torch.from_numpy(scores), G)
    g = np.zeros((n, W), dtype=np.int32)
    for i, c in enumerate(gold):
        g[i, :len(c)] = c
    got = rel_rows_fast(st.classes(rows), st.classes(g), scores, G)
    assert np.array_equal(got, np.asarray(want, dtype=np.uint8))
    assert (got.sum(1) > 1).any()            # rows with the same string twice exist in this sample
