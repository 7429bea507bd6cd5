"""CPU tests of the on-disk formats: text readers, packed dataset cache, CSR trie
file, prediction TSV."""
import os

import numpy as np
import pytest

from helpers import CASES
from gram_b200 import Trie, formats, prefix_allowed_tokens_fn
from gram_b200.data import GramTestData

REF = os.environ.get("GRAM_REFERENCE_ROOT", "/root/reference")


def _write_dataset(d):
    (d / "ids.txt").write_text("A1 |▁red|lip|stick\nB2 |▁red|nail\nC3 |▁blue|nail|▁file|x\n\n", encoding="utf-8")
    (d / "user_sequence.txt").write_text("u1 A1 B2 C3\nu2 C3 A1\nu3\n", encoding="utf-8")
    (d / "similar.txt").write_text("anchor top1 top2 top3\nA1 B2 C3 ZZ\nC3 A1\n", encoding="utf-8")
    (d / "item_plain_text.txt").write_text("A1 title: lipstick; brand: acme\nB2 title: nail polish\n", encoding="utf-8")


def test_text_readers_and_packed_cache(tmp_path):
    _write_dataset(tmp_path)
    asins, pieces = formats.read_item_index(str(tmp_path / "ids.txt"))
    assert asins == ["A1", "B2", "C3"] and pieces[0] == ["▁red", "lip", "stick"] and pieces[2][-1] == "x"
    assert formats.read_user_sequence(str(tmp_path / "user_sequence.txt")) == [("u1", ["A1", "B2", "C3"]), ("u2", ["C3", "A1"])]
    sim = formats.read_similar_items(str(tmp_path / "similar.txt"), top_k=2)
    assert sim == {"A1": ["B2", "C3"], "C3": ["A1"]}
    assert formats.read_item_plain_text(str(tmp_path / "item_plain_text.txt"))["A1"] == "title: lipstick; brand: acme"
    packed = formats.pack_dataset(str(tmp_path / "ids.txt"), str(tmp_path / "user_sequence.txt"), str(tmp_path / "similar.txt"),
                                  top_k=2)
    assert packed["pieces"].tolist() == ["▁red", "lip", "stick", "nail", "▁blue", "▁file", "x"]      # first appearance
    assert packed["item_lex"].tolist() == [[0, 1, 2, -1], [0, 3, -1, -1], [4, 3, 5, 6]]
    assert packed["user_off"].tolist() == [0, 3, 5] and packed["user_items"].tolist() == [0, 1, 2, 2, 0]
    assert packed["similar"].tolist() == [[1, 2], [-1, -1], [0, -1]]                                 # unknown ASIN dropped
    path = str(tmp_path / "toy.npz")
    formats.save_packed(path, packed)
    again = formats.load_packed(path)
    assert all(np.array_equal(packed[k], again[k]) for k in packed)
    # the eval data loads straight from the cache
    d = GramTestData("Beauty", packed_path=path, max_his=4, item_prompt_max_len=16, top_k_similar=2)
    assert d.n_items == 3 and d.n_users == 2
    b = d.collate([0, 1])
    assert b["item_text_ids"].shape[:2] == (2, 4) and b["target_items"] == [2, 0]
    assert d.collate_cached([0, 1])["item_index"].tolist() == [[1, 0, -1], [2, -1, -1]]              # most recent first


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "rec_datasets", "Toys")), reason="reference tree not mounted")
def test_shipped_assets_equal_a_fresh_parse_of_the_reference_files():
    d = os.path.join(REF, "rec_datasets", "Toys")
    packed = formats.pack_dataset(os.path.join(d, "item_generative_indexing_hierarchy_v1_c32_l5_len32768_split.txt"),
                                  os.path.join(d, "user_sequence.txt"), os.path.join(d, "similar_item_sasrec.txt"))
    asset = formats.load_packed(os.path.join(os.path.dirname(formats.__file__), "assets", "Toys.npz"))
    assert sorted(packed) == sorted(asset)
    assert all(np.array_equal(packed[k], asset[k]) for k in packed)


def test_csr_trie_file_round_trip(tmp_path):
    seqs = CASES["tiny_lp"].build()[3]
    t = Trie(seqs)
    path = str(tmp_path / "trie.npz")
    formats.save_trie_csr(path, t, start_token=0)
    c = formats.load_trie_csr(path)
    assert len(c) == len(t) and sorted(map(tuple, c)) == sorted(map(tuple, seqs))
    csr, ref = c.to_csr(0), t.to_csr(0)
    assert all(np.array_equal(csr[k], ref[k]) for k in ("child_offsets", "child_tokens", "child_nodes"))
    assert all(csr[k] == ref[k] for k in ("n_nodes", "n_edges", "root_node", "max_fanout"))
    for s in seqs:
        for j in range(len(s) + 1):
            assert c.get(s[:j]) == sorted(t.get(s[:j]))
    assert c.get([0, 31999]) == [] and c[[0]] == sorted(t.get([0]))
    fn = prefix_allowed_tokens_fn(c)
    assert fn.candidate_trie is c
    with pytest.raises(ValueError):
        c.to_csr(5)
    # a real item-ID trie
    d = GramTestData("Toys")
    cands = d.encoded_candidates()
    big = Trie(cands)
    formats.save_trie_csr(path, big)
    loaded = formats.load_trie_csr(path)
    assert len(loaded) == len(cands) and loaded.to_csr()["n_nodes"] == big.to_csr()["n_nodes"]
    assert loaded.get(cands[17][:3]) == sorted(big.get(cands[17][:3]))


def test_prediction_tsv_round_trip(tmp_path):
    rows = [(0, "red lipstick", ["red lipstick", "blue nail"], [-0.5, -1.25], 0),
            (7, "nail file", ["x", "y"], [-2.0, -3.5], -1)]
    names = ["hit@5", "hit@10", "ndcg@5", "ndcg@10"]
    per_user = [[1, 1, 1.0, 1.0], [0, 0, 0.0, 0.0]]
    metrics = {"hit@5": 0.5, "hit@10": 0.5, "ndcg@5": 0.5, "ndcg@10": 0.5}
    path = str(tmp_path / "pred.tsv")
    formats.write_predictions_tsv(path, rows, metrics, per_user, names)
    lines = open(path, encoding="utf-8").read().splitlines()
    assert lines[0] == "idx\tH@5\tH@10\tNDCG@5\tNDCG@10\tgold\tpred\tscores"      # single_runner_gram.py:580-588
    assert lines[1].split("\t")[-2] == "red lipstick||blue nail"
    got, m = formats.read_predictions_tsv(path)
    assert m == metrics and [r["idx"] for r in got] == ["u0", "u7"]
    assert got[0]["pred"] == rows[0][2] and got[1]["scores"] == rows[1][3] and got[0]["per_user"]["H@5"] == 1.0
