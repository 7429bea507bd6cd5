"""Pack the reference's shipped preprocessed splits into compact fixtures under gram_b200/assets/.

Run in the build container (reads `/root/reference/rec_datasets/<Dataset>/`, which does not exist on
the GPU box):   python scripts/make_dataset_fixture.py

Inputs (formats: reference `rec_datasets/README.md:21-72`, loaders `src/utils/indexing.py:150-176,
236-246`):
  user_sequence.txt                    `user item1 ... itemn`               (absent for Yelp)
  item_generative_indexing_<type>.txt  `ASIN |piece|piece|...`
  similar_item_sasrec.txt              header line, then `anchor top1 .. top20` (absent for Sports, Yelp)
`item_plain_text.txt` is missing from every dataset (`.MISSING_LARGE_BLOBS`), so no metadata text and
no SentencePiece ids exist offline; gram_b200/data.py assigns surrogate token ids to the pieces.

Output npz: pieces (unicode, first-appearance order), item_asin, item_lex int32 [n_items, max_pieces]
(-1 padded piece indices), user_off / user_items (CSR of item indices per user, chronological),
similar int32 [n_items, 10] (-1 padded).
"""
import os
import sys

import numpy as np

SRC = os.environ.get("GRAM_REFERENCE_ROOT", "/root/reference") + "/rec_datasets"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gram_b200", "assets")

ID_FILES = {
    "Beauty": "item_generative_indexing_hierarchy_v1_c128_l7_len32768_split.txt",
    "Toys": "item_generative_indexing_hierarchy_v1_c32_l5_len32768_split.txt",
    "Sports": "item_generative_indexing_hierarchy_v1_c32_l7_len32768_split.txt",
    "Yelp": "item_generative_indexing_hierarchy_v1_c32_l9_len128_split.txt",
}


def pack(name):
    d = os.path.join(SRC, name)
    pieces, piece_idx = [], {}
    asins, lex = [], []
    with open(os.path.join(d, ID_FILES[name]), encoding="utf-8") as f:
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            asin, rest = line.split(" ", 1)
            ps = [p for p in rest.split("|") if p != ""]
            row = []
            for p in ps:
                if p not in piece_idx:
                    piece_idx[p] = len(pieces)
                    pieces.append(p)
                row.append(piece_idx[p])
            asins.append(asin)
            lex.append(row)
    item_idx = {a: i for i, a in enumerate(asins)}
    width = max(len(r) for r in lex)
    item_lex = np.full((len(lex), width), -1, dtype=np.int32)
    for i, r in enumerate(lex):
        item_lex[i, :len(r)] = r
    out = dict(pieces=np.array(pieces), item_asin=np.array(asins), item_lex=item_lex)
    up = os.path.join(d, "user_sequence.txt")
    if os.path.exists(up):
        off, items = [0], []
        with open(up) as f:
            for line in f:
                parts = line.split()
                if len(parts) < 2:
                    continue
                items.extend(item_idx[a] for a in parts[1:])
                off.append(len(items))
        out["user_off"] = np.asarray(off, dtype=np.int32)
        out["user_items"] = np.asarray(items, dtype=np.int32)
    sp = os.path.join(d, "similar_item_sasrec.txt")
    if os.path.exists(sp):
        sim = np.full((len(asins), 10), -1, dtype=np.int32)
        with open(sp) as f:
            for line in f:
                if line.startswith("anchor"):
                    continue
                parts = line.split()
                if not parts or parts[0] not in item_idx:
                    continue
                row = [item_idx[a] for a in parts[1:11] if a in item_idx]
                sim[item_idx[parts[0]], :len(row)] = row
        out["similar"] = sim
    os.makedirs(DST, exist_ok=True)
    path = os.path.join(DST, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {len(asins)} items, {len(pieces)} pieces, id width {width}, "
          f"{len(out.get('user_off', [0])) - 1} users -> {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(ID_FILES)):
        pack(n)
