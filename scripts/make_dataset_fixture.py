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

import numpy as np  # noqa: F401

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gram_b200 import formats  # noqa: E402

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
    out = formats.pack_dataset(os.path.join(d, ID_FILES[name]), os.path.join(d, "user_sequence.txt"),
                               os.path.join(d, "similar_item_sasrec.txt"), top_k=10)
    os.makedirs(DST, exist_ok=True)
    path = os.path.join(DST, f"{name}.npz")
    formats.save_packed(path, out)
    print(f"{name}: {len(out['item_asin'])} items, {len(out['pieces'])} pieces, id width {out['item_lex'].shape[1]}, "
          f"{len(out.get('user_off', [0])) - 1} users -> {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(ID_FILES)):
        pack(n)
