"""Parity report at the headline configuration (Beauty, T5-small, beam 20, 12,101-item trie): the CUDA path against
the CPU oracle on N real test users.  Test infrastructure (imports the oracle); writes a JSON summary.

    python tests/parity_report.py [--users 32] [--out profiles/r1_parity_beauty.json]

fp32: ranked item ids must be identical (north star); the smallest gap between adjacent oracle scores is reported so
near-ties can be told apart from real mismatches.  bf16: top-10 overlap and score error.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from gram_b200 import GRAM, GramConfig, Trie, prefix_allowed_tokens_fn, synth  # noqa: E402
from gram_b200.data import GramTestData  # noqa: E402
from oracle.gram_oracle import OracleGRAM, OracleTrie  # noqa: E402

K = 20


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=32)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    data = GramTestData("Beauty")
    cfg = GramConfig.t5_small(max_seq_len=data.L, max_item_num=data.max_his)
    sd = synth.make_state_dict(cfg, seed=0)
    cands = data.encoded_candidates()
    ml = max(len(c) for c in cands)
    trie = Trie(cands)
    fn = prefix_allowed_tokens_fn(trie)
    # a spread of history lengths: every 97th user of the test split
    users = [(i * 97) % data.n_users for i in range(args.users)]
    batch = data.collate(users)
    ids, mask = torch.from_numpy(batch["item_text_ids"]), torch.from_numpy(batch["item_text_masks"])
    out = {}
    for dtype in ("fp32", "bf16"):
        m = GRAM(cfg, dtype=dtype, device="cuda:0")
        m.load_state_dict(sd)
        o = m.generate(input_ids=ids.cuda(), attention_mask=mask.cuda(), max_length=ml, prefix_allowed_tokens_fn=fn,
                       num_beams=K, num_return_sequences=K, return_dict_in_generate=True)
        out[dtype] = (o["sequences"].cpu().numpy(), o["sequences_scores"].cpu().numpy())
    torch.set_num_threads(os.cpu_count() or 1)
    ora = OracleGRAM(cfg, sd)
    otrie = OracleTrie(cands)
    exact, gaps, overlaps, top1, score_err32, score_err16 = 0, [], [], 0, [], []
    mism = []
    t0 = time.time()
    for i, u in enumerate(users):
        b1 = data.collate([u])                         # the reference evaluates one user per call
        ref = ora.generate(torch.from_numpy(b1["item_text_ids"]), torch.from_numpy(b1["item_text_masks"]), ml, otrie, K, K, 1.0)
        want = ref["sequences"].numpy()
        wsc = ref["sequences_scores"].numpy()
        w = want.shape[1]
        got32 = out["fp32"][0][i * K:(i + 1) * K, :w]
        got16 = out["bf16"][0][i * K:(i + 1) * K, :w]
        same = bool(np.array_equal(got32, want))
        exact += same
        gap = float(np.abs(np.diff(wsc)).min())
        gaps.append(gap)
        if not same:
            mism.append(dict(user=int(u), min_gap=gap, first_diff_rank=int(np.argmax((got32 != want).any(axis=1)))))
        score_err32.append(float(np.abs(out["fp32"][1][i * K:(i + 1) * K] - wsc).max()))
        score_err16.append(float(np.abs(out["bf16"][1][i * K:(i + 1) * K] - wsc).max()))
        a = {tuple(r) for r in want[:10].tolist()}
        c = {tuple(r) for r in got16[:10].tolist()}
        overlaps.append(len(a & c) / 10)
        top1 += bool(np.array_equal(got16[0], want[0]))
    rep = dict(config="Beauty, T5-small random-init tied weights (seed 0), beam 20, 12101-item trie, max_length %d" % ml,
               users=len(users), history_lengths=[int(len(data.split(u)[0])) for u in users],
               fp32_ranked_ids_identical=f"{exact}/{len(users)}", fp32_mismatches=mism,
               fp32_max_score_abs_err=max(score_err32), oracle_min_adjacent_score_gap=min(gaps),
               bf16_top10_overlap_mean=float(np.mean(overlaps)), bf16_top10_overlap_min=float(np.min(overlaps)),
               bf16_top1_identical=f"{top1}/{len(users)}", bf16_max_score_abs_err=max(score_err16),
               oracle_seconds=time.time() - t0,
               note="oracle = CPU restatement, bit-identical to the reference modules (tests/test_oracle.py); the CUDA path "
                    "ran all users in ONE batched call, the oracle one user per call as the reference does")
    print(json.dumps(rep, indent=1))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rep, f, indent=1)


if __name__ == "__main__":
    main()
