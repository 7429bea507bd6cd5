"""Parity of the bf16 encoder with a bf16 residual stream (the default) next to the fp32 stream (GRAM_FLAG_FP32_RESID), both against
the fp32 oracle: fused memory at 21 x 128 tokens per user, and rankings / scores of Beauty test users (T5-small, beam 20).

    python scripts/parity_bf16_resid.py [n_users]        -> one JSON line
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gram_b200 import GRAM, GramConfig, Trie, _cabi, prefix_allowed_tokens_fn, synth  # noqa: E402
from gram_b200.data import GramTestData  # noqa: E402
from oracle.gram_oracle import OracleGRAM, OracleTrie  # noqa: E402   (checker only)

K20 = 20
VARIANTS = {"fp32_stream": _cabi.GRAM_FLAG_FP32_RESID, "bf16_stream": 0}


def main():
    n_users = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    torch.set_num_threads(os.cpu_count() or 1)
    rep = {}
    # ---- encoder memory at the headline shape
    cfg = GramConfig.t5_small(max_seq_len=128, max_item_num=20)
    sd = synth.make_state_dict(cfg, seed=0)
    ids, mask = synth.make_user_batch(cfg, 3, (21, 21), 128, seed=17, full=True)
    ids, mask = torch.from_numpy(ids), torch.from_numpy(mask)
    want = OracleGRAM(cfg, sd).encode(ids, mask)
    fm = mask.view(3, -1)
    for name, flags in VARIANTS.items():
        m = GRAM(cfg, dtype="bf16", device="cuda:0", flags=flags)
        m.load_state_dict(sd)
        mem = m.encode(ids.cuda(), mask.cuda()).cpu().float()
        d = (mem[fm] - want[fm])
        rep[name] = {"memory_max_rel": float(d.abs().max() / want[fm].abs().max()),
                     "memory_rms_rel": float(d.pow(2).mean().sqrt() / want[fm].pow(2).mean().sqrt())}
        del m
    # ---- rankings on Beauty
    data = GramTestData("Beauty")
    cfg = GramConfig.t5_small(max_seq_len=data.L, max_item_num=data.max_his)
    sd = synth.make_state_dict(cfg, seed=0)
    cands = data.encoded_candidates()
    ml = max(len(c) for c in cands)
    fn = prefix_allowed_tokens_fn(Trie(cands))
    users = [(i * 97) % data.n_users for i in range(n_users)]
    batch = data.collate(users)
    ids, mask = torch.from_numpy(batch["item_text_ids"]), torch.from_numpy(batch["item_text_masks"])
    got = {}
    for name, flags in VARIANTS.items():
        m = GRAM(cfg, dtype="bf16", device="cuda:0", max_users=n_users, flags=flags)
        m.load_state_dict(sd)
        o = m.generate(input_ids=ids.cuda(), attention_mask=mask.cuda(), max_length=ml, prefix_allowed_tokens_fn=fn,
                       num_beams=K20, num_return_sequences=K20, return_dict_in_generate=True)
        got[name] = (o["sequences"].cpu().numpy(), o["sequences_scores"].cpu().numpy())
        del m
    ora, otrie = OracleGRAM(cfg, sd), OracleTrie(cands)
    acc = {n: dict(top1=0, ov10=[], ov20=[], err=0.0) for n in VARIANTS}
    for i, u in enumerate(users):
        b1 = data.collate([u])
        ref = ora.generate(torch.from_numpy(b1["item_text_ids"]), torch.from_numpy(b1["item_text_masks"]), ml, otrie, K20, K20, 1.0)
        w_, wsc = ref["sequences"].numpy(), ref["sequences_scores"].numpy()
        w = w_.shape[1]
        for n in VARIANTS:
            g = got[n][0][i * K20:(i + 1) * K20, :w]
            a = acc[n]
            a["top1"] += bool(np.array_equal(g[0], w_[0]))
            a["ov10"].append(len({tuple(r) for r in w_[:10].tolist()} & {tuple(r) for r in g[:10].tolist()}) / 10)
            a["ov20"].append(len({tuple(r) for r in w_.tolist()} & {tuple(r) for r in g.tolist()}) / 20)
            a["err"] = max(a["err"], float(np.abs(got[n][1][i * K20:(i + 1) * K20] - wsc).max()))
    for n, a in acc.items():
        rep[n].update({"users": n_users, "top1_identical": a["top1"], "top10_overlap_mean": float(np.mean(a["ov10"])),
                       "top10_overlap_min": float(np.min(a["ov10"])), "top20_overlap_mean": float(np.mean(a["ov20"])),
                       "max_score_err": a["err"]})
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
