"""BASELINE.json configs[4] (the scale stressor) on one GPU: T5-base random-init (tied head, seed 0), 32 passages x 256
tokens all valid (S = 8,192), beam 50, synthetic 1,000,000-item trie (per-level branching [64, 25, 25, 5, 5, 1], seed 7).
Not a bench line (bench.py measures configs[1]); run once per round and quoted in DESIGN.md.

    python scripts/scale_config5.py [--users 128] [--steps 3] [--items 1000000]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from gram_b200 import GRAM, GramConfig, Trie, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=128)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--items", type=int, default=1000000)
    args = ap.parse_args()
    K, N, L = 50, 32, 256
    cfg = GramConfig.t5_base(max_seq_len=L, max_item_num=N - 1)
    t0 = time.time()
    seqs = synth.make_item_sequences(args.items, [64, 25, 25, 5, 5, 1], cfg.vocab_size, seed=7)
    trie = Trie(seqs)
    csr = trie.to_csr()
    t_trie = time.time() - t0
    ml = max(len(s) for s in seqs)
    sd = synth.make_state_dict(cfg, seed=0)
    m = GRAM(cfg, dtype="bf16", device="cuda:0")
    m.load_state_dict(sd)
    B = args.users
    m.configure(max_users=B, max_beams=K, max_length=ml, max_passages=N, max_seq_len=L)
    rng = np.random.default_rng(2023)
    batches = []
    for s in range(args.steps + 1):
        ids = rng.integers(2, cfg.vocab_size - 28, size=(B, N, L)).astype(np.int64)
        ids[:, :, -1] = 1
        batches.append((torch.from_numpy(ids).cuda(), torch.ones((B, N, L), dtype=torch.bool, device="cuda")))
    out_seq = torch.zeros((B * K, ml), dtype=torch.int64, device="cuda")
    out_sc = torch.zeros((B * K,), dtype=torch.float32, device="cuda")
    out_w = torch.zeros((1,), dtype=torch.int32, device="cuda")
    m.generate_into(*batches[0], ml, trie, K, K, 1.0, out_seq, out_sc, out_w)         # warm-up
    torch.cuda.synchronize()
    m.profile_begin(None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(1, args.steps + 1):
        m.generate_into(*batches[s], ml, trie, K, K, 1.0, out_seq, out_sc, out_w)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    prof = m.profile_end()
    seq = out_seq.cpu().numpy()
    sc = out_sc.cpu().numpy().reshape(B, K)
    items = {tuple(s) for s in seqs[:200000]} if args.items > 200000 else {tuple(s) for s in seqs}
    first = [tuple(r[:list(r).index(1) + 1]) for r in seq[:K]]
    S = N * L
    d, HD, F, V, Le, Ld, T = cfg.d_model, cfg.inner_dim, cfg.d_ff, cfg.vocab_size, cfg.num_layers, cfg.num_decoder_layers, ml - 1
    flops = B * (S * Le * (8 * d * HD + 4 * d * F + 4 * L * HD) + S * Ld * 4 * d * HD
                 + T * K * (Ld * (12 * d * HD + 4 * d * F) + 2 * d * V))
    print(json.dumps(dict(config="configs[4]: T5-base random-init, 32 x 256 tokens all valid, beam 50, %d-item trie (%d nodes, "
                                 "root fan-out %d), max_length %d" % (len(seqs), csr["n_nodes"], len(trie.get([0])), ml),
                          users_per_step=B, ms_per_step=ms, users_per_sec=B / (ms / 1000.0),
                          algorithmic_tflops=flops / (ms / 1000.0) / 1e12,
                          kernel_classes={c: round(v["ms"] / args.steps, 2) for c, v in prof.items()},
                          sorted_scores=bool(np.all(sc[:, :-1] >= sc[:, 1:])), unique_rankings=len(set(first)) == K,
                          trie_build_seconds=t_trie, workspace_gb=m.stats()["workspace_bytes"] / 1e9)))


if __name__ == "__main__":
    main()
