"""Latency of one `GRAM.generate` call at small user batches, eager launches vs GRAM_FLAG_CUDA_GRAPH (the whole call replayed
as one CUDA graph).  Beauty, T5-small, beam 20, bf16; device tensors in, device tensors out, one synchronisation per call.

    python scripts/latency_small_batch.py [--batches 1,4,16,64,256] [--iters 40]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
from gram_b200 import GRAM, _cabi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,4,16,64,256")
    ap.add_argument("--iters", type=int, default=40)
    args = ap.parse_args()
    wl = bench.make_workload("beauty")
    dev = torch.device("cuda", 0)
    K, ml = wl.K, wl.max_length
    rows = []
    for B in [int(x) for x in args.batches.split(",")]:
        ins = []
        for s in range(args.iters + 3):
            ids, mask = wl.collate(wl.users(s, B, 0, 1))
            full = torch.zeros((B, wl.N, wl.L), dtype=torch.int64)            # fixed shape: one graph per batch size
            fm = torch.zeros((B, wl.N, wl.L), dtype=torch.bool)
            full[:, :ids.shape[1], :ids.shape[2]] = torch.from_numpy(ids)
            fm[:, :mask.shape[1], :mask.shape[2]] = torch.from_numpy(mask)
            ins.append((full.to(dev), fm.to(dev)))
        res = {}
        for label, flags in (("eager", 0), ("graph", _cabi.GRAM_FLAG_CUDA_GRAPH)):
            m = GRAM(wl.cfg, dtype="bf16", device=dev, flags=flags)
            m.load_state_dict(wl.sd)
            m.configure(max_users=B, max_beams=K, max_length=ml, max_passages=wl.N, max_seq_len=wl.L)
            out_seq = torch.zeros((B * K, ml), dtype=torch.int64, device=dev)
            out_sc = torch.zeros((B * K,), device=dev)
            out_w = torch.zeros((1,), dtype=torch.int32, device=dev)
            for i in range(3):
                m.generate_into(*ins[i], ml, wl.trie, K, K, 1.0, out_seq, out_sc, out_w)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(3, 3 + args.iters):
                m.generate_into(*ins[i], ml, wl.trie, K, K, 1.0, out_seq, out_sc, out_w)
                torch.cuda.synchronize()
            res[label] = (time.perf_counter() - t0) / args.iters * 1e3
            res[label + "_launches"] = m.stats()["launches"]
            del m
        rows.append(dict(users=B, eager_ms=round(res["eager"], 3), graph_ms=round(res["graph"], 3),
                         speedup=round(res["eager"] / res["graph"], 3), launches=res["eager_launches"]))
        print(rows[-1], flush=True)
    print(json.dumps(dict(workload="Beauty T5-small beam 20 bf16, one generate call per batch of users, wall clock incl. sync",
                          rows=rows)))


if __name__ == "__main__":
    main()
