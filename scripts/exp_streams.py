"""Experiment: N independent engine handles on N CUDA streams, batches issued round-robin (device-resident inputs)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gram_b200 import GRAM

nstreams = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
wl = bench.make_workload("beauty")
data, cfg, sd, max_length, trie = wl.data, wl.cfg, wl.sd, wl.max_length, wl.trie
B, K, W, S = (int(sys.argv[2]) if len(sys.argv) > 2 else 472), 20, 4, 12
models, streams, outs = [], [], []
for i in range(nstreams):
    m = GRAM(cfg, dtype="bf16", device=dev)
    m.load_state_dict(sd)
    m.configure(max_users=B, max_beams=K, max_length=max_length, max_passages=data.max_his + 1, max_seq_len=data.L,
                max_tokens=int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    models.append(m); streams.append(torch.cuda.Stream(dev))
    outs.append((torch.zeros((B * K, max_length), dtype=torch.int64, device=dev), torch.zeros((B * K,), device=dev),
                 torch.zeros((1,), dtype=torch.int32, device=dev)))
ins = []
for s in range(W + S):
    b = data.collate(wl.users(s, B, 0, 1))
    ins.append((torch.from_numpy(b["item_text_ids"]).to(dev), torch.from_numpy(b["item_text_masks"]).to(dev)))
def step(i):
    j = i % nstreams
    with torch.cuda.stream(streams[j]):
        models[j].generate_into(ins[i][0], ins[i][1], max_length, trie, K, K, 1.0, *outs[j])
for i in range(W): step(i)
torch.cuda.synchronize()
t0 = time.time()
for i in range(W, W + S): step(i)
torch.cuda.synchronize()
dt = time.time() - t0
print(f"streams={nstreams} batch={B}: {B * S / dt:.0f} users/s, {dt / S * 1e3:.2f} ms/step")
