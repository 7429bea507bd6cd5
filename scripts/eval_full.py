"""Full-test-set eval through the drop-in runner (BASELINE.json configs[1]/[2]): the whole leave-one-out split of a
dataset, beam 20, Recall/NDCG@5/10, wall-clock end to end (host collation + H2D + generate + D2H + decode + metrics).

    python scripts/eval_full.py [--dataset Beauty] [--batch 472] [--dtype bf16] [--users N]
    torchrun --nproc-per-node 2 ... scripts/eval_full.py      (user-sharded, one gather at the end)

Weights are random-init (no checkpoint exists offline) so the metric VALUES are chance level; what this shows is
the eval loop itself and that its metrics are identical at any batch size / world size.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from gram_b200 import GRAM, GramConfig, synth  # noqa: E402
from gram_b200.data import GramTestData  # noqa: E402
from gram_b200.runner import GramEvalLoader, GramRunner  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", default="Beauty")
    ap.add_argument("--batch", type=int, default=472)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--users", type=int, default=0, help="evaluate only the first N users (0 = all)")
    ap.add_argument("--no-pipeline", action="store_true")
    ap.add_argument("--item-cache", action="store_true", help="encode every item passage once (GRAM.generate_cached)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # communicator set-up (lazy: it happens at the first collective) belongs to process start-up, not to the eval
        warm_t = torch.zeros(1024, device=f"cuda:{local}", dtype=torch.int32)
        dist.all_reduce(warm_t)
        dist.all_gather_into_tensor(torch.empty(1024 * world, device=f"cuda:{local}", dtype=torch.int32), warm_t)
        torch.cuda.synchronize()
    data = GramTestData(args.dataset, synthetic_users=30000 if args.dataset == "Yelp" else 0)
    cfg = GramConfig.t5_small(max_seq_len=data.L, max_item_num=data.max_his)
    model = GRAM(cfg, dtype=args.dtype, device=f"cuda:{local}")
    model.load_state_dict(synth.make_state_dict(cfg, seed=0))
    model.configure(max_users=args.batch, max_beams=20, max_length=12, max_passages=data.max_his + 1, max_seq_len=data.L)
    model.user_limit = args.batch

    class Args:
        metrics = "hit@5,hit@10,ndcg@5,ndcg@10"
        beam_size = 20
        length_penalty = 1.0
        item_id_type = "split"

    users = list(range(args.users or data.n_users))
    runner = GramRunner(model, data.tokenizer, f"cuda:{local}", Args(), rank, world)
    loader = GramEvalLoader(data, args.batch, rank, world, users=users, item_cache=args.item_cache)
    # warm-up (engine creation, weight upload, trie upload) on one batch, outside the clock
    warm = GramEvalLoader(data, args.batch, 0, 1, users=users[:args.batch], item_cache=args.item_cache)
    runner.world_size = 1
    runner.test_dataset_task(warm, "warmup", pipeline=False)
    runner.world_size = world
    torch.cuda.synchronize()
    t0 = time.time()
    res = runner.test_dataset_task(loader, "test", pipeline=not args.no_pipeline)
    dt = time.time() - t0
    if rank == 0:
        print(json.dumps(dict(dataset=args.dataset, users=res["test_total"], n_gpus=world, item_cache=args.item_cache, batch=args.batch, dtype=args.dtype,
                              seconds=dt, users_per_sec=res["test_total"] / dt, generate_seconds_rank0=res["generate_seconds"],
                              metrics=res["metrics"], inputs="surrogate tokenizer, synthetic metadata, random-init weights")))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
