#!/bin/bash
# multi-GPU: full Beauty split through the runner, N ranks vs 1 (communicator warmed up before the clock)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out; mkdir -p $O; tag=mge$N
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR scripts/eval_full.py --batch 944 > $O/${tag}_eval_full.json 2> $O/${tag}_eval_full.err
timeout 600 $TR scripts/eval_full.py --batch 944 --item-cache > $O/${tag}_eval_full_cached.json 2> $O/${tag}_eval_full_cached.err
timeout 600 python scripts/eval_full.py --batch 944 > $O/${tag}_eval_full_1gpu.json 2> $O/${tag}_eval_full_1gpu.err
echo done > $O/${tag}_done
