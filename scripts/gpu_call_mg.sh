#!/bin/bash
# round 2, multi-GPU call: bench.py --check (N-rank rankings == 1-rank rankings under NCCL), bench with the gather, full split through the runner
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out; mkdir -p $O; tag=mg$N
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --check > $O/${tag}_check.json 2> $O/${tag}_check.err
echo "check rc=$?" >> $O/${tag}_check.err
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --cpu-users 0 --no-item-cache > $O/${tag}_bench.json 2> $O/${tag}_bench.err
timeout 600 $TR scripts/eval_full.py --batch 944 > $O/${tag}_eval_full.json 2> $O/${tag}_eval_full.err
timeout 600 python scripts/eval_full.py --batch 944 > $O/${tag}_eval_full_1gpu.json 2> $O/${tag}_eval_full_1gpu.err
echo done > $O/${tag}_done
