#!/bin/bash
# round 2, call 13: per-launch in-step times (GRAM_PROF_DUMP)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c13
rm -f $O/${tag}_dump.txt
GRAM_PROF_DUMP=$O/${tag}_dump.txt timeout 300 python bench.py --steps 3 --warmup 2 --no-item-cache --cpu-users 0 --no-e2e > $O/${tag}_bench.json 2> $O/${tag}_bench.err
echo done > $O/${tag}_done
