#!/bin/bash
# round 2, call 36: ncu --set full capture of the vocabulary head at a full decode step (63rd gemm_tc launch of a generate call)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
P="python bench.py --steps 1 --warmup 1 --no-item-cache --cpu-users 0 --no-e2e"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 62 -c 1 -o $O/prof_lmhead_r2d -f $P > $O/c36_ncu_head.log 2>&1
echo done > $O/c36_done
