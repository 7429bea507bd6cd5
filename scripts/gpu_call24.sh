#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c24
timeout 400 python scripts/exp_hot_memory.py > $O/${tag}_hot_memory.log 2>&1
echo done > $O/${tag}_done
