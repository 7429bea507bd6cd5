#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c22
for d in 0 1 2 3; do
  GRAM_XATTN_DEBUG=$d timeout 200 python scripts/exp_xattn_hot.py > $O/${tag}_dbg$d.log 2>&1
done
echo done > $O/${tag}_done
