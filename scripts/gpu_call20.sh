#!/bin/bash
# round 2, call 20: cross-attention alone, cold vs between tensor-core bursts, 4 vs 8 consumer warps
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; tag=c20
timeout 300 python scripts/exp_xattn_hot.py > $O/${tag}_w4.log 2>&1
GRAM_XATTN_WARPS=8 timeout 300 python scripts/exp_xattn_hot.py > $O/${tag}_w8.log 2>&1
echo done > $O/${tag}_done
